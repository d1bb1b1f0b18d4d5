mkdir -p gpurun_out
rm -f gpurun_out/${1}_ab.txt
for v in 0 1 0 1; do
  echo "== shared rcp $v" >> gpurun_out/${1}_ab.txt
  timeout -k 5 100 tests/cuda/ws_trace_$v 2>&1 | grep "^# L=\|^# value_and_grad\|tiles/SM 40\|SM clock" >> gpurun_out/${1}_ab.txt
done
cat gpurun_out/${1}_ab.txt
