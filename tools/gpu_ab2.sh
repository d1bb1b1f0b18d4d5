mkdir -p gpurun_out
rm -f gpurun_out/${1}_ab.txt
for v in 32 64 256 1024; do
  echo "== wait ns $v" >> gpurun_out/${1}_ab.txt
  timeout -k 5 100 tests/cuda/ws_trace_$v 2>&1 | grep "^# L=\|^# value_and_grad\|tiles/SM 40" >> gpurun_out/${1}_ab.txt
done
cat gpurun_out/${1}_ab.txt
