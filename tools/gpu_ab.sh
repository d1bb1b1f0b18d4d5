mkdir -p gpurun_out
rm -f gpurun_out/${1}_ab.txt
for v in a c d e f g; do
  echo "== variant $v" >> gpurun_out/${1}_ab.txt
  timeout -k 5 100 tests/cuda/fw_bench_$v 2000 8 2>&1 | tail -2 >> gpurun_out/${1}_ab.txt
  timeout -k 5 100 tests/cuda/fw_bench_$v 5000 4 2>&1 | tail -1 >> gpurun_out/${1}_ab.txt
done
cat gpurun_out/${1}_ab.txt
