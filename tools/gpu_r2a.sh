set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
( time python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2a_bench.err
export MOLANN_BENCH_MIN_MS=0
for W in C3 C5; do
  python bench.py --workload $W --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-layers > gpurun_out/r2a_plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_dram_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-layers > gpurun_out/r2a_ncu_$W.log 2>&1
  echo "ncu $W rc=$?"
done
