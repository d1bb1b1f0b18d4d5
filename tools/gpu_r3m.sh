set -x
mkdir -p gpurun_out
T=${1:-r3m}
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
( time timeout -k 10 900 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/${T}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
export MOLANN_BENCH_MIN_MS=0
COMMON="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-layers --no-workloads"
for W in C3 C5; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_dram_$W.csv python bench.py --workload $W $COMMON > gpurun_out/${T}_ncu_$W.log 2>&1
  echo "ncu $W rc=$?"
done
