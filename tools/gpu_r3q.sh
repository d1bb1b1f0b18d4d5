set -x
mkdir -p gpurun_out
T=${1:-r3q}
timeout -k 10 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -k 10 600 python tests/cuda/wide_probe.py C3 C5 > gpurun_out/${T}_probe.log 2>&1; echo "probe rc=$?"
for W in C3 C5; do
  timeout -k 10 600 python bench.py --workload $W --no-e2e --no-cpu-baseline --no-layers --no-workloads > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err; echo "bench $W rc=$?"
done
