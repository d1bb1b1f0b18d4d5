set -x
mkdir -p gpurun_out
T=${1:-r3k}
timeout -k 10 600 python -m pytest tests/test_gpu_wide.py -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout -k 5 200 tests/cuda/fw_bench 5000 4 > gpurun_out/${T}_bench_c5.txt 2>&1; echo "bench rc=$?"
timeout -k 5 200 tests/cuda/fw_trace 5000 4 > gpurun_out/${T}_trace_c5.txt 2>&1; echo "trace rc=$?"
timeout -k 10 600 python tests/cuda/wide_probe.py C3 C5 > gpurun_out/${T}_probe.log 2>&1; echo "probe rc=$?"
