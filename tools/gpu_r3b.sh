set -x
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_wide.py -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3e_pytest.log
tail -5 gpurun_out/r3e_pytest.log
timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/r3e_trace.txt 2>&1; echo "trace rc=$?"
timeout -k 5 200 tests/cuda/fw_bench 2000 8 > gpurun_out/r3e_bench.txt 2>&1; echo "bench rc=$?"
timeout -k 10 600 python tests/cuda/wide_probe.py C3 > gpurun_out/r3e_probe.log 2>&1; echo "probe rc=$?"
