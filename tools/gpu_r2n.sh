set -x
mkdir -p gpurun_out
timeout -k 10 420 python -m pytest tests/test_gpu_jacobian.py -q --timeout 120 --timeout-method thread > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -25 gpurun_out/r2n_pytest.log
