set -x
mkdir -p gpurun_out
timeout -k 10 420 python -m pytest tests/test_gpu_wide.py -x -q --timeout 60 --timeout-method thread > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -30 gpurun_out/r2c_pytest.log
if grep -q "passed" gpurun_out/r2c_pytest.log && ! grep -q "failed" gpurun_out/r2c_pytest.log; then
  timeout -k 10 300 python tests/cuda/wide_probe.py C3 C5 > gpurun_out/r2c_probe.log 2>&1; echo "probe rc=$?"
  cat gpurun_out/r2c_probe.log
fi
