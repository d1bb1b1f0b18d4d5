set -x
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_wide.py -x -q --timeout 120 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -30 gpurun_out/r2c_pytest.log
nvidia-smi --query-gpu=name,memory.used --format=csv
