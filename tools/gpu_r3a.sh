set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3a_pytest.log
tail -5 gpurun_out/r3a_pytest.log
timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/r3a_trace_s6.txt 2>&1; echo "trace rc=$?"
MOLANN_B200_WIDE_SLOTS=8 timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/r3a_trace_s8.txt 2>&1; echo "trace rc=$?"
MOLANN_B200_WIDE_SLOTS=5 MOLANN_B200_WIDE_RING=3 timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/r3a_trace_s5r3.txt 2>&1; echo "trace rc=$?"
