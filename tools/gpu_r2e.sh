set -x
mkdir -p gpurun_out
timeout -k 5 120 tests/cuda/fw_trace 2000 > gpurun_out/r2m_trace_c3.txt 2>&1
head -22 gpurun_out/r2m_trace_c3.txt; tail -4 gpurun_out/r2m_trace_c3.txt
timeout -k 10 300 python -m pytest tests/test_gpu_wide.py -x -q --timeout 60 --timeout-method thread > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -5 gpurun_out/r2m_pytest.log
timeout -k 10 200 python tests/cuda/wide_probe.py C3 C5 > gpurun_out/r2m_probe.log 2>&1; cat gpurun_out/r2m_probe.log
