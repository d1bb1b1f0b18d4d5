set -x
mkdir -p gpurun_out
T=${1:-r3g}
timeout -k 10 600 python -m pytest tests/test_gpu_wide.py -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/${T}_trace.txt 2>&1; echo "trace rc=$?"
timeout -k 5 200 tests/cuda/fw_bench 2000 8 > gpurun_out/${T}_bench.txt 2>&1; echo "bench rc=$?"
timeout -k 10 600 python tests/cuda/wide_probe.py C3 > gpurun_out/${T}_probe.log 2>&1; echo "probe rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 1 -c 1 -o gpurun_out/${T}_wide tests/cuda/fw_bench 2000 8 > gpurun_out/${T}_wide_ncu.log 2>&1
echo "ncu wide rc=$?"
