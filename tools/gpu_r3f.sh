set -x
mkdir -p gpurun_out
timeout -k 5 200 tests/cuda/fw_trace 2000 8 > gpurun_out/r3f_trace.txt 2>&1; echo "trace rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 1 -c 1 -o gpurun_out/r3f_wide tests/cuda/fw_bench 2000 8 > gpurun_out/r3f_wide_ncu.log 2>&1
echo "ncu wide rc=$?"
