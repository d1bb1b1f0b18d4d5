set -x
mkdir -p gpurun_out
timeout -k 5 120 tests/cuda/fw_bench 2000 > gpurun_out/r2i_plain.log 2>&1 &&
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 1 -c 1 -o gpurun_out/r2i_wide tests/cuda/fw_bench 2000 > gpurun_out/r2i_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2i_ncu.log; ls -la gpurun_out/
