mkdir -p gpurun_out
rm -f gpurun_out/${1}_matrix.txt
for KU in 4 2; do for CD in 2 3 4; do for ST in 2 3 4; do for RG in 4 3; do
  echo -n "KU=$KU CDEPTH=$CD STAGES=$ST RING=$RG : " >> gpurun_out/${1}_matrix.txt
  MOLANN_B200_WIDE_KU=$KU MOLANN_B200_WIDE_CDEPTH=$CD MOLANN_B200_WIDE_STAGES=$ST MOLANN_B200_WIDE_RING=$RG timeout -k 5 60 tests/cuda/fw_bench 2000 8 2>&1 | tail -1 >> gpurun_out/${1}_matrix.txt
done; done; done; done
cat gpurun_out/${1}_matrix.txt
