set -x
mkdir -p gpurun_out
export MOLANN_BENCH_MIN_MS=0
COMMON="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-layers"
for W in C3 C5 C2; do
  python bench.py --workload $W $COMMON --no-workloads > gpurun_out/r2t_plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_dram_$W.csv python bench.py --workload $W $COMMON --no-workloads > gpurun_out/r2t_ncu_$W.log 2>&1
  echo "ncu $W rc=$?"
done
python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_plain_C4.log 2>&1 &&
MOLANN_BENCH_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_launches_C4.csv python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_ncu_C4.log 2>&1
echo "ncu C4 rc=$?"
timeout -k 5 120 tests/cuda/fw_bench 2000 > gpurun_out/r2t_fw_plain.log 2>&1 &&
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fused_wide -s 1 -c 1 -o gpurun_out/r2t_wide tests/cuda/fw_bench 2000 > gpurun_out/r2t_wide_ncu.log 2>&1
echo "ncu wide rc=$?"
ls -la gpurun_out | tail -12
