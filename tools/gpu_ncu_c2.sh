set -x
mkdir -p gpurun_out /tmp/rep
export MOLANN_BENCH_MIN_MS=0
COMMON="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-layers --no-workloads"
cap() {  # name regex workload skip
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o /tmp/rep/$1 python bench.py --workload $3 $COMMON > gpurun_out/r3y_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  ncu -i /tmp/rep/$1.ncu-rep --page raw --csv > gpurun_out/r3y_$1_raw.csv 2>/dev/null
  ncu -i /tmp/rep/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/$1_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/rep/$1_src.csv 40 > gpurun_out/r3y_$1_lines.txt 2>&1
}
cap c2_fwd fused_ws_forward C2 2
cap c2_vg fused_tc_value_grad C2 2
cap c3_blockb preprocess_backward_block C3 1
ls -la gpurun_out/r3y*
