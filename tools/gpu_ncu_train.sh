set -x
mkdir -p gpurun_out /tmp/rep
TAG=${1:-r5a}
COMMON="--workload C4 --steps 1 --warmup 3 --no-cpu-baseline"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fused_train -s 2 -c 1 -o /tmp/rep/train python bench.py $COMMON > gpurun_out/${TAG}_ncu_train.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/rep/train.ncu-rep --page raw --csv > gpurun_out/${TAG}_train_raw.csv 2>/dev/null
ncu -i /tmp/rep/train.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/train_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/rep/train_src.csv 60 > gpurun_out/${TAG}_train_lines.txt 2>&1
python profiles/summarize_ncu.py gpurun_out/${TAG}_train_raw.csv > gpurun_out/${TAG}_train_summary.md
cat gpurun_out/${TAG}_train_summary.md
