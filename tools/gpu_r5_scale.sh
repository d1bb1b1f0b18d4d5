set -x
mkdir -p gpurun_out
T=${1:-r5s}
N=${2:-8}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload C4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_C4_n$N.json 2> gpurun_out/${T}_bench_C4_n$N.err; echo "bench C4 rc=$?"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${T}_bench_default_n$N.json 2> gpurun_out/${T}_bench_default_n$N.err; echo "bench default rc=$?"
python - <<PY
import json
for w in ("C4", "default"):
    try:
        d = json.loads(open("gpurun_out/${T}_bench_%s_n$N.json" % w).read().strip().splitlines()[-1])
        print(w, d["metric"], d["value"], d["ms_per_step"], d.get("allreduce_ms"), "e2e", d["e2e"]["value"], d.get("train_path"))
        for k, v in d.get("workloads", {}).items():
            print("  ", k, v.get("value"), (v.get("fwd_dx") or {}).get("value"), v.get("allreduce_ms"))
    except Exception as e:
        print(w, "unreadable", e)
PY
