set -x
mkdir -p gpurun_out
T=${1:-r5p}
N=${2:-8}
for P in 1 0; do
  MOLANN_B200_TRAIN_P2P=$P timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$P bench.py --gpus $N --workload C4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_C4_n${N}_p2p$P.json 2> gpurun_out/${T}_bench_C4_n${N}_p2p$P.err; echo "bench C4 p2p=$P rc=$?"
done
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tests/cuda/train_p2p_check.py 2>&1 | grep -E "P2P_OK|Error|assert" | head -5
python - <<PY
import json
for p in (1, 0):
    try:
        d = json.loads(open("gpurun_out/${T}_bench_C4_n${N}_p2p%d.json" % p).read().strip().splitlines()[-1])
        print("p2p=%d" % p, d["value"], d["ms_per_step"], "eager", d["cuda_graph"]["eager_ms_per_step"], "launches", d["gpu_launches"], "allreduce_ms", d.get("allreduce_ms"), "e2e", d["e2e"]["value"], d.get("collective"))
    except Exception as e:
        print(p, "unreadable", e)
PY
