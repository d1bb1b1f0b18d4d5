set -x
mkdir -p gpurun_out
T=${1:-r3s}
N=${2:-8}
for W in C3 C5; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --no-cpu-baseline --no-layers --no-workloads > gpurun_out/${T}_bench_${W}_n$N.json 2> gpurun_out/${T}_bench_${W}_n$N.err; echo "bench $W rc=$?"
done
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline --no-workloads > gpurun_out/${T}_bench_C2_n$N.json 2> gpurun_out/${T}_bench_C2_n$N.err; echo "bench C2 rc=$?"
timeout -k 10 600 python -m pytest tests -m gpu -q -k "sharded" --timeout 300 > gpurun_out/${T}_pytest_sharded.log 2>&1; tail -3 gpurun_out/${T}_pytest_sharded.log
