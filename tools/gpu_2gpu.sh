set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -k "sharded" --timeout 200 --timeout-method thread > gpurun_out/r2s_pytest2.log 2>&1; tail -3 gpurun_out/r2s_pytest2.log
( time timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2s_bench2.json 2> gpurun_out/r2s_bench2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r2s_bench2.err
