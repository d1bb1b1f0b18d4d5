# fused training kernel: parity tests, then the C4 bench line
TAG=${1:-r5x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -x -q -m gpu -k "train or c4_training" 2>&1 | tail -4
python bench.py --workload C4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_train.json 2> gpurun_out/${TAG}_train.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_train.json')); print('C4', d['value'], d['ms_per_step'], d['final_loss'], 'e2e', d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/${TAG}_train.err
