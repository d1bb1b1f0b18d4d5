set -x
mkdir -p gpurun_out
TAG=${1:-r5z}
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
( time timeout -k 10 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/${TAG}_bench.err
( time timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py --workload C4 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_C4.json 2> gpurun_out/${TAG}_bench_C4.err; echo "C4 rc=$?"
# launch list of the default command's own kernels (+ the training kernels of the C4 workload entry)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:fused_|_tile_kernel|train_|decode_frames|gemm_tc|preprocess_|narrow_' -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/${TAG}_launches.csv
