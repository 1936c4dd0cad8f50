set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sfm" 2>&1 | tail -25 > gpurun_out/r2_t1_parity.log
cat gpurun_out/r2_t1_parity.log | tail -8
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_t1_bench.json 2> gpurun_out/r2_t1_bench.err
tail -c 1500 gpurun_out/r2_t1_bench.json; tail -5 gpurun_out/r2_t1_bench.err
