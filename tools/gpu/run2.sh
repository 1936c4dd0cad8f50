set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t2_parity.log
tail -4 gpurun_out/r2_t2_parity.log
for v in "" "--code-sigma 0.5" "--identity-pose" "--fused-depth"; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 3 $v 2>>gpurun_out/r2_t2_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'roofline', d['roofline'], 'single', d.get('single_launch'))
" | tee -a gpurun_out/r2_t2_bench.log
done
DFK_LIB=$PWD/tools/variants/libdfk_timers.so DFK_TC_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | grep "dfk tc dbg" | head -3 | tee gpurun_out/r2_t2_timers.log
tail -3 gpurun_out/r2_t2_bench.err
