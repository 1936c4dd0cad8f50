cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t4_parity.log
tail -4 gpurun_out/r2_t4_parity.log
for v in "" "--code-sigma 0.5" "--identity-pose"; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 3 --sustain-seconds 0.3 $v 2>>gpurun_out/r2_t4_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'single', round(d['single_launch']['ms_per_eval'],4), 'parity', d.get('parity'), 'sus', d.get('sustained'))
" | tee -a gpurun_out/r2_t4_bench.log
done
for v in timers noop; do
DFK_LIB=$PWD/tools/variants/libdfk_$v.so DFK_TC_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-verify --sustain-seconds 0 2>&1 | grep "dfk tc dbg" | sed -n 3p | tee -a gpurun_out/r2_t4_timers.log
done
tail -3 gpurun_out/r2_t4_bench.err
