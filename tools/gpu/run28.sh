cd $GRAFT_REPO_ROOT
export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_wd.so
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sfm_run_step_matches_oracle and (320 or 640)" 2>&1 | tail -25
timeout 60 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0.1 > gpurun_out/r2_t28.out 2>gpurun_out/r2_t28.err
echo rc $?
grep watchdog gpurun_out/r2_t28.out | head -5
tail -3 gpurun_out/r2_t28.err
unset DFK_LIB
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 10 --sustain-seconds 0.3 2>>gpurun_out/r2_t28.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', 'value', round(d['value']), 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'single us', round(d['single_launch']['ms_per_eval']*1e3,1), 'parity', d['parity'], 'e2e', round(d['e2e']['value']))"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sfm" 2>&1 | tail -8
