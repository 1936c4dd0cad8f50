cd $GRAFT_REPO_ROOT
for v in "" s4 s6 nopf; do
if [ -n "$v" ]; then export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_$v.so; else unset DFK_LIB; fi
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 4 --sustain-seconds 0.2 2>>gpurun_out/r2_t20.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'value', round(d['value']), 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'single us', round(d['single_launch']['ms_per_eval']*1e3,1), 'parity', d['parity']['ok'])"
done
unset DFK_LIB
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sfm_step_tc -s 4 -c 1 -f -o gpurun_out/prof_r02_v4_sfm_step_tc python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t20_ncu.err
ls -la gpurun_out/prof_r02_v4_sfm_step_tc.ncu-rep
tail -2 gpurun_out/r2_t20.err
