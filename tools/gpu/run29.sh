cd $GRAFT_REPO_ROOT
for v in NOMMA NODRAIN f16 s7 s11; do
export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_$v.so
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-verify --e2e-steps 4 --sustain-seconds 0.2 2>>gpurun_out/r2_t29.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'single', round(d['single_launch']['ms_per_eval']*1e3,1))"
done
tail -2 gpurun_out/r2_t29.err
