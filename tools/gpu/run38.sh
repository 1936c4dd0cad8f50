cd $GRAFT_REPO_ROOT
for v in NOGATHER; do
export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_$v.so
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-verify --e2e-steps 4 --sustain-seconds 0.2 2>>gpurun_out/r2_t38.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'kernel ms', round(d['roofline']['avg_launch_ms'],4))"
done
