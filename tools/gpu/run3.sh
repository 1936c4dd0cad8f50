cd $GRAFT_REPO_ROOT
for v in noop nogather noop8 noopnog; do
  echo "=== $v"
  DFK_LIB=$PWD/tools/variants/libdfk_$v.so DFK_TC_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/r2_t3_$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'kernel ms', round(d['roofline']['avg_launch_ms'],4))
"
  grep "dfk tc dbg" gpurun_out/r2_t3_$v.err | tail -1
done 2>&1 | tee gpurun_out/r2_t3.log
