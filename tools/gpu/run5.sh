cd $GRAFT_REPO_ROOT
for v in timers noop nogather; do
echo "== $v"
DFK_LIB=$PWD/tools/variants/libdfk_$v.so DFK_TC_DEBUG=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-verify --sustain-seconds 0 > gpurun_out/r2_t5_$v.out 2> gpurun_out/r2_t5_$v.err
grep "dfk tc dbg" gpurun_out/r2_t5_$v.err | grep "blocks=102240" | sed -n 4p
python -c "
import json
d=json.loads(open('gpurun_out/r2_t5_$v.out').read().strip().splitlines()[-1]); print('kernel ms', d['roofline']['avg_launch_ms'])"
done 2>&1 | tee gpurun_out/r2_t5.log
