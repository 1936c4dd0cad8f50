cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc or tensor or c32 or C32 or sfm" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in "" "--fused-depth"; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 20 --sustain-seconds 0.5 $v 2>>gpurun_out/r2_t16.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', '$v', 'value', round(d['value']), 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'single us', round(d['single_launch']['ms_per_eval']*1e3,1), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']))"
done
tail -3 gpurun_out/r2_t16.err
