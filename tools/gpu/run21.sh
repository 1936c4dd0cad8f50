cd $GRAFT_REPO_ROOT
timeout 60 tools/variants/umma_probe_mn 2>&1 | grep "M=64" | cut -c1-1500
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 4 --sustain-seconds 0.2 2>>gpurun_out/r2_t21.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH', 'value', round(d['value']), 'kernel ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'single us', round(d['single_launch']['ms_per_eval']*1e3,1), 'parity', d['parity']['ok'])"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sfm" 2>&1 | tail -3
tail -2 gpurun_out/r2_t21.err
