cd $GRAFT_REPO_ROOT
./tools/variants/ldlt_probe
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_t6_parity.log
tail -6 gpurun_out/r2_t6_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 6 --e2e-steps 5 > gpurun_out/r2_t6_bench.json 2> gpurun_out/r2_t6_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_t6_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'kernel ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac']); print('parity',d['parity']); print('sus',d['sustained']); print('cpu',d['cpu_baseline']); print('e2e',d['e2e']['value'], 'clocks', d['clocks'])"
tail -3 gpurun_out/r2_t6_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-1200
