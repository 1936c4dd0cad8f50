cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t37_tests.log; tail -12 gpurun_out/r2_t37_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']), round(d['e2e']['gb_per_s_h2d'],1), 'single', round(d['single_launch']['ms_per_eval']*1e3,1),'us', 'launches', d['gpu_launches'], d['roofline']['launches_timed'])"
tail -2 gpurun_out/r02_bench_n1.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 4 --sustain-seconds 0.3 > gpurun_out/r02_bench_n1_steps20.json 2>> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_steps20.json').read().strip().splitlines()[-1])
print('steps20 value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'launches', d['gpu_launches'])"
timeout 200 python bench.py --steps 50 --warmup 5 --code-sigma 0.5 --no-cpu-baseline --e2e-steps 4 --sustain-seconds 0.3 > gpurun_out/r02_bench_n1_code_sigma05.json 2>> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_code_sigma05.json').read().strip().splitlines()[-1])
print('sigma0.5 value',round(d['value']),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'parity', d['parity']['ok'])"
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t37_ncu1.err
wc -l gpurun_out/r02_launches_bench_steps2.csv
