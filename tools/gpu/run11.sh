cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t11_tests.log; tail -4 gpurun_out/r2_t11_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']), round(d['e2e']['gb_per_s_h2d'],1), 'single', round(d['single_launch']['ms_per_eval']*1e3,1),'us', 'launches', d['gpu_launches']); print(d['cpu_baseline'])"
tail -2 gpurun_out/r02_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t11_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfm_step_tc -s 4 -c 1 -f -o gpurun_out/prof_r02_sfm_step_tc python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t11_ncu2.err
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench_steps2.csv
