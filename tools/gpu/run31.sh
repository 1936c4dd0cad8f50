cd $GRAFT_REPO_ROOT
export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_NOMMA.so
timeout 100 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-verify --e2e-steps 4 --sustain-seconds 0.2 2>>gpurun_out/r2_t31.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('BENCH NOMMA', 'kernel ms', round(d['roofline']['avg_launch_ms'],4))"
unset DFK_LIB
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t31_tests.log; tail -4 gpurun_out/r2_t31_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']), round(d['e2e']['gb_per_s_h2d'],1), 'single', round(d['single_launch']['ms_per_eval']*1e3,1),'us', 'launches', d['gpu_launches']); print(d['cpu_baseline'])"
tail -2 gpurun_out/r02_bench_n1.err
for c in window200 c128; do
timeout 300 python bench.py --steps 50 --warmup 5 --config $c --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_$c.json').read().strip().splitlines()[-1])
print('$c value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']))"
done
timeout 300 python bench.py --steps 50 --warmup 5 --fused-depth --no-cpu-baseline > gpurun_out/r02_bench_n1_fused_depth.json 2> gpurun_out/r02_bench_fused.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1_fused_depth.json').read().strip().splitlines()[-1])
print('fused value',round(d['value']),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'parity', d['parity']['ok'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t31_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfm_step_tc -s 4 -c 1 -f -o gpurun_out/prof_r02_sfm_step_tc python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t31_ncu2.err
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench_steps2.csv
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sfm_run_step_matches_oracle and 320-240-32 or fused_depth or mixed_sizes" > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/r02_compute_sanitizer_memcheck.log
