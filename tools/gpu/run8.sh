cd $GRAFT_REPO_ROOT
./tools/variants/ldlt_probe | head -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t8_tests.log; tail -5 gpurun_out/r2_t8_tests.log
for c in pair8 window200 c128; do
timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 30 --sustain-seconds 0.5 > gpurun_out/r2_t8_$c.json 2> gpurun_out/r2_t8_$c.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_t8_$c.json').read().strip().splitlines()[-1])
print('$c value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3)); print('  parity',d['parity']['ok'], d['parity']['max_rel_err_JtJ_vs_f64'], d['parity']['window_buffer_max_rel_err_vs_host_mirror']); print('  e2e',round(d['e2e']['value']), round(d['e2e']['gb_per_s_h2d'],1),'GB/s', 'single', d['single_launch'])"
tail -2 gpurun_out/r2_t8_$c.err
done
