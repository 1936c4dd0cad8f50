cd $GRAFT_REPO_ROOT
for n in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 60 > gpurun_out/r2_t10_n$n.json 2> gpurun_out/r2_t10_n$n.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_t10_n$n.json').read().strip().splitlines()[-1])
print('N=$n value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']), round(d['e2e']['gb_per_s_h2d'],1))"
tail -3 gpurun_out/r2_t10_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config ba2k --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 --sustain-seconds 0 > gpurun_out/r2_t10_ba2k_n2.json 2> gpurun_out/r2_t10_ba2k_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_t10_ba2k_n2.json').read().strip().splitlines()[-1])
print('ba2k N=2 value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4),'frac',round(d['roofline']['frac'],3), 'parity', d['parity']['ok'], d['config']['window'])"
tail -3 gpurun_out/r2_t10_ba2k_n2.err
