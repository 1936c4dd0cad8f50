cd $GRAFT_REPO_ROOT
for K in 100 20; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $K --warmup 5 --no-cpu-baseline --e2e-steps 20 --sustain-seconds 0.6 > gpurun_out/r02_bench_n2_steps$K.json 2> gpurun_out/r02_bench_n2_steps$K.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2_steps$K.json').read().strip().splitlines()[-1])
print('N=2 steps $K value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']))"
done
