cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 60 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4), 'sus', round(d['sustained']['value']), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']), 'allreduce', d.get('allreduce'))"
tail -2 gpurun_out/r02_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config ba2k --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 --sustain-seconds 0 > gpurun_out/r02_bench_ba2k_n$N.json 2> gpurun_out/r02_bench_ba2k_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_ba2k_n$N.json').read().strip().splitlines()[-1])
print('ba2k N=$N value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'kernel ms',round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity']['ok'], 'e2e', round(d['e2e']['value']))"
tail -2 gpurun_out/r02_bench_ba2k_n$N.err
