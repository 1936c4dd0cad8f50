cd $GRAFT_REPO_ROOT
run() { # N reserve
python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu-baseline --no-verify --e2e-steps 4 --sustain-seconds 0.5 > gpurun_out/r2_t35_n$1.json 2> gpurun_out/r2_t35_n$1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_t35_n$1.json').read().strip().splitlines()[-1])
print('N=$1 value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['avg_launch_ms'],4), 'sus', round(d['sustained']['value']), 'sus ms', round(d['sustained']['ms_per_step'],4))"
tail -2 gpurun_out/r2_t35_n$1.err | grep -v "^\*\|OMP_NUM"
}
run 2
run 2
