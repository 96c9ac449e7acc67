# wave-size sweep at N ranks (torchrun), cfg2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
N=${N:-2}
for W in ${WAVES:-256 512 1024}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 1 --cpu-sample 16 --wave $W 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N $N wave $W value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step']), {k:round(v,2) for k,v in d['host_s_per_step'].items()}, 'edges', d['edges'])
"
done
