cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --cpu-sample 64 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'edges',d['edges'])
print(d['gpu_stage_ms_per_step']); print(d['host_s_per_step'])
PY
