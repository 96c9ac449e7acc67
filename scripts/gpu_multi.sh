cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
for N in ${NLIST:-1 2}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 2 --warmup 1 --cpu-sample 64 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --cpu-sample 64 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  fi
  echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'edges',d['edges'], 'launches', d['gpu_launches'])
print({k:round(v) for k,v in d['gpu_stage_ms_per_step'].items()}, d['host_s_per_step'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['host_counters'].items() if k in ('waves','astar_runs','sec_astar')})
PY
done
