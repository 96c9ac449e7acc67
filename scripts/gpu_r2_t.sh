cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2; nproc
make -C oracle -s 2>&1 | tail -3
N=2
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$?"; tail -2 gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$name.json'))
    print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'edges',d['edges'], d['config'].get('wave'))
    print(d['host_s_per_step'], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['host_counters'].items() if k in ('waves','astar_runs','sec_astar','floor_retries','stale_spared')}, 'verify', json.dumps(d.get('verify'))[:300])
except Exception as e:
    print('no json', e)
PY
}
run t_cfg2_n2 --config cfg2_300v --steps 2 --warmup 1 --cpu-sample 64 --verify 2000
run t_cfg3_n2 --config cfg3_1000v --steps 1 --warmup 1 --cpu-sample 64
