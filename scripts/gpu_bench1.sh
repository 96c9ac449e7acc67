set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
timeout 1500 python bench.py --steps 2 --warmup 1 --cpu-sample 64 $BENCH_ARGS > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
tail -5 gpurun_out/bench_first.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_first.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'launches',d['gpu_launches'])
print({k:round(v) for k,v in d['gpu_stage_ms_per_step'].items()}); print(d['host_s_per_step']); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['host_counters'].items() if k in ('waves','astar_runs','sec_astar','sec_visibility')})
print('K5',{k:d['roofline_k5_fallback'][k] for k in ('achieved','peak','frac','traffic','algorithmic_bytes_per_launch')}); print('K1',{k:d['roofline_k1_scoring'][k] for k in ('achieved','peak','frac','gcorr_evals_per_s')}); print(d['cpu_baseline']); print(d['clocks'])
PY
