# A/B of the fallback prefetch pipelining: number of background contexts x pairs per prefetch wave
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "1 2048" "2 2048" "2 1024" "2 1536"; do
  set -- $cfg
  timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 16 --fb-streams $1 --fb-wave $2 > gpurun_out/fbs_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/fbs_$1_$2.json'))
h=d['host_s_per_step']
print('streams $1 wave $2: value',round(d['value']),'e2e',round(d['e2e']['value']),'waves_s %.3f engine %.3f host %.3f wait_pf %.3f'%(h['waves_s'],h['engine_s'],h['host_s'],h['wait_prefetch_s']), {k:round(v) for k,v in d['gpu_stage_ms_per_step'].items()}, 'edges', d['edges'])
PY
done
