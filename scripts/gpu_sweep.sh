cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
for CFG in "--wave 1024 --window 100000" "--wave 1024 --window 256" "--wave 1024 --window 128" "--wave 2048 --window 256" "--wave 4096 --window 512"; do
  timeout 600 python bench.py --steps 1 --warmup 1 --cpu-sample 16 $CFG 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$CFG', 'ms/step', round(d['ms_per_step']), 'launches', d['gpu_launches'], {k:round(v) for k,v in d['gpu_stage_ms_per_step'].items()}, {k:round(v,2) for k,v in d['host_s_per_step'].items()}, {k:(round(d['host_counters'][k],2) if isinstance(d['host_counters'][k],float) else d['host_counters'][k]) for k in ('waves','astar_runs','sec_astar','sec_commit','sec_visibility')})
"
done
