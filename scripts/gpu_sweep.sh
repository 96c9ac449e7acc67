# wave-size sweep of the host loop (pairs per speculative wave) on cfg2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
for CFG in ${SWEEP:-"--wave 256" "--wave 512" "--wave 768" "--wave 1024" "--wave 1536"}; do
  timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 16 $CFG 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$CFG', 'ms/step', round(d['ms_per_step']), {k:round(v,2) for k,v in d['host_s_per_step'].items()}, {k:(round(d['host_counters'][k],2) if isinstance(d['host_counters'][k],float) else d['host_counters'][k]) for k in ('waves','astar_runs','sec_astar','sec_commit','sec_visibility')})
"
done
