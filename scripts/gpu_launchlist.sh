# ncu launch list of one bench step (the first LIMIT kernel launches of `bench.py --steps 1 --warmup 0 --no-overlap`:
# FP64-peak probes, then the whole resident step), summarised on the box.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/summaries
make -C oracle -s 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c ${LIMIT:-2600} --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --no-overlap --wave 1024 --fb-wave 2048 --cpu-sample 16 > gpurun_out/bench_under_ncu.json 2>/dev/null
python scripts/summarize_profiles.py ${TAG:-r01} gpurun_out/summaries
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_under_ncu.json'))
print({k:round(v) for k,v in d['gpu_stage_ms_per_step'].items()}, d['gpu_launches'])
PY
rm -f gpurun_out/launches_bench.csv
cat gpurun_out/summaries/*launches*
