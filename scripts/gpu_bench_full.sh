cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
timeout 1700 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'launches',d['gpu_launches'])
print(d['gpu_stage_ms_per_step']); print(d['host_s_per_step']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks']); print(d['e2e'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference.json
