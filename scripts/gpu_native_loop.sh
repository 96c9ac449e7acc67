cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
timeout 120 python -m pytest tests/test_gpu_scene.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python bench.py --steps 2 --warmup 1 --cpu-sample 16 > gpurun_out/bench_native.json 2>gpurun_out/bench_native.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_native.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'edges',d['edges'],d['host_s_per_step'])
PY
tail -3 gpurun_out/bench_native.err
