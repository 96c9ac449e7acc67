# round 2, third GPU pass: ncu on K6, new bench.py shake-out on cfg2 with --verify
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_astar.py -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k6_astar -s 120 -c 2 -o gpurun_out/prof_k6 python scripts/astar_bench.py 1000 2048 150000 gpu 2>&1 | tail -3
timeout 1200 python bench.py --config cfg2_300v --steps 1 --warmup 1 --host-search --wave 256 --verify 3000 --dump-tuples gpurun_out/tuples_cfg2_300v.npz > gpurun_out/bench_new_cfg2.json 2> gpurun_out/bench_new_cfg2.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_new_cfg2.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_new_cfg2.json"))
for k in ("value","ms_per_step","e2e","verify","cpu_baseline","roofline","branch_mix","host_s_per_step"):
    print(k, json.dumps(d.get(k))[:900])
P
timeout 600 python bench.py --impl reference --config cfg2_300v --steps 2 --warmup 1 | cut -c1-1200
