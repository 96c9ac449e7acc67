set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
timeout 1500 python bench.py --steps 1 --warmup 1 --cpu-sample 64 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
tail -5 gpurun_out/bench_first.err
cat gpurun_out/bench_first.json
