cd $GRAFT_REPO_ROOT
bash scripts/gpu_parity.sh 2>&1 | grep -v "^+" | tail -3
BENCH_ARGS="--no-overlap" bash scripts/gpu_bench1.sh 2>&1 | grep -v "^+" | tail -8
