cd $GRAFT_REPO_ROOT
bash scripts/gpu_parity.sh 2>&1 | grep -v "^+" | tail -4
bash scripts/gpu_bench_full.sh 2>&1 | grep -v "^+" | tail -12
