cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|L2|L3|L1d|Flags" | cut -c1-400 > gpurun_out/n_lscpu.txt
make -C oracle -s 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 300 python scripts/astar_bench.py 1000 256 60000 host 2>&1 | tail -1
timeout 300 python scripts/astar_bench.py 300 256 1000000 host 2>&1 | tail -1
