set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; free -g | head -2
cd $GRAFT_REPO_ROOT
make -C oracle -s 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
