set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
cd $GRAFT_REPO_ROOT
make -C oracle -s 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
