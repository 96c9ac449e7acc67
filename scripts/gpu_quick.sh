cd $GRAFT_REPO_ROOT
python scripts/profile_wave.py 2368 2>&1 | tail -2
