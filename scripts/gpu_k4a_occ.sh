cd $GRAFT_REPO_ROOT
for v in 0 61440 102400; do
  echo "== PGI_K4A_SMEM=$v"
  PGI_K4A_SMEM=$v python scripts/profile_wave.py 1770 2>&1 | grep -v Warn | tail -3 | grep -o "ms_fallback_solve.: [0-9.]*\|verdict sha.*"
done
