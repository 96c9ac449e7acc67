# A/B of the split K4 (K4a polynomial / K4b lane-refill root solve / K4c solutions) against the one-kernel K4
cd $GRAFT_REPO_ROOT
for v in 0 1; do
  echo "== PGI_K4_SPLIT=$v"
  PGI_K4_SPLIT=$v python scripts/profile_wave.py 1770 2>&1 | grep -v Warn | tail -3 | cut -c1-420
done
