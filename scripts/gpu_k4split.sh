# A/B of the split K4 (K4a polynomial / K4b lane-refill root solve / K4c solutions) against the one-kernel K4
cd $GRAFT_REPO_ROOT
for v in ${VARIANTS:-"0 4" "1 4" "1 3" "1 2" "1 1"}; do
  set -- $v
  echo "== PGI_K4_SPLIT=$1 PGI_K4B_CTAS=$2"
  PGI_K4_SPLIT=$1 PGI_K4B_CTAS=$2 python scripts/profile_wave.py 1770 2>&1 | grep -v Warn | tail -3 | cut -c1-420
done
