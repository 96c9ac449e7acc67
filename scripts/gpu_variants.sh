# A/B several prebuilt variants of libpgi.so (variants/libpgi_<tag>.so) on the same box: stage times of one
# 1770-pair fallback wave + path wave, and a checksum of the verdicts (must be identical across variants).
cd $GRAFT_REPO_ROOT
cp pose_graph_initialization_b200/libpgi.so /tmp/libpgi_orig.so
for f in variants/libpgi_*.so; do
  cp $f pose_graph_initialization_b200/libpgi.so
  echo "== $f"
  python scripts/profile_wave.py 1770 2>&1 | grep -v Warn | tail -3 | cut -c1-400
done
cp /tmp/libpgi_orig.so pose_graph_initialization_b200/libpgi.so
