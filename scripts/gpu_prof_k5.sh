cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k5_fallback_score -s 9 -c 1 -o gpurun_out/prof_k5b python scripts/profile_wave.py 1184 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
