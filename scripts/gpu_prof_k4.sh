cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k4_fallback_solve -s 2 -c 1 -o gpurun_out/prof_k4b python scripts/profile_wave.py 1184 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_decompose -s 1 -c 1 -o gpurun_out/prof_k3b python scripts/profile_wave.py 1184 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
