cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
PGB_HYBRID_SHARE=0 timeout 1700 ncu --set full --clock-control none --import-source on -k regex:k6_astar -s 800 -c 1 -o gpurun_out/prof_k6_cfg3 python bench.py --config cfg3_1000v --steps 1 --warmup 0 --wave 2048 --cpu-sample 64 > gpurun_out/h_ncu.json 2> gpurun_out/h_ncu.err
echo rc=$?; tail -3 gpurun_out/h_ncu.err; ls -la gpurun_out/prof_k6_cfg3.ncu-rep
