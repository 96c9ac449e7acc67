cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/summaries
make -C oracle -s 2>&1 | tail -3
timeout 200 python scripts/k2_probe.py 2>&1 | tail -4
# launch list of one cfg2 bench step (bounded), summarised on the box
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${LIMIT:-1800} --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --config cfg2_300v --steps 1 --warmup 0 --no-overlap --wave 1024 --fb-wave 2048 --cpu-sample 16 > gpurun_out/bench_under_ncu.json 2>/dev/null
for K in k5_fallback_score:9 k4a_polynomial:2 k4b_roots:2 k4c_solutions:2 k3_decompose:1 k1_score:0 k2_fivept:0; do
  NAME=${K%%:*}; SKIP=${K##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$NAME -s $SKIP -c 1 -o gpurun_out/prof_$NAME python scripts/profile_wave.py 1184 > /dev/null 2>&1
done
python scripts/summarize_profiles.py r02 gpurun_out/summaries 2>&1 | tail -3
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches_bench.csv
ls gpurun_out/summaries
