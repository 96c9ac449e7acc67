# Produces the raw material of profiles/: launch list of the bench command + full captures of the top kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
make -C oracle -s 2>&1 | tail -3
# 1. every launch of one bench step with its device time (serialised, cold-cache: compare SHARES)
if [ -z "$SKIP_LAUNCHES" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --no-overlap --wave 1024 --fb-wave 2048 --cpu-sample 16 > gpurun_out/bench_under_ncu.json 2>/dev/null
fi
# 2. full captures (one launch each) on a 1184-pair slice of the same shape (2000 rows/pair)
# (profile_wave.py runs two fallback waves, then one path wave: launch #2 of K1/K2 is the path wave's)
for K in ${KLIST:-k5_fallback_score:9 k4a_polynomial:2 k4b_roots:2 k4c_solutions:2 k3_decompose:1 k1_score:2 k2_fivept:2}; do
  NAME=${K%%:*}; SKIP=${K##*:}
  ncu --set full --clock-control none --import-source on -k regex:$NAME -s $SKIP -c 1 -o gpurun_out/prof_$NAME python scripts/profile_wave.py 1184 > /dev/null 2>&1
done
# gpurun_out/ is capped at 64 MiB: summarise here, bring back the text only
python scripts/summarize_profiles.py ${TAG:-r01} gpurun_out/summaries
rm -f gpurun_out/prof_*.ncu-rep
ls -la gpurun_out/ gpurun_out/summaries
