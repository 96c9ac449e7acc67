cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/summaries
free -g | head -2
make -C oracle -s 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python scripts/k2_probe.py 2>&1 | tail -3
for K in k1_score:2 k2_fivept:2; do
  NAME=${K%%:*}; SKIP=${K##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$NAME -s $SKIP -c 1 -o gpurun_out/prof_$NAME python scripts/profile_wave.py 1184 > /dev/null 2>&1
done
python scripts/summarize_profiles.py r02 gpurun_out/summaries 2>&1 | tail -3
rm -f gpurun_out/prof_*.ncu-rep
run() { # name, env..., -- args
  name=$1; shift
  timeout 900 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$?"; python - <<P
import json
try:
    d=json.load(open("gpurun_out/$name.json"))
    print({k:d.get(k) for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("host_s_per_step"), {k:d["host_counters"].get(k) for k in ("astar_runs","astar_pushes","sec_astar","waves","floor_retries","stale_spared")}, "cpu", d["cpu_baseline"]["value"], "verify", json.dumps(d.get("verify"))[:400])
except Exception as e:
    print("no json", e); print(open("gpurun_out/$name.err").read()[-2000:])
P
}
run s_cfg2 X=1 python bench.py --config cfg2_300v --steps 2 --warmup 1 --verify 4000
run s_cfg3 X=1 python bench.py --config cfg3_1000v --steps 2 --warmup 1 --verify 4000 --verify-replay 20000
