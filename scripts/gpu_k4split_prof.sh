cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PGI_K4_SPLIT=1 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k4 -c 24 --csv --log-file gpurun_out/k4split.csv python scripts/profile_wave.py 1184 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/k4split.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    agg[r[ki].split('(')[0]][r[mi]].append(float(r[vi].replace(',','')))
for k,m in agg.items():
    print(k, {n:(round(sum(v)/len(v),2), len(v)) for n,v in m.items()})
PY
