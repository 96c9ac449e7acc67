"""Turn gpurun_out/*.ncu-rep (and launch-list CSVs) into the small text summaries committed under profiles/.
Usage: python scripts/summarize_profiles.py <round-tag>   (reads gpurun_out/, writes profiles/<tag>_*.txt)"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def summarize_rep(path, tag):
    name = os.path.splitext(os.path.basename(path))[0]
    raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    if len(raw) < 3:
        return
    hdr, units = raw[0], raw[1]
    lines = [f"# {name}: ncu --set full --clock-control none (one launch; cold-cache, serialised: compare shares)"]
    for row in raw[2:]:
        lines.append(f"kernel: {row[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                lines.append(f"  {k} = {row[hdr.index(k)]} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg = collections.defaultdict(lambda: [0, 0, 0, ""])
    cur = None
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if len(r) >= 10 and r[0].isdigit():
            try:
                s, ex, tex = int(r[4]), int(r[7]), int(r[8])
            except ValueError:
                continue
            a = agg[(cur, int(r[0]))]
            a[0] += s; a[1] += ex; a[2] += tex; a[3] = r[1][:96]
    tot = sum(v[0] for v in agg.values()) or 1
    totex = sum(v[1] for v in agg.values()) or 1
    lines.append(f"top source lines by warp-stall samples (total {tot} samples, {totex} warp instructions):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        lines.append(f"  {k[0]}:{k[1]:<4d} samples {100*v[0]/tot:5.1f}%  instr {100*v[1]/totex:5.1f}%  "
                     f"threads/instr {v[2]/max(v[1],1):5.1f}  | {v[3]}")
    with open(os.path.join(OUT, f"{tag}_{name}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", f"{tag}_{name}.txt")


def summarize_launches(path, tag):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r[4].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += float(r[-1])
    tot = sum(v[1] for v in agg.values()) or 1.0
    name = os.path.splitext(os.path.basename(path))[0]
    with open(os.path.join(OUT, f"{tag}_{name}_summary.txt"), "w") as f:
        f.write(f"# {name}: ncu --metrics gpu__time_duration.sum --clock-control none ({len(rows)} launches; "
                "per-launch times are cold-cache and serialised: the SHARES are what is comparable)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:40s} launches {v[0]:5d}  total {v[1]/1e6:10.3f} ms  share {100*v[1]/tot:5.1f}%\n")
    print("wrote", f"{tag}_{name}_summary.txt")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if len(sys.argv) > 2:  # summarise on the GPU box into a directory that travels back (gpurun_out/ is capped at 64 MiB)
        OUT = os.path.abspath(sys.argv[2])
    os.makedirs(OUT, exist_ok=True)
    for p in sorted(glob.glob(os.path.join(SRC, "prof_*.ncu-rep"))):
        summarize_rep(p, tag)
    for p in sorted(glob.glob(os.path.join(SRC, "launches_*.csv"))):
        summarize_launches(p, tag)
