"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) by CUDA source line, using the line markers
of `nvdisasm --print-line-info` on the same cubin (instructions appear in the same order in both).
usage: ncu_by_line.py source.csv sass_with_lines.txt [kernel_substring] [source_file]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
seen = set()
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    data.append(dict(zip(hdr, r)))
lines = []  # source line per SASS instruction, in order
cur = ("?", 0)
ins_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.+?);")
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if ins_re.match(ln):
        lines.append(cur)
print("ncu instructions", len(data), "nvdisasm instructions", len(lines))
n = min(len(data), len(lines))
stallkeys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = defaultdict(lambda: defaultdict(int))
tot = 0
for i in range(n):
    d = data[i]
    s = int(d["# Samples"] or 0)
    tot += s
    a = agg[lines[i]]
    a["samples"] += s
    a["exec"] += int(d["Instructions Executed"] or 0)
    for k in stallkeys:
        a[k] += int(d[k] or 0)
src = {}
if len(sys.argv) > 4:
    for i, t in enumerate(open(sys.argv[4]), 1):
        src[i] = t.rstrip()
print("total samples", tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    top = sorted(((k, a[k]) for k in stallkeys), key=lambda x: -x[1])[:3]
    text = src.get(key[1], "") if key[0].endswith(sys.argv[4].split("/")[-1]) else "" if len(sys.argv) > 4 else ""
    print(f"{100 * a['samples'] / max(tot, 1):5.1f}%  {key[0]}:{key[1]:<5d} exec={a['exec']:>10d}  {[(k[6:], v) for k, v in top]}  | {text.strip()[:90]}")
