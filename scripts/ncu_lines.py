"""Aggregate an ncu source page by CUDA source line.
usage: python scripts/ncu_lines.py <report.ncu-rep> [top_n]
Prints, per (file, line): stall samples, executed warp instructions, dominant stall reasons."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
    a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "stalls": defaultdict(int)})
    a["samples"] += int(d["# Samples"] or 0)
    a["inst"] += int(d["Instructions Executed"] or 0)
    for k, v in stalls.items():
        a["stalls"][k] += v
tot_s = sum(a["samples"] for a in agg.values())
tot_i = sum(a["inst"] for a in agg.values())
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k}:{v}" for k, v in st if v)
    print(f"{100*a['samples']/tot_s:5.1f}% smp {100*a['inst']/tot_i:5.1f}% ins  {f}:{ln:<5d} {sts:40s} | {a['src'][:90]}")
