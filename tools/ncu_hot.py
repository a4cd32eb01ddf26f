#!/usr/bin/env python
"""Hot instructions (warp-stall samples) of one kernel in an .ncu-rep:  python tools/ncu_hot.py rep kernel_substr [min]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = [b for b in blocks if pat in b["name"]][0]
hdr = b["rows"][0]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
body = b["rows"][1:]
tot = sum(int(r[si]) for r in body if r[si].isdigit())
print(b["name"], "instructions", len(body), "samples", tot)
named = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
for k, r in enumerate(body):
    s = int(r[si]) if r[si].isdigit() else 0
    if s >= thr:
        why = sorted(((int(r[i]), h[6:]) for i, h in named if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
        print(f"{k:5d} {s:5d} {r[ie]:>8s}  {r[1].strip()[:64]:64s} {why}")
