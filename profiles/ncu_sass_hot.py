#!/usr/bin/env python
"""List the hot SASS instructions of the first kernel in an .ncu-rep (source page).

    python profiles/ncu_sass_hot.py rep.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, iws, iie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot_s = sum(float(r[iws]) for r in data) or 1
tot_i = sum(float(r[iie]) for r in data) or 1
print(f"{len(data)} SASS instructions, {tot_i:.0f} warp-instructions executed, {tot_s:.0f} stall samples")
for n, r in enumerate(data):
    s, i = float(r[iws]) / tot_s * 100, float(r[iie]) / tot_i * 100
    if s >= thr or i >= thr:
        print(f"{n:5d} stall {s:5.1f}%  inst {i:5.1f}%  {r[isrc].strip()[:100]}")
