#!/usr/bin/env python
"""Group the SASS of the first kernel in an .ncu-rep into runs of equal execution count and
print, per run, instructions executed per warp and the stall samples that fell on it.

    python profiles/ncu_sass_segments.py rep.ncu-rep [n_warps] [--list lo hi]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, iws, iie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
W = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else float(data[0][iie])
if "--list" in sys.argv:
    k = sys.argv.index("--list")
    for n in range(int(sys.argv[k + 1]), int(sys.argv[k + 2])):
        r = data[n]
        print(n, f"{float(r[iie]) / W:6.2f} {r[iws]:>5s}", r[isrc].strip()[:100])
    sys.exit(0)
tot = sum(float(r[iie]) for r in data)
tot_s = sum(float(r[iws]) for r in data)
prev, start, acc, stall = None, 0, 0.0, 0.0
for n, r in enumerate(data):
    ie = float(r[iie]) / W
    if prev is None:
        prev = ie
    if abs(ie - prev) > 0.15 * max(prev, 1):
        print(f"{start:5d}-{n - 1:5d}  n={n - start:4d}  exec/warp {prev:7.2f}  inst/warp {acc:8.1f} ({100 * acc * W / tot:4.1f}%)  stall {100 * stall / tot_s:5.1f}%")
        prev, start, acc, stall = ie, n, 0.0, 0.0
    acc += ie
    stall += float(r[iws])
print(f"{start:5d}-{n:5d}  n={n - start + 1:4d}  exec/warp {prev:7.2f}  inst/warp {acc:8.1f} ({100 * acc * W / tot:4.1f}%)  stall {100 * stall / tot_s:5.1f}%")
print(f"total {tot / W:.1f} warp-instructions per warp, {tot_s:.0f} stall samples")
