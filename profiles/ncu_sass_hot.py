#!/usr/bin/env python
"""List the hot SASS instructions of every kernel in an .ncu-rep (source page).

    python profiles/ncu_sass_hot.py rep.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one block per captured launch: a kernel-name line, a header line, then one row per instruction
blocks, hdr = [], None
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        blocks.append((hdr, []))
    elif hdr is not None and len(r) == len(hdr):
        blocks[-1][1].append(r)
blocks = [b for k, b in enumerate(blocks) if k == 0 or b[1] != blocks[k - 1][1]]  # (ncu lists each launch twice)
for n_block, (hdr, data) in enumerate(blocks):
    isrc, iws, iie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")

    def num(v):
        try:
            return float(v)
        except ValueError:
            return 0.0

    tot_s = sum(num(r[iws]) for r in data) or 1
    tot_i = sum(num(r[iie]) for r in data) or 1
    print(f"[launch {n_block}] {len(data)} SASS instructions, {tot_i:.0f} warp-instructions executed, {tot_s:.0f} stall samples")
    for n, r in enumerate(data):
        s, i = num(r[iws]) / tot_s * 100, num(r[iie]) / tot_i * 100
        if s >= thr or i >= thr:
            print(f"{n:5d} stall {s:5.1f}%  inst {i:5.1f}%  {r[isrc].strip()[:100]}")
