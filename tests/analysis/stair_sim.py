#!/usr/bin/env python
"""Offline: how many bytes a tile would stage per frame under different box policies (analysis tool).

uniq = distinct source bytes the tile reads; bbox16 = the bounding rectangle in 16-row boxes (what
the kernels stage); stair = per-box extents, with one common width or each box its own, for 16- and
8-row boxes.  The numbers behind the staircase experiment (profiles/experiments/README.md).

    python tests/analysis/stair_sim.py cfg5
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import c_port
from photonbend_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
wl = workloads.WORKLOADS[name]
idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
H, W, NS = idx.shape
SW = wl["src"]["width"]
TW, TH = 32, 64
def units(nb, odd=True):
    u = (nb + 15) >> 4
    u = max(u, 5)
    return u | 1 if odd else u
tot = {}
rng = np.random.default_rng(0)
tiles = [(ty, tx) for ty in range(H // TH) for tx in range(W // TW)]
sel = rng.choice(len(tiles), size=3000, replace=False)
acc = {'one': np.zeros(6), 'rest': np.zeros(6)}
cnt = {'one': 0, 'rest': 0}
for k in sel:
    ty, tx = tiles[k]
    t = idx[ty*TH:(ty+1)*TH, tx*TW:(tx+1)*TW]
    res = np.zeros(6); nvis = 0
    for s in range(NS):
        v = t[..., s]; v = v[v >= 0]
        if v.size == 0: continue
        nvis += 1
        y, x = v // SW, v % SW
        uniq = np.unique(v).size * 3
        # bbox, 16-row boxes
        xb0 = (x.min()*3) & ~15; need = x.max()*3 + 3 - xb0
        nbox = (y.max() - y.min() + 16) // 16
        bbox = nbox * 16 * 16 * units(need)
        out = [uniq, bbox]
        for R in (16, 8):
            nb = (y.max() - y.min() + R) // R
            widths = []
            for b in range(nb):
                m = (y - y.min()) // R == b
                if not m.any(): widths.append(0); continue
                xs = x[m]; x0 = (xs.min()*3) & ~15
                widths.append(xs.max()*3 + 3 - x0)
            common = nb * R * 16 * units(max(widths))
            perbox = sum(R * 16 * units(w) for w in widths if w > 0)
            out += [common, perbox]
        res += np.array(out)
    if nvis == 0: continue
    cls = 'one' if nvis == 1 and NS == 2 else ('rest' if NS == 2 else 'one')
    acc[cls] += res; cnt[cls] += 1
for c in acc:
    if cnt[c]:
        a = acc[c] / cnt[c]
        print(c, cnt[c], 'uniq %.0f bbox16 %.0f | stair16 common %.0f perbox %.0f | stair8 common %.0f perbox %.0f' % tuple(a))
