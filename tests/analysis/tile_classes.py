#!/usr/bin/env python
"""Offline census of a workload's output tiles by what they need from the source (analysis tool).

For every 32 x 64 output tile of a double-fisheye workload: which lenses it sees, the bytes of the
bounding rectangle(s) the tiled kernel stages per frame (16-row TMA boxes, rows an odd number of
16-byte units wide), and whether it touches the blend band.  These are the numbers behind the two
launches by tile class ("Two grids for a double-fisheye source").

    python tests/analysis/tile_classes.py cfg5
"""
from __future__ import annotations

import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from oracle import c_port  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402

TW, TH = 32, 64


def units(n_bytes: int) -> int:
    return max((n_bytes + 15) >> 4, 5) | 1


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    wl = workloads.WORKLOADS[name]
    idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
    h, w, _ = idx.shape
    sw = wl["src"]["width"]
    fov = wl["src"]["fov"]
    ref = fov / 2 - np.pi / 2
    lo, hi, safety = np.pi / 2 - ref, np.pi / 2 + ref, 0.5 / 180 * np.pi
    lat = np.linspace(0, np.pi, h)
    in_band = lambda l: (l >= lo) & (l <= hi + safety)  # noqa: E731
    band_rows = in_band(lat) | in_band(np.pi - lat)
    rows = []
    for ty in range((h + TH - 1) // TH):
        for tx in range((w + TW - 1) // TW):
            t = idx[ty * TH:(ty + 1) * TH, tx * TW:(tx + 1) * TW]
            rect = []
            for s in range(2):
                v = t[..., s]
                v = v[v >= 0]
                if v.size == 0:
                    rect.append(0)
                    continue
                y, x = v // sw, v % sw
                xb0 = (x.min() * 3) & ~15
                u = units(int(x.max() * 3 + 3 - xb0))
                rect.append(int((y.max() - y.min() + 16) // 16) * 16 * 16 * u)
            rows.append((rect[0], rect[1], bool(band_rows[ty * TH:(ty + 1) * TH].any())))
    r = np.array(rows)
    both = (r[:, 0] > 0) & (r[:, 1] > 0)
    blend = r[:, 2] > 0
    total = r[:, 0] + r[:, 1]
    print(f"{name}: {len(r)} tiles")
    for label, m in (("one lens, unit weights", ~both & ~blend), ("both lenses, unit weights", both & ~blend),
                     ("blend band", blend), ("nothing visible", total == 0)):
        if m.any():
            print(f"  {label:26s} {int(m.sum()):6d} tiles ({m.mean():6.1%})  staged bytes per frame: mean "
                  f"{total[m].mean():7.0f}  median {np.median(total[m]):7.0f}  max {total[m].max():7d}")


if __name__ == "__main__":
    main()
