#!/usr/bin/env python
"""Offline: L2->SM fetch cost (128-byte lines, bytes) of staging tile footprints, per tile shape.

    python tests/analysis/shape_sim.py cfg5 --sample 200
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from oracle import c_port  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--sample", type=int, default=200)
    ap.add_argument("--box-rows", type=int, default=16)
    args = ap.parse_args()
    wl = workloads.WORKLOADS[args.workload]
    idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
    H, W, _ = idx.shape
    sw = wl["src"]["width"]
    nslot = 2 if wl["src"]["kind"] == "double" else 1
    rng = np.random.default_rng(0)
    br = args.box_rows
    print(f"{args.workload}: lines = 128-byte lines fetched per 1000 output px; bytes = staged bytes per output px")
    for (tw, th) in [(32, 64), (32, 32), (64, 32), (64, 64), (32, 128), (128, 32), (64, 128), (128, 64), (128, 128), (256, 64), (256,128)]:
        tx, ty = (W + tw - 1) // tw, (H + th - 1) // th
        tiles = rng.choice(tx * ty, size=min(args.sample, tx * ty), replace=False)
        lines = bytes_ = px = 0
        touched = 0
        stage_sizes = []
        for t in tiles:
            y0, x0 = (t // tx) * th, (t % tx) * tw
            blk = idx[y0:y0 + th, x0:x0 + tw]
            px += blk.shape[0] * blk.shape[1]
            for s in range(nslot):
                v = blk[:, :, s]
                ok = v >= 0
                if not ok.any():
                    continue
                sy, sx = v[ok] // sw, v[ok] % sw
                touched += np.unique(v[ok]).size * 3
                xb0 = int(sx.min() * 3) & ~15
                need = (int(sx.max()) * 3 + 3 - xb0 + 15) // 16 * 16
                nrow = (int(sy.max() - sy.min()) + br) // br * br
                # absolute byte position of the row start inside the source row decides the lines
                nl = (xb0 + need - 1) // 128 - xb0 // 128 + 1
                lines += nl * nrow
                bytes_ += need * nrow
                stage_sizes.append(need * nrow)
        ss = np.array(stage_sizes)
        print(f"  tile {tw:3d}x{th:<3d}: lines {1000 * lines / px:7.1f}  bytes {bytes_ / px:6.2f}  touched {touched / px:5.2f}  "
              f"stage KB mean {ss.mean() / 1024:6.1f} p99 {np.percentile(ss, 99) / 1024:6.1f} max {ss.max() / 1024:6.1f}")


if __name__ == "__main__":
    main()
