#!/usr/bin/env python
"""Offline comparison of thread -> pixel mappings of the tiled kernels (analysis tool).

For every sampled tile the oracle's source-index map gives the staged byte offset of each pixel
(per-tile odd-unit pitch, as the kernels stage it); a mapping assigns the 2048 pixels of a
32 x 64 tile to (warp, lane, slot).  Reported: shared-memory wavefronts per gather LDS.32 pair
and per output STS.32 (bank conflicts), averaged over the tiles.

    python tests/analysis/map_sim.py T cfg4 --sample 300
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "analysis"))

from oracle import c_port  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402
from stage_sim import wavefronts  # noqa: E402

TW, TH = 32, 64


def mapping(q_per_warp: int):
    """-> rows[256, 2], quads[256, 2]: the two quads (row, quad column) of each thread, threads in
    warp-major order.  A warp covers (32 / q_per_warp) consecutive rows x q_per_warp quads."""
    tid = np.arange(256)
    warp, lane = tid >> 5, tid & 31
    r_per_warp = 32 // q_per_warp
    lr, lq = lane // q_per_warp, lane % q_per_warp
    if q_per_warp == 8:  # current kernels: second quad 32 rows further down
        rows = np.stack([warp * 4 + lr, warp * 4 + lr + 32], axis=1)
        quads = np.stack([lq, lq], axis=1)
    else:  # warps tile 64 rows x 4 quads; second quad 4 quad columns to the right
        warps_down = 64 // r_per_warp
        wr, wq = warp % warps_down, warp // warps_down
        row = wr * r_per_warp + lr
        quad = wq * q_per_warp + lq
        rows = np.stack([row, row], axis=1)
        quads = np.stack([quad, quad + 4], axis=1)
    return rows, quads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--sample", type=int, default=300)
    args = ap.parse_args()
    for name in args.workloads:
        wl = workloads.WORKLOADS[name]
        idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
        H, W, _ = idx.shape
        sw = wl["src"]["width"]
        nslot = 2 if wl["src"]["kind"] == "double" else 1
        tx, ty = (W + TW - 1) // TW, (H + TH - 1) // TH
        rng = np.random.default_rng(0)
        tiles = rng.choice(tx * ty, size=min(args.sample, tx * ty), replace=False)
        for qpw in (8, 4, 2, 1):
            rows, quads = mapping(qpw)
            # output tile stores: 3 words per quad at row * 96 + quad * 12
            st = 0
            for q2 in range(2):
                for w in range(3):
                    words = (rows[:, q2] * 96 + quads[:, q2] * 12 + 4 * w) >> 2
                    st += wavefronts(words.reshape(8, 32)).sum()
            sts = st / (8 * 6)
            tot = n = 0
            for t in tiles:
                y0, x0 = (t // tx) * TH, (t % tx) * TW
                for s in range(nslot):
                    blk = idx[y0:y0 + TH, x0:x0 + TW, s]
                    ok = blk >= 0
                    if not ok.any():
                        continue
                    sy, sx = blk[ok] // sw, blk[ok] % sw
                    xb0 = (sx.min() * 3) & ~15
                    need = sx.max() * 3 + 3 - xb0
                    units = max(5, (need + 15) // 16) | 1
                    if units > 31:
                        continue
                    pitch = 16 * units
                    by0 = sy.min()
                    for q2 in range(2):
                        for k in range(4):
                            r = np.minimum(y0 + rows[:, q2], H - 1)
                            c = np.minimum(x0 + quads[:, q2] * 4 + k, W - 1)
                            v = idx[r, c, s]
                            loc = (v // sw - by0) * pitch + (v % sw) * 3 - xb0
                            w0 = np.where(v >= 0, loc >> 2, -1).reshape(8, 32)
                            tot += wavefronts(w0).sum() + wavefronts(np.where(w0 >= 0, w0 + 1, -1)).sum()
                            n += 16
            print(f"{name}: warp = {32 // qpw:2d} rows x {qpw} quads: {tot / n:.2f} wavefronts per gather LDS, "
                  f"{sts:.2f} per output STS")


if __name__ == "__main__":
    main()
