#!/usr/bin/env python
"""Offline model of shared-memory layouts for the staged source of a tile (analysis tool).

Compares, on the oracle's source-index map of a workload, the LDS wavefronts per warp-wide gather
instruction (bank conflicts) and the staged bytes of
  rect     bounding rectangle, odd pitch (the TMA box layout of pb_tiled.cuh),
  packed   touched 16-byte chunks packed row-major (pb_chunk.cuh as first written),
  oddrow   packed, every row padded to an odd number of chunks,
  shear    every row starts at its own first touched chunk; uniform odd pitch = widest row.

    python tests/analysis/chunk_sim.py cfg4 --sample 300 [--both]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import c_port  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402
from stage_sim import TH, TW, thread_pixels_quads, wavefronts  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--sample", type=int, default=300)
    ap.add_argument("--both", action="store_true", help="only tiles that see both lenses")
    args = ap.parse_args()
    wl = workloads.WORKLOADS[args.workload]
    idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
    H, W, _ = idx.shape
    sw = wl["src"]["width"]
    nslot = 2 if wl["src"]["kind"] == "double" else 1
    tx, ty = W // TW, H // TH
    rng = np.random.default_rng(0)
    rows, cols = thread_pixels_quads()
    names = ("rect", "packed", "oddrow", "shear")
    tot = {n: [0, 0, 0] for n in names}  # wavefronts, instructions, staged bytes
    n_items = 0
    order = rng.permutation(tx * ty)
    picked = 0
    for t in order:
        y0, x0 = (t // tx) * TH, (t % tx) * TW
        blk = idx[y0:y0 + TH, x0:x0 + TW]
        if args.both and not all((blk[:, :, s] >= 0).any() for s in range(nslot)):
            continue
        picked += 1
        if picked > args.sample:
            break
        for s in range(nslot):
            v = idx[np.minimum(y0 + rows, H - 1), np.minimum(x0 + cols, W - 1), s]  # (256, 8)
            ok = v >= 0
            if not ok.any():
                continue
            n_items += 1
            sy, sb = v // sw, (v % sw) * 3
            by0 = sy[ok].min()
            xb0 = int(sb[ok].min()) & ~15
            r = np.where(ok, sy - by0, 0)
            xb = np.where(ok, sb - xb0, 0)
            nrow = int(r.max()) + 1
            c0, c1 = xb >> 4, (xb + 2) >> 4
            bm = np.zeros((nrow, 40), bool)
            bm[r[ok], c0[ok]] = True
            bm[r[ok], c1[ok]] = True
            cnt = bm.sum(axis=1)
            first = np.where(cnt > 0, bm.argmax(axis=1), 0)
            need = int(sb[ok].max()) + 3 - xb0
            units = max(5, (need + 15) >> 4) | 1
            before = np.cumsum(bm, axis=1) - bm  # chunks of the row before chunk c
            layouts = {}
            layouts["rect"] = (r * units * 16 + xb, ((nrow + 15) // 16 * 16) * units * 16)
            base = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            layouts["packed"] = ((base[r] + before[r, c0]) * 16 + (xb & 15), int(cnt.sum()) * 16)
            cnt_odd = np.where(cnt > 0, cnt | 1, 0)
            base_o = np.concatenate([[0], np.cumsum(cnt_odd)[:-1]])
            layouts["oddrow"] = ((base_o[r] + before[r, c0]) * 16 + (xb & 15), int(cnt_odd.sum()) * 16)
            width = (np.where(cnt > 0, 40 - bm[:, ::-1].argmax(axis=1), 0) - first).max()
            pitch = int(width) | 1
            layouts["shear"] = ((r * pitch + (c0 - first[r])) * 16 + (xb & 15), nrow * pitch * 16)
            for name, (loc, staged) in layouts.items():
                w0 = np.where(ok, loc >> 2, -1)
                hi = ok & ((loc & 3) >= 2)
                w1 = np.where(hi, (loc >> 2) + 1, -1)
                a = w0.reshape(8, 32, 8).transpose(0, 2, 1).reshape(64, 32)
                b = w1.reshape(8, 32, 8).transpose(0, 2, 1).reshape(64, 32)
                tot[name][0] += wavefronts(a).sum() + wavefronts(b).sum()
                tot[name][1] += 128
                tot[name][2] += staged
    print(f"{args.workload}: {picked - 1 if picked > args.sample else picked} tiles, {n_items} (tile, lens) items")
    for name in names:
        w, i, b = tot[name]
        print(f"  {name:7s} wavefronts per LDS {w / i:.2f}   staged bytes per item {b / n_items:.0f}")


if __name__ == "__main__":
    main()
