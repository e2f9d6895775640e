#!/usr/bin/env python
"""Offline model of the tiled kernel's shared-memory staging (analysis tool, not product code).

Uses the oracle's source-index map of a workload to answer, without GPU time:
  * how many bytes a tile stages per slot under a given box policy (over-fetch vs touched bytes),
  * how many shared-memory wavefronts the gather's LDS.32 pairs cost under a given staged-row
    pitch and thread -> pixel mapping (bank conflicts).

    python tests/analysis/stage_sim.py T --pitch 192 208 --sample 400
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

TW, TH = 32, 64


def wavefronts(words: np.ndarray) -> np.ndarray:
    """words: (n, 32) int word addresses of one warp-wide LDS.32 (-1 = inactive lane).
    Returns (n,) wavefront counts = max over banks of distinct words."""
    n = words.shape[0]
    bank = words & 31
    best = np.zeros(n, dtype=np.int64)
    for b in range(32):
        v = np.where((bank == b) & (words >= 0), words, -1)
        v = np.sort(v, axis=1)
        distinct = (np.diff(v, axis=1) != 0) & (v[:, 1:] >= 0)
        cnt = distinct.sum(axis=1) + (v[:, 0] >= 0)
        best = np.maximum(best, cnt)
    return best


def thread_pixels_quads():
    """current mapping: tid -> (qc = tid & 7, rg = tid >> 3); pixel p = q*4+k at (rg + 32 q, 4 qc + k)."""
    tid = np.arange(256)
    qc, rg = tid & 7, tid >> 3
    rows = np.stack([rg + 32 * (p >> 2) for p in range(8)], axis=1)
    cols = np.stack([4 * qc + (p & 3) for p in range(8)], axis=1)
    return rows, cols  # (256, 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--pitch", type=int, nargs="*", default=[0])
    ap.add_argument("--sample", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    wl = workloads.WORKLOADS[args.workload]
    idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
    H, W, _ = idx.shape
    sw = wl["src"]["width"]
    nslot = 2 if wl["src"]["kind"] == "double" else 1
    tx, ty = (W + TW - 1) // TW, (H + TH - 1) // TH
    rng = np.random.default_rng(args.seed)
    tiles = rng.choice(tx * ty, size=min(args.sample, tx * ty), replace=False)
    rows, cols = thread_pixels_quads()
    touched = staged_tight = 0
    fps = []
    for t in tiles:
        y0, x0 = (t // tx) * TH, (t % tx) * TW
        blk = idx[y0:y0 + TH, x0:x0 + TW]
        for s in range(nslot):
            v = blk[:, :, s]
            ok = v >= 0
            if not ok.any():
                continue
            sy, sx = v[ok] // sw, v[ok] % sw
            touched += np.unique(v[ok]).size * 3
            xb0 = (sx.min() * 3) & ~15
            need = sx.max() * 3 + 3 - xb0
            nrow = sy.max() - sy.min() + 1
            fps.append((t, s, sy.min(), xb0, need, nrow))
            staged_tight += ((need + 15) // 16 * 16) * ((nrow + 15) // 16 * 16)
    fps_a = np.array(fps)
    need_all, nrow_all = fps_a[:, 4], fps_a[:, 5]
    print(f"{args.workload}: {len(tiles)} tiles sampled, {len(fps)} (tile, slot) items")
    print(f"  touched bytes/item {touched / len(fps):.0f}; tight-box staged bytes/item {staged_tight / len(fps):.0f} "
          f"(x{staged_tight / touched:.2f})")
    print(f"  need_bytes  mean {need_all.mean():.0f}  p50 {np.percentile(need_all, 50):.0f}  p99.5 {np.percentile(need_all, 99.5):.0f}  max {need_all.max()}")
    print(f"  rows        mean {nrow_all.mean():.0f}  p50 {np.percentile(nrow_all, 50):.0f}  p99.5 {np.percentile(nrow_all, 99.5):.0f}  max {nrow_all.max()}")
    for pitch in args.pitch:
        if pitch == 0:
            continue
        fixed = sum(pitch * ((r + 15) // 16 * 16) for r in nrow_all if True)
        tot_w = tot_i = 0
        for (t, s, by0, xb0, need, nrow) in fps:
            if need > pitch:
                continue
            y0, x0 = (t // tx) * TH, (t % tx) * TW
            r = np.minimum(y0 + rows, H - 1)
            c = np.minimum(x0 + cols, W - 1)
            v = idx[r, c, s]  # (256, 8)
            loc = (v // sw - by0) * pitch + (v % sw) * 3 - xb0
            w0 = np.where(v >= 0, loc >> 2, -1)
            # 8 warps x 8 pixel slots -> (64, 32)
            wv = w0.reshape(8, 32, 8).transpose(0, 2, 1).reshape(64, 32)
            wf = wavefronts(wv) + wavefronts(np.where(wv >= 0, wv + 1, -1))
            tot_w += wf.sum()
            tot_i += 2 * 64
        print(f"  pitch {pitch:4d}: staged bytes/item {fixed / len(fps):.0f} (x{fixed / touched:.2f} of touched); "
              f"LDS wavefronts per warp-instruction {tot_w / max(tot_i, 1):.2f}")


if __name__ == "__main__":
    main()
