#!/usr/bin/env python
"""How many bytes ANY staging scheme has to read from DRAM for a workload (analysis tool, host only).

The roofline's algorithmic source bytes count every referenced source PIXEL once (3 bytes).  DRAM
moves 32-byte sectors: a sector is fetched whole if one of its pixels is referenced.  Where the
output under-samples the source (the rim of a fisheye circle in BASELINE config 4/5: 1.44 x 0.92
source pixels per output pixel) a quarter of the pixels inside a tile's footprint are referenced by
nobody, yet their sectors are read.  This prints, per tile class of a double-fisheye source and for
the whole frame, referenced pixel bytes against distinct 32-byte and 64-byte sectors.

    python tests/analysis/sector_floor.py [cfg4]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from oracle import c_port  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402

TH, TW = 64, 32


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    wl = workloads.WORKLOADS[name]
    idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
    H, W = idx.shape[:2]
    sw = wl["src"]["width"]
    pitch = sw * 3
    ty, tx = H // TH, W // TW
    vis = [(idx[: ty * TH, : tx * TW, s] >= 0).reshape(ty, TH, tx, TW).any(axis=(1, 3)) for s in range(idx.shape[2])]
    if wl["src"]["kind"] == "double":
        lat = np.linspace(0, np.pi, H)
        ref = wl["src"]["fov"] / 2 - np.pi / 2
        lo, hi = np.pi / 2 - ref, np.pi / 2 + ref + np.deg2rad(0.5)
        band = ((lat >= lo) & (lat <= hi)) | ((np.pi - lat >= lo) & (np.pi - lat <= hi))
        band_t = band[: ty * TH].reshape(ty, TH).any(axis=1)
        cls2 = (vis[0] & vis[1]) | ((vis[0] ^ vis[1]) & band_t[:, None])
        masks = {"class 2 (both lenses / blend band)": cls2, "class 1 (one lens)": ~cls2 & (vis[0] | vis[1])}
    else:
        masks = {}
    masks["whole frame"] = np.ones((ty, tx), bool)
    for label, tiles in masks.items():
        px_mask = np.repeat(np.repeat(tiles, TH, axis=0), TW, axis=1)
        bytes_ = []
        for s in range(idx.shape[2]):
            o = idx[: ty * TH, : tx * TW, s][px_mask & (idx[: ty * TH, : tx * TW, s] >= 0)]
            bytes_.append((o // sw) * pitch + (o % sw) * 3)
        b = np.concatenate(bytes_)
        touched = np.unique(b).size * 3
        s32 = np.unique(np.concatenate([b >> 5, (b + 2) >> 5])).size * 32
        s64 = np.unique(np.concatenate([b >> 6, (b + 2) >> 6])).size * 64
        print(f"{label:36s} {int(tiles.sum()):6d} tiles  referenced pixels {touched / 1e6:6.1f} MB  "
              f"32-byte sectors {s32 / 1e6:6.1f} MB ({s32 / touched:.2f}x)  64-byte {s64 / 1e6:6.1f} MB ({s64 / touched:.2f}x)")


if __name__ == "__main__":
    main()
