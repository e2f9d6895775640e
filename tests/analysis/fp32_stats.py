#!/usr/bin/env python
"""Calibration table of the FP32-first tier (csrc/pb_fast32.cuh): pb_debug_fast32_stats per geometry.

    python tests/analysis/fp32_stats.py            # BASELINE configs + the mid-size lens matrix
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import case_matrix  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402


def main():
    rows = []
    for n in ("cfg1", "cfg2", "cfg3", "cfg4", "T"):
        wl = workloads.WORKLOADS[n]
        rows.append((n, wl["out"], wl["rotations"], wl["src"]))
    outs, srcs = T._mid_size_geometries()
    rot = ((0.3, -0.2, 1.0),)
    for og in outs:
        for sg in srcs:
            name = f"{og['kind'][:3]}-{og.get('lens', '')[:6]} <- {sg['kind'][:3]}-{sg.get('lens', '')[:6]}"
            rows.append((name, og, rot, sg))
    worst = 0
    for name, og, rots, sg in rows:
        try:
            st = T.fp32_tier_stats(og, rots, sg)
        except ValueError:
            continue
        worst = max(worst, st["max_ratio_x"], st["max_ratio_y"])
        print(f"{name:34s} ratio x {st['max_ratio_x']:7.2f} y {st['max_ratio_y']:7.2f}  undecided {st['undecided'] / st['pixels']:.4f}  "
              f"wrong {st['wrong']:.0f}  status {st['status_mismatch']:.0f}", flush=True)
    print("largest ratio", worst)


if __name__ == "__main__":
    main()
