#!/usr/bin/env python
"""Quick kernel-only timing of one or more workloads (CUDA events, device-resident frames).

    python tests/analysis/kbench.py cfg5:16 T:16 cfg1:1 [--steps 30] [--tag text]

One line per workload: ms per launch, output Gpix/s, fraction of the measured HBM roofline.
Used for A/B runs of tuning knobs (environment variables read by pb_plan_create).
"""
from __future__ import annotations

import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("items", nargs="+")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch

    import bench
    import helpers
    from photonbend_b200 import workloads
    from photonbend_b200.batch import remap_batch

    peak, _ = bench.measured_peak_gbs()
    for item in args.items:
        name, _, fr = item.partition(":")
        frames = int(fr or 1)
        if frames == 1:
            # one frame per launch: 20 launches as one CUDA graph over a pool of frames larger than L2
            r = bench.quick_kernel_rate(torch, name, 1, steps=args.steps, warmup=args.warmup)
            print(f"{args.tag:24s} {name:5s} x1   {r['ms_per_launch']:8.4f} ms  {r['value']:8.1f} Gpix/s  "
                  f"frac {r['roofline']['frac']:.3f}  (events around each launch: {r['ms_per_launch_events_around_each']:.4f} ms)",
                  flush=True)
            continue
        wl = workloads.WORKLOADS[name]
        batch = bench.make_device_batch(torch, name, frames, 0)
        source = helpers.product_image(wl["src"], batch)
        cmap = helpers.product_map(wl["out"], wl["rotations"])
        out = remap_batch(source, cmap, batch)
        total_ms, launch_ms = bench.timed_kernel_steps(torch, source, cmap, batch, out, args.steps, args.warmup, None)
        ms = float(np.median(launch_ms))
        px = bench.golden_info(name)["out_pixels"] * frames
        gbs = bench.algorithmic_bytes_per_frame(name) * frames / (ms * 1e-3) / 1e9
        print(f"{args.tag:24s} {name:5s} x{frames:<3d} {ms:8.4f} ms  {px / ms / 1e6:8.1f} Gpix/s  frac {gbs / peak:.3f}", flush=True)
        del batch, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
