#!/usr/bin/env python
"""Generate tests/golden/* from the LIVE reference (/root/reference, read-only).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # small matrix + full-size configs
    python tests/golden/make_golden.py --small    # small matrix only

The reference is imported unmodified and driven through its own three-call protocol
(photonbend/core/__init__.py:66-92):

    dst.get_coordinate_map() -> Rotation(p, y, r).rotate_coordinate_map(map)* -> src.process_coordinate_map(map)

Outputs (all committed):
  small_cases.json     sha256 of the reference's output image and of its final float64
                       coordinate map for every case of tests/case_matrix.py
  small_outputs.npz    full reference output images for a subset of those cases
  small_maps.npz       full float64 coordinate maps for a few (output geometry, rotation) pairs
  full_configs.json    for BASELINE configs 1-4 and the 8K target T at FULL size: sha256 of the
                       reference output, shape, and N_touched (distinct source pixels referenced,
                       the roofline's algorithmic read bytes / channels)
  full_rows_<cfg>.npz  every 131st output row (plus the last) of those reference outputs
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"

if not os.path.isdir(os.path.join(REFERENCE, "photonbend")):
    sys.exit("make_golden.py needs the reference checkout at /root/reference")

sys.path.insert(0, REFERENCE)
sys.path.insert(1, REPO)
sys.path.insert(2, os.path.join(REPO, "tests"))
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402

from photonbend.core import lens as ref_lens  # noqa: E402  (the reference)
from photonbend.core.projection import (  # noqa: E402
    CameraImage,
    DoubleCameraImage,
    PanoramaImage,
)
from photonbend.core.rotation import Rotation  # noqa: E402

import case_matrix  # noqa: E402
from photonbend_b200 import workloads  # noqa: E402  (plain data only)

ROW_STRIDE = 131


def reference_object(geom, array):
    if geom["kind"] == "equirect":
        return PanoramaImage(array)
    lens = getattr(ref_lens, geom["lens"])()
    if geom["kind"] == "camera":
        return CameraImage(array, geom["fov"], lens, magnitude=geom.get("magnitude"))
    return DoubleCameraImage(array, geom["fov"], lens)


def run_reference(out_geom, rotations, src_geom, image):
    """-> (final coordinate map as handed to process_coordinate_map, output image)"""
    dst = reference_object(out_geom, np.zeros((out_geom["height"], out_geom["width"], 3), np.uint8))
    cmap = dst.get_coordinate_map()
    for pyr in rotations:
        cmap = Rotation(*pyr).rotate_coordinate_map(cmap)
    final_map = cmap.copy()
    out = reference_object(src_geom, image).process_coordinate_map(cmap)
    return final_map, np.ascontiguousarray(out)


def sha(arr) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def small_matrix():
    cases = case_matrix.all_cases()
    meta, outputs, maps = {}, {}, {}
    seen_maps = set()
    for index, (cid, og, rots, sg, seed) in enumerate(cases):
        image = case_matrix.case_image(sg, seed)
        try:
            cmap, out = run_reference(og, rots, sg, image)
        except ValueError as exc:  # rectilinear fov > 178 etc.: the reference refuses at construction
            meta[cid] = {"raises": "ValueError", "message": str(exc)}
            continue
        meta[cid] = {"out_sha256": sha(out), "map_sha256": sha(cmap), "shape": list(out.shape)}
        if case_matrix.stores_full_output(index, cid):
            outputs[cid] = out
        map_key = cid.split("__")[0] + "__" + cid.split("__")[2]
        if map_key not in seen_maps and (len(seen_maps) % 4 == 0 or "odd" in map_key):
            maps[map_key] = cmap
        seen_maps.add(map_key)
    # a few extra channel layouts through the same protocol (grey HxW and RGBA)
    for cid, og, rots, sg, seed in cases[:: max(1, len(cases) // 24)]:
        if sg["kind"] == "double":
            continue  # the reference's double blend needs a channel axis
        for channels, tag in ((0, "grey"), (4, "rgba")):
            image = case_matrix.case_image(sg, seed, channels)
            try:
                _, out = run_reference(og, rots, sg, image)
            except ValueError:
                continue
            key = f"{cid}__{tag}"
            meta[key] = {"out_sha256": sha(out), "shape": list(out.shape), "channels": channels}
            outputs[key] = out
    with open(os.path.join(HERE, "small_cases.json"), "w") as fh:
        json.dump(meta, fh, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_outputs.npz"), **outputs)
    np.savez_compressed(os.path.join(HERE, "small_maps.npz"), **maps)
    print(f"small matrix: {len(meta)} cases, {len(outputs)} stored outputs, {len(maps)} stored maps")


def full_configs():
    from oracle import c_port  # only to count the distinct source pixels referenced

    info = {}
    for name in ("cfg1", "cfg2", "cfg3", "cfg4", "T"):
        wl = workloads.WORKLOADS[name]
        image = workloads.source_image(wl)
        t0 = time.time()
        _, out = run_reference(wl["out"], wl["rotations"], wl["src"], image)
        dt = time.time() - t0
        idx = c_port.source_index(wl["out"], wl["rotations"], wl["src"])
        touched = np.unique(idx[idx >= 0]).size
        del idx
        rows = sorted(set(range(0, out.shape[0], ROW_STRIDE)) | {out.shape[0] - 1})
        np.savez_compressed(os.path.join(HERE, f"full_rows_{name}.npz"),
                            rows=np.array(rows), pixels=out[rows])
        info[name] = {
            "title": wl["title"],
            "out_sha256": sha(out),
            "src_sha256": sha(image),
            "shape": list(out.shape),
            "out_pixels": int(out.shape[0] * out.shape[1]),
            "n_touched": int(touched),
            "src_pixels": int(image.shape[0] * image.shape[1]),
            "reference_seconds_1core_build_container": round(dt, 2),
        }
        print(name, info[name])
    with open(os.path.join(HERE, "full_configs.json"), "w") as fh:
        json.dump(info, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    if not args.full:
        small_matrix()
    if not args.small:
        full_configs()
