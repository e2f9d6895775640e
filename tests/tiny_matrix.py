"""Degenerate / tiny geometries (pure data, like case_matrix.py): one-pixel and one-row images,
odd double widths, panoramas that are not 2:1."""

from __future__ import annotations

import math

import numpy as np


def rad(deg):
    return deg / 180 * math.pi


OUTS = [
    ("eq1x2", {"kind": "equirect", "height": 1, "width": 2}),
    ("eq2x4", {"kind": "equirect", "height": 2, "width": 4}),
    ("eq3x5", {"kind": "equirect", "height": 3, "width": 5}),
    ("eq2x1", {"kind": "equirect", "height": 2, "width": 1}),
    ("cam1x1", {"kind": "camera", "height": 1, "width": 1, "lens": "equidistant", "fov": rad(180), "magnitude": None}),
    ("cam2x3", {"kind": "camera", "height": 2, "width": 3, "lens": "equisolid", "fov": rad(180), "magnitude": 1.0}),
    ("cam5x1", {"kind": "camera", "height": 5, "width": 1, "lens": "stereographic", "fov": rad(120), "magnitude": 2.0}),
    ("cam1x6", {"kind": "camera", "height": 1, "width": 6, "lens": "rectilinear", "fov": rad(100), "magnitude": 2.5}),
    ("dbl2x2", {"kind": "double", "height": 2, "width": 2, "lens": "equidistant", "fov": rad(190)}),
    ("dbl3x5", {"kind": "double", "height": 3, "width": 5, "lens": "equidistant", "fov": rad(200)}),
]

SRCS = [
    ("eq1x2", {"kind": "equirect", "height": 1, "width": 2}),
    ("eq2x4", {"kind": "equirect", "height": 2, "width": 4}),
    ("eq3x5", {"kind": "equirect", "height": 3, "width": 5}),
    ("cam1x1", {"kind": "camera", "height": 1, "width": 1, "lens": "equidistant", "fov": rad(360), "magnitude": None}),
    ("cam3x2", {"kind": "camera", "height": 3, "width": 2, "lens": "equisolid", "fov": rad(180), "magnitude": 1.0}),
    ("cam4x16", {"kind": "camera", "height": 4, "width": 16, "lens": "orthographic", "fov": rad(180), "magnitude": 7.5}),
    ("dbl2x4", {"kind": "double", "height": 2, "width": 4, "lens": "equidistant", "fov": rad(195)}),
    ("dbl3x5", {"kind": "double", "height": 3, "width": 5, "lens": "equidistant", "fov": rad(195)}),
    ("dbl2x2", {"kind": "double", "height": 2, "width": 2, "lens": "equisolid", "fov": rad(190)}),
]

ROTS = [("r0", ()), ("r1", ((0.3, -0.2, 1.0),))]


def all_cases():
    cases = []
    seed = 5000
    for oname, og in OUTS:
        for sname, sg in SRCS:
            for rname, rots in ROTS:
                cases.append((f"{oname}__{sname}__{rname}", og, rots, sg, seed))
                seed += 1
    return cases


def case_image(src_geom, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (src_geom["height"], src_geom["width"], 3), dtype=np.uint8)
