"""The BASELINE.json configurations as plain data.

Geometry dicts use the same vocabulary everywhere in this repo (tests, bench, oracle):

    {"kind": "camera",   "height": H, "width": W, "lens": name, "fov": radians, "magnitude": M | None}
    {"kind": "double",   "height": H, "width": W, "lens": name, "fov": sensor_fov_radians}
    {"kind": "equirect", "height": H, "width": W}

``rotations`` are (pitch, yaw, roll) triples in radians, applied in order
(reference scripts/commands/make_pano.py:126-129).  ``seed`` feeds
``numpy.random.default_rng(seed).integers(0, 256, shape, dtype=uint8)``; ``input`` names a
committed image file instead.
"""

from __future__ import annotations

import math
import os

import numpy as np


def to_radians(degrees: float) -> float:
    # same expression as the reference's utils/__init__.py:27-37 (``deg / 180 * pi``)
    return degrees / 180 * math.pi


REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKLOADS = {
    # make-pano --type inscribed --lens equidistant --fov 360 (bundled example photo)
    "cfg1": {
        "title": "make-pano: 3072x3072 equidistant 360 inscribed -> equirect 3072x6144",
        "out": {"kind": "equirect", "height": 3072, "width": 6144},
        "rotations": [],
        "src": {
            "kind": "camera",
            "height": 3072,
            "width": 3072,
            "lens": "equidistant",
            "fov": to_radians(360),
            "magnitude": 3072 / 2 - 0.5,
        },
        "input": "tests/golden/equidistant.jpg",
        "seed": None,
    },
    # alter-photo: 360 equidistant -> 180 equisolid inscribed, rotated
    "cfg2": {
        "title": "alter-photo: 4096x4096 equidistant 360 -> equisolid 180 inscribed, rot (10,20,30) deg",
        "out": {
            "kind": "camera",
            "height": 4096,
            "width": 4096,
            "lens": "equisolid",
            "fov": to_radians(180),
            "magnitude": 4096 / 2 - 0.5,
        },
        "rotations": [(to_radians(10), to_radians(20), to_radians(30))],
        "src": {
            "kind": "camera",
            "height": 4096,
            "width": 4096,
            "lens": "equidistant",
            "fov": to_radians(360),
            "magnitude": 4096 / 2 - 0.5,
        },
        "input": None,
        "seed": 1234,
    },
    # make-photo: equirect 8192x4096 -> rectilinear 140 full frame 7680x4320 (core API only)
    "cfg3": {
        "title": "make-photo: equirect 8192x4096 -> rectilinear 140 full-frame 7680x4320, rot (-90,0,195) deg",
        "out": {
            "kind": "camera",
            "height": 4320,
            "width": 7680,
            "lens": "rectilinear",
            "fov": to_radians(140),
            "magnitude": float(np.sqrt((7680 / 2.0 - 0.5) ** 2 + (4320 / 2.0 - 0.5) ** 2)),
        },
        "rotations": [(to_radians(-90), to_radians(0), to_radians(195))],
        "src": {"kind": "equirect", "height": 4096, "width": 8192},
        "input": None,
        "seed": 1234,
    },
    # make-pano --type double: Gear-360 style side-by-side circles -> equirect
    "cfg4": {
        "title": "make-pano: double 7680x3840 equidistant 195 -> equirect 7680x3840",
        "out": {"kind": "equirect", "height": 3840, "width": 7680},
        "rotations": [],
        "src": {
            "kind": "double",
            "height": 3840,
            "width": 7680,
            "lens": "equidistant",
            "fov": to_radians(195),
        },
        "input": None,
        "seed": 1234,
    },
    # the 8K headline target: cfg1 geometry at 8K
    "T": {
        "title": "8K target: 3840x3840 equidistant 360 inscribed -> equirect 7680x3840",
        "out": {"kind": "equirect", "height": 3840, "width": 7680},
        "rotations": [],
        "src": {
            "kind": "camera",
            "height": 3840,
            "width": 3840,
            "lens": "equidistant",
            "fov": to_radians(360),
            "magnitude": 3840 / 2 - 0.5,
        },
        "input": None,
        "seed": 1234,
    },
}
# cfg5 = a stream of cfg4 frames (1024 frames, sharded over the GPUs of one box)
WORKLOADS["cfg5"] = dict(WORKLOADS["cfg4"], title="video: 1024 frames of cfg4 geometry", frames=1024)


def output_shape(out_geom: dict, channels: int = 3):
    w = out_geom["width"]
    if out_geom["kind"] == "double":
        w = 2 * (w // 2)
    return (out_geom["height"], w, channels)


def source_image(workload: dict, frame: int = 0) -> np.ndarray:
    """The uint8 HWC source of a workload (frame k of a stream uses seed + k)."""
    src = workload["src"]
    if workload.get("input"):
        from PIL import Image

        with Image.open(os.path.join(REPO_ROOT, workload["input"])) as im:
            return np.ascontiguousarray(np.asarray(im))
    rng = np.random.default_rng(workload["seed"] + frame)
    return rng.integers(0, 256, (src["height"], src["width"], 3), dtype=np.uint8)
