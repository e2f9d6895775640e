"""Small-size parity matrix: output format x source format x lens x fov x rotations.

Pure data, shared by tests/golden/make_golden.py (which runs the live reference on every
case) and the parity tests (which run the oracles / the CUDA path on the same cases).
Geometry dicts as in photonbend_b200/workloads.py.  Follows SURVEY.md Appendix A: every
NaN / out-of-domain / out-of-bounds / wrap quirk of the reference is reachable from here.
"""

from __future__ import annotations

import math

import numpy as np

LENSES = ("equidistant", "equisolid", "orthographic", "stereographic", "rectilinear", "thoby")


def rad(deg):
    return deg / 180 * math.pi


def _camera_fovs(lens):
    if lens == "rectilinear":
        return (rad(100), rad(170))
    return (rad(120), rad(180), rad(360))


def output_geometries():
    outs = [("eq", {"kind": "equirect", "height": 32, "width": 64}),
            ("eqodd", {"kind": "equirect", "height": 27, "width": 51})]
    for lens in LENSES:
        for fov in _camera_fovs(lens):
            outs.append((f"cam-{lens}-{round(fov * 180 / math.pi)}",
                         {"kind": "camera", "height": 36, "width": 44, "lens": lens, "fov": fov,
                          "magnitude": 21.5}))
        if lens != "rectilinear":
            outs.append((f"dbl-{lens}",
                         {"kind": "double", "height": 32, "width": 64, "lens": lens, "fov": rad(195)}))
    # default magnitude (height / 2), odd sizes, full-frame magnitude
    outs.append(("cam-default-M", {"kind": "camera", "height": 33, "width": 33, "lens": "equisolid",
                                   "fov": rad(180), "magnitude": None}))
    outs.append(("cam-fullframe", {"kind": "camera", "height": 27, "width": 48, "lens": "rectilinear",
                                   "fov": rad(140),
                                   "magnitude": float(np.sqrt(23.5**2 + 13.0**2))}))
    outs.append(("dbl-odd", {"kind": "double", "height": 31, "width": 63, "lens": "equidistant",
                             "fov": rad(200)}))
    return outs


def source_geometries():
    srcs = [("eq", {"kind": "equirect", "height": 40, "width": 80}),
            ("eq48", {"kind": "equirect", "height": 48, "width": 96})]
    for lens in LENSES:
        for fov in _camera_fovs(lens)[-2:]:
            srcs.append((f"cam-{lens}-{round(fov * 180 / math.pi)}",
                         {"kind": "camera", "height": 52, "width": 48, "lens": lens, "fov": fov,
                          "magnitude": 23.5}))
        if lens != "rectilinear":
            srcs.append((f"dbl-{lens}",
                         {"kind": "double", "height": 40, "width": 80, "lens": lens, "fov": rad(190)}))
    srcs.append(("dbl-odd", {"kind": "double", "height": 41, "width": 83, "lens": "equidistant",
                             "fov": rad(195)}))
    srcs.append(("cam-default-M", {"kind": "camera", "height": 45, "width": 45, "lens": "equidistant",
                                   "fov": rad(360), "magnitude": None}))
    return srcs


ROTATION_SETS = (
    ("r0", ()),
    ("r1", ((0.3, -0.2, 1.0),)),
    ("r2", ((rad(-90), rad(0), rad(195)), (0.1, 0.2, 0.3))),
)


def all_cases():
    """[(case_id, out_geom, rotations, src_geom, seed)] -- deterministic order."""
    cases = []
    seed = 0
    for oname, og in output_geometries():
        for sname, sg in source_geometries():
            for rname, rots in ROTATION_SETS:
                cases.append((f"{oname}__{sname}__{rname}", og, rots, sg, 1000 + seed))
                seed += 1
    return cases


def case_image(src_geom, seed, channels=3):
    rng = np.random.default_rng(seed)
    shape = (src_geom["height"], src_geom["width"]) + ((channels,) if channels else ())
    return rng.integers(0, 256, shape, dtype=np.uint8)


# cases whose full reference output (not only its hash) is stored in tests/golden/
def stores_full_output(index: int, case_id: str) -> bool:
    return index % 9 == 0 or "odd" in case_id or "default-M" in case_id or "fullframe" in case_id
