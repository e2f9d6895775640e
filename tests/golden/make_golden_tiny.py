#!/usr/bin/env python
"""Degenerate and tiny sizes, from the LIVE reference (/root/reference, read-only): one-pixel and
one-row images, odd double widths, non-2:1 panoramas.  Run in the build container only.

    python tests/golden/make_golden_tiny.py   ->  tests/golden/tiny_cases.json, tiny_outputs.npz

Every (output geometry x source geometry x {no rotation, one rotation}) combination of the lists
below is pushed through the reference's three-call protocol; combinations the reference itself
raises on are recorded as {"raises": "<exception type>"} so the product can be held to the same
behaviour.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
if not os.path.isdir(os.path.join(REFERENCE, "photonbend")):
    sys.exit("make_golden_tiny.py needs the reference checkout at /root/reference")
sys.path.insert(0, REFERENCE)
sys.path.insert(1, os.path.join(REPO, "tests"))
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402

import tiny_matrix  # noqa: E402
from make_golden import run_reference  # noqa: E402  (same driver as the main golden set)


def main():
    meta, outputs = {}, {}
    for cid, og, rots, sg, seed in tiny_matrix.all_cases():
        image = tiny_matrix.case_image(sg, seed)
        try:
            _, out = run_reference(og, rots, sg, image)
        except Exception as exc:  # the reference's own behaviour on this input
            meta[cid] = {"raises": type(exc).__name__}
            continue
        meta[cid] = {"shape": list(out.shape)}
        outputs[cid] = out
    with open(os.path.join(HERE, "tiny_cases.json"), "w") as fh:
        json.dump(meta, fh, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "tiny_outputs.npz"), **outputs)
    n_raise = sum(1 for v in meta.values() if "raises" in v)
    print(f"{len(meta)} cases, {n_raise} raise in the reference")


if __name__ == "__main__":
    main()
