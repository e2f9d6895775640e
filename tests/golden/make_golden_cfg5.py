#!/usr/bin/env python
"""Golden hashes for BASELINE config 5 (a stream of cfg4-geometry frames) from the LIVE reference.

    python tests/golden/make_golden_cfg5.py [n_frames]

Frame k of the stream is ``default_rng(1234 + k)`` noise (photonbend_b200.workloads.source_image).
Every frame goes through the unmodified reference's three-call protocol at FULL size
(3840x7680 double fisheye -> 3840x7680 equirect, photonbend/core/projection.py:408-462,
487-513); stored per frame: sha256 of the source and of the reference's output.  Build container
only (needs /root/reference); output: tests/golden/cfg5_frames.json (committed).
"""

from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the reference, sets sys.path)

from photonbend_b200 import workloads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    wl = workloads.WORKLOADS["cfg5"]
    frames = []
    for k in range(n):
        image = workloads.source_image(wl, frame=k)
        t0 = time.time()
        _, out = mg.run_reference(wl["out"], wl["rotations"], wl["src"], image)
        frames.append({"frame": k, "seed": wl["seed"] + k, "src_sha256": mg.sha(image), "out_sha256": mg.sha(out),
                       "shape": list(out.shape)})
        print(k, frames[-1]["out_sha256"], f"{time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(HERE, "cfg5_frames.json"), "w") as fh:
        json.dump({"title": wl["title"], "frames": frames}, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
