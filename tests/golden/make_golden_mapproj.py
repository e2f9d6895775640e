#!/usr/bin/env python
"""Golden outputs of the reference's map_projection (photonbend/core/projection.py:550-599) for the
coordinate maps stored in small_maps.npz.  Build container only.  Output: map_projection.npz."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402
from photonbend.core.projection import map_projection  # noqa: E402  (the reference)

maps = np.load(os.path.join(HERE, "small_maps.npz"))
out = {}
for key in maps.files:
    cmap = maps[key].copy()
    if not (cmap[:, :, 2] == 0).any():
        continue
    out[key] = map_projection(cmap)
np.savez_compressed(os.path.join(HERE, "map_projection.npz"), **out)
print(len(out), "map_projection outputs")
