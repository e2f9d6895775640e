"""CoordinateMap -- the lazy stand-in for the reference's float64 (H, W, 3) coordinate map.

In the reference a coordinate map is a dense ndarray of (latitude, longitude, invalid)
(photonbend/core/__init__.py:42-49): 24 bytes per output pixel, written by
get_coordinate_map, rewritten by every rotate_coordinate_map, read by process_coordinate_map.
Here ``get_coordinate_map()`` returns a CoordinateMap that only *describes* the rays (output
geometry + the rotation matrices appended so far); ``process_coordinate_map`` turns that
description into ONE fused kernel launch and the map never exists in memory.

Code that looks at the map still works: ``numpy.asarray(m)``, ``m[...]``, ``m[...] = v``,
``m.copy()``, ``m.shape`` / ``dtype`` / ``ndim`` all behave like the ndarray the reference
returns.  The first such access materialises the map on the GPU (pb_materialize_map_f64) and from
then on the object is an ordinary explicit map: rotation and sampling go through the explicit-map
kernels (pb_rotate_map_f64, pb_gather_from_map_u8), so edits are honoured.
"""

from __future__ import annotations

import numpy as np

from photonbend_b200 import engine


class CoordinateMap:
    __array_priority__ = 100

    def __init__(self, rays: engine.RayPlan):
        self._rays = rays
        self._array = None          # explicit ndarray once materialised
        self._zero_invalid = False  # a consumer already zeroed our invalid entries "in place"

    # ------------------------------------------------------------------ lazy side
    @property
    def is_lazy(self) -> bool:
        return self._array is None

    @property
    def rays(self) -> engine.RayPlan:
        if self._array is not None:
            raise RuntimeError("this coordinate map has been materialised; use its array")
        return self._rays

    def _mark_invalid_zeroed(self) -> None:
        """The reference's rotate / panorama-process zero the invalid (lat, lon) of the map they
        are handed, in place (rotation.py:124-125, projection.py:533-536)."""
        if self._array is None:
            self._zero_invalid = True
        else:
            self._array[self._array[:, :, 2] != 0.0, :2] = 0

    # ------------------------------------------------------------------ ndarray side
    def materialize(self) -> np.ndarray:
        if self._array is None:
            dev = engine.materialize_map_device(self._rays)
            arr = dev.cpu().numpy()
            if self._zero_invalid:
                arr[arr[:, :, 2] != 0.0, :2] = 0
            self._array = arr
        return self._array

    @property
    def shape(self):
        return self._rays.shape

    @property
    def dtype(self):
        return np.dtype(np.float64)

    @property
    def ndim(self) -> int:
        return 3

    @property
    def size(self) -> int:
        h, w, c = self.shape
        return h * w * c

    def __len__(self) -> int:
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        arr = self.materialize()
        if dtype is not None and np.dtype(dtype) != arr.dtype:
            return arr.astype(dtype)
        return arr.copy() if copy else arr

    def __getitem__(self, key):
        return self.materialize()[key]

    def __setitem__(self, key, value):
        self.materialize()[key] = value

    def copy(self) -> np.ndarray:
        return self.materialize().copy()

    def __repr__(self) -> str:
        state = "lazy" if self.is_lazy else "materialised"
        return (f"CoordinateMap({state}, shape={self.shape}, "
                f"rotations={len(self._rays.rotations)})")
