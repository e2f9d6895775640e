"""Rotation of coordinate maps -- same surface as the reference's photonbend/core/rotation.py
(``Rotation(pitch, yaw, roll)``, ``.rotation_matrix``, ``.rotate_coordinate_map(map)``).

The 3x3 matrix is parameter derivation and stays on the host in float64, composed exactly as the
reference composes it (rotation.py:27-62 with the sign flip of :100) so that the nine doubles
handed to the kernel are bit-identical.  Applying it to the rays happens on the GPU: for a lazy
CoordinateMap the matrix is simply appended to the description and the rotation is evaluated in
registers inside the fused remap kernel; an explicit float64 map goes through pb_rotate_map_f64.
"""

from __future__ import annotations

import numpy as np

from photonbend_b200 import engine
from photonbend_b200.core.coordinate_map import CoordinateMap


def _axis_matrix(entries) -> np.ndarray:
    return np.array(entries).reshape((3, 3))


def _calculate_rotation_matrix(pitch: float, yaw: float, roll: float) -> np.ndarray:
    """pitch (about x) . yaw (about y) . roll (about z), angles in radians.

    The polar axis of a coordinate map is +y (the image centre), README.md:29-35 of the
    reference.  Products are taken left to right with ``@`` like the reference does.
    """
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    cr, sr = np.cos(roll), np.sin(roll)
    about_x = _axis_matrix((1, 0, 0,
                            0, cp, sp,
                            0, -sp, cp))
    about_y = _axis_matrix((cy, 0, -sy,
                            0, 1, 0,
                            sy, 0, cy))
    about_z = _axis_matrix((cr, sr, 0,
                            -sr, cr, 0,
                            0, 0, 1))
    return about_x @ about_y @ about_z


class Rotation:
    """A pitch / yaw / roll rotation (radians) that can be applied to coordinate maps.

    Example:
        coordinate_map = destination.get_coordinate_map()
        coordinate_map = Rotation(np.pi / 2, 0, 0).rotate_coordinate_map(coordinate_map)
        rotated = source.process_coordinate_map(coordinate_map)

    Attributes:
        rotation_matrix: the float64 3x3 matrix applied to the unit ray vectors.
    """

    def __init__(self, pitch: float, yaw: float, roll: float) -> None:
        # the camera turns one way, the rays the other: the map is rotated by the negated angles
        self.rotation_matrix = _calculate_rotation_matrix(-pitch, -yaw, -roll)

    def rotate_coordinate_map(self, coordinate_map):
        """Rotate a coordinate map, returning a new one of the same shape.

        ``coordinate_map`` may be the lazy CoordinateMap of this package (nothing is computed:
        the rotation is fused into the remap kernel), a float64 ndarray (H, W, 3), or a float64
        CUDA tensor of that shape.  Like the reference, the invalid entries of the map passed in
        are zeroed in place.
        """
        if isinstance(coordinate_map, CoordinateMap) and coordinate_map.is_lazy:
            rotated = CoordinateMap(coordinate_map.rays.rotated(self.rotation_matrix))
            coordinate_map._mark_invalid_zeroed()
            return rotated

        if engine.is_torch_tensor(coordinate_map):
            if not coordinate_map.is_cuda:
                raise ValueError("a torch coordinate map must live on the CUDA device")
            return engine.rotate_map_device(coordinate_map, self.rotation_matrix)

        host = coordinate_map.materialize() if isinstance(coordinate_map, CoordinateMap) \
            else coordinate_map
        if not isinstance(host, np.ndarray) or host.ndim != 3 or host.shape[2] != 3:
            raise ValueError("coordinate map must be a float64 array of shape (H, W, 3)")
        torch = engine._torch()
        dev = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64)).cuda()
        rotated = engine.rotate_map_device(dev, self.rotation_matrix).cpu().numpy()
        # side effect the reference has on its argument (rotation.py:119-125)
        host[host[:, :, 2] != 0.0, :2] = 0
        return rotated
