"""Angle helpers and output-size helper -- same surface as the reference's
photonbend/utils/__init__.py (to_radians :27-37, to_degrees :40-50,
calculate_size_panorama_to_photo :53-118).  Host-side scalar math only."""

from __future__ import annotations

import math
from typing import Callable, Tuple

import numpy as np


def to_radians(degrees: float) -> float:
    """Degrees -> radians, evaluated as ``degrees / 180 * pi`` (that order keeps
    ``to_radians(360) == 2*pi`` exactly, and every fov constant bit-identical to the
    reference's)."""
    return degrees / 180 * np.pi


def to_degrees(radians: float) -> float:
    """Radians -> degrees, evaluated as ``radians / pi * 180``."""
    return radians / np.pi * 180.0


def calculate_size_panorama_to_photo(
    panorama_size: Tuple[int, int],
    lens_function: Callable[[float], float],
    preserve_vertical_resolution: bool = False,
) -> Tuple[int, int]:
    """Side of the inscribed (square) photo that keeps the pixel density of an equirectangular
    panorama of ``panorama_size`` = (width, height) when seen through ``lens_function``.

    Horizontal rule: the panorama spends width/pi pixels per radian at the horizon; the photo's
    diameter is that density times the lens radius ratio r(pi)/r(pi/2).  With
    ``preserve_vertical_resolution`` the larger of the horizontal and vertical estimates wins.
    """
    width, height = panorama_size
    if width != 2 * height:
        raise AssertionError(
            "Equirectangular panoramas should have width and height in a 2:1 proportion"
        )
    ratio = lens_function(np.pi) / lens_function(np.pi / 2)
    side = int(math.ceil(width / np.pi * ratio))
    if preserve_vertical_resolution:
        scale = 1.0 / (1.0 - ratio if ratio > 0.5 else ratio)
        side = max(side, abs(int(math.ceil(height * scale))))
    return (side, side)


__all__ = ["to_radians", "to_degrees", "calculate_size_panorama_to_photo"]
