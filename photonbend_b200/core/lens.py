"""Lens models: the ``Lens`` pair of callables plus the built-in factories, same surface as the
reference's photonbend/core/lens.py (Lens :48-64, factories :341-401).

The host callables below serve two purposes only: deriving the focal distance of an image
(``f = magnitude / forward(fov / 2)``, a scalar) and being handed to users who call them.  The
per-pixel evaluation happens in the CUDA kernel, which recognises the six built-in models by
the identity of these function objects (``lens_id``).  A user-supplied callable (the reference's
Lens wraps any pair of callables, lens.py:48-64) cannot be compiled into the kernel; it is sampled
on the host once per image (``lens_table``) and the kernel interpolates the table (PB_LENS_TABLE) --
the one case whose pixels follow the reference to the interpolation error (~1e-7 px for smooth
functions) instead of bit for bit.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Callable

import numpy as np

from photonbend_b200 import _native
from photonbend_b200.utils import to_radians

_THOBY_K1 = 1.47
_THOBY_K2 = 0.713


@dataclass
class Lens:
    """A lens as a pair of functions.

    Attributes:
        forward_function: incidence angle (radians) -> distance from the projection centre in
            focal units.  Accepts a float or a float64 array.
        reverse_function: the inverse mapping.
    """

    forward_function: Callable
    reverse_function: Callable


def _is_scalar(v) -> bool:
    return isinstance(v, float)  # numpy.float64 included, ints and arrays not


# --- rectilinear: r = tan(theta) -------------------------------------------------------------


def _rectilinear(theta):
    limit = to_radians(89)
    if _is_scalar(theta):
        if theta < 0:
            raise ValueError("The angle theta cannot be negative")
        if theta > limit:
            raise ValueError("The Rectilinear lens can't handle FoV larger than 179 degrees")
        return np.tan(theta)
    theta = np.asarray(theta)
    r = np.tan(theta)
    r[(theta < 0) | (theta > limit)] = np.nan
    return r


def _rectilinear_inverse(r):
    return np.arctan(r)


# --- stereographic: r = 2 tan(theta / 2) -----------------------------------------------------


def _stereographic(theta):
    return 2.0 * np.tan(theta / 2.0)


def _stereographic_inverse(r):
    return 2.0 * np.arctan(r / 2.0)


# --- equidistant: r = theta ------------------------------------------------------------------


def _equidistant(theta):
    return theta


def _equidistant_inverse(r):
    return r


# --- equisolid: r = 2 sin(theta / 2) ---------------------------------------------------------


def _equisolid(theta):
    return 2 * np.sin(theta / 2.0)


def _equisolid_inverse(r):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        theta = 2.0 * np.arcsin(r / 2.0)
    if _is_scalar(theta):
        return 0.0 if np.isnan(theta) else theta
    theta[np.isnan(theta)] = 0.0  # radii beyond the lens' domain collapse onto the optical axis
    return theta


# --- orthographic: r = sin(theta) ------------------------------------------------------------


def _orthographic(theta):
    return np.sin(theta)


def _orthographic_inverse(r):
    return np.arcsin(r)


# --- thoby: r = 1.47 sin(0.713 theta) --------------------------------------------------------


def _thoby(theta):
    return _THOBY_K1 * np.sin(_THOBY_K2 * theta)


def _thoby_inverse(r):
    return np.arcsin(r / _THOBY_K1) / _THOBY_K2


_BUILTIN = {
    _native.LENS_EQUIDISTANT: (_equidistant, _equidistant_inverse),
    _native.LENS_EQUISOLID: (_equisolid, _equisolid_inverse),
    _native.LENS_ORTHOGRAPHIC: (_orthographic, _orthographic_inverse),
    _native.LENS_STEREOGRAPHIC: (_stereographic, _stereographic_inverse),
    _native.LENS_RECTILINEAR: (_rectilinear, _rectilinear_inverse),
    _native.LENS_THOBY: (_thoby, _thoby_inverse),
}


def lens_id(forward_function, reverse_function) -> int:
    """The kernel's id of a lens pair: one of the six built-in models, or LENS_TABLE for
    user-defined callables (sampled with ``lens_table``)."""
    for ident, (fwd, inv) in _BUILTIN.items():
        if forward_function is fwd and reverse_function is inv:
            return ident
    return _native.LENS_TABLE


LENS_TABLE_SAMPLES = 65537


def lens_table(function, x_max: float, samples: int = LENS_TABLE_SAMPLES) -> np.ndarray:
    """float64 samples of a user-defined lens function on [0, x_max] (it is called with an array,
    as the reference calls it: lens.py docstrings, projection.py:189, 251).  With 65537 samples
    the kernel's linear interpolation is within x_max^2 / 2^35 * max|f''| of the function."""
    if not (np.isfinite(x_max) and x_max > 0):
        raise ValueError("lens_table needs a positive, finite range")
    xs = np.linspace(0.0, float(x_max), int(samples))
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ys = np.asarray(function(xs), dtype=np.float64)
    if ys.shape != xs.shape:
        raise ValueError("a lens function must map an array of angles / radii to an array of the same shape")
    return np.ascontiguousarray(ys)


def rectilinear() -> Lens:
    r"""Rectilinear lens: $r = \tan\theta$, $\theta = \arctan r$."""
    return Lens(_rectilinear, _rectilinear_inverse)


def equisolid() -> Lens:
    r"""Equisolid lens: $r = 2\sin(\theta/2)$, $\theta = 2\arcsin(r/2)$."""
    return Lens(_equisolid, _equisolid_inverse)


def equidistant() -> Lens:
    r"""Equidistant lens: $r = \theta$."""
    return Lens(_equidistant, _equidistant_inverse)


def orthographic() -> Lens:
    r"""Orthographic lens: $r = \sin\theta$, $\theta = \arcsin r$."""
    return Lens(_orthographic, _orthographic_inverse)


def stereographic() -> Lens:
    r"""Stereographic lens: $r = 2\tan(\theta/2)$, $\theta = 2\arctan(r/2)$."""
    return Lens(_stereographic, _stereographic_inverse)


def thoby() -> Lens:
    r"""Thoby lens: $r = 1.47\sin(0.713\,\theta)$, $\theta = \arcsin(r/1.47)/0.713$."""
    return Lens(_thoby, _thoby_inverse)


__all__ = [
    "Lens",
    "equisolid",
    "equidistant",
    "rectilinear",
    "stereographic",
    "orthographic",
    "thoby",
]
