"""make_complex -- API-compatible helper (reference photonbend/core/_shared.py:25-55).

The device path never builds complex arrays (the reference only ever feeds them to
``np.log(...).imag``, i.e. atan2, which the kernel evaluates directly); this host helper exists
for library users who import it."""

from __future__ import annotations

import numpy as np


def make_complex(x, y, sparse: bool = True):
    """Broadcast two float64 arrays (possibly a sparse mesh pair) against each other and return
    them as one complex128 array ``x + 1j*y``."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    re = x + y * 0
    im = y + x * 0
    out = np.empty(re.shape, dtype=np.complex128)
    out.real = re
    out.imag = im
    return out
