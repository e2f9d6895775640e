"""CPU model of the guarded fixed-point blend (csrc/pb_tiled.cuh: fix_weights / blend_px_fix) against
the reference's float64 expression (projection.py:439-459, restated in oracle/numpy_port.py).

The kernel evaluates, per channel, t = a*ia + b*ib with the row's weights rounded to 24 fractional
bits and takes the top byte of the unsigned word as trunc(a*wa + b*wb) wherever the 24-bit fraction
is more than 256 units away from an integer; everything else goes through the float64 expression.
This test re-states that integer arithmetic in NumPy and checks, for EVERY byte pair and every
blend-band row of several geometries, that a decided channel equals the reference -- i.e. that
the guard is wide enough.  No kernel runs here.
"""

import numpy as np
import pytest

from oracle import numpy_port

SHIFT = 24
GUARD = 257


def fix_weights(wa, wb):
    """-> (ok, ia, ib) as the kernel derives them (round to nearest, ties to even)."""
    ok = (wa >= 0.0) & (wb >= 0.0) & (wa + wb <= 1.00390625)
    with np.errstate(invalid="ignore"):
        ia = np.where(ok, np.rint(wa * float(1 << SHIFT)), 0).astype(np.uint64)
        ib = np.where(ok, np.rint(wb * float(1 << SHIFT)), 0).astype(np.uint64)
    return ok, ia, ib


def band_weights(height, fov_deg):
    """Blend weights (left, right) of every output row of an equirect map, as the reference computes them."""
    fov = numpy_port.deg2rad(fov_deg)
    ref = (fov / 2) - (np.pi / 2)
    lo, hi, span = np.pi / 2 - ref, np.pi / 2 + ref, 2.0 * ref
    safety = numpy_port.deg2rad(0.5)
    lat = np.linspace(0, np.pi, height)
    lat_r = lat * -1 + np.pi

    def weight(l):
        with np.errstate(all="ignore"):
            band = np.logical_and(l >= lo, l <= (hi + safety))
            w = (l - hi) / span * -1
        w[np.logical_not(band)] = 1.0
        return w

    return weight(lat), weight(lat_r)


@pytest.mark.parametrize("height,fov_deg", [(3840, 195), (352, 195), (352, 181), (352, 210), (1080, 250), (352, 180), (352, 170)])
def test_decided_channels_equal_the_float64_expression(height, fov_deg):
    wa, wb = band_weights(height, fov_deg)
    rows = np.nonzero(~((wa == 1.0) & (wb == 1.0)))[0]
    ok, ia, ib = fix_weights(wa[rows], wb[rows])
    # rows that do not qualify (negative weight of the safety strip, NaN / inf of a 180-degree pair)
    # never take the short cut; the others must be exact wherever the guard says "decided"
    a = np.arange(256, dtype=np.uint64).reshape(1, 256, 1)
    b = np.arange(256, dtype=np.uint64).reshape(1, 1, 256)
    n_decided = n_total = 0
    for chunk in np.array_split(np.nonzero(ok)[0], max(1, ok.sum() // 32)):
        if chunk.size == 0:
            continue
        t = a * ia[chunk].reshape(-1, 1, 1) + b * ib[chunk].reshape(-1, 1, 1)
        assert t.max() < (1 << 32)  # the unsigned word never overflows
        frac = t & ((1 << SHIFT) - 1)
        decided = (frac >= GUARD) & (frac <= (1 << SHIFT) - GUARD - 1)
        fast = (t >> SHIFT) & 0xFF
        w0 = wa[rows][chunk].reshape(-1, 1, 1)
        w1 = wb[rows][chunk].reshape(-1, 1, 1)
        exact = (a.astype(np.float64) * w0 + b.astype(np.float64) * w1).astype(np.int64) & 0xFF
        assert np.array_equal(fast[decided], exact[decided].astype(np.uint64))
        n_decided += int(decided.sum())
        n_total += decided.size
    if ok.any():
        # the short cut decides nearly every channel; the undecided ones are exact integers: a == b
        # where the weights add up to 1, black pixels, and whole rows whose weights are (1, 0)
        assert n_decided / n_total > 0.95
    # every qualifying row has both weights in [0, 1 + 2^-8]; a 195-degree pair has ~340 band rows at 8K
    if fov_deg == 195 and height == 3840:
        assert 300 < rows.size < 400 and ok.sum() > 300


def test_rows_that_do_not_qualify():
    wa, wb = band_weights(352, 195)
    neg = (wa < 0) | (wb < 0)
    assert neg.any()  # the half-degree safety strip: one weight is negative there
    ok, _, _ = fix_weights(wa[neg], wb[neg])
    assert not ok.any()
    ok, _, _ = fix_weights(np.array([np.nan, np.inf, 0.5]), np.array([0.5, 0.5, np.nan]))
    assert not ok.any()
