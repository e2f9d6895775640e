#!/usr/bin/env python
"""Coefficients of csrc/pb_fast.cuh::kAtanPoly and their error (analysis tool, host only).

atan(t)/t on t^2 in [0, tan(pi/8)^2] by Chebyshev interpolation of degree 8; the error of
t * P(t^2) against a long-double atan is printed (it must stay below 1e-13: the short cut's
decisions carry a guard band of 2^-19 px, i.e. ~1e-10 rad on a 64K-wide panorama).

    python tests/analysis/atan_fit.py
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

T = np.sqrt(2) - 1
UMAX = T * T * 1.0001
DEGREE = 8


def g(u):
    u = np.asarray(u, dtype=np.longdouble)
    t = np.sqrt(u)
    return np.where(u > 0, np.arctan(t) / np.where(t == 0, 1, t), 1.0)


def main():
    k = np.arange(DEGREE + 1)
    x = np.cos(np.pi * (k + 0.5) / (DEGREE + 1))
    c = C.chebfit(x, g((x + 1) / 2 * UMAX).astype(np.float64), DEGREE)
    coef = P.Polynomial(C.cheb2poly(c))(P.Polynomial([-1, 2 / UMAX])).coef
    tt = np.linspace(1e-9, T, 400001).astype(np.longdouble)
    uu = (tt * tt).astype(np.float64)
    val = np.zeros_like(uu)
    for a in coef[::-1]:
        val = val * uu + a
    err = np.abs((tt.astype(np.float64) * val).astype(np.longdouble) - np.arctan(tt))
    print("coefficients:", ", ".join(repr(float(a)) for a in coef))
    print("max |t P(t^2) - atan t| on [0, tan(pi/8)]:", float(err.max()))


if __name__ == "__main__":
    main()
