"""The transfer-function formulation of qo_tf.cuh / qo_tf_core.h, restated in numpy and pinned -- without a GPU --
against the 40-digit evaluations of tests/golden/appendix_b.json (an implementation independent of oracle and product).

Each lumped branch is a ratio of real polynomials N(s)/D(s) of degree <= 2; the cascade [P; Q]/D = M1 ... MN [Rl; 1] is
expanded by polynomial products from the load end, and
    S21 = 2 sqrt(Rs Rl) D / (P + Rs Q)        S11 = (P - Rs Q) / (P + Rs Q)
are evaluated at s = j x by Horner on the even / odd coefficients in y = -x^2 -- exactly what the device does per
sample and per point.  This pins the element -> (N, D) table and the algebra; the device twin is checked against the
oracle and against the chain kernels in tests/test_gpu_parity.py.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

SER_R, SHUNT_R, SER_L, SHUNT_L, SER_C, SHUNT_C, SER_LCS, SER_LCP, SHUNT_LCS, SHUNT_LCP = range(1, 11)


def branch(kind, p, wr):
    """(N, D, series) with the coefficient of sn^k carrying wr^k (qo_tf_core.h::qo_tf_element)."""
    p = list(p) + [0.0] * 3
    if kind == SER_R:
        return [p[0], 0, 0], [1, 0, 0], True
    if kind == SHUNT_R:
        return [1, 0, 0], [p[0], 0, 0], False
    if kind in (SER_L, SHUNT_L):          # (L, R, Cp): Z = (R + sL) / (1 + s R Cp + s^2 L Cp)
        z_n, z_d = [p[1], p[0] * wr, 0], [1, p[1] * p[2] * wr, p[0] * p[2] * wr * wr]
        return (z_n, z_d, True) if kind == SER_L else (z_d, z_n, False)
    if kind in (SER_C, SHUNT_C):          # (C, R, Ls): Z = (1 + s R C + s^2 Ls C) / (s C)
        z_n, z_d = [1, p[1] * p[0] * wr, p[2] * p[0] * wr * wr], [0, p[0] * wr, 0]
        return (z_n, z_d, True) if kind == SER_C else (z_d, z_n, False)
    lc = p[0] * p[1] * wr * wr
    if kind == SER_LCS:
        return [1, 0, lc], [0, p[1] * wr, 0], True
    if kind == SER_LCP:
        return [0, p[0] * wr, 0], [1, 0, lc], True
    if kind == SHUNT_LCS:
        return [0, p[1] * wr, 0], [1, 0, lc], False
    if kind == SHUNT_LCP:
        return [1, 0, lc], [0, p[0] * wr, 0], False
    raise ValueError(kind)


def expand(elements, rl, wr):
    p, q, d = np.array([float(rl)]), np.array([1.0]), np.array([1.0])
    for kind, par in reversed(elements):
        n, dd, series = branch(kind, par, wr)
        n, dd = np.array(n, float), np.array(dd, float)
        if series:
            p, q = np.polynomial.polynomial.polyadd(np.convolve(p, dd), np.convolve(q, n)), np.convolve(q, dd)
        else:
            q, p = np.polynomial.polynomial.polyadd(np.convolve(q, dd), np.convolve(p, n)), np.convolve(p, dd)
        d = np.convolve(d, dd)
    return p, q, d


def at_jx(c, x):
    """real-coefficient polynomial at sn = j x: Horner in y = -x^2 on the even and odd coefficients"""
    y = -x * x
    ev, od = c[0::2], c[1::2]
    re = 0.0
    for a in ev[::-1]:
        re = re * y + a
    im = 0.0
    for a in od[::-1]:
        im = im * y + a
    return complex(re, im * x)


@pytest.mark.parametrize("name", ["if_bpf", "gpsdo_10m", "gpsdo_15m", "gpsdo_40m", "gpsdo_60m", "lol_hpf", "cheby11_ideal", "cfg2_nominal"])
def test_polynomial_formulation_vs_40_digit_goldens(name):
    g = json.load(open(os.path.join(GOLDEN, "appendix_b.json")))[name]
    rs, rl = float(g["rs"]), float(g["rl"])
    fs = np.array([r["f"] for r in g["rows"]])
    wr = 2 * np.pi * np.sqrt(fs.min() * fs.max())
    p, q, d = expand(g["elements"], rl, wr)
    worst21 = worst11 = 0.0
    for r in g["rows"]:
        x = 2 * np.pi * r["f"] / wr
        P, Qv, D = at_jx(p, x), at_jx(q, x), at_jx(d, x)
        s21 = 2 * np.sqrt(rs * rl) * D / (P + rs * Qv)
        s11 = (P - rs * Qv) / (P + rs * Qv)
        g21 = complex(float(r["s21"][0]), float(r["s21"][1]))
        g11 = complex(float(r["s11"][0]), float(r["s11"][1]))
        worst21 = max(worst21, abs(s21 - g21) / abs(g21))
        worst11 = max(worst11, abs(s11 - g11) / max(abs(g11), 0.02))
    # the monomial basis loses log10(kappa) digits (kappa ~ 5e3 at the pass-band edge of the 11th-order Chebyshev ladder,
    # more next to the notches of the tank / trap filters): a few 1e-13 ... 7e-12 on these rows.  1e-10 is the tolerance
    # of the plan-time self-check that gates the kernel; north_star asks 1e-9.
    assert worst21 < 1e-10, (name, worst21)
    assert worst11 < 1e-10, (name, worst11)
