"""CPU model of the kernels' reciprocal / reciprocal-square-root refinements (csrc/pdeq_blockops.cuh: fast_rcp,
fast_rsqrt): a seed with ~20 correct bits (what `rcp.approx.ftz.f64` / `rsqrt.approx.ftz.f64` deliver) followed by ONE
third-order correction, r (1 + e + e^2) and y (1 + e + 3/2 e^2). The operations are replayed here with exactly
rounded fused multiply-adds (rational arithmetic), for seeds at and inside the +-2^-20 error bound, and the results are
compared with the exact values: the claim in DESIGN.md is "within an ulp of the correctly rounded result"."""

import math
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np
import pytest

getcontext().prec = 80


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))  # float(Fraction) rounds to nearest even: a true fma


def ulp(x):
    return math.ulp(abs(x))


def rcp_refined(x, seed):
    e = fma(-x, seed, 1.0)
    e2 = fma(e, e, e)
    return fma(seed, e2, seed)


def rsqrt_refined(x, seed):
    h = (0.5 * x) * seed
    e = fma(-h, seed, 0.5)
    c = fma(1.5, e, 1.0)
    return fma(seed * e, c, seed)


def _samples():
    rng = np.random.Generator(np.random.PCG64(11))
    xs = list(rng.uniform(1.0, 2.0, size=150) * 2.0 ** rng.integers(-40, 41, size=150))
    xs += [1.0, 1.5, 2.0 - 2.0**-52, 3.0, 1e-300, 1e300]
    deltas = [2.0**-20, -(2.0**-20), 0.0] + list(rng.uniform(-1, 1, size=3) * 2.0**-20)
    return xs, deltas


def test_third_order_reciprocal_is_within_an_ulp():
    xs, deltas = _samples()
    worst = 0.0
    for x in xs:
        exact = Fraction(1) / Fraction(x)
        for dl in deltas:
            seed = float(exact) * (1.0 + dl)
            got = rcp_refined(x, seed)
            err = abs(Fraction(got) - exact) / Fraction(ulp(float(exact)))
            worst = max(worst, float(err))
    assert worst <= 1.0, worst


def test_third_order_reciprocal_square_root_is_within_an_ulp():
    xs, deltas = _samples()
    worst = 0.0
    for x in xs:
        exact = 1 / Decimal(x).sqrt()
        for dl in deltas:
            seed = float(exact) * (1.0 + dl)
            got = rsqrt_refined(x, seed)
            err = abs(Decimal(got) - exact) / Decimal(ulp(float(exact)))
            worst = max(worst, float(err))
    assert worst <= 1.0, worst


@pytest.mark.parametrize("dl", [2.0**-20, -(2.0**-20)])
def test_one_second_order_step_would_not_be_enough(dl):
    """Why the correction is third order: a single Newton step from a 20-bit seed leaves ~2^-40."""
    x = 1.2345678901234567
    exact = Fraction(1) / Fraction(x)
    seed = float(exact) * (1.0 + dl)
    newton = fma(seed, fma(-x, seed, 1.0), seed)
    assert abs(Fraction(newton) - exact) / Fraction(ulp(float(exact))) > 1000
