"""Known answers for the pieces of oracle/refshim (the NumPy implementation of the reference's array backend, on which
the reference's own modules are run to produce the parity fixtures) that are not thin NumPy wrappers: Taylor-mode
`jet` in exact rational arithmetic, JAX's index clamping and functional updates on arrays, `vmap` as a loop, and the
complex-step derivatives."""

import math

import numpy as np
import pytest

from oracle.refshim.backend import _array, func


def test_jet_of_a_polynomial_path_gives_the_exact_derivatives():
    # f(x) = x^2 along x(s) = 1 + s + s^2 (derivatives 1, 2, 0): f = 1 + 2s + 3s^2 + 2s^3 + s^4
    primal, series = func.jet(lambda x: x * x, (np.asarray([1.0]),), ([np.asarray([1.0]), np.asarray([2.0]), np.asarray([0.0])],))
    assert primal.tolist() == [1.0]
    assert [float(s[0]) for s in series] == [2.0, 6.0, 12.0]  # k! times the coefficients 2, 3, 2
    # the same path given as Taylor COEFFICIENTS (the reference's factorial_scaled=False)
    primal, series = func.jet(lambda x: x * x, (np.asarray([1.0]),), ([np.asarray([1.0]), np.asarray([1.0]), np.asarray([0.0])],),
                              is_tcoeff=True)  # fmt: skip
    assert [float(s[0]) for s in series] == [2.0, 3.0, 2.0]


def test_jet_recursion_reproduces_a_known_series():
    """u' = u^2, u(0) = 1 has u(t) = 1 / (1 - t): every normalised Taylor coefficient is 1, the k-th derivative is k!.
    Built the way the reference's `jetexpand_ode_coefficient_increment` does (jet_expansion_algorithms.py:155-178):
    push the known derivatives through f, append the new one."""
    derivs = [np.asarray([1.0]), np.asarray([1.0])]  # u, u' = f(u)
    for _ in range(5):
        p, s_new = func.jet(lambda x: x * x, (derivs[0],), (derivs[1:],))
        derivs = [derivs[0], p, *s_new]
    assert [float(x[0]) for x in derivs] == [float(math.factorial(k)) for k in range(len(derivs))]


def test_jet_is_exact_where_float_arithmetic_is_not():
    # (0.1 + s)^3: the coefficient of s is 3 * 0.1^2 evaluated exactly on the double 0.1, then rounded once
    import fractions

    primal, series = func.jet(lambda x: x * x * x, (np.asarray(0.1),), ([np.asarray(1.0)],))
    exact = 3 * fractions.Fraction(0.1) ** 2
    assert float(series[0]) == float(exact) and float(primal) == float(fractions.Fraction(0.1) ** 3)


def test_arrays_clamp_out_of_range_integer_indices_like_jax_and_still_iterate():
    a = np.arange(12.0).reshape(4, 3).view(_array.Arr)
    assert np.array_equal(a[7], a[3]) and np.array_equal(a[7, ...], a[3]) and np.array_equal(a[-9], a[0])
    assert a[np.asarray(5), 1] == a[3, 1]
    assert [row.tolist() for row in a] == a.tolist()  # iteration ends (ndarray iterates through __getitem__)
    assert len([*a]) == 4
    with pytest.raises(IndexError):
        a[0, 0, 0]
    b = a.at[1, 2].set(-1.0)
    assert b[1, 2] == -1.0 and a[1, 2] == 5.0  # functional update: the original is untouched
    assert np.array_equal(a.at[0].add(1.0)[0], a[0] + 1.0)


def test_vmap_loops_over_the_mapped_axes():
    f = func.vmap(lambda x, y: x @ y, in_axes=(0, None))
    x, y = np.arange(6.0).reshape(2, 3), np.arange(3.0)
    assert np.array_equal(np.asarray(f(x, y)), x @ y)
    s, d = func.vmap(lambda x: (x.sum(), 2 * x))(x)
    assert np.array_equal(np.asarray(s), x.sum(axis=1)) and np.array_equal(np.asarray(d), 2 * x)
    assert np.array_equal(np.asarray(func.vmap(lambda x: 2 * x, out_axes=1)(x)), (2 * x).T)


def test_complex_step_derivatives_are_exact_to_rounding_for_polynomials():
    def f(u):
        return np.stack([u[0] * u[1] - 2.0 * u[0], u[1] * u[1] * u[0]])

    u, v = np.asarray([0.7, -1.3]), np.asarray([0.2, 0.5])
    _, tangent = func.jvp(f, (u,), (v,))
    jac = np.asarray([[u[1] - 2.0, u[0]], [u[1] ** 2, 2 * u[0] * u[1]]])
    assert np.allclose(np.asarray(tangent), jac @ v, rtol=1e-15, atol=1e-16)
    assert np.allclose(np.asarray(func.jacfwd(f)(u)), jac, rtol=1e-15, atol=1e-16)


def test_the_references_own_tests_pass_on_this_backend():
    """Where the reference's sources exist (the build container): the modules of the reference's OWN test-suite that
    concern this path -- solver vs an independent integrator on a pytree-valued Lotka-Volterra problem, save_at (55
    tests), fixed grid == adaptive grid, dense == isotropic == block-diagonal, filter vs smoother, fixed-interval vs
    fixed-point smoother, dynamic and MLE calibration, both log-marginal-likelihood losses, sampling, dense output,
    priors (is_exact, output scales, diffuse derivatives), the second-order harmonic oscillator, the IWP known
    answers, the controllers, the Cholesky utilities -- run on the NumPy backend the parity fixtures were produced on
    (oracle/refshim/run_reference_tests.py; the record is profiles/r3k_reference_own_tests_on_numpy_backend.txt)."""
    import os
    import subprocess
    import sys

    from oracle import refshim

    if not refshim.available():
        pytest.skip("the reference sources are not on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = subprocess.run([sys.executable, "-m", "oracle.refshim.run_reference_tests"], cwd=root, capture_output=True,
                         text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"), timeout=600)  # fmt: skip
    last = run.stdout.strip().splitlines()[-1]
    assert run.returncode == 0 and "224 passed" in last and "failed" not in last, run.stdout[-2000:]
