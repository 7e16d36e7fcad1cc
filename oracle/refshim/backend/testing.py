"""pytest glue of the reference's own test-suite (reference: probdiffeq/backend/testing.py), so that those of the
reference's tests that do not need `pytest_cases` (not installed here) can be run on this backend -- which is how the
backend itself is validated (oracle/refshim/run_reference_tests.py). Not needed to run the step loop."""
import numpy as _np
import pytest as _pytest

from oracle.refshim.backend import tree as _tree

filterwarnings = _pytest.mark.filterwarnings
parametrize = _pytest.mark.parametrize
raises = _pytest.raises
warns = _pytest.warns
xfail = _pytest.xfail


def skip(reason):
    return _pytest.skip(reason=reason)


def _needs_pytest_cases(*_args, **_kwargs):
    raise ImportError("pytest_cases is not installed: this test module cannot be collected on the NumPy backend")


case = fixture = parametrize_with_cases = _needs_pytest_cases


def _allclose(a, b, /, *, atol, rtol, strict_shapes):
    a, b = _np.asarray(1.0 * _np.asarray(a)), _np.asarray(1.0 * _np.asarray(b))
    # The reference derives its default tolerances from the dtype of the leaves, and its suite runs in JAX's default
    # SINGLE precision (no x64 switch in its test configuration): sqrt(eps_float32). The same acceptance thresholds
    # are used here although the arithmetic is float64 -- several of those tests compare a solver's output with an
    # exact solution, where the threshold has to cover the discretisation error, not rounding.
    tol = _np.sqrt(_np.finfo(_np.float32).eps)
    atol = tol if atol is None else atol
    rtol = 10 * tol if rtol is None else rtol
    close = bool(_np.allclose(a, b, atol=atol, rtol=rtol))
    return close and (a.shape == b.shape or not strict_shapes)


def allclose(tree1, tree2, /, *, atol=None, rtol=None, strict_shapes=True):
    """Pytree-aware allclose with tolerances proportional to sqrt(machine epsilon), like the reference's."""
    flags = _tree.tree_map(lambda a, b: _allclose(a, b, atol=atol, rtol=rtol, strict_shapes=strict_shapes), tree1, tree2)
    return _tree.tree_all(flags)


def marginals_allclose(m1, m2, /, *, atol=None, rtol=None, strict_shapes=True):
    mean1, cov1 = m1.to_multivariate_normal()
    mean2, cov2 = m2.to_multivariate_normal()
    return (_allclose(mean1, mean2, atol=atol, rtol=rtol, strict_shapes=strict_shapes)
            and _allclose(cov1, cov2, atol=atol, rtol=rtol, strict_shapes=strict_shapes))  # fmt: skip
