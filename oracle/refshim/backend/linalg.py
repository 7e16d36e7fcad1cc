"""`probdiffeq.backend.linalg` on NumPy / SciPy (LAPACK through SciPy instead of through XLA)."""
import numpy as _np
import scipy.linalg as _sla

from oracle.refshim.backend._array import wrap_module as _wrap_module


def qr_r(arr, /):
    return _np.linalg.qr(arr, mode="r")


def vector_norm(arr, /, *, order=None):
    return _np.linalg.norm(arr, ord=order)


def matrix_norm(arr, /, *, order=None):
    return _np.linalg.norm(arr, ord=order)


def _substitute(matrix, rhs, lower):
    """Forward / back substitution like BLAS trsm (what XLA calls): a zero pivot divides by zero (inf / nan), it does
    not raise -- the reference relies on that for the singular factor of an exactly known initial condition."""
    a = _np.asarray(matrix, dtype=_np.result_type(matrix, rhs, _np.float64))
    b = _np.array(rhs, dtype=a.dtype, copy=True)
    vec = b.ndim == 1
    if vec:
        b = b[:, None]
    n = a.shape[0]
    order = range(n) if lower else range(n - 1, -1, -1)
    with _np.errstate(all="ignore"):
        for i in order:
            js = slice(0, i) if lower else slice(i + 1, n)
            b[i] = (b[i] - a[i, js] @ b[js]) / a[i, i]
    return b[:, 0] if vec else b


def solve_triu(matrix, rhs, /, *, trans=0):
    m = _np.asarray(matrix)
    return _substitute(m.T, rhs, lower=True) if trans in (1, "T") else _substitute(m, rhs, lower=False)


def solve_tril(matrix, rhs, /, *, trans=0):
    m = _np.asarray(matrix)
    return _substitute(m.T, rhs, lower=False) if trans in (1, "T") else _substitute(m, rhs, lower=True)


def solve_lu(matrix, rhs, /):
    return _np.linalg.solve(matrix, rhs)


def lstsq_svd(matrix, rhs, /):
    return _np.linalg.lstsq(matrix, rhs, rcond=None)[0]


def lstsq_lsmr(vecmat_fun, rhs, /, *, x0, damp, tol, **lsmr_kwargs):
    raise NotImplementedError("matrix-free least squares is outside the path this shim serves")


def inv(matrix, /):
    return _np.linalg.inv(matrix)


def pinv(matrix, /):
    return _np.linalg.pinv(matrix)


def vector_dot(a, b, /):
    return _np.dot(a, b)


def diagonal_along_axis(arr, /, *, axis1, axis2):
    return _np.diagonal(arr, axis1=axis1, axis2=axis2)


def diagonal(arr, /, *, axis1=0, axis2=1):
    return _np.diagonal(arr, axis1=axis1, axis2=axis2)


def trace(arr, /, *, axis1=0, axis2=1):
    return _np.trace(arr, axis1=axis1, axis2=axis2)


def diagonal_matrix(arr, /, k=0):
    return _np.diag(arr, k=k)


def triu(arr, /):
    return _np.triu(arr)


def expm(arr, /):
    return _sla.expm(arr)


def einsum(expression, *args):
    return _np.einsum(expression, *args)


_wrap_module(globals())
