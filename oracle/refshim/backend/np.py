"""`probdiffeq.backend.np` on NumPy (float64 throughout, as the reference is run with jax_enable_x64)."""
import numpy as _np
import scipy.linalg as _sla
import scipy.special as _sp

from oracle.refshim.backend._array import wrap_module as _wrap_module


def factorial(n, /):
    return _np.exp(_sp.gammaln(_np.asarray(n) + 1.0))


def arange(start, stop, *, step=1):
    return _np.arange(start, stop, step)


ndim, shape, minimum, maximum = _np.ndim, _np.shape, _np.minimum, _np.maximum
amax, amin, argmin, sign, atleast_1d, squeeze = _np.amax, _np.amin, _np.argmin, _np.sign, _np.atleast_1d, _np.squeeze
sqrt, log, log2, ceil, exp, cos, hypot = _np.sqrt, _np.log, _np.log2, _np.ceil, _np.exp, _np.cos, _np.hypot
tril, triu, kron, tile, repeat, cumsum = _np.tril, _np.triu, _np.kron, _np.tile, _np.repeat, _np.cumsum
logical_not, logical_and, isinf, isnan, meshgrid, hstack = (_np.logical_not, _np.logical_and, _np.isinf, _np.isnan,
                                                            _np.meshgrid, _np.hstack)  # fmt: skip
searchsorted, block, dtype, power = _np.searchsorted, _np.block, _np.dtype, _np.power


def where(cond, /, if_true, if_false):
    return _np.where(cond, if_true, if_false)


def abs(arr, /):  # noqa: A001
    return _np.abs(arr)


def diff(arr, /, axis=-1):
    return _np.diff(arr, axis=axis)


def reshape(arr, /, new_shape, order="C"):
    return _np.reshape(arr, new_shape, order=order)


def flip(arr, /, axis=None):
    return _np.flip(arr, axis=axis)


def asarray(x, /, dtype=None):
    a = _np.asarray(x, dtype=dtype)
    if dtype is None and a.dtype.kind in "iub" and not isinstance(x, _np.ndarray):
        return a  # index lists stay integer
    return a


def finfo_eps(arr_or_dtype, /):
    d = arr_or_dtype.dtype if hasattr(arr_or_dtype, "dtype") else arr_or_dtype
    return _np.finfo(d).eps


def concatenate(list_of_arrays, /, axis=0):
    return _np.concatenate(list_of_arrays, axis=axis)


def ones(shape, /, dtype=None):
    return _np.ones(shape, dtype=dtype or _np.float64)


def zeros(shape, /, dtype=None):
    return _np.zeros(shape, dtype=dtype or _np.float64)


def empty(shape, /):
    return _np.zeros(shape)


def _like(arr, fill):
    shp = arr.shape if hasattr(arr, "shape") else _np.shape(arr)
    dt = arr.dtype if hasattr(arr, "dtype") else _np.asarray(arr).dtype
    return _np.full(shp, fill, dtype=dt)


def empty_like(arr, /):
    return _like(arr, 0)


def ones_like(arr, /):
    return _like(arr, 1)


def zeros_like(arr, /):
    return _like(arr, 0)


def block_diag(list_of_arrays, /):
    return _sla.block_diag(*list_of_arrays)


def inf():
    return _np.inf


def pi():
    return _np.pi


def eye(n, m=None, /, dtype=None):
    return _np.eye(n, M=m, dtype=dtype or _np.float64)


def save(path, arr, /):
    return _np.save(path, arr, allow_pickle=True)


def load(path, /):
    return _np.load(path, allow_pickle=True)


def stack(list_of_arrays, /, *, axis=0):
    return _np.stack(list_of_arrays, axis=axis)


def transpose(arr, /, *, axes):
    return _np.transpose(arr, axes=axes)


def einsum(how, *operands):
    return _np.einsum(how, *operands)


def any(arr, /):  # noqa: A001
    return _np.any(arr)


def all(arr, /):  # noqa: A001
    return _np.all(arr)


def sum(arr, /):  # noqa: A001
    return _np.sum(arr)


def linspace(start, stop, *, num=50, endpoint=True):
    return _np.linspace(start, stop, num=num, endpoint=endpoint)


def mean(arr, /, axis=None, keepdims=False):
    return _np.mean(arr, axis=axis, keepdims=keepdims)


def std(arr, /, *, axis=None, ddof=0):
    return _np.std(arr, ddof=ddof, axis=axis)


def comb(N, k):
    return _sp.comb(N, k)


_wrap_module(globals())
