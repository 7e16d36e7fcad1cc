"""Function transformations on NumPy: jit is the identity, vmap is a loop over the mapped axis, and derivatives come
from the complex-step method (exact to rounding for the analytic, NumPy-generic functions the step loop differentiates:
derivative selectors, which are linear, and polynomial vector fields); Taylor-mode differentiation (`jet`) is provided
for polynomial functions in exact rational arithmetic."""
import functools

import numpy as _np

from oracle.refshim.backend import structs as _structs
from oracle.refshim.backend import tree as _tree


def partial(func, *args, **kwargs):
    return functools.partial(func, *args, **kwargs)


def jit(func, /, static_argnums=None, static_argnames=None):
    return func


def stop_gradient(x, /):
    return x


def eval_shape(func, *args, **kwargs):
    def concrete(x):
        return _np.zeros(x.shape, x.dtype) if isinstance(x, _structs.ShapeDtypeStruct) else x

    out = func(*_tree.tree_map(concrete, args), **_tree.tree_map(concrete, kwargs))
    return _tree.tree_map(lambda s: _structs.ShapeDtypeStruct(_np.shape(s), _np.asarray(s).dtype), out)


def _axes_for(arg, ax):
    """Broadcast an in_axes prefix over the leaves of `arg`."""
    if ax is None or isinstance(ax, int):
        return _tree.tree_map(lambda _: ax, arg)
    s_arg, s_ax = _tree._split(arg), _tree._split(ax)
    if s_arg is None or s_ax is None or len(s_arg[2]) != len(s_ax[2]):
        raise ValueError("vmap: in_axes is not a prefix of the argument")
    return _tree._build(s_arg[0], s_arg[1], [_axes_for(c, a) for c, a in zip(s_arg[2], s_ax[2])])


def vmap(func, /, in_axes=0, out_axes=0):
    def mapped(*args):
        plain = isinstance(in_axes, (tuple, list)) and len(in_axes) == len(args) and not hasattr(in_axes, "_fields")
        axes = list(in_axes) if plain else [in_axes] * len(args)
        per_arg = [(_tree.tree_flatten(a)[0], _tree.tree_flatten(a)[1], _tree.tree_flatten(_axes_for(a, ax), is_leaf=lambda x: x is None)[0])
                   for a, ax in zip(args, axes)]  # fmt: skip
        size = None
        for leaves, _, lax in per_arg:
            for leaf, ax in zip(leaves, lax):
                if ax is not None:
                    size = _np.shape(leaf)[ax]
                    break
            if size is not None:
                break
        if size is None:
            raise ValueError("vmap: nothing to map over")
        outs = []
        for i in range(size):
            call = []
            for leaves, structure, lax in per_arg:
                sl = [leaf if ax is None else _np.take(_np.asarray(leaf), i, axis=ax) for leaf, ax in zip(leaves, lax)]
                call.append(_tree.tree_unflatten(structure, sl))
            outs.append(func(*call))
        stacked = _tree.tree_array_stack(outs)
        if out_axes == 0:
            return stacked
        if isinstance(out_axes, int):
            return _tree.tree_map(lambda s: _np.moveaxis(s, 0, out_axes), stacked)
        raise NotImplementedError("vmap: only integer out_axes")

    return mapped


_H = 1e-200  # complex-step size: no subtractive cancellation, the derivative is exact to rounding


def _jvp_flat(fun_flat, x, v):
    out = fun_flat(x.astype(_np.complex128) + 1j * _H * v)
    return _np.imag(out) / _H


def jvp(func, /, primals, tangents):
    flat_p, unravel = _tree.ravel_pytree(primals)
    flat_t, _ = _tree.ravel_pytree(tangents)
    out = func(*primals)
    _, unravel_out = _tree.ravel_pytree(out)

    def fun_flat(z):
        return _tree.ravel_pytree(func(*unravel(z)))[0]

    return out, unravel_out(_jvp_flat(fun_flat, flat_p, flat_t))


def linearize(func, *args):
    out = func(*args)

    def lin(*tangents):
        return jvp(func, args, tangents)[1]

    return out, lin


def jacfwd(func):
    def jac(x):
        x = _np.asarray(x, dtype=_np.float64)
        out = _np.asarray(func(x))
        cols = []
        for k in range(x.size):
            e = _np.zeros(x.size)
            e[k] = 1.0
            z = x.astype(_np.complex128) + 1j * _H * e.reshape(x.shape)
            cols.append(_np.imag(_np.asarray(func(z))) / _H)
        return _np.stack(cols, axis=-1).reshape(out.shape + x.shape)

    return jac


jacrev = jacfwd  # the same matrix


def grad(func):
    def g(x):
        return jacfwd(func)(x)

    return g


def vjp(func, *args):
    out = func(*args)
    if len(args) != 1:
        raise NotImplementedError
    J = jacfwd(func)(args[0])

    def pullback(ct):
        ct = _np.asarray(ct)
        return (_np.tensordot(ct, J, axes=ct.ndim),)

    return out, pullback


def linear_transpose(func, *args):
    raise NotImplementedError("outside the path this shim serves")


class _ExactSeries:
    """A truncated power series in one variable with EXACT rational coefficients (every float64 is a rational number):
    what `jet` below pushes through a polynomial vector field.  Independent of the oracle's float64 series arithmetic
    (oracle/problems.py:Series) -- the results are the correctly rounded values of the exact Taylor coefficients."""

    __array_ufunc__ = None  # NumPy scalars defer to the reflected operators below
    __slots__ = ("c",)

    def __init__(self, coeffs):
        self.c = list(coeffs)

    @staticmethod
    def _lift(x, order):
        import fractions

        if isinstance(x, _ExactSeries):
            return x
        return _ExactSeries([fractions.Fraction(float(x))] + [fractions.Fraction(0)] * order)

    # scalar-like array methods a right-hand side may call on a component
    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        out = _np.empty(tuple(shape), dtype=object)
        for idx in _np.ndindex(out.shape):
            out[idx] = self
        return out

    def squeeze(self):
        return self

    @staticmethod
    def _broadcast(op, array):
        """`series (op) array`: element by element, as an object array."""
        out = _np.empty(array.shape, dtype=object)
        for idx in _np.ndindex(array.shape):
            out[idx] = op(array[idx])
        return out

    def __add__(self, other):
        if isinstance(other, _np.ndarray) and other.ndim > 0:
            return self._broadcast(lambda o: self + o, other)
        other = self._lift(other, len(self.c) - 1)
        return _ExactSeries([a + b for a, b in zip(self.c, other.c)])

    __radd__ = __add__

    def __neg__(self):
        return _ExactSeries([-a for a in self.c])

    def __sub__(self, other):
        if isinstance(other, _np.ndarray) and other.ndim > 0:
            return self._broadcast(lambda o: self - o, other)
        return self + (-self._lift(other, len(self.c) - 1))

    def __rsub__(self, other):
        if isinstance(other, _np.ndarray) and other.ndim > 0:
            return self._broadcast(lambda o: o - self, other)
        return self._lift(other, len(self.c) - 1) - self

    def __mul__(self, other):
        if isinstance(other, _np.ndarray) and other.ndim > 0:
            return self._broadcast(lambda o: self * o, other)
        other = self._lift(other, len(self.c) - 1)
        n = len(self.c)
        return _ExactSeries([sum(self.c[i] * other.c[k - i] for i in range(k + 1)) for k in range(n)])

    __rmul__ = __mul__

    def __truediv__(self, other):  # by a constant only
        import fractions

        if isinstance(other, _ExactSeries):
            return NotImplemented
        return _ExactSeries([a / fractions.Fraction(float(other)) for a in self.c])


def jet(func, /, primals, series, *, is_tcoeff=False):
    """`jax.experimental.jet.jet` for POLYNOMIAL functions, in exact rational arithmetic: push the paths
    x_i(s) = primal_i + sum_k series_i[k] s^(k+1) / (k+1)!  (without the factorials if `is_tcoeff`, the reference's
    `factorial_scaled=False`) through `func` and return (func(primals), [k-th derivative (or coefficient) of the image
    path, k = 1..K]) rounded to float64.  Division, roots and transcendental functions are not provided: the Taylor
    coefficients of the non-polynomial benchmark right-hand sides (Pleiades) stay an input of the fixtures."""
    import fractions
    import math

    K = len(series[0])

    def path(primal, terms):
        primal = _np.asarray(primal, dtype=_np.float64)
        terms = [_np.asarray(x, dtype=_np.float64) for x in terms]
        out = _np.empty(primal.shape, dtype=object)
        for idx in _np.ndindex(primal.shape):
            coeffs = [fractions.Fraction(float(primal[idx]))]
            for k, x in enumerate(terms, start=1):
                scale = 1 if is_tcoeff else math.factorial(k)
                coeffs.append(fractions.Fraction(float(x[idx])) / scale)
            out[idx] = _ExactSeries(coeffs)
        return out if out.ndim else out[()]

    image = _np.asarray(func(*[path(p, s) for p, s in zip(primals, series)]), dtype=object)
    lifted = _np.empty(image.shape, dtype=object)
    for idx in _np.ndindex(image.shape):
        lifted[idx] = _ExactSeries._lift(image[idx], K)

    def coefficient(k):
        scale = 1 if (is_tcoeff or k == 0) else math.factorial(k)
        out = _np.empty(image.shape, dtype=_np.float64)
        for idx in _np.ndindex(image.shape):
            out[idx] = float(lifted[idx].c[k] * scale)
        return out

    return coefficient(0), [coefficient(k) for k in range(1, K + 1)]
