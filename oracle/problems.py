"""Vector fields, exact Jacobians and Taylor-mode initialisation for the oracle.

Test infrastructure (see oracle/__init__.py).  Restates

* the benchmark vector fields: benchmarks/A0_work-precision-lotka-volterra.py:104-109,
  A1_work-precision-pleiades.py:148-159 (first-order form of :105-119), A2_work-precision-hires.py:157-173,
  A3_work-precision-vanderpol.py:98-101, A4_work-precision-linear-ode.py:114-118,
  A5_work-precision-burgers-pde.py:123-134;
* _probdiffeq/problems.py:283-312 (`ode`, `ode_order_two`) as the `Ode` container;
* _probdiffeq/jet_expansion_algorithms.py:49-177 (`jetexpand_ode_padded_scan` / `_unroll`): both
  return the unnormalised derivatives [u, u', ..., u^(q-1+num)] at t0.  The reference obtains them
  from `jax.experimental.jet`; here the same truncated-series arithmetic is written out (`Series`).
* _probdiffeq/jacobians.py:93-98 (`materialize_dense`: exact Jacobian).  The reference uses
  forward-mode autodiff; the oracle uses complex-step differentiation, which is exact to rounding
  for these analytic right-hand sides.

Every vector field is written once against a tiny array namespace `xp` so that the same code runs
on float arrays, complex arrays (Jacobians) and `Series` (Taylor mode).
"""

import math

import numpy as np


# ------------------------------------------------------------------------------------------------
# Truncated Taylor series arithmetic (normalised coefficients c[k] = x^(k)/k!), vectorised over
# trailing array axes.
# ------------------------------------------------------------------------------------------------


class Series:
    __array_priority__ = 1000
    __array_ufunc__ = None  # make ndarray <op> Series defer to Series.__r<op>__

    def __init__(self, c):
        self.c = np.asarray(c, dtype=np.float64)  # (K, ...)

    @property
    def K(self):
        return self.c.shape[0]

    @staticmethod
    def _lift(x, like):
        if isinstance(x, Series):
            return x
        x = np.asarray(x, dtype=np.float64)
        c = np.zeros((like.K,) + np.broadcast_shapes(x.shape, like.c.shape[1:]))
        c[0] = x
        return Series(c)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Series(self.c[(slice(None),) + idx])

    @property
    def shape(self):
        return self.c.shape[1:]

    def __neg__(self):
        return Series(-self.c)

    def __add__(self, o):
        o = Series._lift(o, self)
        return Series(self.c + o.c)

    __radd__ = __add__

    def __sub__(self, o):
        o = Series._lift(o, self)
        return Series(self.c - o.c)

    def __rsub__(self, o):
        o = Series._lift(o, self)
        return Series(o.c - self.c)

    def __mul__(self, o):
        if not isinstance(o, Series):
            return Series(self.c * np.asarray(o, dtype=np.float64))
        K = self.K
        out = np.zeros((K,) + np.broadcast_shapes(self.shape, o.shape))
        for k in range(K):
            for j in range(k + 1):
                out[k] = out[k] + self.c[j] * o.c[k - j]
        return Series(out)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if not isinstance(o, Series):
            return Series(self.c / np.asarray(o, dtype=np.float64))
        K = self.K
        q = np.zeros((K,) + np.broadcast_shapes(self.shape, o.shape))
        for k in range(K):
            acc = self.c[k] + 0.0 * q[k]
            for j in range(k):
                acc = acc - q[j] * o.c[k - j]
            q[k] = acc / o.c[0]
        return Series(q)

    def __rtruediv__(self, o):
        return Series._lift(o, self) / self

    def __pow__(self, r):
        if isinstance(r, int) and r >= 0:
            out = Series._lift(1.0, self)
            for _ in range(r):
                out = out * self
            return out
        K = self.K
        y = np.zeros_like(self.c)
        y[0] = self.c[0] ** r
        for k in range(1, K):
            acc = 0.0
            for j in range(1, k + 1):
                acc = acc + (r * j - (k - j)) * self.c[j] * y[k - j]
            y[k] = acc / (k * self.c[0])
        return Series(y)


class _NumpyOps:
    @staticmethod
    def stack(xs):
        return np.stack(xs)

    @staticmethod
    def concatenate(xs):
        return np.concatenate(xs)

    @staticmethod
    def pad1(x):
        return np.pad(x, 1)

    @staticmethod
    def sum_last(x):
        return np.sum(x, axis=-1)

    @staticmethod
    def outer_diff(x):  # x_j - x_i, shape (i, j)
        return x[None, :] - x[:, None]


class _SeriesOps:
    @staticmethod
    def stack(xs):
        like = next(x for x in xs if isinstance(x, Series))
        return Series(np.stack([Series._lift(x, like).c for x in xs], axis=1))

    @staticmethod
    def concatenate(xs):
        like = next(x for x in xs if isinstance(x, Series))
        return Series(np.concatenate([Series._lift(x, like).c for x in xs], axis=1))

    @staticmethod
    def pad1(x):
        return Series(np.pad(x.c, [(0, 0), (1, 1)]))

    @staticmethod
    def sum_last(x):
        return Series(np.sum(x.c, axis=-1))

    @staticmethod
    def outer_diff(x):
        return Series(x.c[:, None, :] - x.c[:, :, None])


def _ops_for(x):
    return _SeriesOps if isinstance(x, Series) else _NumpyOps


# ------------------------------------------------------------------------------------------------
# The benchmark right-hand sides.  Signature f(xp, p, t, u[, du]) -> f
# ------------------------------------------------------------------------------------------------


def _lotka_volterra(xp, p, t, u):
    a, b, c, d = p
    return xp.stack([a * u[0] - b * u[0] * u[1], -c * u[1] + d * u[0] * u[1]])


def _pleiades(xp, p, t, u):
    x, y, vx, vy = u[0:7], u[7:14], u[14:21], u[21:28]
    dx, dy = xp.outer_diff(x), xp.outer_diff(y)  # (i, j): x_j - x_i
    eye = np.eye(7)
    r = (dx * dx + dy * dy + eye) ** 1.5  # +eye: the i == j terms are skipped (numerator is 0)
    mj = np.arange(1.0, 8.0)[None, :]
    ddx = xp.sum_last(mj * dx / r)
    ddy = xp.sum_last(mj * dy / r)
    return xp.concatenate([vx, vy, ddx, ddy])


def _hires(xp, p, t, u):
    du1 = -1.71 * u[0] + 0.43 * u[1] + 8.32 * u[2] + 0.0007
    du2 = 1.71 * u[0] - 8.75 * u[1]
    du3 = -10.03 * u[2] + 0.43 * u[3] + 0.035 * u[4]
    du4 = 8.32 * u[1] + 1.71 * u[2] - 1.12 * u[3]
    du5 = -1.745 * u[4] + 0.43 * u[5] + 0.43 * u[6]
    du6 = -280.0 * u[5] * u[7] + 0.69 * u[3] + 1.71 * u[4] - 0.43 * u[5] + 0.69 * u[6]
    du7 = 280.0 * u[5] * u[7] - 1.81 * u[6]
    du8 = -280.0 * u[5] * u[7] + 1.81 * u[6]
    return xp.stack([du1, du2, du3, du4, du5, du6, du7, du8])


def _vanderpol(xp, p, t, u, du):
    (stiffness,) = p
    return stiffness * ((1.0 - u * u) * du - u)


def _linear(xp, p, t, u):
    (scale,) = p
    return scale * u


def _burgers(xp, p, t, u):
    (nu,) = p
    d = u.shape[-1]
    dx = 1.0 / (d + 1)
    u_bc = xp.pad1(u)
    u_left, u_right = u_bc[:-2], u_bc[2:]
    flux = u_bc * u_bc / 2.0
    fluxterm = (flux[2:] - flux[:-2]) / (2.0 * dx)
    laplacian = (u_right - 2.0 * u + u_left) / dx**2
    return -fluxterm + nu * laplacian


def _logistic(xp, p, t, u):
    """u' = r u (1 - u / K): the right-hand side the plug-in test compiles at run time (not a benchmark problem)."""
    r, cap = p
    return r * u * (1.0 - u * (1.0 / cap))


_REGISTRY = {
    # name: (function, ode order, number of parameters, default parameters)
    "logistic": (_logistic, 1, 2, (1.0, 1.0)),
    "lotka_volterra": (_lotka_volterra, 1, 4, (0.5, 0.05, 0.5, 0.05)),
    "pleiades": (_pleiades, 1, 0, ()),
    "hires": (_hires, 1, 0, ()),
    "vanderpol": (_vanderpol, 2, 1, (1e3,)),
    "linear": (_linear, 1, 1, (1.5,)),
    "burgers": (_burgers, 1, 1, (0.01,)),
}


class Ode:
    """u^(order) = f(u, ..., u^(order-1), t).  _probdiffeq/problems.py:213-312."""

    def __init__(self, name, params=None):
        fn, order, num_params, default = _REGISTRY[name]
        self.name = name
        self.fn = fn
        self.order = order
        self.params = tuple(default if params is None else np.asarray(params, dtype=np.float64).reshape(-1))
        if len(self.params) != num_params:
            raise ValueError(f"{name} expects {num_params} parameters.")

    def __repr__(self):
        return f"Ode({self.name!r}, order={self.order}, params={self.params})"

    def vector_field(self, jet_coords, t):
        """jet_coords: sequence of `order` arrays (d,). Returns f (d,)."""
        us = [np.asarray(u) for u in jet_coords]
        return self.fn(_NumpyOps, self.params, t, *us)

    def jacobians(self, jet_coords, t):
        """[df/du, df/du', ...], each (d, d), by complex-step differentiation (exact to rounding)."""
        us = [np.asarray(u, dtype=np.float64) for u in jet_coords]
        d = us[0].shape[-1] if us[0].ndim else 1
        h = 1e-200
        out = []
        for k in range(self.order):
            J = np.zeros((d, d))
            for j in range(d):
                args = [u.astype(np.complex128) for u in us]
                if args[k].ndim:
                    args[k][j] += 1j * h
                else:
                    args[k] = args[k] + 1j * h
                col = self.fn(_NumpyOps, self.params, t, *args)
                J[:, j] = np.imag(col) / h
            out.append(J)
        return out

    def taylor_coefficients(self, inits, t, num):
        """[u, u', ..., u^(order-1+num)] at t.  jet_expansion_algorithms.py:49-152."""
        inits = [np.atleast_1d(np.asarray(u, dtype=np.float64)) for u in inits]
        q = self.order
        if len(inits) != q:
            raise ValueError("Number of initial values must equal the ODE order.")
        if num == 0:
            return np.stack(inits)
        total = q + num
        shape = inits[0].shape
        U = np.zeros((total,) + shape)  # normalised coefficients U_k = u^(k)/k!
        for k in range(q):
            U[k] = inits[k] / math.factorial(k)
        for k in range(num):
            # Series of u, u', ..., u^(q-1) truncated to order k (enough for F_k).
            args = []
            for j in range(q):
                cj = np.stack(
                    [U[i + j] * (math.factorial(i + j) / math.factorial(i)) for i in range(k + 1)]
                )
                args.append(Series(cj))
            tser = np.zeros(k + 1)
            tser[0] = t
            if k >= 1:
                tser[1] = 1.0
            F = self.fn(_SeriesOps, self.params, Series(tser), *args)
            Fk = F.c[k]
            U[k + q] = Fk * (math.factorial(k) / math.factorial(k + q))
        fact = np.asarray([float(math.factorial(k)) for k in range(total)])
        return U * fact.reshape((-1,) + (1,) * len(shape))


def ode(name, params=None):
    return Ode(name, params)


def taylor_coefficients_batched(name, params, inits, t, num):
    """Taylor coefficients for an ensemble at once: params (B, P), inits = sequence of (B, d) -> (B, q+num, d).

    Same recursion as `Ode.taylor_coefficients`, with the ensemble as a trailing array axis (works for the
    right-hand sides that only index components: lotka_volterra, hires, vanderpol, linear).
    """
    vf = Ode(name, None if _REGISTRY[name][2] == 0 else np.zeros(_REGISTRY[name][2]))
    params = np.asarray(params, dtype=np.float64)
    vf.params = tuple(params[:, k] for k in range(params.shape[1])) if params.size else ()
    out = vf.taylor_coefficients([np.asarray(u, dtype=np.float64).T for u in inits], t, num)  # (n, d, B)
    return np.ascontiguousarray(np.transpose(out, (2, 0, 1)))


# Initial values of the benchmark problems ---------------------------------------------------------


def pleiades_u0():
    """benchmarks/A1_work-precision-pleiades.py:95-104 (positions then velocities, d=28)."""
    # fmt: off
    return np.asarray([
        3.0, 3.0, -1.0, -3.00, 2.0, -2.00, 2.0,
        3.0, -3.0, 2.0, 0.00, 0.0, -4.00, 4.0,
        0.0, 0.0, 0.0, 0.00, 0.0, 1.75, -1.5,
        0.0, 0.0, 0.0, -1.25, 1.0, 0.00, 0.0,
    ])
    # fmt: on


def hires_u0():
    """benchmarks/A2_work-precision-hires.py:173."""
    return np.asarray([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0057])


def burgers_u0(d):
    """benchmarks/A5_work-precision-burgers-pde.py:136-137 with N = d + 1."""
    x = np.linspace(0.0, 1.0, d + 2, endpoint=True)[1:-1]
    return np.sin(3 * np.pi * x) ** 3 * (1 - x) ** 1.5
