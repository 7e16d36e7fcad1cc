"""Linear-algebra primitives of the oracle (test infrastructure, see oracle/__init__.py).

Restates probdiffeq/backend/linalg.py, probdiffeq/backend/np.py:8-9,
probdiffeq/util/cholesky_util.py and probdiffeq/_probdiffeq/utilities.py:57-97.
All functions broadcast over leading batch axes, which is how the block-diagonal
factorisation's ``vmap`` is restated.
"""

import numpy as np
import scipy.special


def qr_r(arr):
    """R-factor of a (batched) QR. backend/linalg.py:8-10 (LAPACK geqrf, no sign fix)."""
    arr = np.asarray(arr, dtype=np.float64)
    if arr.shape[-2] < arr.shape[-1]:
        # geqrf on a wide matrix returns an (m, n) trapezoid; numpy does too.
        return np.linalg.qr(arr, mode="r")
    return np.linalg.qr(arr, mode="r")


def solve_triu(matrix, rhs):
    """Back substitution, batched. backend/linalg.py:48-49 (trsm, upper, no transpose)."""
    matrix = np.asarray(matrix, dtype=np.float64)
    rhs = np.asarray(rhs, dtype=np.float64)
    vec = rhs.ndim == matrix.ndim - 1
    if vec:
        rhs = rhs[..., None]
    n = matrix.shape[-1]
    x = np.zeros(np.broadcast_shapes(rhs.shape, matrix.shape[:-2] + rhs.shape[-2:]))
    for i in range(n - 1, -1, -1):
        acc = rhs[..., i, :] - np.einsum(
            "...j,...jk->...k", matrix[..., i, i + 1 :], x[..., i + 1 :, :]
        )
        x[..., i, :] = acc / matrix[..., i, i, None]
    return x[..., 0] if vec else x


def solve_tril(matrix, rhs):
    """Forward substitution, batched. backend/linalg.py:52-53."""
    matrix = np.asarray(matrix, dtype=np.float64)
    rhs = np.asarray(rhs, dtype=np.float64)
    vec = rhs.ndim == matrix.ndim - 1
    if vec:
        rhs = rhs[..., None]
    n = matrix.shape[-1]
    x = np.zeros(np.broadcast_shapes(rhs.shape, matrix.shape[:-2] + rhs.shape[-2:]))
    for i in range(n):
        acc = rhs[..., i, :] - np.einsum(
            "...j,...jk->...k", matrix[..., i, :i], x[..., :i, :]
        )
        x[..., i, :] = acc / matrix[..., i, i, None]
    return x[..., 0] if vec else x


def factorial(n):
    """exp(lgamma(n+1)) -- deliberately inexact like backend/np.py:8-9."""
    return np.exp(scipy.special.gammaln(np.asarray(n, dtype=np.float64) + 1.0))


def _T(m):
    return np.swapaxes(m, -1, -2)


def triu_via_qr(R):
    """util/cholesky_util.py:98-103."""
    return qr_r(R)


def sum_of_sqrtm_factors(R_stack):
    """R with R^T R = sum_i R_i^T R_i. util/cholesky_util.py:89-95."""
    return triu_via_qr(np.concatenate(R_stack, axis=-2))


def lstsq_triu(matrix, rhs):
    """backend/linalg.py:60-61 (`jnp.linalg.lstsq`): minimum-norm least squares, batched over leading axes.

    Equals `solve_triu` for a non-singular matrix; a zero pivot (noise-free observation of an exactly known
    state) yields a zero gain instead of NaN."""
    matrix = np.asarray(matrix, dtype=np.float64)
    rhs = np.asarray(rhs, dtype=np.float64)
    if matrix.ndim == 2:
        return np.linalg.lstsq(matrix, rhs, rcond=None)[0]
    out = np.empty(np.broadcast_shapes(matrix.shape[:-2], rhs.shape[:-2]) + rhs.shape[-2:])
    for idx in np.ndindex(out.shape[:-2]):
        out[idx] = np.linalg.lstsq(matrix[idx], rhs[idx], rcond=None)[0]
    return out


def revert_conditional(R_X_F, R_X, R_YX, solve=solve_triu):
    """Square-root change of parametrisation p(Y|X)p(X) -> p(X|Y)p(Y).

    util/cholesky_util.py:27-82. Returns (R_Y, (R_XY, G)).
    """
    k = R_YX.shape[-1]
    n = R_X.shape[-1]
    batch = R_X.shape[:-2]
    top = np.concatenate([R_YX, np.zeros(batch + (R_YX.shape[-2], n))], axis=-1)
    bot = np.concatenate([R_X_F, R_X], axis=-1)
    R = triu_via_qr(np.concatenate([top, bot], axis=-2))
    R_Y = R[..., :k, :k]
    R12 = R[..., :k, k:]
    G = _T(solve(R_Y, R12))
    R_XY = R[..., k:, k:]
    return R_Y, (R_XY, G)


def cholesky_hilbert(n, K=0):
    """Kahan's recurrence for the Cholesky factor of a Hilbert matrix.

    util/cholesky_util.py:106-176. Returns the lower-triangular factor.
    """
    Kf = float(K)
    dr = np.sqrt(np.arange(K + 1, K + 2 * n, 2, dtype=np.float64))
    f = np.ones(n) * (1.0 + Kf)
    for i in range(1, n):
        fi = float(i)
        f[i] = (((f[i - 1] / fi) * (Kf + 2.0 * fi)) / (Kf + fi)) * (Kf + 2.0 * fi + 1.0)
    f = 1.0 / f
    U = np.eye(n)
    for j in range(1, n):
        g = U[:, j].copy()
        for kk in range(j):
            i = j - 1 - kk
            g[i] = (g[i + 1] / float(j - i)) * (Kf + float(i + 1) + float(j + 1))
        U[:, j] = g
    U = U * (dr[:, None] * f[None, :])
    return np.tril(U.T)


def system_matrices_1d_iwp(num_derivatives):
    """Flipped Pascal A and Cholesky factor of the flipped Hilbert matrix.

    _probdiffeq/utilities.py:57-71, 87-97.
    """
    x = np.arange(0, num_derivatives + 1, dtype=np.float64)
    nn, kk = x[:, None], x[None, :]
    with np.errstate(invalid="ignore", divide="ignore"):
        pascal = factorial(nn) / (factorial(nn - kk) * factorial(kk))
    # lgamma of non-positive integers is +inf => factorial = inf => binom = 0 there.
    pascal = np.where(np.isfinite(pascal), pascal, 0.0)
    A = np.flip(pascal)
    Q = cholesky_hilbert(num_derivatives + 1)
    Qf = np.flip(Q, axis=0)
    Q = qr_r(Qf.T).T
    s = np.sign(np.diagonal(Q))
    s = np.where(s == 0.0, 1.0, s)
    return A, Q * s[None, :]


def preconditioner_taylor(num_derivatives):
    """_probdiffeq/utilities.py:74-84."""
    powers = np.arange(num_derivatives, -1.0, -1.0)
    scales = factorial(powers)

    def precon(dt):
        return np.power(dt, powers) / scales, np.power(dt, -powers) * scales

    return precon
