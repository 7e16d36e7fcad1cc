"""Host-side constants of the integrated Wiener process prior.

These are computed once per solver construction, on the host, with the same formulas the reference
uses (probdiffeq/_probdiffeq/utilities.py:57-97, probdiffeq/util/cholesky_util.py:106-176,
probdiffeq/backend/np.py:8-9) and handed to the kernels inside ``pdeq_config``.
"""

from __future__ import annotations

import functools

import numpy as np
from scipy.special import gammaln


def factorial(k):
    """exp(lgamma(k+1)): inexact on purpose, like the reference (backend/np.py:8-9)."""
    return np.exp(gammaln(np.asarray(k, dtype=np.float64) + 1.0))


def _hilbert_cholesky(n: int) -> np.ndarray:
    """Kahan's recurrence (util/cholesky_util.py:106-176); lower-triangular factor of hilbert(n)."""
    roots = np.sqrt(np.arange(1, 2 * n, 2, dtype=np.float64))
    f = np.ones(n)
    for i in range(1, n):
        f[i] = (((f[i - 1] / i) * (2.0 * i)) / i) * (2.0 * i + 1.0)
    f = 1.0 / f
    upper = np.eye(n)
    for j in range(1, n):
        for i in range(j - 1, -1, -1):
            upper[i, j] = (upper[i + 1, j] / (j - i)) * (i + j + 2.0)
    upper = upper * (roots[:, None] * f[None, :])
    return np.tril(upper.T)


@functools.lru_cache(maxsize=None)
def error_constants(num_derivatives: int, order: int) -> np.ndarray:
    """Constants of `error_state_std` for the ts0 constraint and damp = 0 (see `pdeq_config.err_const`).

    The estimator's Bayes rule (reference solvers.py:1070-1086) triangularises [[0, 0], [(h L)^T, L^T]] with
    L = diag(|p|) sqrt(dt) lambda q and h = e_order; column scalings commute with the triangularisation, so only
    R = qr_r([[0, 0], [q[order, :]^T, q^T]]) is needed: entry 0 is R[0, 0], entry 1 + i the norm of R[1:2+i, 1+i]."""
    _, q, _ = system_matrices(num_derivatives)
    n = num_derivatives + 1
    m = np.zeros((n + 1, n + 1))
    m[1:, 0] = q[order, :]
    m[1:, 1:] = q.T
    r = np.linalg.qr(m, mode="r")
    out = np.zeros(n + 1)
    out[0] = r[0, 0]
    for i in range(n):
        out[1 + i] = np.linalg.norm(r[1 : 2 + i, 1 + i])
    return out


@functools.lru_cache(maxsize=None)
def system_matrices(num_derivatives: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(A, Q, factorials): flipped Pascal matrix, Cholesky factor of the flipped Hilbert matrix, k!."""
    n = num_derivatives + 1
    idx = np.arange(n, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        binom = factorial(idx[:, None]) / (factorial(idx[:, None] - idx[None, :]) * factorial(idx[None, :]))
    binom = np.where(np.isfinite(binom), binom, 0.0)  # 1/Gamma at non-positive integers is 0
    a = np.flip(binom)
    q = np.flip(_hilbert_cholesky(n), axis=0)
    q = np.linalg.qr(q.T, mode="r").T
    sign = np.sign(np.diagonal(q))
    q = q * np.where(sign == 0.0, 1.0, sign)[None, :]
    facts = factorial(np.arange(n + 1))
    return a, q, facts
