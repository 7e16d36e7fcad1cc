"""Factorised state-space algebra of the oracle (test infrastructure, see oracle/__init__.py).

Restates, with plain NumPy arrays instead of pytrees,

* _probdiffeq/ssm_impl_isotropic.py  (mean (n,d), Cholesky (n,n) shared over d)
* _probdiffeq/ssm_impl_blockdiag.py  (mean (d,n), Cholesky (d,n,n))
* _probdiffeq/ssm_impl_dense.py      (mean (n*d,), Cholesky (n*d,n*d), coefficient-major)
* _probdiffeq/ssm_impl_api.py:35-132 (composite Bayes rules)

Left square roots everywhere: cov = L L^T.  `linalg` works with right factors R = L^T.
"""

import numpy as np

from oracle import linalg

_T = linalg._T


class Normal:
    """(mean_flat, cholesky_flat) + the algebra it belongs to. ssm_impl_api.py:203-283."""

    def __init__(self, mean, chol, alg):
        self.mean = np.asarray(mean, dtype=np.float64)
        self.chol = np.asarray(chol, dtype=np.float64)
        self.alg = alg

    # convenience pass-throughs --------------------------------------------------------------
    @property
    def std(self):
        return self.alg.std(self)

    @property
    def tcoeffs(self):
        """Mean as an (n, d) array of Taylor coefficients (the reference's `rv.mean` list)."""
        return self.alg.mean_nd(self)

    def rescale_cholesky(self, factor):
        return self.alg.rescale(self, factor)

    def cov_dense(self):
        return self.alg.cov_dense(self)


class Cond:
    """A latent conditional x -> N(to_observed * (A (to_latent * x)) + noise). ssm_impl_api.py:35-50."""

    def __init__(self, A, noise, to_latent, to_observed):
        self.A = np.asarray(A, dtype=np.float64)
        self.noise = noise
        self.to_latent = np.asarray(to_latent, dtype=np.float64)
        self.to_observed = np.asarray(to_observed, dtype=np.float64)
        self.alg = noise.alg

    def apply_flat(self, x):
        return self.alg.apply_flat(self, x)

    def marginalise(self, rv):
        return self.alg.marginalise(self, rv)

    def revert(self, rv, solve=None):
        return self.alg.revert(self, rv) if solve is None else self.alg.revert(self, rv, solve=solve)

    def merge(self, other):
        return self.alg.merge(self, other)

    def rescale_noise(self, factor):
        """ssm_impl_api.py:78-86."""
        return Cond(self.A, self.noise.rescale_cholesky(factor), self.to_latent, self.to_observed)

    # ssm_impl_api.py:112-132 ---------------------------------------------------------------
    def bayes_rule(self, data, rv):
        _, reverted = self.revert(rv)
        return reverted.apply_flat(data)

    def bayes_rule_and_residual_whitened_rms(self, data, rv):
        observed, reverted = self.revert(rv)
        rms = self.alg.residual_whitened_rms(observed, data)
        return rms, reverted.apply_flat(data)

    def bayes_rule_and_logpdf(self, data, rv, solve=None):
        observed, reverted = self.revert(rv, solve=solve)
        return self.alg.logpdf(observed, data), reverted.apply_flat(data)


# ================================================================================================
# Isotropic
# ================================================================================================


class Isotropic:
    name = "isotropic"

    # ----- LatentCond ops: ssm_impl_isotropic.py:72-141 ---------------------------------------
    def apply_flat(self, c, x):
        x = c.to_latent[:, None] * x
        mean = (c.to_observed[:, None] * c.A) @ x + c.noise.mean  # :77 (noise added unscaled)
        chol = np.abs(c.to_observed[:, None]) * c.noise.chol
        return Normal(mean, chol, self)

    def marginalise(self, c, rv):
        mean = c.to_latent[:, None] * rv.mean
        chol = c.to_latent[:, None] * rv.chol  # :83 (no abs here)
        R = linalg.sum_of_sqrtm_factors(((c.A @ chol).T, c.noise.chol.T))
        mean_new = c.to_observed[:, None] * (c.A @ mean + c.noise.mean)
        return Normal(mean_new, np.abs(c.to_observed[:, None]) * R.T, self)

    def merge(self, outer, inner):
        T = outer.to_latent * inner.to_observed
        g = outer.A @ (T[:, None] * inner.A)
        xi = outer.A @ (T[:, None] * inner.noise.mean) + outer.noise.mean
        R1 = (outer.A @ (np.abs(T[:, None]) * inner.noise.chol)).T
        Xi = linalg.sum_of_sqrtm_factors((R1, outer.noise.chol.T))
        return Cond(g, Normal(xi, Xi.T, self), inner.to_latent, outer.to_observed)

    def revert(self, c, rv, solve=linalg.solve_triu):
        mean = c.to_latent[:, None] * rv.mean
        chol = np.abs(c.to_latent[:, None]) * rv.chol
        r_obs, (r_cor, gain) = linalg.revert_conditional(
            R_X_F=(c.A @ chol).T, R_X=chol.T, R_YX=c.noise.chol.T, solve=solve
        )
        m_obs = c.A @ mean + c.noise.mean
        corrected = Normal(mean - gain @ m_obs, r_cor.T, self)
        cond = Cond(gain, corrected, 1.0 / c.to_observed, 1.0 / c.to_latent)
        observed = Normal(
            c.to_observed[:, None] * m_obs, np.abs(c.to_observed[:, None]) * r_obs.T, self
        )
        return observed, cond

    def preconditioner_apply(self, c):
        A = c.to_observed[:, None] * c.A * c.to_latent[None, :]
        noise = Normal(
            c.to_observed[:, None] * c.noise.mean,
            np.abs(c.to_observed[:, None]) * c.noise.chol,
            self,
        )
        return Cond(A, noise, np.ones(A.shape[1]), np.ones(A.shape[0]))

    # ----- Normal ops: ssm_impl_isotropic.py:146-295 -------------------------------------------
    def from_mean_and_std(self, mean_nd, std_n):
        mean_nd = np.asarray(mean_nd, dtype=np.float64)
        std_n = np.asarray(std_n, dtype=np.float64)
        if std_n.shape != (mean_nd.shape[0],):
            raise ValueError("'std' must hold one scalar per Taylor coefficient.")
        return Normal(mean_nd, np.diag(std_n), self)

    def from_dirac(self, mean_nd, damp):
        mean_nd = np.atleast_2d(np.asarray(mean_nd, dtype=np.float64))
        return self.from_mean_and_std(mean_nd, damp * np.ones(mean_nd.shape[0]))

    def mean_nd(self, rv):
        return rv.mean

    def from_nd(self, x_nd):
        return np.asarray(x_nd, dtype=np.float64)

    def std(self, rv):
        return np.linalg.norm(rv.chol, axis=-1)  # (n,) one scalar per coefficient

    def residual_whitened_rms(self, rv, u):
        w = linalg.solve_tril(rv.chol, rv.mean - u)
        return np.linalg.norm(w.reshape(-1)) / np.sqrt(rv.mean.size)

    def rescale(self, rv, factor):
        return Normal(rv.mean, np.asarray(factor)[..., None, None] * rv.chol, self)

    def logpdf(self, rv, u):
        chol = linalg.qr_r(rv.chol.T).T
        dx = u - rv.mean  # (n, d)
        w = linalg.solve_tril(chol, dx)
        slogdet = np.sum(np.log(np.abs(np.diagonal(chol))))
        n, d = rv.mean.shape
        return np.sum(-0.5 * (2.0 * slogdet + np.sum(w * w, axis=0) + n * np.log(2 * np.pi)))

    def cov_dense(self, rv):
        n, d = rv.mean.shape
        cov = rv.chol @ rv.chol.T
        full = np.einsum("nm,dt->ndmt", cov, np.eye(d)).reshape(n * d, n * d)
        return rv.mean.reshape(-1), full

    def identity_conditional(self, rv):
        n, d = rv.mean.shape
        return Cond(np.eye(n), Normal(np.zeros((n, d)), np.zeros((n, n)), self), np.ones(n), np.ones(n))

    def prototype_output_scale(self, rv):
        return np.ones(())

    def to_derivative(self, rv, i, std):
        n, d = rv.mean.shape
        linop = np.zeros((1, n))
        linop[0, i] = 1.0
        noise = self.from_mean_and_std(np.zeros((1, d)), np.asarray([std], dtype=np.float64).reshape(1))
        return Cond(linop, noise, np.ones(n), np.ones(1))

    # ----- linearisation: ssm_impl_isotropic.py:298-355 -----------------------------------------
    def linearize_ts0(self, ode, rv, damp, t):
        f = ode.vector_field(rv.mean[: ode.order], t)
        n = rv.mean.shape[0]
        H = np.zeros((1, n))
        H[0, ode.order] = 1.0
        return Cond(H, self.from_dirac(-f[None, :], damp), np.ones(n), np.ones(1))

    def linearize_ts1(self, ode, rv, damp, t):
        n, d = rv.mean.shape
        r = rv.mean[ode.order] - ode.vector_field(rv.mean[: ode.order], t)
        jacs = ode.jacobians(rv.mean[: ode.order], t)  # list over input coefficient, each (d, d)
        H = np.zeros((1, n))
        for k, Jk in enumerate(jacs):
            H[0, k] = -np.trace(Jk)
        H[0, ode.order] += float(d)
        H = H / d
        fx = r[None, :] - H @ rv.mean
        return Cond(H, self.from_dirac(fx, damp), np.ones(n), np.ones(1))

    # ----- prior: ssm_impl_isotropic.py:358-482 --------------------------------------------------
    def prior_wiener_integrated(self, tcoeffs, tcoeffs_std=None, output_scale=None):
        tcoeffs = np.asarray(tcoeffs, dtype=np.float64)
        n, d = tcoeffs.shape
        std = np.zeros(n) if tcoeffs_std is None else np.asarray(tcoeffs_std, dtype=np.float64)
        init = self.from_mean_and_std(tcoeffs, std)
        if output_scale is None:
            output_scale = np.ones(())
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != ():
            raise ValueError("The base-scale has the wrong shape.")
        return Prior(self, init, output_scale, n - 1, d)

    def transition(self, prior, dt, output_scale):
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != ():
            raise ValueError("The base-scale has the wrong shape.")
        scale = np.sqrt(np.abs(dt)) * prior.output_scale * output_scale
        noise = Normal(np.zeros((prior.n, prior.d)), scale * prior.q_sqrtm, self)
        p, p_inv = prior.precon(dt)
        return Cond(prior.a, noise, p_inv, p)


# ================================================================================================
# Dense
# ================================================================================================


class Dense:
    name = "dense"

    def __init__(self, d):
        self.d = d  # ODE dimension; the flat state is coefficient-major (index k*d + i)

    # ----- LatentCond ops: ssm_impl_dense.py:15-84 ------------------------------------------------
    def apply_flat(self, c, x):
        x = c.to_latent * x
        mean = c.to_observed * (c.A @ x + c.noise.mean)
        return Normal(mean, np.abs(c.to_observed[:, None]) * c.noise.chol, self)

    def marginalise(self, c, rv):
        mean = c.to_latent * rv.mean
        chol = c.to_latent[:, None] * rv.chol
        R = linalg.sum_of_sqrtm_factors(((c.A @ chol).T, c.noise.chol.T))
        mean_new = c.to_observed * (c.A @ mean + c.noise.mean)
        return Normal(mean_new, np.abs(c.to_observed[:, None]) * R.T, self)

    def merge(self, outer, inner):
        T = outer.to_latent * inner.to_observed
        g = outer.A @ (T[:, None] * inner.A)
        xi = outer.A @ (T * inner.noise.mean) + outer.noise.mean
        R1 = (outer.A @ (np.abs(T[:, None]) * inner.noise.chol)).T
        Xi = linalg.sum_of_sqrtm_factors((R1, outer.noise.chol.T))
        return Cond(g, Normal(xi, Xi.T, self), inner.to_latent, outer.to_observed)

    def revert(self, c, rv, solve=linalg.solve_triu):
        mean = c.to_latent * rv.mean
        chol = np.abs(c.to_latent[:, None]) * rv.chol
        r_obs, (r_cor, gain) = linalg.revert_conditional(
            R_X_F=(c.A @ chol).T, R_X=chol.T, R_YX=c.noise.chol.T, solve=solve
        )
        m_obs = c.A @ mean + c.noise.mean
        corrected = Normal(mean - gain @ m_obs, r_cor.T, self)
        cond = Cond(gain, corrected, 1.0 / c.to_observed, 1.0 / c.to_latent)
        observed = Normal(c.to_observed * m_obs, np.abs(c.to_observed[:, None]) * r_obs.T, self)
        return observed, cond

    def preconditioner_apply(self, c):
        A = c.to_observed[:, None] * c.A * c.to_latent[None, :]
        noise = Normal(c.to_observed * c.noise.mean, np.abs(c.to_observed[:, None]) * c.noise.chol, self)
        return Cond(A, noise, np.ones(A.shape[1]), np.ones(A.shape[0]))

    # ----- Normal ops: ssm_impl_dense.py:108-234 ----------------------------------------------------
    def from_mean_and_std(self, mean_nd, std_nd):
        mean_nd = np.atleast_2d(np.asarray(mean_nd, dtype=np.float64))
        std_nd = np.broadcast_to(np.asarray(std_nd, dtype=np.float64), mean_nd.shape)
        return Normal(mean_nd.reshape(-1), np.diag(std_nd.reshape(-1)), self)

    def from_dirac(self, mean_nd, damp):
        mean_nd = np.atleast_2d(np.asarray(mean_nd, dtype=np.float64))
        return self.from_mean_and_std(mean_nd, damp * np.ones_like(mean_nd))

    def mean_nd(self, rv):
        return rv.mean.reshape(-1, self.d)

    def from_nd(self, x_nd):
        return np.asarray(x_nd, dtype=np.float64).reshape(-1)

    def std(self, rv):
        # qr_r of each row as a column == +-norm; abs. ssm_impl_dense.py:144-150
        r = linalg.qr_r(rv.chol[..., None])
        return np.abs(r.reshape(-1)).reshape(-1, self.d)

    def residual_whitened_rms(self, rv, u):
        w = linalg.solve_tril(rv.chol, np.asarray(u).reshape(-1) - rv.mean)
        return np.linalg.norm(w) / np.sqrt(rv.mean.size)

    def rescale(self, rv, factor):
        return Normal(rv.mean, np.asarray(factor)[..., None, None] * rv.chol, self)

    def logpdf(self, rv, u):
        chol = linalg.qr_r(rv.chol.T).T
        slogdet = np.sum(np.log(np.abs(np.diagonal(chol))))
        w = linalg.solve_tril(chol, np.asarray(u).reshape(-1) - rv.mean)
        return -0.5 * (w @ w) - rv.mean.size / 2 * np.log(2 * np.pi) - slogdet

    def cov_dense(self, rv):
        return rv.mean, rv.chol @ rv.chol.T

    def identity_conditional(self, rv):
        N = rv.mean.size
        return Cond(np.eye(N), Normal(np.zeros(N), np.zeros((N, N)), self), np.ones(N), np.ones(N))

    def prototype_output_scale(self, rv):
        return np.ones(())

    def to_derivative(self, rv, i, std):
        d = self.d
        n = rv.mean.size // d
        linop = np.zeros((d, n * d))
        linop[:, i * d : (i + 1) * d] = np.eye(d)
        noise = self.from_mean_and_std(np.zeros((1, d)), np.broadcast_to(std, (1, d)))
        return Cond(linop, noise, np.ones(n * d), np.ones(d))

    # ----- linearisation: ssm_impl_dense.py:237-334 ---------------------------------------------------
    def linearize_ts0(self, ode, rv, damp, t):
        d = self.d
        n = rv.mean.size // d
        m = self.mean_nd(rv)
        f = ode.vector_field(m[: ode.order], t)
        H = np.zeros((d, n * d))
        H[:, ode.order * d : (ode.order + 1) * d] = np.eye(d)
        return Cond(H, self.from_dirac(-f[None, :], damp), np.ones(n * d), np.ones(d))

    def linearize_ts1(self, ode, rv, damp, t):
        d = self.d
        n = rv.mean.size // d
        xi = rv.mean  # taylor_point_prior: _probdiffeq/taylor_points.py:150-156
        m = xi.reshape(n, d)
        r = m[ode.order] - ode.vector_field(m[: ode.order], t)
        J = np.zeros((d, n * d))
        for k, Jk in enumerate(ode.jacobians(m[: ode.order], t)):
            J[:, k * d : (k + 1) * d] = -Jk
        J[:, ode.order * d : (ode.order + 1) * d] += np.eye(d)
        fx = r - J @ xi
        return Cond(J, self.from_dirac(fx[None, :], damp), np.ones(n * d), np.ones(d))

    # ----- prior: ssm_impl_dense.py:337-389, 455-514, 686-715 ----------------------------------------
    def prior_wiener_integrated(self, tcoeffs, tcoeffs_std=None, output_scale=None):
        tcoeffs = np.asarray(tcoeffs, dtype=np.float64)
        n, d = tcoeffs.shape
        std = np.zeros((n, d)) if tcoeffs_std is None else np.asarray(tcoeffs_std, dtype=np.float64)
        init = self.from_mean_and_std(tcoeffs, std)
        if output_scale is None:
            output_scale = np.ones(d)
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != (d,):
            raise ValueError("The base-scale has the wrong shape.")
        prior = Prior(self, init, np.diag(output_scale), n - 1, d)
        prior.A_full = np.kron(prior.a, np.eye(d))
        prior.Q_full = np.kron(prior.q_sqrtm, np.diag(output_scale))
        return prior

    def transition(self, prior, dt, output_scale):
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != ():
            raise ValueError("The output-scale has the wrong shape.")
        p, p_inv = prior.precon(dt)
        p, p_inv = np.repeat(p, prior.d), np.repeat(p_inv, prior.d)
        noise = Normal(
            np.zeros(prior.n * prior.d), np.sqrt(np.abs(dt)) * output_scale * prior.Q_full, self
        )
        return Cond(prior.A_full, noise, p_inv, p)


# ================================================================================================
# Block-diagonal
# ================================================================================================


class BlockDiag:
    name = "blockdiag"

    # ----- LatentCond ops: ssm_impl_blockdiag.py:16-113 ---------------------------------------------
    def apply_flat(self, c, x):
        s = c.to_latent * x
        mean = c.to_observed * (np.einsum("ijk,ik->ij", c.A, s) + c.noise.mean)
        return Normal(mean, np.abs(c.to_observed[:, :, None]) * c.noise.chol, self)

    def marginalise(self, c, rv):
        mean = c.to_latent * rv.mean
        chol = np.abs(c.to_latent[:, :, None]) * rv.chol
        mean_marg = np.einsum("ijk,ik->ij", c.A, mean) + c.noise.mean
        R = linalg.sum_of_sqrtm_factors((_T(c.A @ chol), _T(c.noise.chol)))
        return Normal(c.to_observed * mean_marg, np.abs(c.to_observed[:, :, None]) * _T(R), self)

    def merge(self, outer, inner):
        T = outer.to_latent * inner.to_observed
        A1, A2 = outer.A, T[:, :, None] * inner.A
        g = A1 @ A2
        xi = np.einsum("ijk,ik->ij", A1, T * inner.noise.mean) + outer.noise.mean
        C2 = np.abs(T[:, :, None]) * inner.noise.chol
        Xi = linalg.sum_of_sqrtm_factors((_T(A1 @ C2), _T(outer.noise.chol)))
        return Cond(g, Normal(xi, _T(Xi), self), inner.to_latent, outer.to_observed)

    def revert(self, c, rv, solve=linalg.solve_triu):
        mean = c.to_latent * rv.mean
        chol = np.abs(c.to_latent[:, :, None]) * rv.chol
        r_obs, (r_cor, gain) = linalg.revert_conditional(
            R_X_F=_T(c.A @ chol), R_X=_T(chol), R_YX=_T(c.noise.chol), solve=solve
        )
        m_obs = np.einsum("ijk,ik->ij", c.A, mean) + c.noise.mean
        m_cor = mean - np.einsum("ijk,ik->ij", gain, m_obs)
        cond = Cond(gain, Normal(m_cor, _T(r_cor), self), 1.0 / c.to_observed, 1.0 / c.to_latent)
        observed = Normal(c.to_observed * m_obs, np.abs(c.to_observed[:, :, None]) * _T(r_obs), self)
        return observed, cond

    def preconditioner_apply(self, c):
        A = c.to_observed[:, :, None] * c.A * c.to_latent[:, None, :]
        noise = Normal(c.to_observed * c.noise.mean, np.abs(c.to_observed[:, :, None]) * c.noise.chol, self)
        return Cond(A, noise, np.ones_like(c.to_latent), np.ones_like(c.to_observed))

    # ----- Normal ops: ssm_impl_blockdiag.py:215-394 ---------------------------------------------------
    def from_mean_and_std(self, mean_nd, std_nd):
        mean_nd = np.atleast_2d(np.asarray(mean_nd, dtype=np.float64))
        std_nd = np.broadcast_to(np.asarray(std_nd, dtype=np.float64), mean_nd.shape)
        n = mean_nd.shape[0]
        return Normal(mean_nd.T, std_nd.T[..., None] * np.eye(n)[None], self)

    def from_dirac(self, mean_nd, damp):
        mean_nd = np.atleast_2d(np.asarray(mean_nd, dtype=np.float64))
        return self.from_mean_and_std(mean_nd, damp * np.ones_like(mean_nd))

    def mean_nd(self, rv):
        return rv.mean.T

    def from_nd(self, x_nd):
        return np.asarray(x_nd, dtype=np.float64).T

    def std(self, rv):
        return np.linalg.norm(rv.chol, axis=-1).T  # (n, d)

    def residual_whitened_rms(self, rv, u):
        w = linalg.solve_tril(rv.chol, u - rv.mean)
        return np.linalg.norm(w, axis=-1) / np.sqrt(rv.mean.shape[-1])  # (d,)

    def rescale(self, rv, factor):
        return Normal(rv.mean, np.asarray(factor)[..., None, None] * rv.chol, self)

    def logpdf(self, rv, u):
        chol = _T(linalg.qr_r(_T(rv.chol)))
        w = linalg.solve_tril(chol, u - rv.mean)
        slogdet = np.sum(np.log(np.abs(np.diagonal(chol, axis1=-1, axis2=-2))), axis=-1)
        k = rv.mean.shape[-1]
        return np.sum(-0.5 * (2.0 * slogdet + np.sum(w * w, axis=-1) + k * np.log(2 * np.pi)))

    def cov_dense(self, rv):
        d, n = rv.mean.shape
        cov = rv.chol @ _T(rv.chol)
        full = np.einsum("dnm,dt->ndmt", cov, np.eye(d)).reshape(n * d, n * d)
        return rv.mean.T.reshape(-1), full

    def identity_conditional(self, rv):
        d, n = rv.mean.shape
        noise = Normal(np.zeros((d, n)), np.zeros((d, n, n)), self)
        return Cond(np.ones((d, 1, 1)) * np.eye(n)[None], noise, np.ones((d, n)), np.ones((d, n)))

    def prototype_output_scale(self, rv):
        return np.ones(rv.mean.shape[0])

    def to_derivative(self, rv, i, std):
        d, n = rv.mean.shape
        linop = np.zeros((d, 1, n))
        linop[:, 0, i] = 1.0
        noise = self.from_mean_and_std(np.zeros((1, d)), np.broadcast_to(std, (1, d)))
        return Cond(linop, noise, np.ones((d, n)), np.ones((d, 1)))

    # ----- linearisation: ssm_impl_blockdiag.py:123-183 -------------------------------------------------
    def linearize_ts0(self, ode, rv, damp, t):
        d, n = rv.mean.shape
        f = ode.vector_field(rv.mean.T[: ode.order], t)
        H = np.zeros((d, 1, n))
        H[:, 0, ode.order] = 1.0
        return Cond(H, self.from_dirac(-f[None, :], damp), np.ones((d, n)), np.ones((d, 1)))

    def linearize_ts1(self, ode, rv, damp, t):
        d, n = rv.mean.shape
        m = rv.mean.T
        r = m[ode.order] - ode.vector_field(m[: ode.order], t)
        H = np.zeros((d, 1, n))
        for k, Jk in enumerate(ode.jacobians(m[: ode.order], t)):
            H[:, 0, k] = -np.diagonal(Jk)
        H[:, 0, ode.order] += 1.0
        fx = r[:, None] - np.einsum("din,dn->di", H, rv.mean)
        return Cond(H, self.from_dirac(fx.T, damp), np.ones((d, n)), np.ones((d, 1)))

    # ----- prior: ssm_impl_blockdiag.py:397-547 -----------------------------------------------------------
    def prior_wiener_integrated(self, tcoeffs, tcoeffs_std=None, output_scale=None):
        tcoeffs = np.asarray(tcoeffs, dtype=np.float64)
        n, d = tcoeffs.shape
        std = np.zeros((n, d)) if tcoeffs_std is None else np.asarray(tcoeffs_std, dtype=np.float64)
        init = self.from_mean_and_std(tcoeffs, std)
        if output_scale is None:
            output_scale = np.ones(d)
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != (d,):
            raise ValueError("The output-scale has the wrong shape.")
        return Prior(self, init, output_scale, n - 1, d)

    def transition(self, prior, dt, output_scale):
        p, p_inv = prior.precon(dt)
        output_scale = np.asarray(output_scale, dtype=np.float64)
        if output_scale.shape != prior.output_scale.shape:
            raise ValueError("The output-scale has the wrong shape.")
        scale = prior.output_scale * output_scale
        d = prior.d
        chol = np.sqrt(np.abs(dt)) * scale[:, None, None] * prior.q_sqrtm[None]
        noise = Normal(np.zeros((d, prior.n)), chol, self)
        A = np.ones((d, 1, 1)) * prior.a[None]
        return Cond(A, noise, np.ones((d, 1)) * p_inv[None], np.ones((d, 1)) * p[None])


class Prior:
    """Integrated Wiener process prior (init rv + system matrices). ssm_impl_api.py:286-297."""

    def __init__(self, alg, init, output_scale, num_derivatives, d):
        self.alg = alg
        self.init = init
        self.output_scale = output_scale
        self.num_derivatives = num_derivatives
        self.n = num_derivatives + 1
        self.d = d
        self.a, self.q_sqrtm = linalg.system_matrices_1d_iwp(num_derivatives)
        self.precon = linalg.preconditioner_taylor(num_derivatives)

    def transition(self, dt, output_scale):
        return self.alg.transition(self, dt, output_scale)
