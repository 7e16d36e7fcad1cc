"""GPU parity for SURVEY §8(f) rank 1: the smoothing posterior (terminal marginal + backward conditionals) and
`loss_lml_timeseries` (probdiffeq/_probdiffeq/estimators_and_losses.py:53-105, 180-218), against the oracle.

The setup follows tests/test_probdiffeq/test_losses/test_lml_timeseries.py:19-46 (Lotka-Volterra, ts0, fixed-point
smoother, save_at grid of 13 points, data = solution mean + noise); the fixed-interval smoother on a fixed grid is the
README's recommendation for parameter estimation (README.md:200)."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from oracle import probdiffeq as o_pdq

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _oracle_natural_conditionals(post):
    out = []
    for c in post.conditional:
        cn = c.alg.preconditioner_apply(c)
        out.append((cn.A, cn.noise.mean, cn.noise.chol))
    return out


def _data(rng, osol_mean, fact, T, d):
    data = np.asarray(osol_mean)[:, 0] + 0.05 * rng.normal(size=(T, d))
    sd = 0.05 + 0.01 * np.arange(T)
    std = sd if fact == "isotropic" else np.stack([sd * (1 + 0.5 * j) for j in range(d)], axis=1)
    return data, std


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
@pytest.mark.parametrize("solver", ["solver", "solver_mle"])
def test_fixedpoint_posterior_and_lml_timeseries(cuda, fact, solver):
    import torch

    s = H.spec(fact=fact, strategy="fixedpoint", solver=solver, error="residual_std", control="i", clip_dt=False)
    B, T, d = 5, 13, 2
    params, u0 = H.lv_ensemble(B, seed=41)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 4.0, T)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-4, rtol=1e-4
    )
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    post = sol.solution_full.posterior
    assert isinstance(post, p_pdq.MarkovSequence)
    tc = tcoeffs.cpu().numpy()
    rng = np.random.Generator(np.random.PCG64(5))
    datas, stds, expected = [], [], {True: [], False: []}
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-4, 1e-4)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps))
        # the conditionals, in natural coordinates (the reference compares them the same way,
        # test_smoother_fixedinterval_vs_fixedpoint.py:78-86)
        for k, (A, xi, Xi) in enumerate(_oracle_natural_conditionals(opost), start=1):
            g = post.conditional.gain[b, k].cpu().numpy()
            m = post.conditional.mean[b, k].cpu().numpy()
            L = post.conditional.cholesky[b, k].cpu().numpy()
            assert _rel(g, A) < 1e-7, (k, _rel(g, A))
            assert _rel(m, xi if fact == "isotropic" else xi.T) < 1e-6 or np.max(np.abs(m - (xi if fact == "isotropic" else xi.T))) < 1e-9
            assert _rel(L @ np.swapaxes(L, -1, -2), Xi @ np.swapaxes(Xi, -1, -2)) < 1e-6 or np.max(np.abs(Xi)) < 1e-12
        data, std = _data(rng, osol.u_mean, fact, T, d)
        datas.append(data)
        stds.append(std)
        for avg in (True, False):
            expected[avg].append(o_pdq.loss_lml_timeseries(average_pdfs=avg)(data, posterior=opost, std=std))
    datas, stds = np.stack(datas), np.stack(stds)
    for avg in (True, False):
        got = p_pdq.loss_lml_timeseries(average_pdfs=avg)(datas, posterior=post, std=stds).cpu().numpy()
        assert got.shape == (B,)
        assert np.allclose(got, expected[avg], rtol=1e-7, atol=1e-9), (avg, got, expected[avg])
    # shared data / std broadcast over the ensemble; a wrong container raises like the reference
    got = p_pdq.loss_lml_timeseries()(datas[0], posterior=post, std=stds[0]).cpu().numpy()
    assert np.isclose(got[0], expected[True][0], rtol=1e-7)
    with pytest.raises(ValueError, match="container differs"):
        p_pdq.loss_lml_timeseries()(datas[0], posterior=post, std=stds[0][:-1])
    with pytest.raises(TypeError, match="datatype"):
        p_pdq.loss_lml_timeseries()(datas[0], posterior=sol.u, std=stds[0])


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
@pytest.mark.parametrize("strategy", ["fixedinterval", "fixedinterval_aligned"])
def test_fixedinterval_lml_timeseries_on_a_fixed_grid(cuda, fact, strategy):
    import torch

    s = H.spec(fact=fact, strategy=strategy, solver="solver_mle")
    B, T, d = 3, 21, 2
    params, u0 = H.lv_ensemble(B, seed=42)
    p_pdq, p_ivp, vf, ssm, slv, _e, _c = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    grid = np.linspace(0.0, 1.0, T)
    sol = p_ivp.solve_fixed_grid(solver=slv)(ssm.prior_wiener_integrated(tcoeffs), grid=grid)
    torch.cuda.synchronize()
    post = sol.solution_full.posterior
    tc = tcoeffs.cpu().numpy()
    rng = np.random.Generator(np.random.PCG64(6))
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tc[b], params[b], grid)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        data, std = _data(rng, osol.u_mean, fact, T, d)
        ref = o_pdq.loss_lml_timeseries()(data, posterior=opost, std=std)
        one = p_pdq.MarkovSequence(
            marginal=p_pdq.Normal(post.marginal.factorisation, post.marginal.mean_flat[b], post.marginal.cholesky_flat[b]),
            conditional=p_pdq.BackwardConditional(post.conditional.gain[b], post.conditional.mean[b], post.conditional.cholesky[b]),
        )  # fmt: skip
        got = p_pdq.loss_lml_timeseries()(data, posterior=one, std=std)  # unbatched posterior -> scalar
        assert got.shape == ()
        assert np.isclose(float(got), ref, rtol=1e-7, atol=1e-9), (float(got), ref)


def test_unbatched_solve_returns_an_unbatched_posterior(cuda):
    s = H.spec(fact="isotropic", strategy="fixedpoint", clip_dt=False, error="residual_std", control="i")
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, H.BASE_LV)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (np.asarray([20.0, 20.0]),), t=0.0)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0, 2, 5), atol=1e-3, rtol=1e-3
    )
    post = sol.solution_full.posterior
    assert post.marginal.mean_flat.shape == (5, 2) and post.conditional.gain.shape == (5, 5, 5)
    lml = p_pdq.loss_lml_timeseries()(sol.u.mean[0], posterior=post, std=np.ones(5))
    assert lml.shape == () and np.isfinite(float(lml))
    filt = H.product_build(H.spec(fact="isotropic"), H.BASE_LV)
    sol_f = filt[1].solve_adaptive_save_at(solver=filt[4], error=filt[5], control=filt[6])(
        ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0, 2, 5), atol=1e-3, rtol=1e-3
    )
    assert sol_f.solution_full is None


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
def test_posterior_samples_match_the_oracle_given_the_same_draws(cuda, fact):
    """MarkovSequence.sample (estimators_and_losses.py:233-271; shapes as in tests/test_probdiffeq/test_sample.py:
    36-45): with the standard-normal draws supplied, the sampled trajectories equal the oracle's."""
    import torch

    s = H.spec(fact=fact, strategy="fixedpoint", solver="solver_mle", error="residual_std", control="i", clip_dt=False)
    B, T, d, n = 3, 12, 2, 5
    params, u0 = H.lv_ensemble(B, seed=43)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, T)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-3, rtol=1e-3
    )
    post = sol.solution_full.posterior
    rng = np.random.Generator(np.random.PCG64(7))
    core = (T, n) if fact == "isotropic" else (T, d, n)
    base = rng.normal(size=(B, 2, 3, *core))
    smp = post.sample(base=base)
    torch.cuda.synchronize()
    assert len(smp) == n and smp[0].shape == (B, 2, 3, T, d)
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-3, 1e-3)
        opost = osol.solution_full.posterior
        # A Cholesky factor is unique only up to the signs of its columns (LAPACK's reflector signs depend on the
        # row order of the stacked matrix, and the product triangularises a row-permuted stack), and L eps depends
        # on them. Feed the oracle the same draws with the product's column signs.
        onat = _oracle_natural_conditionals(opost.remove_filtering_distributions())
        omarg = opost.remove_filtering_distributions().marginal.chol
        sign = np.ones((T, *core[1:]))

        def colsign(Lp, Lo):
            sp, so = np.sign(np.diagonal(Lp, axis1=-1, axis2=-2)), np.sign(np.diagonal(Lo, axis1=-1, axis2=-2))
            return np.where(sp * so == 0, 1.0, sp * so)

        sign[T - 1] = colsign(post.marginal.cholesky_flat[b].cpu().numpy(), omarg)
        for k in range(1, T):
            sign[k - 1] = colsign(post.conditional.cholesky[b, k].cpu().numpy(), onat[k - 1][2])
        for idx in ((0, 0), (1, 2)):
            ref = opost.sample(base[b][idx] * sign)  # list over grid points of (n, d) / (d, n)
            ref = np.stack([x if fact == "isotropic" else x.T for x in ref])  # (T, n, d)
            got = smp.flat[b][idx].cpu().numpy()
            for i in range(n):
                assert _rel(got[:, i], ref[:, i]) < (1e-7 if i <= 1 else 1e-4), (b, idx, i)
    # seeded draws: reproducible, of the requested shape, and centred on the smoothing mean
    s1 = post.sample(3, shape=(256,))
    s2 = post.sample(3, shape=(256,))
    assert s1[0].shape == (B, 256, T, d) and torch.equal(s1.flat, s2.flat)
    assert not torch.equal(s1.flat, post.sample(4, shape=(256,)).flat)
    dev = (s1[0].mean(dim=1) - sol.u.mean[0]).abs() / (sol.u.std[0].reshape(B, T, -1) + 1e-12)
    assert float(dev.max()) < 0.5  # |sample mean - mean| well within one standard deviation
    one = post.sample(1)
    assert one[0].shape == (B, T, d)
