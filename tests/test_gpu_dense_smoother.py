"""GPU parity for the smoothers of the DENSE factorisation (csrc/pdeq_smooth_dense.cuh; reference:
probdiffeq/_probdiffeq/ssm_impl_dense.py:24-84 under estimators_and_losses.py:437-717): the tests of the isotropic /
block-diagonal smoothers (test_gpu_group_and_smoother.py, test_gpu_lml_timeseries.py) with fact="dense", same
tolerance policy, plus one case at the size of BASELINE config 4a's state (HIRES, N = 48)."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from test_gpu_group_and_smoother import _cov, _rel, _run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("combo", [dict(solver="solver", error="residual_std", control="i"),
                                   dict(solver="solver_dynamic", error="residual_std", control="i"),
                                   dict(solver="solver_mle", error="state_std", control="pi", constraint="ts1")],
                         ids=["plain", "dynamic", "mle-ts1"])  # fmt: skip
def test_dense_fixedpoint_smoother_lotka_volterra(cuda, combo):
    s = H.spec(fact="dense", strategy="fixedpoint", clip_dt=False, **combo)
    params, u0 = H.lv_ensemble(6, seed=11)
    _run_case(s, params, (u0,), 4, np.linspace(0.0, 4.0, 13), 1e-7, 1e-5)


def test_dense_fixedpoint_smoother_with_clipping(cuda):
    s = H.spec(fact="dense", strategy="fixedpoint", clip_dt=True, solver="solver_dynamic", error="residual_std",
               control="i")  # fmt: skip
    params, u0 = H.lv_ensemble(4, seed=12)
    _run_case(s, params, (u0,), 4, np.linspace(0.0, 3.0, 4), 1e-7, 1e-5)


@pytest.mark.parametrize("solver", ["solver", "solver_mle"])
@pytest.mark.parametrize("strategy", ["fixedinterval", "fixedinterval_aligned"])
def test_dense_fixedinterval_smoother_on_a_fixed_grid(cuda, solver, strategy):
    import torch

    s = H.spec(fact="dense", strategy=strategy, solver=solver)
    B = 4
    params, u0 = H.lv_ensemble(B, seed=31)
    p_pdq, p_ivp, vf, ssm, slv, _e, _c = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    grid = np.linspace(0.0, 1.0, 26)
    sol = p_ivp.solve_fixed_grid(solver=slv)(ssm.prior_wiener_integrated(tcoeffs), grid=grid)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tc[b], params[b], grid)
        pert = H.oracle_solve_fixed(s, tc[b] * (1.0 + 2.3e-16), params[b], grid)
        got, ref, prt = sol.u.mean_flat[b].cpu().numpy(), np.asarray(osol.u_mean), np.asarray(pert.u_mean)
        for i in range(ref.shape[1]):
            tol = max(1e-10, 100 * _rel(prt[:, i], ref[:, i]))
            assert tol < (1e-8 if i <= 2 else 1e-4), (i, tol)
            assert _rel(got[:, i], ref[:, i]) < tol, (i, _rel(got[:, i], ref[:, i]), tol)
        L = sol.u.cholesky_flat[b].cpu().numpy()
        for k in range(len(grid)):
            assert _rel(_cov(L[k]), _cov(osol.u_chol[k])) < 1e-8, k


@pytest.mark.parametrize("solver", ["solver", "solver_mle"])
def test_dense_fixedpoint_posterior_conditionals(cuda, solver):
    """The posterior a dense smoother returns: terminal marginal + backward conditionals in natural coordinates
    (compared like the reference compares them, test_smoother_fixedinterval_vs_fixedpoint.py:78-86)."""
    import torch

    s = H.spec(fact="dense", strategy="fixedpoint", solver=solver, error="residual_std", control="i", clip_dt=False)
    B, T = 4, 9
    params, u0 = H.lv_ensemble(B, seed=41)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, T)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-4, rtol=1e-4
    )
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    post = sol.solution_full.posterior
    assert post.conditional.gain.shape == (B, T, 10, 10) and post.conditional.mean.shape == (B, T, 5, 2)
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-4, 1e-4)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps))
        for k, c in enumerate(opost.conditional, start=1):
            cn = c.alg.preconditioner_apply(c)
            g = post.conditional.gain[b, k].cpu().numpy()
            m = post.conditional.mean[b, k].cpu().numpy().reshape(-1)
            L = post.conditional.cholesky[b, k].cpu().numpy()
            assert _rel(g, cn.A) < 1e-7, (k, _rel(g, cn.A))
            assert _rel(m, cn.noise.mean) < 1e-6 or np.max(np.abs(m - cn.noise.mean)) < 1e-9
            assert _rel(_cov(L), _cov(cn.noise.chol)) < 1e-6 or np.max(np.abs(cn.noise.chol)) < 1e-12
        om = opost.marginal
        assert _rel(post.marginal.mean_flat[b].cpu().numpy().reshape(-1), om.mean) < 1e-7


def test_dense_smoother_at_the_size_of_config_4a(cuda):
    """HIRES (d = 8, nu = 5: N = 48, the state of BASELINE config 4a), ts1, solver_dynamic + error_residual_std, with
    the fixed-point smoother: the 96 x 96 reverted transition and the 190 KB of shared memory per instance."""
    from oracle import problems as o_problems

    s = H.spec(vf="hires", fact="dense", constraint="ts1", strategy="fixedpoint", solver="solver_dynamic",
               error="residual_std", control="pi", clip_dt=False)  # fmt: skip
    u0 = np.repeat(o_problems.hires_u0()[None, :], 2, axis=0)
    u0[1, 0] *= 0.95
    _run_case(s, None, (u0,), 5, np.asarray([0.0, 0.205, 0.5]), 1e-8, 1e-5, dt0=1e-4, min_stable=1)


# ------------------------------------------------------------------------------------------------------
# What hangs off a dense smoothing solution (pdeq_aux_dense.cuh): loss_lml_timeseries, posterior.sample,
# solver.offgrid_marginals -- the isotropic / block-diagonal tests of test_gpu_lml_timeseries.py and test_gpu_offgrid.py
# with fact="dense" (the oracle's dense state is flat, coefficient-major: (n d,) means, (n d, n d) factors).
# ------------------------------------------------------------------------------------------------------
def _data(rng, osol_mean, T, d):
    data = np.asarray(osol_mean)[:, 0] + 0.05 * rng.normal(size=(T, d))
    sd = 0.05 + 0.01 * np.arange(T)
    return data, np.stack([sd * (1 + 0.5 * j) for j in range(d)], axis=1)


@pytest.mark.parametrize("solver", ["solver", "solver_mle"])
def test_dense_lml_timeseries_fixedpoint(cuda, solver):
    import torch

    from oracle import probdiffeq as o_pdq

    s = H.spec(fact="dense", strategy="fixedpoint", solver=solver, error="residual_std", control="i", clip_dt=False)
    B, T, d = 4, 11, 2
    params, u0 = H.lv_ensemble(B, seed=61)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, T)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-4, rtol=1e-4
    )
    torch.cuda.synchronize()
    post = sol.solution_full.posterior
    tc = tcoeffs.cpu().numpy()
    rng = np.random.Generator(np.random.PCG64(8))
    datas, stds, expected = [], [], {True: [], False: []}
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-4, 1e-4)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps))
        data, std = _data(rng, np.asarray(osol.u_mean).reshape(T, 5, d), T, d)
        datas.append(data)
        stds.append(std)
        for avg in (True, False):
            expected[avg].append(o_pdq.loss_lml_timeseries(average_pdfs=avg)(data, posterior=opost, std=std))
    datas, stds = np.stack(datas), np.stack(stds)
    for avg in (True, False):
        got = p_pdq.loss_lml_timeseries(average_pdfs=avg)(datas, posterior=post, std=stds).cpu().numpy()
        assert got.shape == (B,)
        assert np.allclose(got, expected[avg], rtol=1e-7, atol=1e-9), (avg, got, expected[avg])
    got1 = p_pdq.loss_lml_timeseries(tcoeff_index=1)(datas, posterior=post, std=stds).cpu().numpy()
    assert np.all(np.isfinite(got1)) and not np.allclose(got1, expected[True])


def test_dense_lml_timeseries_fixedinterval_on_a_fixed_grid(cuda):
    import torch

    from oracle import probdiffeq as o_pdq

    s = H.spec(fact="dense", strategy="fixedinterval", solver="solver_mle")
    B, T, d = 3, 17, 2
    params, u0 = H.lv_ensemble(B, seed=62)
    p_pdq, p_ivp, vf, ssm, slv, _e, _c = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    grid = np.linspace(0.0, 1.0, T)
    sol = p_ivp.solve_fixed_grid(solver=slv)(ssm.prior_wiener_integrated(tcoeffs), grid=grid)
    torch.cuda.synchronize()
    post = sol.solution_full.posterior
    tc = tcoeffs.cpu().numpy()
    rng = np.random.Generator(np.random.PCG64(9))
    datas, stds, refs = [], [], []
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tc[b], params[b], grid)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        data, std = _data(rng, np.asarray(osol.u_mean).reshape(T, 5, d), T, d)
        datas.append(data)
        stds.append(std)
        refs.append(o_pdq.loss_lml_timeseries()(data, posterior=opost, std=std))
    got = p_pdq.loss_lml_timeseries()(np.stack(datas), posterior=post, std=np.stack(stds)).cpu().numpy()
    assert np.allclose(got, refs, rtol=1e-7, atol=1e-9), (got, refs)


def test_dense_posterior_samples_match_the_oracle_given_the_same_draws(cuda):
    import torch

    s = H.spec(fact="dense", strategy="fixedpoint", solver="solver_mle", error="residual_std", control="i",
               clip_dt=False)  # fmt: skip
    B, T, d, n = 3, 10, 2, 5
    N = n * d
    params, u0 = H.lv_ensemble(B, seed=63)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, T)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-3, rtol=1e-3
    )
    post = sol.solution_full.posterior
    rng = np.random.Generator(np.random.PCG64(10))
    base = rng.normal(size=(B, 2, T, N))
    smp = post.sample(base=base)
    torch.cuda.synchronize()
    assert len(smp) == n and smp[0].shape == (B, 2, T, d)
    tc = tcoeffs.cpu().numpy()

    def colsign(Lp, Lo):  # a factor is unique up to the signs of its columns; feed the oracle the product's
        sp, so = np.sign(np.diagonal(Lp)), np.sign(np.diagonal(Lo))
        return np.where(sp * so == 0, 1.0, sp * so)

    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-3, 1e-3)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        sign = np.ones((T, N))
        sign[T - 1] = colsign(post.marginal.cholesky_flat[b].cpu().numpy(), opost.marginal.chol)
        for k, c in enumerate(opost.conditional, start=1):
            sign[k - 1] = colsign(post.conditional.cholesky[b, k].cpu().numpy(), c.alg.preconditioner_apply(c).noise.chol)
        for idx in (0, 1):
            ref = np.stack(opost.sample(base[b][idx] * sign)).reshape(T, n, d)
            got = smp.flat[b][idx].cpu().numpy()
            for i in range(n):
                assert _rel(got[:, i], ref[:, i]) < (1e-7 if i <= 1 else 1e-4), (b, idx, i, _rel(got[:, i], ref[:, i]))
    s1 = post.sample(3, shape=(64,))
    assert s1[0].shape == (B, 64, T, d) and torch.equal(s1.flat, post.sample(3, shape=(64,)).flat)


@pytest.mark.parametrize("solver,constraint", [("solver_dynamic", "ts1"), ("solver_mle", "ts0")])
def test_dense_filter_offgrid_marginals(cuda, solver, constraint):
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq

    s = H.spec(fact="dense", solver=solver, constraint=constraint, error="residual_std", control="i", clip_dt=False)
    B = 3
    params, u0 = H.lv_ensemble(B, seed=64)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, 9)
    scale = np.asarray([1.5, 0.75])
    prior = ssm.prior_wiener_integrated(tcoeffs, output_scale=scale)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(prior, save_at=save_at, atol=1e-5, rtol=1e-4)
    ts = np.concatenate([0.5 * (save_at[1:] + save_at[:-1]), [0.01, 2.99, 1.2345]])
    rv = slv.offgrid_marginals(ts, solution=sol)
    torch.cuda.synchronize()
    assert rv.mean_flat.shape == (B, len(ts), 5, 2) and rv.cholesky_flat.shape == (B, len(ts), 10, 10)
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-5, 1e-4, output_scale=scale)
        _, oslv, _, _ = H._build(o_pdq, o_ivp, s, H.oracle_vf(s, params[b]))
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps))
        for k, t in enumerate(ts):
            orv = oslv.offgrid_marginals(t, solution=osol)
            m, got_m = orv.mean.reshape(5, 2), rv.mean_flat[b, k].cpu().numpy()
            for i in range(5):
                assert _rel(got_m[i], m[i]) < (1e-7 if i <= 1 else 1e-4), (b, k, i, _rel(got_m[i], m[i]))
            assert _rel(_cov(rv.cholesky_flat[b, k].cpu().numpy()), _cov(orv.chol)) < 1e-5, (b, k)


@pytest.mark.parametrize("strategy", ["fixedinterval", "fixedinterval_aligned"])
def test_dense_fixedinterval_offgrid_marginals(cuda, strategy):
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq

    s = H.spec(fact="dense", strategy=strategy, solver="solver_mle")
    B = 3
    params, u0 = H.lv_ensemble(B, seed=65)
    p_pdq, p_ivp, vf, ssm, slv, _e, _c = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    grid = np.linspace(0.0, 1.0, 17)
    sol = p_ivp.solve_fixed_grid(solver=slv)(ssm.prior_wiener_integrated(tcoeffs), grid=grid)
    assert sol.solution_full.filtering is not None
    ts = np.asarray([0.01, 0.33, 0.5 * (grid[7] + grid[8]), 0.97])
    rv = slv.offgrid_marginals(ts, solution=sol)
    torch.cuda.synchronize()
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tc[b], params[b], grid)
        _, oslv, _, _ = H._build(o_pdq, o_ivp, s, H.oracle_vf(s, params[b]))
        for k, t in enumerate(ts):
            orv = oslv.offgrid_marginals(t, solution=osol)
            m, got_m = orv.mean.reshape(5, 2), rv.mean_flat[b, k].cpu().numpy()
            for i in range(5):
                assert _rel(got_m[i], m[i]) < (1e-7 if i <= 2 else 1e-4), (b, k, i, _rel(got_m[i], m[i]))
            assert _rel(_cov(rv.cholesky_flat[b, k].cpu().numpy()), _cov(orv.chol)) < 1e-5, (b, k)
