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
