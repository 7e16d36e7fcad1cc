"""GPU parity for the dense factorisation (K3: CTA per instance, cooperative shared-memory Householder)."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from oracle import problems as o_problems
from test_gpu_group_and_smoother import _run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("combo", [
    dict(constraint="ts0", solver="solver", error="residual_std", control="i"),
    dict(constraint="ts1", solver="solver", error="state_std", control="pi"),
    dict(constraint="ts1", solver="solver_dynamic", error="residual_std", control="pi"),
    dict(constraint="ts1", solver="solver_mle", error="residual_std", control="i"),
    dict(constraint="ts0", solver="solver_dynamic", error="state_std", control="i", error_norm="rms_then_scale"),
], ids=lambda c: "-".join(str(v) for v in c.values()))  # fmt: skip
def test_dense_lotka_volterra_terminal(cuda, combo):
    s = H.spec(fact="dense", clip_dt=True, **combo)
    params, u0 = H.lv_ensemble(5, seed=21)
    _run_case(s, params, (u0,), 4, np.asarray([0.0, 5.0]), 1e-8, 1e-6, terminal=True)


def test_dense_lotka_volterra_save_at_with_interpolation(cuda):
    s = H.spec(fact="dense", clip_dt=False, constraint="ts1", solver="solver_dynamic", error="residual_std",
               control="i")  # fmt: skip
    params, u0 = H.lv_ensemble(4, seed=22)
    _run_case(s, params, (u0,), 3, np.linspace(0.0, 3.0, 16), 1e-7, 1e-5)


def test_dense_hires_ts1(cuda):
    """BASELINE config 4a wiring (short horizon): HIRES d = 8, nu = 5, dense ts1 with the exact Jacobian, filter,
    solver_dynamic + error_residual_std + PI control, terminal values."""
    s = H.spec(vf="hires", fact="dense", constraint="ts1", solver="solver_dynamic", error="residual_std",
               control="pi", clip_dt=True)  # fmt: skip
    B = 3
    rng = np.random.Generator(np.random.PCG64(2))
    u0 = np.repeat(o_problems.hires_u0()[None, :], B, axis=0)
    scale = rng.uniform(0.9, 1.1, size=(B, 2))
    u0[:, 0] *= scale[:, 0]
    u0[:, 7] *= scale[:, 1]
    _run_case(s, None, (u0,), 5, np.asarray([0.0, 2.0]), 1e-9, 1e-6, dt0=1e-3, terminal=True)


def test_dense_fixed_grid(cuda):
    import torch

    s = H.spec(fact="dense", constraint="ts1", solver="solver")
    B = 3
    params, u0 = H.lv_ensemble(B, seed=23)
    p_pdq, p_ivp, vf, ssm, slv, _e, _c = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    grid = np.linspace(0.0, 1.0, 21)
    sol = p_ivp.solve_fixed_grid(solver=slv)(ssm.prior_wiener_integrated(tcoeffs), grid=grid)
    torch.cuda.synchronize()
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tc[b], params[b], grid)
        got_m = sol.u.mean_flat[b].cpu().numpy()
        assert np.max(np.abs(got_m - osol.u_mean)) / np.max(np.abs(osol.u_mean)) < 1e-10
        L = sol.u.cholesky_flat[b].cpu().numpy()
        cov, ocov = L @ np.swapaxes(L, -1, -2), osol.u_chol @ np.swapaxes(osol.u_chol, -1, -2)
        for k in range(len(grid)):
            assert np.max(np.abs(cov[k] - ocov[k])) <= 1e-10 * max(np.max(np.abs(ocov[k])), 1e-300), k


def test_dense_lml_terminal_values(cuda):
    """loss_lml_terminal_values on a dense marginal (estimators_and_losses.py:20-50 with ssm_impl_dense.py:108-234)."""
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq

    s = H.spec(fact="dense", constraint="ts1", solver="solver_mle", error="residual_std", control="i")
    B = 4
    params, u0 = H.lv_ensemble(B, seed=61)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    sol = p_ivp.solve_adaptive_terminal_values(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=1.0, atol=1e-5, rtol=1e-3
    )
    rng = np.random.Generator(np.random.PCG64(8))
    data = sol.u.mean[0].cpu().numpy() + 1e-2 * rng.normal(size=(B, 2))
    std = np.asarray([1e-2, 3e-2])
    for idx in (0, 1):
        obs = data if idx == 0 else sol.u.mean[1].cpu().numpy() + 1e-2 * rng.normal(size=(B, 2))
        got = p_pdq.loss_lml_terminal_values(tcoeff_index=idx)(obs, marginals=sol.u, std=std).cpu().numpy()
        torch.cuda.synchronize()
        for b in range(B):
            osol, _ = H.oracle_solve_save_at(s, tcoeffs[b].cpu().numpy(), params[b], np.asarray([0.0, 1.0]), 1e-5, 1e-3)
            ref = o_pdq.loss_lml_terminal_values(tcoeff_index=idx)(obs[b], marginals=osol.u[-1], std=std)
            assert np.isclose(got[b], ref, rtol=1e-7, atol=1e-9), (idx, b, got[b], ref)
