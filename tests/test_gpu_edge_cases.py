"""GPU edge cases and size-independent properties: empty / ragged / unbatched ensembles, single checkpoints,
inexact initial conditions and prior scales, the attempt guard, and BASELINE's full ensemble size."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from oracle import ivpsolve as o_ivp
from oracle import probdiffeq as o_pdq

pytestmark = pytest.mark.gpu


def _headline(params, u0, **solve_kw):
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(H.spec(), params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    solve = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl)
    return solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, atol=1e-8, rtol=1e-6, **solve_kw), tcoeffs


def test_empty_ensemble(cuda):
    params, u0 = H.lv_ensemble(4, seed=0)
    sol, _ = _headline(params[:0], u0[:0], t1=1.0)
    assert sol.t.shape == (0,) and sol.u.mean_flat.shape == (0, 5, 2) and sol.num_steps.shape == (0,)


def test_unbatched_inputs_are_squeezed(cuda):
    import torch

    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(H.spec(), H.BASE_LV)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (np.asarray([20.0, 20.0]),), t=0.0)
    assert tcoeffs.shape == (5, 2)
    sol = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6
    )
    torch.cuda.synchronize()
    # BASELINE config 1: the reference's own single-IVP case; the oracle takes 331 accepted steps
    assert sol.u.mean_flat.shape == (5, 2) and sol.t.shape == () and int(sol.status) == 0
    ovf = o_pdq.ode("lotka_volterra")
    ossm = o_pdq.state_space_model_isotropic()
    ots0 = ossm.constraint_ode_ts0(ovf)
    osolve = o_ivp.solve_adaptive_terminal_values(
        solver=o_pdq.solver(strategy=o_pdq.strategy_filter(), constraint=ots0),
        error=o_pdq.error_state_std(constraint=ots0), control=o_ivp.control_proportional_integral(),
    )  # fmt: skip
    osol = osolve(ossm.prior_wiener_integrated(tcoeffs.cpu().numpy()), t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6)
    assert int(sol.num_steps) == int(osol.num_steps) == 331
    assert np.allclose(sol.u.mean[0].cpu().numpy(), osol.u.tcoeffs[0], rtol=1e-6)


def test_single_checkpoint_returns_the_initial_condition(cuda):
    params, u0 = H.lv_ensemble(5, seed=1)
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(H.spec(clip_dt=False), params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    sol = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=np.asarray([0.25]), atol=1e-6, rtol=1e-4
    )
    assert sol.t.shape == (5, 1) and np.all(sol.t.cpu().numpy() == 0.25)
    assert np.array_equal(sol.u.mean_flat[:, 0].cpu().numpy(), tcoeffs.cpu().numpy())
    assert int(sol.num_attempts.sum()) == 0 and int(sol.status.abs().max()) == 0


def test_ragged_ensemble_and_independence_of_instances(cuda):
    """An ensemble whose size is not a multiple of the CTA size; every instance is solved exactly as it is alone."""
    params, u0 = H.lv_ensemble(1037, seed=2)
    full, _ = _headline(params, u0, t1=5.0)
    part, _ = _headline(params[1000:], u0[1000:], t1=5.0)
    assert int(full.status.abs().max()) == 0
    assert np.array_equal(full.u.mean_flat[1000:].cpu().numpy(), part.u.mean_flat.cpu().numpy())
    assert np.array_equal(full.u.cholesky_flat[1000:].cpu().numpy(), part.u.cholesky_flat.cpu().numpy())
    assert np.array_equal(full.num_steps[1000:].cpu().numpy(), part.num_steps.cpu().numpy())


def test_full_baseline_size_properties(cuda):
    """BASELINE config 2 at its full 2^20 instances: no failures, permutation equivariance, bitwise agreement with a
    sub-ensemble, step counts and terminal times in range, agreement of a sample with the oracle."""
    import torch

    B = 1 << 20
    params, u0 = H.lv_ensemble(B, seed=0)
    sol, tcoeffs = _headline(params, u0, t1=50.0, want_cholesky=False)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    steps = sol.num_steps.cpu().numpy()
    assert steps.min() > 100 and steps.max() < 2000
    assert np.all(np.abs(sol.t.cpu().numpy() - 50.0) < 1e-9)
    assert np.all(sol.num_attempts.cpu().numpy() >= steps)
    # permutation equivariance + independence: a shuffled 4096-instance subsample reproduces its rows bit for bit
    idx = np.random.Generator(np.random.PCG64(9)).permutation(B)[:4096]
    sub, _ = _headline(params[idx], u0[idx], t1=50.0, want_cholesky=False)
    assert np.array_equal(sub.u.mean_flat.cpu().numpy(), sol.u.mean_flat.cpu().numpy()[idx])
    assert np.array_equal(sub.num_steps.cpu().numpy(), steps[idx])
    # a sample against the oracle (step counts; terminal values within the oracle's own conditioning at t = 50)
    tc = tcoeffs.cpu().numpy()
    for b in (0, 12345, B - 1):
        osol, trace = H.oracle_solve_save_at(H.spec(), tc[b], params[b], np.asarray([0.0, 50.0]), 1e-8, 1e-6)
        assert int(steps[b]) == int(osol.num_steps[-1]) and int(sol.num_attempts[b]) == len(trace)
        assert np.allclose(sol.u.mean[0][b].cpu().numpy(), osol.u_mean[-1][0], rtol=1e-5)


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
def test_inexact_initial_condition_and_prior_scale(cuda, fact):
    """prior_wiener_integrated(is_exact=False, output_scale=...): non-zero initial Cholesky factor and base scale."""
    import torch

    B = 4
    params, u0 = H.lv_ensemble(B, seed=6)
    s = H.spec(fact=fact, solver="solver_mle", error="residual_std", control="i")
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    scale = 2.5 if fact == "isotropic" else np.asarray([2.5, 0.5])
    prior = ssm.prior_wiener_integrated(tcoeffs, is_exact=False, inexact_eps=1e-3, output_scale=scale)
    sol = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl)(
        prior, t0=0.0, t1=3.0, atol=1e-7, rtol=1e-5
    )
    torch.cuda.synchronize()
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        ovf = o_pdq.ode("lotka_volterra", params[b])
        ossm = getattr(o_pdq, "state_space_model_" + fact)()
        ocons = ossm.constraint_ode_ts0(ovf)
        oprior = ossm.prior_wiener_integrated(tc[b], is_exact=False, inexact_eps=1e-3, output_scale=scale)
        osolve = o_ivp.solve_adaptive_terminal_values(
            solver=o_pdq.solver_mle(strategy=o_pdq.strategy_filter(), constraint=ocons),
            error=o_pdq.error_residual_std(constraint=ocons), control=o_ivp.control_integral(),
        )  # fmt: skip
        osol = osolve(oprior, t0=0.0, t1=3.0, atol=1e-7, rtol=1e-5)
        assert int(sol.num_steps[b]) == int(osol.num_steps)
        assert np.allclose(sol.u.mean_flat[b].cpu().numpy(), osol.u.tcoeffs, rtol=1e-7, atol=1e-9)
        L, Lo = sol.u.cholesky_flat[b].cpu().numpy(), osol.u.chol
        assert np.allclose(L @ np.swapaxes(L, -1, -2), Lo @ np.swapaxes(Lo, -1, -2), rtol=1e-5, atol=1e-16)
        assert np.allclose(sol.output_scale[b].cpu().numpy(), osol.output_scale, rtol=1e-6)


def test_attempt_guard_reports_status(cuda):
    params, u0 = H.lv_ensemble(8, seed=7)
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(H.spec(), params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    solve = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl, max_attempts=5)
    sol = solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6)
    assert np.all(sol.status.cpu().numpy() == 2) and np.all(sol.num_attempts.cpu().numpy() == 5)


@pytest.mark.parametrize("fact,strategy", [("isotropic", "filter"), ("blockdiag", "fixedpoint"), ("dense", "filter"),
                                           ("dense", "fixedpoint")])  # fmt: skip
def test_abandoned_checkpoints_read_nan_not_garbage(cuda, fact, strategy):
    """An instance that hits max_attempts reports status 2, and the checkpoints it never reached are NaN in every
    kernel (thread-per-instance: written by the kernel; lane-per-dimension and dense: the host pre-fills)."""
    s = H.spec(fact=fact, strategy=strategy, clip_dt=False, error="residual_std", control="i")
    params, u0 = H.lv_ensemble(4, seed=9)
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    solve = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl, max_attempts=6, warn=False)
    sol = solve(ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0.0, 50.0, 6), atol=1e-8, rtol=1e-6)
    assert np.all(sol.status.cpu().numpy() == 2) and np.all(sol.num_attempts.cpu().numpy() == 6)
    mean = sol.u.mean_flat.cpu().numpy()
    assert np.all(np.isfinite(mean[:, 0])) and np.all(np.isnan(mean[:, -1]))
    assert np.all(np.isnan(sol.t[:, -1].cpu().numpy()))


def test_dt0_adaptive_matches_oracle(cuda):
    params, u0 = H.lv_ensemble(16, seed=8)
    p_pdq, p_ivp, vf, *_ = H.product_build(H.spec(), params)
    got = p_ivp.dt0_adaptive(vf, (u0,), 0.0, error_contraction_rate=5, rtol=1e-6, atol=1e-8).cpu().numpy()
    for b in range(16):
        ref = o_ivp.dt0_adaptive(o_pdq.ode("lotka_volterra", params[b]), (u0[b],), 0.0, error_contraction_rate=5,
                                 rtol=1e-6, atol=1e-8)  # fmt: skip
        assert abs(got[b] - ref) <= 1e-12 * ref


def test_unsupported_configurations_raise(cuda):
    from probdiffeq_b200 import ivpsolve, probdiffeq

    vf = probdiffeq.ode("pleiades")
    ssm = probdiffeq.state_space_model_dense()  # dense Pleiades is not instantiated
    tcoeffs = np.zeros((2, 6, 28))
    ts0 = ssm.constraint_ode_ts0(vf)
    solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=probdiffeq.error_residual_std(constraint=ts0))
    with pytest.raises(ValueError, match="no kernel"):
        solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=1.0, atol=1e-3, rtol=1e-3)
    with pytest.raises(ValueError):
        ssm.prior_wiener_integrated(tcoeffs, output_scale=np.ones(3))


def test_taylor_coefficient_increment(cuda):
    """jetexpand_ode_coefficient_increment (jet_expansion_algorithms.py:155-177): one more coefficient per call."""
    from probdiffeq_b200 import probdiffeq

    params, u0 = H.lv_ensemble(6, seed=9)
    vf = probdiffeq.ode("lotka_volterra", params=params)
    full, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    inc = probdiffeq.jetexpand_ode_coefficient_increment(num_arguments=1)
    tc = full[:, :2]
    for k in range(3, 6):
        tc = inc(vf, tc, t=0.0)
        assert tc.shape == (6, k, 2)
        assert np.array_equal(tc.cpu().numpy(), full[:, :k].cpu().numpy())
    for b in range(2):
        ref = o_pdq.ode("lotka_volterra", params[b]).taylor_coefficients((u0[b],), 0.0, 4)
        assert np.allclose(tc[b].cpu().numpy(), ref, rtol=1e-13)
    with pytest.raises(ValueError):
        probdiffeq.jetexpand_ode_coefficient_increment(num_arguments=2)(vf, tc)
