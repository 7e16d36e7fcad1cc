"""user-supplied vector fields, compiled at run time into a plug-in (probdiffeq_b200/plugins.py).

(1) A clone of the built-in Lotka-Volterra functor must reproduce the built-in kernels bit for bit.
(2) A right-hand side that is not built in (logistic growth) is checked against the oracle like every other path."""

import numpy as np
import pytest

import pdeq_test_helpers as H

pytestmark = pytest.mark.gpu


def _solve(vf, u0, *, constraint="ts0", t1=2.0):
    from probdiffeq_b200 import ivpsolve, probdiffeq

    ssm = probdiffeq.state_space_model_isotropic()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    cons = getattr(ssm, "constraint_ode_" + constraint)(vf)
    solver = probdiffeq.solver_mle(strategy=probdiffeq.strategy_filter(), constraint=cons)
    error = probdiffeq.error_residual_std(constraint=cons)
    solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, control=ivpsolve.control_integral())
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    sol = solve(ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0.0, t1, 9), atol=1e-6, rtol=1e-4, dt0=dt0)
    return tcoeffs, dt0, sol


@pytest.mark.parametrize("constraint", ["ts0", "ts1"])
def test_plugin_clone_of_lotka_volterra_is_bitwise_the_builtin(cuda, constraint):
    import torch

    from probdiffeq_b200 import plugins, probdiffeq

    params, u0 = H.lv_ensemble(64, seed=71)
    spec = {k: v for k, v in plugins.LOTKA_VOLTERRA_CLONE.items()}
    vf_plugin = plugins.ode_from_cuda("lotka_volterra_user", params=params, **spec)
    vf_builtin = probdiffeq.ode("lotka_volterra", params=params)
    assert vf_plugin.vf_id >= 6 and vf_plugin.order == 1 and vf_plugin.num_params == 4
    tc_p, dt0_p, sol_p = _solve(vf_plugin, u0, constraint=constraint)
    tc_b, dt0_b, sol_b = _solve(vf_builtin, u0, constraint=constraint)
    torch.cuda.synchronize()
    assert torch.equal(tc_p, tc_b) and torch.equal(dt0_p, dt0_b)
    assert int(sol_p.status.abs().max()) == 0
    assert torch.equal(sol_p.num_steps, sol_b.num_steps)
    assert torch.equal(sol_p.u.mean_flat, sol_b.u.mean_flat)
    assert torch.equal(sol_p.u.cholesky_flat, sol_b.u.cholesky_flat)
    assert torch.equal(sol_p.output_scale, sol_b.output_scale)


@pytest.mark.parametrize("constraint", ["ts0", "ts1"])
def test_plugin_logistic_matches_the_oracle(cuda, constraint):
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq
    from probdiffeq_b200 import plugins

    B = 6
    rng = np.random.Generator(np.random.PCG64(72))
    params = np.stack([rng.uniform(0.5, 3.0, size=B), rng.uniform(2.0, 5.0, size=B)], axis=1)  # rate, capacity
    u0 = rng.uniform(0.1, 1.0, size=(B, 1))
    vf = plugins.ode_from_cuda("logistic", params=params, **plugins.LOGISTIC)
    tcoeffs, dt0, sol = _solve(vf, u0, constraint=constraint, t1=3.0)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    save_at = np.linspace(0.0, 3.0, 9)
    for b in range(B):
        ovf = o_pdq.ode("logistic", params[b])
        otc, _ = o_pdq.jetexpand_ode_padded_scan(num=4)(ovf, (u0[b],), t=0.0)
        assert np.allclose(tcoeffs[b].cpu().numpy(), otc, rtol=1e-13, atol=1e-15)
        ossm = o_pdq.state_space_model_isotropic()
        ocons = getattr(ossm, "constraint_ode_" + constraint)(ovf)
        osolver = o_pdq.solver_mle(strategy=o_pdq.strategy_filter(), constraint=ocons)
        oerr = o_pdq.error_residual_std(constraint=ocons)
        odt0 = o_ivp.dt0(ovf, (u0[b],), t=0.0)
        assert abs(float(dt0[b]) - odt0) <= 1e-13 * odt0
        osol = o_ivp.solve_adaptive_save_at(solver=osolver, error=oerr, control=o_ivp.control_integral(), warn=False)(
            ossm.prior_wiener_integrated(np.asarray(otc)), save_at=save_at, atol=1e-6, rtol=1e-4, dt0=odt0
        )
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps)), b
        got, ref = sol.u.mean_flat[b].cpu().numpy(), np.asarray(osol.u_mean)
        assert np.allclose(got[:, 0], ref[:, 0], rtol=1e-9, atol=1e-12)  # the solution itself
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-8)  # all Taylor coefficients
        exact = params[b, 1] / (1.0 + (params[b, 1] / u0[b, 0] - 1.0) * np.exp(-params[b, 0] * save_at))
        assert np.allclose(got[:, 0, 0], exact, rtol=1e-3)  # and the closed-form logistic curve


def test_plugin_with_runtime_dimension_and_smoother(cuda):
    """Lane-per-dimension kernels (run-time d) through a plug-in clone of `linear`: the filter is bitwise the built-in
    kernel; the fixed-point smoother with ts1 (not instantiated for the built-in) is checked against the oracle."""
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq
    from probdiffeq_b200 import ivpsolve, plugins, probdiffeq

    B, d = 3, 100
    rng = np.random.Generator(np.random.PCG64(73))
    params = rng.uniform(-1.5, -0.5, size=(B, 1))
    u0 = rng.uniform(0.5, 1.5, size=(B, d))
    save_at = np.linspace(0.0, 1.0, 5)
    vf_user = plugins.ode_from_cuda("linear_user", params=params, **plugins.LINEAR_CLONE)
    out = []
    for vf in (vf_user, probdiffeq.ode("linear", params=params)):
        ssm = probdiffeq.state_space_model_blockdiag()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0,), t=0.0)
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts0)
        error = probdiffeq.error_residual_std(constraint=ts0)
        sol = ivpsolve.solve_adaptive_save_at(solver=solver, error=error)(
            ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-6, rtol=1e-4
        )
        assert int(sol.status.abs().max()) == 0
        out.append(sol)
    torch.cuda.synchronize()
    assert torch.equal(out[0].num_steps, out[1].num_steps)
    assert torch.equal(out[0].u.mean_flat, out[1].u.mean_flat)
    assert torch.equal(out[0].u.cholesky_flat, out[1].u.cholesky_flat)
    # smoother + ts1, against the oracle and the exact solution
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf_user, (u0,), t=0.0)
    ts1 = ssm.constraint_ode_ts1(vf_user)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts1)
    error = probdiffeq.error_residual_std(constraint=ts1)
    sol = ivpsolve.solve_adaptive_save_at(solver=solver, error=error)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-6, rtol=1e-4
    )
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    exact = u0[:, None, :] * np.exp(params[:, :, None] * save_at[None, :, None])
    assert np.allclose(sol.u.mean[0].cpu().numpy(), exact, rtol=1e-4)
    ovf = o_pdq.ode("linear", params[0])
    ossm = o_pdq.state_space_model_blockdiag()
    ocons = ossm.constraint_ode_ts1(ovf)
    osol = o_ivp.solve_adaptive_save_at(
        solver=o_pdq.solver_dynamic(strategy=o_pdq.strategy_smoother_fixedpoint(), constraint=ocons),
        error=o_pdq.error_residual_std(constraint=ocons),
    )(ossm.prior_wiener_integrated(tcoeffs[0].cpu().numpy()), save_at=save_at, atol=1e-6, rtol=1e-4)
    assert np.array_equal(sol.num_steps[0, 1:].cpu().numpy(), np.asarray(osol.num_steps))
    assert np.allclose(sol.u.mean_flat[0, :, 0].cpu().numpy(), np.asarray(osol.u_mean)[:, 0], rtol=1e-9, atol=1e-12)


def test_plugin_dense_and_second_order_clones_are_bitwise_the_builtins(cuda):
    import torch

    from probdiffeq_b200 import ivpsolve, plugins, probdiffeq

    # dense factorisation (CTA per instance), Lotka-Volterra, nu = 3
    params, u0 = H.lv_ensemble(5, seed=74)
    sols = []
    for vf in (plugins.ode_from_cuda("lotka_volterra_dense_user", params=params, **plugins.LOTKA_VOLTERRA_DENSE_CLONE),
               probdiffeq.ode("lotka_volterra", params=params)):  # fmt: skip
        ssm = probdiffeq.state_space_model_dense()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0,), t=0.0)
        ts1 = ssm.constraint_ode_ts1(vf)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
        error = probdiffeq.error_residual_std(constraint=ts1)
        solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error)
        sols.append(solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=2.0, atol=1e-6, rtol=1e-4))
    torch.cuda.synchronize()
    assert int(sols[0].status.abs().max()) == 0 and torch.equal(sols[0].num_steps, sols[1].num_steps)
    assert torch.equal(sols[0].u.mean_flat, sols[1].u.mean_flat)
    assert torch.equal(sols[0].u.cholesky_flat, sols[1].u.cholesky_flat)
    # second-order right-hand side (Van der Pol), thread per instance
    rng = np.random.Generator(np.random.PCG64(75))
    params = np.full((7, 1), 5.0)
    u0, du0 = 2.0 * rng.uniform(0.9, 1.1, size=(7, 1)), np.zeros((7, 1))
    sols = []
    for vf in (plugins.ode_from_cuda("vanderpol_user", params=params, **plugins.VANDERPOL_CLONE),
               probdiffeq.ode("vanderpol", params=params)):  # fmt: skip
        ssm = probdiffeq.state_space_model_isotropic()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0, du0), t=0.0)
        ts1 = ssm.constraint_ode_ts1(vf)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
        error = probdiffeq.error_state_std(constraint=ts1)
        solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error)
        sols.append(solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=3.0, atol=1e-7, rtol=1e-5))
    torch.cuda.synchronize()
    assert int(sols[0].status.abs().max()) == 0 and torch.equal(sols[0].num_steps, sols[1].num_steps)
    assert torch.equal(sols[0].u.mean_flat, sols[1].u.mean_flat)


def test_plugin_jacobian_by_forward_mode_differentiation(cuda):
    """Without a `jacobian` body the ts1 derivatives come from `component` evaluated on dual numbers (the reference's
    jacfwd, jacobians.py:93-98): same solution as with the hand-written Jacobian, and as the oracle's."""
    import torch

    from probdiffeq_b200 import plugins

    B = 5
    rng = np.random.Generator(np.random.PCG64(76))
    params = np.stack([rng.uniform(0.5, 3.0, size=B), rng.uniform(2.0, 5.0, size=B)], axis=1)
    u0 = rng.uniform(0.1, 1.0, size=(B, 1))
    vf_auto = plugins.ode_from_cuda("logistic_autodiff", params=params, **plugins.LOGISTIC_AUTODIFF)
    vf_hand = plugins.ode_from_cuda("logistic", params=params, **plugins.LOGISTIC)
    _, _, sol_auto = _solve(vf_auto, u0, constraint="ts1", t1=3.0)
    _, _, sol_hand = _solve(vf_hand, u0, constraint="ts1", t1=3.0)
    torch.cuda.synchronize()
    assert int(sol_auto.status.abs().max()) == 0
    assert torch.equal(sol_auto.num_steps, sol_hand.num_steps)
    a, h = sol_auto.u.mean_flat.cpu().numpy(), sol_hand.u.mean_flat.cpu().numpy()
    assert np.allclose(a[:, :, 0], h[:, :, 0], rtol=1e-11, atol=1e-13)
    assert np.allclose(a, h, rtol=1e-6, atol=1e-9)
    assert np.allclose(sol_auto.u.std_flat.cpu().numpy(), sol_hand.u.std_flat.cpu().numpy(), rtol=1e-6, atol=1e-12)


def test_reregistering_a_name_with_new_code_replaces_every_kernel(cuda):
    """The same name registered twice with different bodies (same signature): the second plug-in's kernels must serve
    the step loop AND the Taylor initialisation / dt0 -- register_loop replaces an entry with an equal key, as
    register_aux does (the loop used to keep the first functor while the auxiliaries took the second)."""
    import torch

    from probdiffeq_b200 import plugins

    rng = np.random.Generator(np.random.PCG64(75))
    B = 4
    params = np.stack([rng.uniform(0.5, 1.5, size=B), rng.uniform(2.0, 5.0, size=B)], axis=1)
    u0 = rng.uniform(0.1, 1.0, size=(B, 1))
    save_at = np.linspace(0.0, 2.0, 9)

    def exact(rate_factor):
        k, r = params[:, 1:2], rate_factor * params[:, 0:1]
        return k / (1.0 + (k / u0 - 1.0) * np.exp(-r * save_at[None, :]))

    vf1 = plugins.ode_from_cuda("logistic_swap", params=params, **plugins.LOGISTIC)
    _, _, sol1 = _solve(vf1, u0)
    vf2 = plugins.ode_from_cuda("logistic_swap", params=params, **plugins.LOGISTIC_DOUBLED)
    assert vf2.vf_id == vf1.vf_id
    tc2, _, sol2 = _solve(vf2, u0)
    torch.cuda.synchronize()
    got1, got2 = sol1.u.mean_flat[:, :, 0, 0].cpu().numpy(), sol2.u.mean_flat[:, :, 0, 0].cpu().numpy()
    assert np.allclose(got1, exact(1.0), rtol=1e-3)
    assert np.allclose(got2, exact(2.0), rtol=1e-3) and not np.allclose(got2, exact(1.0), rtol=1e-2)
    f0 = 2.0 * params[:, 0] * u0[:, 0] * (1.0 - u0[:, 0] / params[:, 1])  # first Taylor coefficient of the new body
    assert np.allclose(tc2[:, 1, 0].cpu().numpy(), f0, rtol=1e-13)
    # and back again: the first plug-in is already loaded, its registrars must run again
    vf3 = plugins.ode_from_cuda("logistic_swap", params=params, **plugins.LOGISTIC)
    _, _, sol3 = _solve(vf3, u0)
    assert torch.equal(sol3.u.mean_flat, sol1.u.mean_flat)
