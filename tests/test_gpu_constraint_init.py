"""`solver(..., constraint_init=constraint)`: the Bayes update at t0 (reference: _probdiffeq/solvers.py:361-372,
526-537, 670-680; the diffuse-derivative start of ssm_impl_isotropic.py:556-567) on the device, against the oracle
(whose update is pinned against plain Gaussian conditioning in tests/test_oracle_kats.py)."""

import itertools

import numpy as np
import pytest

import pdeq_test_helpers as H

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _diffuse_start(B, n, d, fact, seed):
    params, u0 = H.lv_ensemble(B, seed=seed)
    tcoeffs = np.zeros((B, n, d))
    tcoeffs[:, 0] = u0
    std = np.ones((n,) if fact == "isotropic" else (n, d))
    std[0] = 0.0  # exact initial value, diffuse derivatives
    return params, tcoeffs, std


@pytest.mark.parametrize("fact,constraint,solver", list(itertools.product(
    ["isotropic", "blockdiag", "dense"], ["ts0", "ts1"], ["solver", "solver_mle", "solver_dynamic"])))  # fmt: skip
def test_fixed_grid_with_constraint_init_matches_oracle(cuda, fact, constraint, solver):
    import torch

    s = H.spec(fact=fact, constraint=constraint, solver=solver, constraint_init=True)
    B, n, d = 4, 4, 2
    params, tcoeffs, std = _diffuse_start(B, n, d, fact, seed=5)
    p_pdq, p_ivp, vf, ssm, slv, _err, _ctrl = H.product_build(s, params)
    prior = ssm.prior_wiener_integrated_diffuse(torch.as_tensor(tcoeffs, device="cuda"), torch.as_tensor(std, device="cuda"))
    grid = np.linspace(0.0, 0.5, 21)
    sol = p_ivp.solve_fixed_grid(solver=slv)(prior, grid=grid)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    for b in range(B):
        osol = H.oracle_solve_fixed(s, tcoeffs[b], params[b], grid, init_std=std)
        got_m = sol.u.mean_flat[b].cpu().numpy().reshape(np.asarray(osol.u_mean).shape)
        # the state at t0 is the updated one: coefficient 1 equals f(u0)
        p = params[b]
        u = tcoeffs[b, 0]
        f0 = np.asarray([p[0] * u[0] - p[1] * u[0] * u[1], -p[2] * u[1] + p[3] * u[0] * u[1]])
        assert _rel(got_m[0, 1], f0) < 1e-12
        tol = 1e-10
        L = sol.u.cholesky_flat[b].cpu().numpy()
        Lo = np.asarray(osol.u_chol).reshape(L.shape)
        if solver == "solver_dynamic":
            # the calibrated scale is a whitened residual of nearly cancelling terms: hold the comparison to the
            # oracle's own response to a 1-ulp change of its input (as tests/test_gpu_lv_parity.py does)
            pert = H.oracle_solve_fixed(s, tcoeffs[b] * (1.0 + 2.3e-16), params[b], grid, init_std=std)
            Lp = np.asarray(pert.u_chol).reshape(L.shape)
            sens = max(_rel(pert.u_mean, osol.u_mean),
                       max(_rel(H.cov_from_chol(Lp[k]), H.cov_from_chol(Lo[k])) for k in range(len(grid))))
            tol = max(tol, 100 * sens)
            assert tol < 1e-3
        assert _rel(got_m, osol.u_mean) < tol
        # the diffuse coefficients make the covariance entries span many orders of magnitude: compare per grid point
        for k in range(len(grid)):
            assert _rel(H.cov_from_chol(L[k]), H.cov_from_chol(Lo[k])) < max(tol, 1e-9), k


@pytest.mark.parametrize("fact,constraint", [("isotropic", "ts0"), ("blockdiag", "ts1"), ("dense", "ts1")])
def test_adaptive_with_constraint_init_matches_oracle(cuda, fact, constraint):
    import torch

    s = H.spec(fact=fact, constraint=constraint, solver="solver_dynamic", error="residual_std", control="i",
               constraint_init=True)  # fmt: skip
    B, n, d = 6, 4, 2
    params, tcoeffs, std = _diffuse_start(B, n, d, fact, seed=6)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    prior = ssm.prior_wiener_integrated_diffuse(torch.as_tensor(tcoeffs, device="cuda"), torch.as_tensor(std, device="cuda"))
    solve = p_ivp.solve_adaptive_terminal_values(solver=slv, error=err, control=ctrl)
    sol = solve(prior, t0=0.0, t1=2.0, atol=1e-8, rtol=1e-6, dt0=1e-3)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    same = 0
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tcoeffs[b], params[b], np.asarray([0.0, 2.0]), 1e-8, 1e-6, dt0=1e-3, init_std=std)
        if int(sol.num_steps[b]) == int(np.asarray(osol.num_steps)[-1]):
            same += 1
            assert _rel(sol.u.mean[0][b].cpu().numpy(), np.asarray(osol.u_mean)[-1, 0]) < 1e-7
        else:  # a tie-flip in the chaotic step-size feedback: still the same IVP solution
            assert _rel(sol.u.mean[0][b].cpu().numpy(), np.asarray(osol.u_mean)[-1, 0]) < 1e-4
    assert same >= B - 1


def test_constraint_init_must_be_the_solvers_constraint(cuda):
    from probdiffeq_b200 import probdiffeq as p_pdq

    vf = p_pdq.ode("lotka_volterra", params=np.asarray([[0.5, 0.05, 0.5, 0.05]]))
    ssm = p_pdq.state_space_model_isotropic()
    ts0, ts1 = ssm.constraint_ode_ts0(vf), ssm.constraint_ode_ts1(vf)
    p_pdq.solver(strategy=p_pdq.strategy_filter(), constraint=ts0, constraint_init=ssm.constraint_ode_ts0(vf))
    with pytest.raises(NotImplementedError):
        p_pdq.solver(strategy=p_pdq.strategy_filter(), constraint=ts0, constraint_init=ts1)
