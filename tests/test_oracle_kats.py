"""Pin the oracle against the reference's own known-answer tests and cross-implementation identities.

Each test names the reference test it restates (paths relative to /root/reference/tests). The reference
cannot be imported here (no JAX), so these identities -- which do not depend on JAX-generated data -- are what
pins the restatement.
"""

import numpy as np
import pytest
import scipy.integrate
import scipy.stats

from oracle import ivpsolve, linalg, ssm
from oracle import probdiffeq as pdq

FACTORIES = [pdq.state_space_model_isotropic, pdq.state_space_model_dense, pdq.state_space_model_blockdiag]


@pytest.mark.parametrize("factory", FACTORIES)
@pytest.mark.parametrize("dt", [1.234, -1.234])
def test_iwp_transitions_are_correct_in_1d(factory, dt):
    """test_probdiffeq/test_priors/test_wiener_integrated.py:12-43,49-81."""
    model = factory()
    iwp = model.prior_wiener_integrated(np.asarray([2.0, 3.0, 4.0, 5.0]))
    scale = np.ones(1) if model.kind == "blockdiag" else 1.0
    cond = iwp.transition(dt=dt, output_scale=scale)
    cond = cond.alg.preconditioner_apply(cond)
    A = np.asarray([[1.0, dt, dt**2 / 2, dt**3 / 6], [0, 1.0, dt, dt**2 / 2], [0, 0, 1.0, dt], [0, 0, 0, 1.0]])
    h = abs(dt)
    Q = np.asarray(
        [
            [h**7 / 252, h**6 / 72, h**5 / 30, h**4 / 24],
            [h**6 / 72, h**5 / 20, h**4 / 8, h**3 / 6],
            [h**5 / 30, h**4 / 8, h**3 / 3, h**2 / 2],
            [h**4 / 24, h**3 / 6, h**2 / 2, h],
        ]
    )
    assert np.allclose(np.squeeze(cond.A), A, rtol=1e-14, atol=1e-14)
    mean, cov = cond.noise.cov_dense()
    assert np.allclose(mean, 0.0)
    assert np.allclose(cov, Q, rtol=1e-13, atol=0)


@pytest.mark.parametrize("shapes", [[(4, 3), (3, 3), (4, 4)], [(2, 3), (3, 3), (2, 2)]])
def test_revert_conditional(shapes):
    """test_util/test_cholesky_util.py:14-33."""
    rng = np.random.default_rng(1)
    HC = rng.normal(size=shapes[0]) + 1.0
    C = rng.normal(size=shapes[1]) + 2.0
    X = rng.normal(size=shapes[2]) + 3.0 + np.eye(shapes[2][0])
    S = HC @ HC.T + X @ X.T
    K = C @ HC.T @ np.linalg.inv(S)
    C1 = C @ C.T - K @ S @ K.T
    r_obs, (r_cor, gain) = linalg.revert_conditional(R_X_F=HC.T, R_X=C.T, R_YX=X.T)
    assert np.allclose(r_obs.T @ r_obs, S)
    assert np.allclose(gain, K)
    assert np.allclose(r_cor.T @ r_cor, C1)


def test_revert_conditional_zero_covariance_is_finite():
    """test_util/test_cholesky_util.py:41-55 (values, not gradients): zero prior covariance is admissible."""
    HC, C = np.zeros((2, 3)), np.zeros((3, 3))
    X = np.asarray([[4.0, 0.3], [0.1, 3.5]])
    r_obs, (r_cor, gain) = linalg.revert_conditional(R_X_F=HC.T, R_X=C.T, R_YX=X.T)
    assert np.all(np.isfinite(r_obs)) and np.all(np.isfinite(gain)) and np.allclose(r_cor, 0.0)


@pytest.mark.parametrize("factory", FACTORIES)
def test_logpdf_matches_dense_mvn(factory):
    """test_probdiffeq/test_logpdf.py:18-29."""
    rng = np.random.default_rng(3)
    n, d = 3, 2
    model = factory()
    prior = model.prior_wiener_integrated(rng.normal(size=(n, d)))
    tr = prior.transition(dt=0.3, output_scale=np.ones(d) if model.kind == "blockdiag" else 1.0)
    rv = tr.marginalise(prior.init)
    u_nd = rng.normal(size=(n, d))
    mean, cov = rv.cov_dense()
    expected = scipy.stats.multivariate_normal(mean, cov).logpdf(u_nd.reshape(-1))
    assert np.isclose(rv.alg.logpdf(rv, rv.alg.from_nd(u_nd)), expected, rtol=1e-10)


def test_controllers_pi_equals_i_for_trivial_exponents():
    """test_ivpsolve/test_controllers.py:10-26."""
    pi = ivpsolve.control_proportional_integral(exponent_integral=1.0, exponent_proportional=0.0)
    i = ivpsolve.control_integral()
    dt_pi, x_pi, dt_i, x_i = 0.1428, pi.init(0.1428), 0.1428, i.init(0.1428)
    for _ in range(4):
        dt_pi, x_pi = pi.apply(dt_pi, x_pi, error_power=3.142)
        dt_i, x_i = i.apply(dt_i, x_i, error_power=3.142)
    assert np.isclose(dt_pi, dt_i, rtol=1e-15)


def test_cholesky_hilbert_is_a_cholesky_factor():
    """util/cholesky_util.py:106-176 against the definition."""
    for n in (2, 5, 8):
        L = linalg.cholesky_hilbert(n)
        H = 1.0 / (np.arange(1, n + 1)[:, None] + np.arange(1, n + 1)[None, :] - 1.0)
        assert np.allclose(L @ L.T, H, rtol=1e-9)


# ---------------------------------------------------------------------------------------------------
# end-to-end: accuracy against an independent integrator + cross-implementation identities
# ---------------------------------------------------------------------------------------------------

LV = pdq.ode("lotka_volterra")
U0 = np.asarray([20.0, 20.0])


def _setup(kind, constraint, solver_name, strategy, error_name, num=3):
    model = getattr(pdq, "state_space_model_" + kind)()
    tcoeffs, _ = pdq.jetexpand_ode_padded_scan(num=num)(LV, (U0,), t=0.0)
    prior = model.prior_wiener_integrated(tcoeffs)
    cons = getattr(model, "constraint_ode_" + constraint)(LV)
    slv = getattr(pdq, solver_name)(strategy=strategy(), constraint=cons)
    err = getattr(pdq, error_name)(constraint=cons)
    return prior, slv, err


@pytest.fixture(scope="module")
def truth():
    sol = scipy.integrate.solve_ivp(
        lambda t, y: LV.vector_field((y,), t), (0.0, 2.0), U0, rtol=1e-12, atol=1e-12, method="DOP853",
        t_eval=np.linspace(0.0, 2.0, 5),
    )  # fmt: skip
    return sol.y.T


@pytest.mark.parametrize("kind", ["isotropic", "blockdiag", "dense"])
@pytest.mark.parametrize("constraint", ["ts0", "ts1"])
@pytest.mark.parametrize("solver_name", ["solver", "solver_mle", "solver_dynamic"])
@pytest.mark.parametrize("strategy", [pdq.strategy_filter, pdq.strategy_smoother_fixedpoint])
def test_save_at_solution_is_accurate(kind, constraint, solver_name, strategy, truth):
    """test_ivpsolve/test_solve_adaptive_save_at.py:211-247 with scipy instead of jax's odeint."""
    prior, slv, err = _setup(kind, constraint, solver_name, strategy, "error_residual_std")
    solve = ivpsolve.solve_adaptive_save_at(solver=slv, error=err)
    sol = solve(prior, save_at=np.linspace(0.0, 2.0, 5), atol=1e-6, rtol=1e-4)
    assert np.allclose(sol.u_mean[:, 0], truth, rtol=2e-3)
    assert np.allclose(sol.t, np.linspace(0.0, 2.0, 5))


@pytest.mark.parametrize("error_name", ["error_residual_std", "error_state_std"])
def test_terminal_values_equal_last_save_at_entry(error_name):
    """solve_adaptive_terminal_values is save_at=[t0, t1] (solvers_via_adaptive_steps.py:16-43)."""
    prior, slv, err = _setup("isotropic", "ts0", "solver", pdq.strategy_filter, error_name)
    a = ivpsolve.solve_adaptive_terminal_values(solver=slv, error=err)(prior, t0=0.0, t1=2.0, atol=1e-6, rtol=1e-4)
    b = ivpsolve.solve_adaptive_save_at(solver=slv, error=err, clip_dt=True)(
        prior, save_at=np.asarray([0.0, 2.0]), atol=1e-6, rtol=1e-4
    )
    assert np.array_equal(a.u.tcoeffs, b.u_mean[-1])


@pytest.mark.parametrize("solver_name", ["solver", "solver_dynamic"])
def test_dense_equals_isotropic_for_ts0(solver_name):
    """test_probdiffeq/test_calibration/test_dynamic_across_factorisations.py:12: for ts0 the dense and
    isotropic models describe the same process when the dynamics are calibrated with a shared scale."""
    grid = np.linspace(0.0, 1.0, 11)
    out = {}
    for kind in ("isotropic", "dense"):
        prior, slv, _ = _setup(kind, "ts0", solver_name, pdq.strategy_filter, "error_residual_std")
        out[kind] = ivpsolve.solve_fixed_grid(solver=slv)(prior, grid=grid)
    assert np.allclose(out["isotropic"].u_mean, out["dense"].u_mean, rtol=1e-9, atol=1e-12)
    iso_cov = np.stack([r.cov_dense()[1] for r in out["isotropic"].u])
    dense_cov = np.stack([r.cov_dense()[1] for r in out["dense"].u])
    # entries that vanish exactly in the isotropic model are rounding noise in the dense one
    assert np.allclose(iso_cov, dense_cov, rtol=1e-7, atol=1e-13 * np.abs(dense_cov).max())


def test_blockdiag_equals_independent_scalar_dense_models():
    """test_dynamic_across_factorisations.py:66 (vmap(dense) == blockdiag) on a decoupled problem."""
    lin = pdq.ode("linear", [1.5])
    u0 = np.asarray([1.0, -2.0, 0.5])
    grid = np.linspace(0.0, 1.0, 9)
    tc, _ = pdq.jetexpand_ode_padded_scan(num=3)(lin, (u0,), t=0.0)
    bd = pdq.state_space_model_blockdiag()
    slv = pdq.solver_dynamic(strategy=pdq.strategy_filter(), constraint=bd.constraint_ode_ts0(lin))
    sol_bd = ivpsolve.solve_fixed_grid(solver=slv)(bd.prior_wiener_integrated(tc), grid=grid)
    for i in range(3):
        dn = pdq.state_space_model_dense()
        slv_i = pdq.solver_dynamic(strategy=pdq.strategy_filter(), constraint=dn.constraint_ode_ts0(lin))
        sol_i = ivpsolve.solve_fixed_grid(solver=slv_i)(dn.prior_wiener_integrated(tc[:, i : i + 1]), grid=grid)
        assert np.allclose(sol_bd.u_mean[:, :, i], sol_i.u_mean[:, :, 0], rtol=1e-10, atol=1e-14)
        assert np.allclose(sol_bd.u_std[:, :, i], sol_i.u_std[:, :, 0], rtol=1e-8, atol=1e-16)


def test_fixed_grid_on_the_adaptive_grid_reproduces_the_adaptive_solution():
    """test_ivpsolve/test_solve_fixed_grid.py:17-54."""
    prior, slv, err = _setup("isotropic", "ts0", "solver", pdq.strategy_filter, "error_residual_std")
    adaptive = ivpsolve.solve_adaptive_save_every_step(solver=slv, error=err, clip_dt=True)(
        prior, t0=0.0, t1=1.0, atol=1e-5, rtol=1e-3
    )
    fixed = ivpsolve.solve_fixed_grid(solver=slv)(prior, grid=adaptive.t)
    assert np.allclose(adaptive.u_mean, fixed.u_mean, rtol=1e-10)
    assert np.allclose(adaptive.u_std, fixed.u_std, rtol=1e-8, atol=1e-16)


def test_fixedpoint_smoother_equals_fixedinterval_smoother_on_the_same_grid():
    """test_probdiffeq/test_strategies/test_smoother_fixedinterval_vs_fixedpoint.py:50-86."""
    for kind in ("isotropic", "blockdiag", "dense"):
        prior, slv_fi, err = _setup(kind, "ts0", "solver", pdq.strategy_smoother_fixedinterval, "error_residual_std", num=2)
        every = ivpsolve.solve_adaptive_save_every_step(solver=slv_fi, error=err)(
            prior, t0=0.0, t1=2.0, atol=1e-3, rtol=1e-3
        )
        _, slv_fp, err_fp = _setup(kind, "ts0", "solver", pdq.strategy_smoother_fixedpoint, "error_residual_std", num=2)
        fp = ivpsolve.solve_adaptive_save_at(solver=slv_fp, error=err_fp)(prior, save_at=every.t, atol=1e-3, rtol=1e-3)
        assert np.allclose(fp.t, every.t)
        assert np.array_equal(fp.num_steps, every.num_steps)
        assert np.allclose(fp.u_mean, every.u_mean, rtol=1e-9, atol=1e-12)
        assert np.allclose(fp.u_std, every.u_std, rtol=1e-7, atol=1e-14)


def test_fixedinterval_smoother_on_a_fixed_grid_reference_vs_aligned():
    """solve_fixed_grid hands Smoother.finalize the last grid state as `solution1` (solvers_via_fixed_steps.py:30-32),
    and finalize marginalises it through its own conditional (estimators_and_losses.py:453-454). Literally restated,
    the last entry is the smoothed marginal one grid point earlier; with an identity conditional ("aligned") the
    pass ends in the filtering marginal and equals a hand-rolled Rauch-Tung-Striebel recursion."""
    for kind in ("isotropic", "blockdiag", "dense"):
        prior, slv_fi, _ = _setup(kind, "ts0", "solver", pdq.strategy_smoother_fixedinterval, "error_residual_std", num=2)
        _, slv_f, _ = _setup(kind, "ts0", "solver", pdq.strategy_filter, "error_residual_std", num=2)
        grid = np.linspace(0.0, 1.0, 11)
        lit = ivpsolve.solve_fixed_grid(solver=slv_fi)(prior, grid=grid)
        ali = ivpsolve.solve_fixed_grid(solver=slv_fi, terminal="aligned")(prior, grid=grid)
        filt = ivpsolve.solve_fixed_grid(solver=slv_f)(prior, grid=grid)
        assert np.allclose(ali.u_mean[-1], filt.u_mean[-1], rtol=1e-9, atol=1e-12)
        assert np.allclose(ali.u_std[-1], filt.u_std[-1], rtol=1e-7, atol=1e-14)
        # the literal terminal entry is the aligned entry one grid point earlier
        assert np.allclose(lit.u_mean[-1], ali.u_mean[-2], rtol=1e-9, atol=1e-12)
        assert not np.allclose(lit.u_mean[-1], ali.u_mean[-1], rtol=1e-3)
        # hand-rolled backward recursion over the stored conditionals
        conds = lit.solution_full.posterior.conditional  # one per step, t_k -> t_{k-1}
        rv = filt.u[-1] if isinstance(filt.u, list) else None
        if rv is not None:
            for k in range(len(conds) - 1, -1, -1):
                rv = conds[k].marginalise(rv)
                assert np.allclose(rv.tcoeffs, ali.u_mean[k], rtol=1e-8, atol=1e-10), (kind, k)
        # smoothing never widens the filter's uncertainty
        assert np.all(ali.u_std <= filt.u_std * (1 + 1e-9) + 1e-14)


@pytest.mark.parametrize("kind", ["isotropic", "blockdiag", "dense"])
@pytest.mark.parametrize("smoother", [False, True])
def test_offgrid_marginals_match_the_save_at_solution(kind, smoother):
    """test_probdiffeq/test_dense_output/test_offgrid_marginals_vs_solve_and_save_at.py:51-85 (ts1, solver_dynamic)."""
    every_strategy = pdq.strategy_smoother_fixedinterval if smoother else pdq.strategy_filter
    at_strategy = pdq.strategy_smoother_fixedpoint if smoother else pdq.strategy_filter
    prior, slv_every, err = _setup(kind, "ts1", "solver_dynamic", every_strategy, "error_residual_std", num=2)
    _, slv_at, err_at = _setup(kind, "ts1", "solver_dynamic", at_strategy, "error_residual_std", num=2)
    ts = np.linspace(0.0, 2.0, 5)
    every = ivpsolve.solve_adaptive_save_every_step(solver=slv_every, error=err)(
        prior, t0=0.0, t1=2.0, atol=1e-2, rtol=1e-2
    )
    at = ivpsolve.solve_adaptive_save_at(solver=slv_at, error=err_at)(prior, save_at=ts, atol=1e-2, rtol=1e-2)
    for k, t in enumerate(ts[1:-1], start=1):
        rv = slv_every.offgrid_marginals(t, solution=every)
        m1, C1 = rv.cov_dense()
        m2, C2 = at.u[k].cov_dense()
        assert np.allclose(m1, m2, rtol=1e-8, atol=1e-10), (kind, smoother, k)
        assert np.allclose(C1, C2, rtol=1e-6, atol=1e-12), (kind, smoother, k)


def test_save_at_is_invariant_to_cutting_the_grid():
    """test_probdiffeq/test_dense_output/test_behaviour_close_to_t1.py:59-94: solving to [t0, t1] equals solving
    on a grid that contains t1 and reading off t1 (filter, no clipping)."""
    prior, slv, err = _setup("isotropic", "ts0", "solver", pdq.strategy_filter, "error_residual_std")
    solve = ivpsolve.solve_adaptive_save_at(solver=slv, error=err)
    a = solve(prior, save_at=np.asarray([0.0, 0.7, 1.4]), atol=1e-6, rtol=1e-4)
    b = solve(prior, save_at=np.asarray([0.0, 1.4]), atol=1e-6, rtol=1e-4)
    assert np.allclose(a.u_mean[-1], b.u_mean[-1], rtol=1e-9)


def test_taylor_coefficients_match_hand_derivatives():
    """jet_expansion_algorithms.py:49-177 against derivatives of u' = f(u) written out by hand."""
    a, b, c, d = LV.params
    u = U0
    f = np.asarray([a * u[0] - b * u[0] * u[1], -c * u[1] + d * u[0] * u[1]])
    J = np.asarray([[a - b * u[1], -b * u[0]], [d * u[1], -c + d * u[0]]])
    d2 = J @ f
    tc, _ = pdq.jetexpand_ode_padded_scan(num=2)(LV, (U0,), t=0.0)
    assert np.allclose(tc[1], f) and np.allclose(tc[2], d2)
    vdp = pdq.ode("vanderpol", [1e3])
    tc2, _ = pdq.jetexpand_ode_padded_scan(num=2)(vdp, (np.asarray([2.0]), np.asarray([0.0])), t=0.0)
    # u'' = s((1-u^2)u' - u) = -2s ; u''' = s(-2uu'u' + (1-u^2)u'' - u') = s(-3)(-2s) = 6 s^2
    assert np.allclose(tc2[:, 0], [2.0, 0.0, -2e3, 6e6])


def test_jacobians_are_exact():
    """jacobians.py:93-98: complex-step Jacobians against closed forms."""
    a, b, c, d = LV.params
    (J,) = LV.jacobians((U0,), 0.0)
    assert np.allclose(J, [[a - b * U0[1], -b * U0[0]], [d * U0[1], -c + d * U0[0]]], rtol=1e-15)
    hires = pdq.ode("hires")
    from oracle import problems

    u = problems.hires_u0() + 0.1
    (Jh,) = hires.jacobians((u,), 0.0)
    assert np.isclose(Jh[5, 5], -280.0 * u[7] - 0.43) and np.isclose(Jh[6, 7], 280.0 * u[5])


def test_lml_terminal_values_matches_dense_gaussian():
    """estimators_and_losses.py:20-50 on all three factorisations."""
    rng = np.random.default_rng(0)
    for kind in ("isotropic", "blockdiag", "dense"):
        prior, slv, err = _setup(kind, "ts0", "solver", pdq.strategy_filter, "error_residual_std")
        sol = ivpsolve.solve_adaptive_terminal_values(solver=slv, error=err)(prior, t0=0.0, t1=1.0, atol=1e-5, rtol=1e-3)
        data = sol.u.tcoeffs[0] + 1e-2 * rng.normal(size=2)
        std = 1e-2 if kind == "isotropic" else 1e-2 * np.ones(2)
        lml = pdq.loss_lml_terminal_values()(data, marginals=sol.u, std=std)
        mean, cov = sol.u.cov_dense()
        if kind == "dense":
            m0, c0 = mean[:2], cov[:2, :2]
        else:
            n = sol.u.tcoeffs.shape[0]
            m0 = mean.reshape(n, 2)[0]
            c0 = cov.reshape(n, 2, n, 2)[0, :, 0, :]
        expected = scipy.stats.multivariate_normal(m0, c0 + 1e-4 * np.eye(2)).logpdf(data)
        assert np.isclose(lml, expected, rtol=1e-9), kind


def test_lml_timeseries_matches_joint_gaussian():
    """estimators_and_losses.py:53-105, 180-218 (setup of test_losses/test_lml_timeseries.py:19-46). The backward
    scan must equal the log-density of all observations under the joint Gaussian that the Markov sequence
    defines; isotropic == dense (ts0) as in test_dynamic_across_factorisations.py."""
    rng = np.random.default_rng(1)
    save_at = np.linspace(0.0, 2.0, 7)
    values = {}
    for kind in ("isotropic", "blockdiag", "dense"):
        prior, slv, err = _setup(kind, "ts0", "solver_mle", pdq.strategy_smoother_fixedpoint, "error_residual_std", num=2)
        sol = ivpsolve.solve_adaptive_save_at(solver=slv, error=err)(prior, save_at=save_at, atol=1e-3, rtol=1e-3)
        post = sol.solution_full.posterior
        data = np.asarray(sol.u_mean)[:, 0] + 0.05 * rng.normal(size=(7, 2))
        sd = 0.05 + 0.01 * np.arange(7)
        std = sd if kind == "isotropic" else np.stack([sd, 2 * sd], axis=1)
        total = pdq.loss_lml_timeseries(average_pdfs=False)(data, posterior=post, std=std)
        mean_of = pdq.loss_lml_timeseries(average_pdfs=True)(data, posterior=post, std=std)
        assert np.isclose(mean_of, total / 7, rtol=1e-12)
        values[kind] = (total, data, std)
        with pytest.raises(ValueError, match="container differs"):
            pdq.loss_lml_timeseries()(data, posterior=post, std=std[:-1])
        with pytest.raises(TypeError, match="datatype"):
            pdq.loss_lml_timeseries()(data, posterior=sol.u, std=std)
        if kind != "dense":
            continue
        # joint Gaussian over the stacked states, built from the terminal marginal and the natural-form conditionals
        post = post.remove_filtering_distributions()
        m, C = post.marginal.cov_dense()
        N, T = m.size, len(save_at)
        means, covs = {T - 1: m}, {(T - 1, T - 1): C}
        for k in range(T - 1, 0, -1):
            c = post.conditional[k - 1].alg.preconditioner_apply(post.conditional[k - 1])
            G, xi, Xi = c.A, c.noise.mean.reshape(-1), c.noise.chol
            means[k - 1] = G @ means[k] + xi
            for l in range(k, T):
                covs[(k - 1, l)] = G @ covs[(k, l)]
                covs[(l, k - 1)] = covs[(k - 1, l)].T
            covs[(k - 1, k - 1)] = G @ covs[(k, k)] @ G.T + Xi @ Xi.T
        H = np.zeros((2, N))
        H[0, 0] = H[1, 1] = 1.0  # coefficient-major: the state is the first d entries
        mu = np.concatenate([H @ means[k] for k in range(T)])
        S = np.block([[H @ covs[(k, l)] @ H.T for l in range(T)] for k in range(T)])
        S = S + np.diag(np.concatenate([std[k] ** 2 for k in range(T)]))
        expected = scipy.stats.multivariate_normal(mu, S, allow_singular=True).logpdf(data.reshape(-1))
        assert np.isclose(total, expected, rtol=1e-8), (total, expected)


def test_posterior_sample_with_zero_draws_is_the_smoothing_mean_and_has_the_right_spread():
    """estimators_and_losses.py:233-271 with explicit draws: zero draws walk the means backwards; unit draws, averaged,
    reproduce the smoothing covariance of the state at a grid point."""
    rng = np.random.default_rng(3)
    for kind in ("isotropic", "blockdiag", "dense"):
        prior, slv, err = _setup(kind, "ts0", "solver", pdq.strategy_smoother_fixedpoint, "error_residual_std", num=2)
        sol = ivpsolve.solve_adaptive_save_at(solver=slv, error=err)(prior, save_at=np.linspace(0, 2, 5), atol=1e-2, rtol=1e-2)
        post = sol.solution_full.posterior
        shape = {"isotropic": (5, 3), "blockdiag": (5, 2, 3), "dense": (5, 6)}[kind]
        zero = post.sample(np.zeros(shape))
        for k in range(5):
            assert np.allclose(zero[k], sol.u[k].mean, rtol=1e-9, atol=1e-12), (kind, k)
        if kind == "blockdiag":
            draws = np.stack([np.stack(post.sample(rng.normal(size=shape)))[2] for _ in range(4000)])  # (S, d, n)
            emp = np.einsum("sdi,sdj->dij", draws - draws.mean(0), draws - draws.mean(0)) / 4000
            cov = sol.u[2].chol @ np.swapaxes(sol.u[2].chol, -1, -2)
            assert np.allclose(emp, cov, rtol=0.2, atol=0.05 * np.abs(cov).max())


def test_unused_import_guard():
    assert ssm.Normal is not None


@pytest.mark.parametrize("nu", [1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("dt", [1.234, 0.037, -0.5])
def test_preconditioned_iwp_mean_transition_is_the_taylor_shift(nu, dt):
    """The identity the kernels' `predict_mean` rests on (csrc/pdeq_blockops.cuh): with A = flipped Pascal
    (utilities.py:59-61) and p_k = dt^(nu-k)/(nu-k)! (utilities.py:74-84),  p_i A_ik / p_k = dt^(k-i)/(k-i)!  for
    k >= i, so p * (A (p^-1 * m)) -- LatentCond.marginalise's mean, ssm_impl_isotropic.py:81-89 -- is the Taylor
    shift  m_i + sum_{k>i} p[nu-(k-i)] m_k.  Checked against the oracle's own system matrices and preconditioner."""
    from oracle import linalg as ol

    n = nu + 1
    A, _ = ol.system_matrices_1d_iwp(nu)
    p, pinv = ol.preconditioner_taylor(nu)(dt)
    rng = np.random.Generator(np.random.PCG64(nu))
    m = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-3, 4, size=(n, 1))
    reference = p[:, None] * (A @ (pinv[:, None] * m))
    shifted = m.copy()
    for i in range(n):
        for k in range(i + 1, n):
            shifted[i] += p[n - 1 - (k - i)] * m[k]
    scale = np.abs(p[:, None] * (np.abs(A) @ np.abs(pinv[:, None] * m)))  # size of the terms that are summed
    assert np.all(np.abs(shifted - reference) <= 1e-14 * scale)
    # the coefficient form of the same statement
    M = p[:, None] * A * pinv[None, :]
    for i in range(n):
        for k in range(n):
            expect = 0.0 if k < i else dt ** (k - i) / float(ol.factorial(k - i))
            assert abs(M[i, k] - expect) <= 1e-13 * max(abs(expect), 1e-300) or (k < i and M[i, k] == 0.0)


def _relinearize_case(fact, constraint, strategy, error_name, relin_solver, relin_error):
    draw = np.random.Generator(np.random.PCG64(3)).uniform(0.8, 1.2, size=(1, 6))
    params, u0 = np.asarray([0.5, 0.05, 0.5, 0.05])[None, :] * draw[:, :4], 20.0 * draw[:, 4:]
    vf = pdq.ode("lotka_volterra", params[0])
    tc = pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0[0],), t=0.0)[0]
    ssm = getattr(pdq, "state_space_model_" + fact)()
    cons = getattr(ssm, "constraint_ode_" + constraint)(vf)
    strat = {"filter": pdq.strategy_filter, "fixedpoint": pdq.strategy_smoother_fixedpoint}[strategy]()
    solver = pdq.solver_dynamic(strategy=strat, constraint=cons, re_linearize_after_calibration=relin_solver)
    err = getattr(pdq, "error_" + error_name)(constraint=cons, re_linearize_before_error=relin_error)
    prior = ssm.prior_wiener_integrated(np.asarray(tc))
    solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=err, control=ivpsolve.control_integral(), clip_dt=False, warn=False)
    return solve(prior, save_at=np.linspace(0.0, 3.0, 4), atol=1e-8, rtol=1e-6, dt0=0.1)


def _fields(sol):
    return [np.asarray(x) for x in (sol.t, sol.u_mean, sol.u_chol, sol.num_steps, sol.output_scale)]


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag", "dense"])
@pytest.mark.parametrize("constraint", ["ts0", "ts1"])
@pytest.mark.parametrize("strategy", ["filter", "fixedpoint"])
@pytest.mark.parametrize("error_name", ["state_std", "residual_std"])
def test_re_linearize_flags_change_nothing_for_prior_taylor_points(fact, constraint, strategy, error_name):
    """Why the product accepts `re_linearize_before_error` / `re_linearize_after_calibration` and runs the same
    kernels (reference: solvers.py:578-582, 946-952, 1061-1067): the second linearisation is taken at the mean of a
    state whose mean the first linearisation already used, so with the prior Taylor point (taylor_points.py:150-156)
    it reproduces the first. In the oracle this is bitwise for the error estimators in every factorisation and for
    solver_dynamic in the block-diagonal and dense models; the isotropic model extrapolates the mean through two code
    paths that differ by an ulp (ssm_impl_isotropic.py:75-79 vs :81-89), which the adaptive loop turns into ~1e-10 --
    the same accepted steps, values far inside the 1e-8 tolerance."""
    base = _fields(_relinearize_case(fact, constraint, strategy, error_name, False, False))
    with_error_flag = _fields(_relinearize_case(fact, constraint, strategy, error_name, False, True))
    for a, b in zip(base, with_error_flag):
        assert np.array_equal(a, b)
    with_solver_flag = _fields(_relinearize_case(fact, constraint, strategy, error_name, True, False))
    if fact == "isotropic":
        t0, m0, L0, n0, s0 = base
        t1, m1, L1, n1, s1 = with_solver_flag
        assert np.array_equal(n0, n1)  # accepted steps per checkpoint
        rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300)  # noqa: E731
        assert rel(t0, t1) <= 1e-12 and rel(m0, m1) <= 1e-8
        cov = lambda L: L @ np.swapaxes(L, -1, -2)  # noqa: E731  (factor signs are not unique)
        assert rel(cov(L0), cov(L1)) <= 1e-6 and rel(s0, s1) <= 1e-6
    else:
        for a, b in zip(base, with_solver_flag):
            assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------------
# constraint_init: the Bayes update at t0 (solvers.py:361-372, 526-537, 670-680)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["isotropic", "blockdiag", "dense"])
@pytest.mark.parametrize("solver_name", ["solver", "solver_mle", "solver_dynamic"])
def test_constraint_init_pins_the_first_derivative_of_a_diffuse_prior(kind, solver_name):
    """What tests/test_probdiffeq/test_solver_uses_constraint_init.py checks for jet-lifted residuals, for the ODE
    constraint itself: with an exact u0 and diffuse derivatives, conditioning on u' - f(u0) = 0 at t0 makes the mean
    of coefficient 1 equal f(u0) and removes its uncertainty; the other diffuse coefficients stay diffuse."""
    params = np.asarray([0.5, 0.05, 0.5, 0.05])
    u0 = np.asarray([20.0, 19.0])
    vf = pdq.ode("lotka_volterra", params)
    model = getattr(pdq, "state_space_model_" + kind)()
    ts0 = model.constraint_ode_ts0(vf)
    n, d, big = 4, 2, 1e3
    tcoeffs = np.zeros((n, d))
    tcoeffs[0] = u0
    std = np.full((n,) if kind == "isotropic" else (n, d), big)
    std[0] = 0.0
    prior = model.prior_wiener_integrated_diffuse(tcoeffs, std)
    slv = getattr(pdq, solver_name)(strategy=pdq.strategy_filter(), constraint=ts0, constraint_init=ts0)
    state = slv.init(0.0, prior, damp=0.0)
    f0 = np.asarray([0.5 * u0[0] - 0.05 * u0[0] * u0[1], -0.5 * u0[1] + 0.05 * u0[0] * u0[1]])
    mean = state.u.tcoeffs
    sd = np.asarray(state.u.std).reshape(n, -1)
    np.testing.assert_allclose(mean[0], u0, rtol=0, atol=0)
    np.testing.assert_allclose(mean[1], f0, rtol=1e-13)
    assert np.all(sd[0] == 0.0) and np.all(sd[1] <= 1e-9 * big)
    np.testing.assert_allclose(sd[2:], big, rtol=1e-12)
    # without constraint_init nothing happens at t0
    plain = getattr(pdq, solver_name)(strategy=pdq.strategy_filter(), constraint=ts0).init(0.0, prior, damp=0.0)
    np.testing.assert_array_equal(plain.u.tcoeffs, tcoeffs)


def test_constraint_init_is_gaussian_conditioning_dense_ts1():
    """Dense model, ts1, inexact u0: the update equals m - S H^T (H S H^T + damp^2 I)^-1 r in covariance form."""
    params = np.asarray([0.5, 0.05, 0.5, 0.05])
    vf = pdq.ode("lotka_volterra", params)
    model = pdq.state_space_model_dense()
    ts1 = model.constraint_ode_ts1(vf)
    n, d = 3, 2
    rng = np.random.default_rng(3)
    tcoeffs = rng.normal(size=(n, d)) + np.asarray([[20.0, 19.0], [0.0, 0.0], [0.0, 0.0]])
    std = rng.uniform(0.1, 2.0, size=(n, d))
    prior = model.prior_wiener_integrated_diffuse(tcoeffs, std)
    damp = 0.3
    state = pdq.solver(strategy=pdq.strategy_filter(), constraint=ts1, constraint_init=ts1).init(0.0, prior, damp=damp)
    fx, _ = ts1.linearize(prior.init, None, damp=damp, t=0.0)
    H, bias = fx.A, np.asarray(fx.noise.mean).reshape(-1)
    m0 = prior.init.mean.reshape(-1)
    S0 = prior.init.cov_dense()[1]
    r = H @ m0 + bias
    K = S0 @ H.T @ np.linalg.inv(H @ S0 @ H.T + damp**2 * np.eye(d))
    np.testing.assert_allclose(state.u.mean.reshape(-1), m0 - K @ r, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(state.u.cov_dense()[1], S0 - K @ (H @ S0 @ H.T + damp**2 * np.eye(d)) @ K.T, rtol=1e-9, atol=1e-11)


def test_constraint_init_solve_is_accurate(truth):
    """A diffuse-derivative start with constraint_init still solves the IVP (adaptive, save_at)."""
    model = pdq.state_space_model_isotropic()
    ts0 = model.constraint_ode_ts0(LV)
    n = 4
    tcoeffs = np.zeros((n, 2))
    tcoeffs[0] = U0
    std = np.asarray([0.0, 1.0, 1.0, 1.0])
    prior = model.prior_wiener_integrated_diffuse(tcoeffs, std)
    slv = pdq.solver_dynamic(strategy=pdq.strategy_filter(), constraint=ts0, constraint_init=ts0)
    err = pdq.error_residual_std(constraint=ts0)
    solve = ivpsolve.solve_adaptive_save_at(solver=slv, error=err, warn=False)
    sol = solve(prior, save_at=np.linspace(0.0, 2.0, 5), atol=1e-8, rtol=1e-6, dt0=1e-3)
    np.testing.assert_allclose(np.asarray(sol.u_mean)[:, 0], truth, rtol=2e-3)
