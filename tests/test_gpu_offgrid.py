"""GPU parity for dense output, SURVEY §8(f) rank 2: `solver.offgrid_marginals` (probdiffeq/_probdiffeq/solvers.py:149-203)
for filters and the fixed-interval smoother, against the oracle -- whose own offgrid marginals are pinned on CPU by the
reference's identity "offgrid marginals of the every-step solution == save_at solution"
(tests/test_probdiffeq/test_dense_output/test_offgrid_marginals_vs_solve_and_save_at.py:51-85)."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from oracle import ivpsolve as o_ivp
from oracle import probdiffeq as o_pdq

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _cov(L):
    return L @ np.swapaxes(L, -1, -2)


def _oracle_mean_chol(rv, fact):
    """Oracle Normal -> the product's (n, d) mean and ([d,] n, n) factor."""
    return (rv.mean, rv.chol) if fact == "isotropic" else (rv.mean.T, rv.chol)


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
@pytest.mark.parametrize("solver,constraint", [("solver_dynamic", "ts1"), ("solver_mle", "ts0"), ("solver", "ts0")])
def test_filter_offgrid_marginals_of_a_save_at_solution(cuda, fact, solver, constraint):
    import torch

    s = H.spec(fact=fact, solver=solver, constraint=constraint, error="residual_std", control="i", clip_dt=False)
    B = 4
    params, u0 = H.lv_ensemble(B, seed=51)
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, params)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    save_at = np.linspace(0.0, 3.0, 9)
    scale = 1.5 if fact == "isotropic" else np.asarray([1.5, 0.75])
    prior = ssm.prior_wiener_integrated(tcoeffs, output_scale=scale)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(prior, save_at=save_at, atol=1e-5, rtol=1e-4)
    ts = np.concatenate([0.5 * (save_at[1:] + save_at[:-1]), [0.01, 2.99, 1.2345]])
    rv = slv.offgrid_marginals(ts, solution=sol)
    one = slv.offgrid_marginals(float(ts[3]), solution=sol)
    torch.cuda.synchronize()
    assert rv.mean_flat.shape == (B, len(ts), 5, 2) and one.mean_flat.shape == (B, 5, 2)
    assert np.array_equal(one.mean_flat.cpu().numpy(), rv.mean_flat[:, 3].cpu().numpy())
    tc = tcoeffs.cpu().numpy()
    for b in range(B):
        osol, _ = H.oracle_solve_save_at(s, tc[b], params[b], save_at, 1e-5, 1e-4, output_scale=scale)
        _, oslv, _, _ = H._build(o_pdq, o_ivp, s, H.oracle_vf(s, params[b]))
        assert np.array_equal(sol.num_steps[b, 1:].cpu().numpy(), np.asarray(osol.num_steps))
        for k, t in enumerate(ts):
            m, L = _oracle_mean_chol(oslv.offgrid_marginals(t, solution=osol), fact)
            got_m = rv.mean_flat[b, k].cpu().numpy()
            got_L = rv.cholesky_flat[b, k].cpu().numpy()
            for i in range(5):  # per Taylor coefficient (the high ones carry the solve's own conditioning)
                assert _rel(got_m[i], m[i]) < (1e-7 if i <= 1 else 1e-4), (b, k, i, _rel(got_m[i], m[i]))
            assert _rel(_cov(got_L), _cov(L)) < 1e-5, (b, k)


@pytest.mark.parametrize("fact", ["isotropic", "blockdiag"])
@pytest.mark.parametrize("strategy", ["fixedinterval", "fixedinterval_aligned"])
@pytest.mark.parametrize("solver", ["solver_mle", "solver_dynamic"])
def test_fixedinterval_offgrid_marginals_on_a_fixed_grid(cuda, fact, strategy, solver):
    import torch

    s = H.spec(fact=fact, strategy=strategy, solver=solver)
    B = 3
    params, u0 = H.lv_ensemble(B, seed=52)
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
        # the filtering marginals the smoother keeps (SmoothingSolution.filtering)
        filt = sol.solution_full.filtering
        of = osol.solution_full.filtering
        for k in (0, 5, 16):
            m, L = _oracle_mean_chol(of[k], fact)
            assert _rel(filt.mean_flat[b, k, :2].cpu().numpy(), m[:2]) < 1e-9
            assert _rel(_cov(filt.cholesky_flat[b, k].cpu().numpy()), _cov(L)) < 1e-6 or np.max(np.abs(L)) == 0.0
        for k, t in enumerate(ts):
            m, L = _oracle_mean_chol(oslv.offgrid_marginals(t, solution=osol), fact)
            got_m = rv.mean_flat[b, k].cpu().numpy()
            got_L = rv.cholesky_flat[b, k].cpu().numpy()
            for i in range(5):
                assert _rel(got_m[i], m[i]) < (1e-7 if i <= 1 else 1e-4), (b, k, i, _rel(got_m[i], m[i]))
            assert _rel(_cov(got_L), _cov(L)) < 1e-5, (b, k)


def test_offgrid_marginals_reject_the_fixedpoint_smoother(cuda):
    s = H.spec(fact="isotropic", strategy="fixedpoint", clip_dt=False, error="residual_std", control="i")
    p_pdq, p_ivp, vf, ssm, slv, err, ctrl = H.product_build(s, H.BASE_LV)
    tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (np.asarray([20.0, 20.0]),), t=0.0)
    sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0, 2, 5), atol=1e-3, rtol=1e-3
    )
    with pytest.raises(NotImplementedError):
        slv.offgrid_marginals(0.3, solution=sol)
    # ... and an unbatched filter solve returns an unbatched marginal
    f = H.product_build(H.spec(fact="isotropic"), H.BASE_LV)
    sol_f = f[1].solve_adaptive_save_at(solver=f[4], error=f[5], control=f[6])(
        ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0, 2, 5), atol=1e-3, rtol=1e-3
    )
    rv = f[4].offgrid_marginals(np.asarray([0.3, 1.7]), solution=sol_f)
    assert rv.mean_flat.shape == (2, 5, 2) and rv.cholesky_flat.shape == (2, 5, 5)
