"""Golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py from the pinned oracle)."""

import pathlib
import sys

import numpy as np
import pytest

import pdeq_test_helpers as H

GOLDEN = pathlib.Path(__file__).parent / "golden" / "oracle_golden.npz"


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN))


def test_oracle_reproduces_its_golden_vectors(golden):
    sys.path.insert(0, str(GOLDEN.parent))
    import make_golden

    fresh = make_golden.build()
    assert set(fresh) == set(golden)
    for key, value in golden.items():
        assert np.allclose(fresh[key], value, rtol=1e-9, atol=1e-300), key
    assert int(golden["config1_num_steps"]) == 331 and int(golden["config1_num_attempts"]) == 344


def _cov(L):
    return L @ np.swapaxes(L, -1, -2)


@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors(cuda, golden):
    import torch

    # config 1
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(H.spec(), H.BASE_LV)
    sol = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(golden["config1_tcoeffs"]), t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6
    )
    torch.cuda.synchronize()
    assert int(sol.num_steps) == int(golden["config1_num_steps"])
    assert int(sol.num_attempts) == int(golden["config1_num_attempts"])
    # t = 50: the oracle's own 1-ulp sensitivity is 1e-8..1e-7 here
    assert np.allclose(sol.u.mean[0].cpu().numpy(), golden["config1_mean"][0], rtol=1e-6)
    # fixed grids: 1e-10
    params, u0, grid = golden["fixed_params"], golden["fixed_u0"], golden["fixed_grid"]
    for fact in ("isotropic", "blockdiag"):
        for slv in ("solver", "solver_mle"):
            p_pdq, p_ivp, vf, ssm, solver, _e, _c = H.product_build(H.spec(fact=fact, solver=slv), params)
            tc, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
            s = p_ivp.solve_fixed_grid(solver=solver)(ssm.prior_wiener_integrated(tc), grid=grid)
            torch.cuda.synchronize()
            gm, gc = golden[f"fixed_{fact}_{slv}_mean"], golden[f"fixed_{fact}_{slv}_chol"]
            assert np.max(np.abs(s.u.mean_flat.cpu().numpy() - gm)) <= 1e-10 * np.max(np.abs(gm))
            cov, gcov = _cov(s.u.cholesky_flat.cpu().numpy()), _cov(gc)
            for k in range(1, len(grid)):
                assert np.max(np.abs(cov[:, k] - gcov[:, k])) <= 1e-10 * np.max(np.abs(gcov[:, k])), (fact, slv, k)
    # fixed-point smoother
    s = H.spec(fact="blockdiag", strategy="fixedpoint", solver="solver", error="residual_std", control="i", clip_dt=False)
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params[:1])
    tc, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0[:1],), t=0.0)
    sol = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tc), save_at=golden["smoother_save_at"], atol=1e-6, rtol=1e-4
    )
    torch.cuda.synchronize()
    assert np.array_equal(sol.num_steps[0, 1:].cpu().numpy(), golden["smoother_num_steps"])
    assert np.allclose(sol.u.mean_flat[0].cpu().numpy(), golden["smoother_mean"], rtol=1e-7, atol=1e-9)
    assert np.allclose(_cov(sol.u.cholesky_flat[0].cpu().numpy()), _cov(golden["smoother_chol"]), rtol=1e-5, atol=1e-14)
    # its time-series log-marginal-likelihood
    post = sol.solution_full.posterior
    for avg, key in ((True, "lml_mean_of_pdfs"), (False, "lml_sum_of_pdfs")):
        lml = p_pdq.loss_lml_timeseries(average_pdfs=avg)(golden["lml_data"], posterior=post, std=golden["lml_std"])
        assert np.isclose(float(lml[0]), float(golden[key]), rtol=1e-8), (key, float(lml[0]), float(golden[key]))
    # dense output of a filter solution
    s = H.spec(fact="isotropic", solver="solver_mle", error="residual_std", control="i", clip_dt=False)
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params[:1])
    sol = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl)(
        ssm.prior_wiener_integrated(tc), save_at=golden["smoother_save_at"], atol=1e-6, rtol=1e-4
    )
    rv = solver.offgrid_marginals(golden["offgrid_t"], solution=sol)
    torch.cuda.synchronize()
    assert np.allclose(rv.mean_flat[0, :, :2].cpu().numpy(), golden["offgrid_mean"][:, :2], rtol=1e-8, atol=1e-10)
    assert np.allclose(rv.mean_flat[0].cpu().numpy(), golden["offgrid_mean"], rtol=1e-4, atol=1e-7)
    assert np.allclose(_cov(rv.cholesky_flat[0].cpu().numpy()), _cov(golden["offgrid_chol"]), rtol=1e-5, atol=1e-16)
