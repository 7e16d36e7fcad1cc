"""Generate tests/golden/oracle_golden.npz from the oracle.

These vectors are produced by the ORACLE (not by the reference): they freeze the restatement as regression
fixtures -- a CPU test re-derives them, and a GPU test compares the CUDA path with them. Outputs of the REFERENCE's own
code are a separate set: reference_numpy_backend*.npz, written by make_reference_golden*.py (the unmodified reference
on the NumPy backend of oracle/refshim); those are what pin the oracle.

    python tests/golden/make_golden.py
"""

import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import pdeq_test_helpers as H  # noqa: E402
from oracle import probdiffeq as o_pdq  # noqa: E402


def build():
    out = {}
    # BASELINE config 1: the reference's own single IVP (benchmarks/A0 wiring with nu = 4)
    vf = o_pdq.ode("lotka_volterra")
    tc, _ = o_pdq.jetexpand_ode_padded_scan(num=4)(vf, (np.asarray([20.0, 20.0]),), t=0.0)
    sol, trace = H.oracle_solve_save_at(H.spec(), tc, H.BASE_LV, np.asarray([0.0, 50.0]), 1e-8, 1e-6)
    out["config1_tcoeffs"] = tc
    out["config1_mean"] = sol.u_mean[-1]
    out["config1_chol"] = sol.u_chol[-1]
    out["config1_num_steps"] = np.asarray(sol.num_steps[-1])
    out["config1_num_attempts"] = np.asarray(len(trace))
    # a fixed-grid solve (1e-10 parity class) for each solver on a small randomised ensemble
    params, u0 = H.lv_ensemble(3, seed=42)
    grid = np.linspace(0.0, 1.0, 11)
    out["fixed_params"], out["fixed_u0"], out["fixed_grid"] = params, u0, grid
    for fact in ("isotropic", "blockdiag"):
        for solver in ("solver", "solver_mle"):
            means, chols = [], []
            for b in range(3):
                tcb = o_pdq.ode("lotka_volterra", params[b]).taylor_coefficients((u0[b],), 0.0, 4)
                s = H.oracle_solve_fixed(H.spec(fact=fact, solver=solver), tcb, params[b], grid)
                means.append(s.u_mean)
                chols.append(s.u_chol)
            out[f"fixed_{fact}_{solver}_mean"] = np.stack(means)
            out[f"fixed_{fact}_{solver}_chol"] = np.stack(chols)
    # fixed-point smoother on a save_at grid (smoothed marginals)
    save_at = np.linspace(0.0, 2.0, 6)
    s = H.spec(fact="blockdiag", strategy="fixedpoint", solver="solver", error="residual_std", control="i", clip_dt=False)
    tcb = o_pdq.ode("lotka_volterra", params[0]).taylor_coefficients((u0[0],), 0.0, 4)
    sol, trace = H.oracle_solve_save_at(s, tcb, params[0], save_at, 1e-6, 1e-4)
    out["smoother_save_at"] = save_at
    out["smoother_mean"] = sol.u_mean
    out["smoother_chol"] = sol.u_chol
    out["smoother_num_steps"] = np.asarray(sol.num_steps)
    # ... its time-series log-marginal-likelihood for fixed data
    rng = np.random.Generator(np.random.PCG64(2024))
    data = np.asarray(sol.u_mean)[:, 0] + 0.1 * rng.normal(size=(6, 2))
    std = 0.1 + 0.02 * np.arange(6)[:, None] * np.ones((1, 2))
    out["lml_data"], out["lml_std"] = data, std
    post = sol.solution_full.posterior
    out["lml_mean_of_pdfs"] = np.asarray(o_pdq.loss_lml_timeseries()(data, posterior=post, std=std))
    out["lml_sum_of_pdfs"] = np.asarray(o_pdq.loss_lml_timeseries(average_pdfs=False)(data, posterior=post, std=std))
    # dense output of a filter solution between the checkpoints
    sf = H.spec(fact="isotropic", solver="solver_mle", error="residual_std", control="i", clip_dt=False)
    solf, _ = H.oracle_solve_save_at(sf, tcb, params[0], save_at, 1e-6, 1e-4)
    import pdeq_test_helpers as H2
    from oracle import ivpsolve as o_ivp

    _, oslv, _, _ = H2._build(o_pdq, o_ivp, sf, H.oracle_vf(sf, params[0]))
    ts = np.asarray([0.1, 0.9, 1.7])
    rvs = [oslv.offgrid_marginals(t, solution=solf) for t in ts]
    out["offgrid_t"] = ts
    out["offgrid_mean"] = np.stack([rv.mean for rv in rvs])
    out["offgrid_chol"] = np.stack([rv.chol for rv in rvs])
    return out


if __name__ == "__main__":
    data = build()
    path = pathlib.Path(__file__).with_name("oracle_golden.npz")
    np.savez_compressed(path, **data)
    print(path, {k: v.shape for k, v in data.items()})
