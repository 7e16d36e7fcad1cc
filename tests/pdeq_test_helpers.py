"""Shared helpers: build the same solver in the oracle and in the product from one spec."""

from __future__ import annotations

import numpy as np

from oracle import ivpsolve as o_ivp
from oracle import probdiffeq as o_pdq

BASE_LV = np.asarray([0.5, 0.05, 0.5, 0.05])


def lv_ensemble(B, seed=0):
    """one (B, 6) uniform draw, params then u0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    draw = rng.uniform(0.8, 1.2, size=(B, 6))
    params = BASE_LV[None, :] * draw[:, :4]
    u0 = 20.0 * draw[:, 4:]
    return params, u0


def spec(**kw):
    s = dict(vf="lotka_volterra", fact="isotropic", constraint="ts0", solver="solver", strategy="filter",
             error="state_std", control="pi", clip_dt=True, error_norm="scale_then_rms", derivative_idx=0,
             error_per_unit_step=False)  # fmt: skip
    s.update(kw)
    return s


def _build(mod_pdq, mod_ivp, s, vf):
    ssm = getattr(mod_pdq, "state_space_model_" + s["fact"])()
    cons = getattr(ssm, "constraint_ode_" + s["constraint"])(vf)
    if s["strategy"] == "fixedinterval_aligned":  # product-only option; the oracle takes it in solve_fixed_grid
        aligned = {"terminal": "aligned"} if mod_pdq is not o_pdq else {}
        strat = mod_pdq.strategy_smoother_fixedinterval(**aligned)
    else:
        strat = {"filter": mod_pdq.strategy_filter, "fixedpoint": mod_pdq.strategy_smoother_fixedpoint,
                 "fixedinterval": mod_pdq.strategy_smoother_fixedinterval}[s["strategy"]]()
    extra = {"constraint_init": cons} if s.get("constraint_init") else {}
    solver = getattr(mod_pdq, s["solver"])(strategy=strat, constraint=cons, **extra, **s.get("solver_kwargs", {}))
    norm = getattr(mod_pdq, "error_norm_" + s["error_norm"])()
    if s["error"] == "state_std":
        err = mod_pdq.error_state_std(constraint=cons, error_norm=norm, derivative_idx=s["derivative_idx"],
                                      error_per_unit_step=s["error_per_unit_step"], **s.get("error_kwargs", {}))  # fmt: skip
    else:
        err = mod_pdq.error_residual_std(constraint=cons, error_norm=norm,
                                         error_per_unit_step=s["error_per_unit_step"], **s.get("error_kwargs", {}))  # fmt: skip
    make_control = mod_ivp.control_proportional_integral if s["control"] == "pi" else mod_ivp.control_integral
    ctrl = make_control(**s.get("control_kwargs", {}))
    return ssm, solver, err, ctrl


def oracle_vf(s, params):
    return o_pdq.ode(s["vf"], params if params is not None and len(params) else None)


def oracle_solve_save_at(s, tcoeffs, params, save_at, atol, rtol, dt0=0.1, init_std=None, output_scale=None,
                         **solve_kwargs):  # fmt: skip
    """Run the oracle on ONE instance. Returns (solution, trace)."""
    vf = oracle_vf(s, params)
    ssm, solver, err, ctrl = _build(o_pdq, o_ivp, s, vf)
    if init_std is None:
        prior = ssm.prior_wiener_integrated(tcoeffs, output_scale=output_scale)
    else:
        prior = ssm.prior_wiener_integrated_diffuse(tcoeffs, init_std, output_scale=output_scale)
    trace = []
    solve = o_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl, clip_dt=s["clip_dt"], warn=False,
                                         trace=trace)  # fmt: skip
    return solve(prior, save_at=save_at, atol=atol, rtol=rtol, dt0=dt0, **solve_kwargs), trace


def oracle_solve_fixed(s, tcoeffs, params, grid, output_scale=None, init_std=None):
    vf = oracle_vf(s, params)
    ssm, solver, _err, _ctrl = _build(o_pdq, o_ivp, s, vf)
    if init_std is None:
        prior = ssm.prior_wiener_integrated(tcoeffs, output_scale=output_scale)
    else:
        prior = ssm.prior_wiener_integrated_diffuse(tcoeffs, init_std, output_scale=output_scale)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        terminal = "aligned" if s["strategy"] == "fixedinterval_aligned" else "reference"
        return o_ivp.solve_fixed_grid(solver=solver, terminal=terminal)(prior, grid=grid)


def product_build(s, params):
    from probdiffeq_b200 import ivpsolve as p_ivp
    from probdiffeq_b200 import probdiffeq as p_pdq

    vf = p_pdq.ode(s["vf"], params=params)
    return (p_pdq, p_ivp, vf, *_build(p_pdq, p_ivp, s, vf))


def oracle_chol_to_bnn(sol_chol, fact, d):
    """Oracle Cholesky (per checkpoint) -> the product's layout for one instance."""
    return np.asarray(sol_chol)


def cov_from_chol(L):
    return L @ np.swapaxes(L, -1, -2)
