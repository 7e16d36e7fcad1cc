"""Tiny invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from probdiffeq_b200 import ivpsolve as p_ivp  # noqa: E402
from probdiffeq_b200 import probdiffeq as p_pdq  # noqa: E402
from probdiffeq_b200 import problems as pb  # noqa: E402


def spec(**kw):
    s = dict(vf="lotka_volterra", fact="isotropic", constraint="ts0", solver="solver", strategy="filter",
             error="state_std", control="pi", clip_dt=True)  # fmt: skip
    s.update(kw)
    return s


def build(s, params):
    vf = p_pdq.ode(s["vf"], params=params)
    ssm = getattr(p_pdq, "state_space_model_" + s["fact"])()
    cons = getattr(ssm, "constraint_ode_" + s["constraint"])(vf)
    strat = {"filter": p_pdq.strategy_filter, "fixedpoint": p_pdq.strategy_smoother_fixedpoint,
             "fixedinterval": p_pdq.strategy_smoother_fixedinterval}[s["strategy"]]()  # fmt: skip
    slv = getattr(p_pdq, s["solver"])(strategy=strat, constraint=cons)
    err = getattr(p_pdq, "error_" + s["error"])(constraint=cons)
    ctrl = p_ivp.control_integral() if s["control"] == "i" else p_ivp.control_proportional_integral()
    return vf, ssm, slv, err, ctrl


class H:  # the two helpers this script used from the test suite
    spec = staticmethod(spec)
    lv_ensemble = staticmethod(pb.lotka_volterra_ensemble)


def run(s, params, u0, num, *, save_at=None, grid=None, t1=None, atol=1e-4, rtol=1e-3, dt0=0.1):
    vf, ssm, slv, err, ctrl = build(s, params)
    tc, _ = p_pdq.jetexpand_ode_padded_scan(num=num)(vf, u0, t=0.0)
    prior = ssm.prior_wiener_integrated(tc)
    if grid is not None:
        sol = p_ivp.solve_fixed_grid(solver=slv)(prior, grid=grid)
    elif save_at is not None:
        sol = p_ivp.solve_adaptive_save_at(solver=slv, error=err, control=ctrl, clip_dt=s["clip_dt"], warn=False)(
            prior, save_at=save_at, atol=atol, rtol=rtol, dt0=dt0)
    else:
        sol = p_ivp.solve_adaptive_terminal_values(solver=slv, error=err, control=ctrl, clip_dt=s["clip_dt"])(
            prior, t0=0.0, t1=t1, atol=atol, rtol=rtol, dt0=dt0)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0, sol.status
    return p_pdq, slv, sol


params, u0 = H.lv_ensemble(5, seed=0)
run(H.spec(), params, (u0,), 4, t1=1.0)  # K1
run(H.spec(fact="blockdiag", clip_dt=False), params, (u0,), 4, save_at=np.linspace(0, 1, 4))  # K1 blockdiag + interpolation
s = H.spec(fact="blockdiag", strategy="fixedpoint", solver="solver_dynamic", error="residual_std", control="i", clip_dt=False)
p_pdq, slv, sol = run(s, params, (u0,), 4, save_at=np.linspace(0, 1, 5))  # K2 warp mode, fixed-point smoother
post = sol.solution_full.posterior
lml = p_pdq.loss_lml_timeseries()(sol.u.mean[0], posterior=post, std=np.ones((5, 2)))
smp = post.sample(0, shape=(3,))
s = H.spec(fact="isotropic", strategy="fixedinterval", solver="solver_mle")
p_pdq, slv, sol = run(s, params, (u0,), 4, grid=np.linspace(0, 0.5, 6))  # K2 fixed grid, fixed-interval smoother
rv = slv.offgrid_marginals(np.asarray([0.05, 0.33]), solution=sol)
rng = np.random.Generator(np.random.PCG64(3))
d = 48
run(H.spec(vf="burgers", fact="blockdiag", solver="solver", error="state_std", control="pi"),
    0.01 * rng.uniform(0.5, 2.0, size=(2, 1)), (np.repeat(pb.burgers_u0(d)[None, :], 2, axis=0),), 3,
    t1=0.01, atol=1e-7, rtol=1e-4, dt0=1e-3)  # K2 CTA mode, one dimension per lane
d = 300
run(H.spec(vf="burgers", fact="blockdiag", solver="solver", error="state_std", control="pi"),
    0.01 * rng.uniform(0.5, 2.0, size=(2, 1)), (np.repeat(pb.burgers_u0(d)[None, :], 2, axis=0),), 3,
    t1=0.002, atol=1e-7, rtol=1e-4, dt0=1e-3)  # K2 CTA mode, two dimensions per lane
run(H.spec(fact="dense", constraint="ts1", solver="solver_dynamic", error="residual_std", control="pi"), params[:2], (u0[:2],), 4,
    t1=0.5)  # K3
run(H.spec(fact="dense", constraint="ts1", solver="solver", error="state_std", control="pi", clip_dt=False), params[:2], (u0[:2],), 4,
    save_at=np.linspace(0, 0.5, 3))  # K3 with interpolation and the state-std estimator
torch.cuda.synchronize()
print("sanitize smoke ok", float(lml.sum()), tuple(smp.flat.shape), tuple(rv.mean_flat.shape))
