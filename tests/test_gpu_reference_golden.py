"""The CUDA path against outputs of the REFERENCE's own code (tests/golden/reference_numpy_backend.npz, produced by
running the unmodified reference on a NumPy backend: oracle/refshim, tests/golden/make_reference_golden.py): every
strategy x factorisation combination on Lotka-Volterra, the headline configuration at its full horizon, HIRES dense
ts1 and the Pleiades fixed-point smoother.  Accepted-step counts and checkpoint times must be the reference's, the ODE
solution agrees to 1e-8, higher Taylor coefficients and covariances to their conditioning."""

import numpy as np
import pytest

import pdeq_test_helpers as H
import test_reference_golden_aux as AUX
from test_reference_golden import CASES, check_against_reference, diffuse_std

pytestmark = pytest.mark.gpu


def product_solve(c):
    """Run the case's solve on the device; returns (solution, solver, probdiffeq module, ivpsolve module, vf)."""
    import torch

    s, prob = c["spec"], c["problem"]
    params = np.asarray(prob["params"])[None, :] if prob["params"] else None
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params)
    tcoeffs = torch.from_numpy(c["tcoeffs"][None]).cuda()
    scale = None if c.get("output_scale") is None else np.asarray(c["output_scale"])
    if c.get("diffuse_start"):
        prior = ssm.prior_wiener_integrated_diffuse(tcoeffs, torch.from_numpy(diffuse_std(c)).cuda(), output_scale=scale)
    else:
        prior = ssm.prior_wiener_integrated(tcoeffs, output_scale=scale, **(c.get("prior_kwargs") or {}))
    kw = c.get("solve_kwargs") or {}  # eps, damp
    grid = np.asarray(c["grid"])
    if c["kind"] == "terminal":
        solve = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl, clip_dt=s["clip_dt"])
        sol = solve(prior, t0=grid[0], t1=grid[1], atol=c["atol"], rtol=c["rtol"], dt0=c["dt0"], **kw)
    elif c["kind"] == "save_at":
        solve = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl, clip_dt=s["clip_dt"], warn=False)
        sol = solve(prior, save_at=grid, atol=c["atol"], rtol=c["rtol"], dt0=c["dt0"], **kw)
    else:
        sol = p_ivp.solve_fixed_grid(solver=solver)(prior, grid=grid)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    return sol, solver, p_pdq, p_ivp, vf


def product_run(c):
    sol = product_solve(c)[0]
    s = c["spec"]
    mean = sol.u.mean_flat[0].cpu().numpy()  # ([T,] n, d)
    L = sol.u.cholesky_flat[0].cpu().numpy()
    if s["fact"] == "blockdiag":
        mean = np.swapaxes(mean, -1, -2)  # the reference keeps (d, n)
    elif s["fact"] == "dense":
        mean = mean.reshape(*mean.shape[:-2], -1)  # coefficient-major flat state
    ref = c["ref"]

    def trimmed(x, like):
        x = np.asarray(x)
        # the product reports the initial point too (num_steps 0, output scale 1); the reference starts after it
        return x[1:] if x.ndim >= 1 and like.ndim >= 1 and x.shape[0] == like.shape[0] + 1 else x

    return dict(t=sol.t[0].cpu().numpy(), num_steps=trimmed(sol.num_steps[0].cpu().numpy(), ref["num_steps"]),
                output_scale=trimmed(sol.output_scale[0].cpu().numpy(), ref["output_scale"]), mean=mean,
                cov=L @ np.swapaxes(L, -1, -2))  # fmt: skip


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_reproduces_the_reference(cuda, c):
    check_against_reference(c, product_run(c))


def product_aux_outputs(c):
    """dt0 / dt0_adaptive / loss_lml_* / offgrid_marginals of the CUDA path, in the reference's layouts."""
    import torch

    prob, params, u0 = AUX.problem_of(c)
    a = c["arrays"]
    if c["aux"] == "taylor":
        from probdiffeq_b200 import probdiffeq as p_pdq

        vf = p_pdq.ode(prob["vf"], params=None if params is None else params[None, :])
        inits = (u0[None, :],) if c["du0"] is None else (u0[None, :], np.asarray(c["du0"])[None, :])
        tcoeffs, _ = getattr(p_pdq, c["alg"])(num=c["num"])(vf, inits, t=c["t"])
        return dict(tcoeffs=tcoeffs[0].cpu().numpy())
    if c["aux"] in ("dt0", "dt0_adaptive"):
        from probdiffeq_b200 import ivpsolve as p_ivp
        from probdiffeq_b200 import probdiffeq as p_pdq

        vf = p_pdq.ode(prob["vf"], params=None if params is None else params[None, :])
        if c["aux"] == "dt0":
            return dict(value=p_ivp.dt0(vf, (u0[None, :],), t=0.0).cpu().numpy()[0])
        kw = {k: c[k] for k in ("error_contraction_rate", "rtol", "atol")}
        return dict(value=p_ivp.dt0_adaptive(vf, (u0[None, :],), 0.0, **kw).cpu().numpy()[0])
    b = dict(c["base"], tcoeffs=a["tcoeffs"])
    sol, solver, p_pdq, _p_ivp, _vf = product_solve(b)
    fact = b["spec"]["fact"]
    steps, want = sol.num_steps[0].cpu().numpy(), a["num_steps"]
    if steps.ndim == 1 and want.ndim == 1 and steps.shape[0] == want.shape[0] + 1:
        steps = steps[1:]  # the product reports the initial point too
    assert np.array_equal(np.ravel(steps), np.ravel(want)), (steps, want)
    if c["aux"] == "lml_terminal":
        return {f"lml{i}": p_pdq.loss_lml_terminal_values(tcoeff_index=i)(
            a[f"data{i}"][None], marginals=sol.u, std=a[f"std{i}"].reshape(1, -1)).cpu().numpy()[0] for i in (0, 1)}  # fmt: skip
    if c["aux"] == "lml_timeseries":
        post = sol.solution_full.posterior
        return {key: p_pdq.loss_lml_timeseries(average_pdfs=avg)(
            a["data"][None], posterior=post, std=a["std"][None]).cpu().numpy()[0]
            for key, avg in (("lml_avg", True), ("lml_sum", False))}  # fmt: skip
    if c["aux"] == "sample":
        post = sol.solution_full.posterior

        def stacked(base):  # list over Taylor coefficients of (B, T, d) -> (T, n, d)
            return np.stack([x[0].cpu().numpy() for x in post.sample(base=base[None])], axis=1)

        # diagonals of the factors the device draws with: conditional k (k = 1 .. T-1) belongs to grid point k - 1
        own = torch.cat([torch.diagonal(post.conditional.cholesky[0, 1:], dim1=-2, dim2=-1),
                         torch.diagonal(post.marginal.cholesky_flat[0], dim1=-2, dim2=-1)[None]]).cpu().numpy()  # fmt: skip
        base = AUX.draws_for(a, own)
        return dict(samples=stacked(base), samples_zero_draws=stacked(0.0 * base))
    rv = solver.offgrid_marginals(a["ts"], solution=sol)
    torch.cuda.synchronize()
    mean = rv.mean_flat[0].cpu().numpy()  # (K, n, d)
    L = rv.cholesky_flat[0].cpu().numpy()
    if fact == "blockdiag":
        mean = np.swapaxes(mean, -1, -2)
    elif fact == "dense":
        mean = mean.reshape(mean.shape[0], -1)
    return dict(mean=mean, cov=AUX.cov(L))


@pytest.mark.parametrize("c", AUX.CASES, ids=AUX.IDS)
def test_cuda_path_reproduces_the_reference_either_side_of_the_loop(cuda, c):
    AUX.check(c, product_aux_outputs(c))
