"""The CUDA path against outputs of the REFERENCE's own code (tests/golden/reference_numpy_backend.npz, produced by
running the unmodified reference on a NumPy backend: oracle/refshim, tests/golden/make_reference_golden.py): every
strategy x factorisation combination on Lotka-Volterra, the headline configuration at its full horizon, HIRES dense
ts1 and the Pleiades fixed-point smoother.  Accepted-step counts and checkpoint times must be the reference's, the ODE
solution agrees to 1e-8, higher Taylor coefficients and covariances to their conditioning."""

import numpy as np
import pytest

import pdeq_test_helpers as H
from test_reference_golden import CASES, check_against_reference

pytestmark = pytest.mark.gpu


def product_run(c):
    import torch

    s, prob = c["spec"], c["problem"]
    params = np.asarray(prob["params"])[None, :] if prob["params"] else None
    p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params)
    prior = ssm.prior_wiener_integrated(torch.from_numpy(c["tcoeffs"][None]).cuda())
    grid = np.asarray(c["grid"])
    if c["kind"] == "terminal":
        solve = p_ivp.solve_adaptive_terminal_values(solver=solver, error=err, control=ctrl, clip_dt=s["clip_dt"])
        sol = solve(prior, t0=grid[0], t1=grid[1], atol=c["atol"], rtol=c["rtol"], dt0=c["dt0"])
    elif c["kind"] == "save_at":
        solve = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl, clip_dt=s["clip_dt"], warn=False)
        sol = solve(prior, save_at=grid, atol=c["atol"], rtol=c["rtol"], dt0=c["dt0"])
    else:
        sol = p_ivp.solve_fixed_grid(solver=solver)(prior, grid=grid)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
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
