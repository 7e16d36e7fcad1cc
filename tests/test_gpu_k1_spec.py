"""The compile-time specialised builds of the headline loop kernel (ThreadLoop SPEC = 1..2,
csrc/pdeq_loop_thread.cuh) against the general kernel (SPEC = 0), which the oracle parity tests pin.

The specialisation removes run-time branches and moves the accepted state; it does not change a single floating-
point operation, so the bar is BITWISE equality of every output, on an ensemble large enough to fill the GPU
several times over (instances are pulled from a global counter, so lanes see different instance sequences).
The launcher reads PDEQ_K1_SPEC on every launch.
"""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(B, control, trace_capacity=0, t1=20.0, seed=5):
    import torch

    from probdiffeq_b200 import ivpsolve, probdiffeq, problems

    params_np, u0_np = problems.lotka_volterra_ensemble(B, seed=seed)
    dev = torch.device("cuda", 0)
    params, u0 = torch.from_numpy(params_np).to(dev), torch.from_numpy(u0_np).to(dev)
    ssm = probdiffeq.state_space_model_isotropic()
    vf = probdiffeq.ode("lotka_volterra", params=params)
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
    prior = ssm.prior_wiener_integrated(tcoeffs)
    ts0 = ssm.constraint_ode_ts0(vf)
    solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
    error = probdiffeq.error_state_std(constraint=ts0)
    ctrl = ivpsolve.control_proportional_integral() if control == "pi" else ivpsolve.control_integral()
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ctrl)
    sol = solve(prior, t0=0.0, t1=t1, atol=1e-8, rtol=1e-6, trace_capacity=trace_capacity)
    torch.cuda.synchronize()
    outs = [sol.t, sol.u.mean_flat, sol.u.cholesky_flat, sol.output_scale, sol.num_steps, sol.num_attempts, sol.status]
    if trace_capacity:
        outs.append(sol.trace)
    return [o.cpu().numpy().copy() for o in outs]


@pytest.fixture
def spec_env():
    old = os.environ.get("PDEQ_K1_SPEC")
    yield
    if old is None:
        os.environ.pop("PDEQ_K1_SPEC", None)
    else:
        os.environ["PDEQ_K1_SPEC"] = old


@pytest.mark.parametrize("control", ["pi", "i"])
def test_specialised_builds_are_bitwise_the_general_kernel(cuda, spec_env, control):
    B = 100_003  # ragged: not a multiple of the CTA size, several waves of the persistent grid
    os.environ["PDEQ_K1_SPEC"] = "0"
    ref = _solve(B, control)
    assert int(np.abs(ref[-1]).max()) == 0
    assert int(ref[4].min()) > 10
    for spec in (1, 2):
        os.environ["PDEQ_K1_SPEC"] = str(spec)
        got = _solve(B, control)
        for a, b in zip(got, ref):
            assert a.shape == b.shape and a.dtype == b.dtype
            assert a.tobytes() == b.tobytes(), f"spec {spec}: output differs from the general kernel"


def test_specialised_builds_emit_the_same_attempt_trace(cuda, spec_env):
    B = 257
    os.environ["PDEQ_K1_SPEC"] = "0"
    ref = _solve(B, "pi", trace_capacity=256, t1=10.0)
    n_att = ref[5]
    assert int(n_att.max()) <= 256
    for spec in (1, 2):
        os.environ["PDEQ_K1_SPEC"] = str(spec)
        got = _solve(B, "pi", trace_capacity=256, t1=10.0)
        for a, b in zip(got[:-1], ref[:-1]):
            assert a.tobytes() == b.tobytes()
        for b_ in range(B):  # rows past num_attempts are uninitialised scratch
            k = int(n_att[b_])
            assert got[-1][b_, :k].tobytes() == ref[-1][b_, :k].tobytes()
