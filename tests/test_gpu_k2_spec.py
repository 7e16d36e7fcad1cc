"""The specialised builds of the group kernel (GroupLoop SPEC = 1: solver_dynamic + error_residual_std fixed at
compile time; SPEC = 2: in addition the smoother's backward conditional is computed for accepted steps only --
csrc/pdeq_loop_group.cuh) against the general kernel, which the oracle parity tests pin: neither changes a floating-
point operation on an accepted step, so every output must agree BITWISE.
The launcher reads PDEQ_K2_SPEC on every launch."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(B, strategy):
    import torch

    from probdiffeq_b200 import ivpsolve, probdiffeq
    from probdiffeq_b200 import problems as pb

    u0 = pb.pleiades_ensemble(B, seed=7)
    vf = probdiffeq.ode("pleiades")
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
    ts0 = ssm.constraint_ode_ts0(vf)
    strat = probdiffeq.strategy_filter() if strategy == "filter" else probdiffeq.strategy_smoother_fixedpoint()
    solver = probdiffeq.solver_dynamic(strategy=strat, constraint=ts0)
    error = probdiffeq.error_residual_std(constraint=ts0)
    solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, control=ivpsolve.control_integral())
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    prior = ssm.prior_wiener_integrated(tcoeffs)
    sol = solve(prior, save_at=np.linspace(0.0, 1.0, 5), atol=1e-9, rtol=1e-6, dt0=dt0)
    torch.cuda.synchronize()
    outs = [sol.t, sol.u.mean_flat, sol.u.cholesky_flat, sol.output_scale, sol.num_steps, sol.num_attempts, sol.status]
    return [o.cpu().numpy().copy() for o in outs if o is not None]


@pytest.mark.parametrize("strategy", ["filter", "fixedpoint"])
def test_specialised_group_kernel_is_bitwise_the_general_kernel(cuda, strategy):
    old = os.environ.get("PDEQ_K2_SPEC")
    try:
        os.environ["PDEQ_K2_SPEC"] = "0"
        ref = _solve(1500, strategy)
        assert int(np.abs(ref[-1]).max()) == 0
        for spec in ("1", "2"):
            os.environ["PDEQ_K2_SPEC"] = spec
            got = _solve(1500, strategy)
            assert len(got) == len(ref)
            for a, b in zip(got, ref):
                assert a.shape == b.shape and a.tobytes() == b.tobytes(), f"PDEQ_K2_SPEC={spec}"
    finally:
        if old is None:
            os.environ.pop("PDEQ_K2_SPEC", None)
        else:
            os.environ["PDEQ_K2_SPEC"] = old
