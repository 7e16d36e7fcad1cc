"""A/B of the group kernel's specialised build (PDEQ_K2_SPEC=0/1, GroupLoop SPEC in csrc/pdeq_loop_group.cuh) on
BASELINE config 3 (Pleiades, fixed-point smoother) and on the Pleiades filter: time per pass and a BITWISE
comparison of all outputs with the general kernel. usage: python scripts/sweep_k2_spec.py [instances] [specs, e.g. 0,1,2]
(PDEQ_K2_SPEC=2: the smoother build that computes the backward conditional for accepted steps only.)"""

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import bench_configs as bc  # noqa: E402
from probdiffeq_b200 import ivpsolve, probdiffeq  # noqa: E402
from probdiffeq_b200 import problems as pb  # noqa: E402


def pleiades_filter(B):
    u0 = pb.pleiades_ensemble(B, seed=1)
    vf = probdiffeq.ode("pleiades")
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
    ts0 = ssm.constraint_ode_ts0(vf)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts0)
    error = probdiffeq.error_residual_std(constraint=ts0)
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ivpsolve.control_integral())
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    prior = ssm.prior_wiener_integrated(tcoeffs)
    return lambda: (solve(prior, t0=0.0, t1=3.0, atol=1e-9, rtol=1e-6, dt0=dt0), None)


def outputs(sol):
    outs = [sol.t, sol.u.mean_flat, sol.u.cholesky_flat, sol.output_scale, sol.num_steps, sol.num_attempts, sol.status]
    return [o.clone() for o in outs if o is not None]


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    specs = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 1, 0, 1)
    torch.cuda.set_device(0)
    for name, make in (("3:pleiades-bd-fixedpoint", bc.config3), ("pleiades-bd-filter", pleiades_filter)):
        run = make(B)
        ref = None
        for spec in specs:
            os.environ["PDEQ_K2_SPEC"] = str(spec)
            run()
            secs, (sol, _) = bc.timed(run)
            outs = outputs(sol)
            if ref is None:
                ref = outs
            same = all(torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8)) for a, b in zip(outs, ref))
            print(json.dumps(dict(config=name, spec=spec, instances=B, seconds=secs,
                                  attempts=int(sol.num_attempts.sum().item()), failed=int((sol.status != 0).sum().item()),
                                  bitwise_equal_to_first=bool(same))), flush=True)
