"""Time BASELINE.json configs 3, 4a, 4b, 5 on one GPU (configs[1] is bench.py's workload).

Each config is first run on a small ensemble; the full ensemble is only run when the extrapolated time fits the
budget given on the command line (seconds, default 60). Prints one JSON line per run.
"""

import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from probdiffeq_b200 import ivpsolve, probdiffeq
from probdiffeq_b200 import problems as pb

BUDGET, ONLY = 60.0, None


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3, out


def report(name, B, secs, sol, extra=None):
    steps = int(sol.num_steps.reshape(B, -1)[:, -1].sum().item())
    att = int(sol.num_attempts.sum().item())
    bad = int((sol.status != 0).sum().item())
    line = dict(config=name, instances=B, seconds=secs, accepted_steps=steps, attempts=att,
                steps_per_s=steps / secs, attempts_per_s=att / secs, failed=bad)  # fmt: skip
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)
    return secs


def ladder(name, make, sizes):
    """Run increasing ensemble sizes while the extrapolated time stays within the budget."""
    prev = None
    for B in sizes:
        if prev is not None and prev[1] / prev[0] * B > BUDGET:
            print(json.dumps(dict(config=name, skipped_instances=B, estimated_seconds=prev[1] / prev[0] * B)), flush=True)
            break
        run = make(B)
        run()  # warm-up (module load, allocator)
        secs, (sol, extra) = timed(run)
        report(name, B, secs, sol, extra)
        prev = (B, secs)


def config3(B):
    u0 = pb.pleiades_ensemble(B, seed=1)
    vf = probdiffeq.ode("pleiades")
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
    ts0 = ssm.constraint_ode_ts0(vf)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts0)
    error = probdiffeq.error_residual_std(constraint=ts0)
    solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, control=ivpsolve.control_integral())
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    prior = ssm.prior_wiener_integrated(tcoeffs)
    save_at = np.linspace(0.0, 3.0, 33)
    return lambda: (solve(prior, save_at=save_at, atol=1e-9, rtol=1e-6, dt0=dt0), None)


def config4a(B):
    rng = np.random.Generator(np.random.PCG64(2))
    u0 = np.repeat(pb.HIRES_U0[None, :], B, axis=0)
    sc = rng.uniform(0.9, 1.1, size=(B, 2))
    u0[:, 0] *= sc[:, 0]
    u0[:, 7] *= sc[:, 1]
    vf = probdiffeq.ode("hires")
    ssm = probdiffeq.state_space_model_dense()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
    ts1 = ssm.constraint_ode_ts1(vf)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
    error = probdiffeq.error_residual_std(constraint=ts1)
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ivpsolve.control_proportional_integral())
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    prior = ssm.prior_wiener_integrated(tcoeffs)
    return lambda: (solve(prior, t0=0.0, t1=321.8122, atol=1e-11, rtol=1e-8, dt0=dt0, want_cholesky=False), None)


def config4b(B):
    rng = np.random.Generator(np.random.PCG64(2))
    rng.uniform(0.9, 1.1, size=(16384, 2))  # continue the stream of config 4a
    u0 = 2.0 * rng.uniform(0.9, 1.1, size=(B, 1))
    du0 = np.zeros((B, 1))
    vf = probdiffeq.ode("vanderpol", params=np.full((B, 1), 1e3))
    ssm = probdiffeq.state_space_model_dense()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0, du0), t=0.0)
    ts1 = ssm.constraint_ode_ts1(vf)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
    error = probdiffeq.error_state_std(constraint=ts1)
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ivpsolve.control_integral())
    prior = ssm.prior_wiener_integrated(tcoeffs)
    return lambda: (solve(prior, t0=0.0, t1=6.3, atol=1e-11, rtol=1e-8, want_cholesky=False), None)


def config5(B, d=1024, constraint="ts0"):
    rng = np.random.Generator(np.random.PCG64(3))
    nu = 0.01 * rng.uniform(0.5, 2.0, size=(B, 1))
    u0 = np.repeat(pb.burgers_u0(d)[None, :], B, axis=0)
    vf = probdiffeq.ode("burgers", params=nu)
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0,), t=0.0)
    cons = ssm.constraint_ode_ts0(vf) if constraint == "ts0" else ssm.constraint_ode_ts1(vf)
    solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=cons)
    error = probdiffeq.error_state_std(constraint=cons)
    solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ivpsolve.control_proportional_integral())
    prior = ssm.prior_wiener_integrated(tcoeffs)
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    data = torch.zeros((1, d), dtype=torch.float64, device="cuda")
    lml = probdiffeq.loss_lml_terminal_values()

    def run():
        sol = solve(prior, t0=0.0, t1=1.0, atol=1e-7, rtol=1e-4, dt0=dt0)
        ll = lml(data, marginals=sol.u, std=torch.full((1, d), 1e-2, dtype=torch.float64, device="cuda"))
        return sol, dict(ensemble_lml=float(ll.sum().item()))

    return run


CONFIGS = {
    "3:pleiades-bd-fixedpoint": (config3, [64, 1024, 8192, 65536]),
    "4a:hires-dense-ts1": (config4a, [16, 256, 2048, 16384]),
    "4b:vanderpol-dense-ts1": (config4b, [256, 4096, 16384]),
    "5:burgers-bd-ts0-d1024+lml": (config5, [8, 128, 1024, 4096]),
    "5b:burgers-bd-ts1-d1024+lml": (lambda B: config5(B, constraint="ts1"), [8, 128, 1024, 4096]),
    "5c:burgers-bd-ts1-d256+lml": (lambda B: config5(B, d=256, constraint="ts1"), [8, 128, 1024, 4096]),
}

if __name__ == "__main__":
    BUDGET = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    SIZES = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None  # override the size ladder
    torch.cuda.set_device(0)
    for name, (make, sizes) in CONFIGS.items():
        if ONLY and not any(name.startswith(o) for o in ONLY):
            continue
        t0 = time.time()
        try:
            ladder(name, make, SIZES or sizes)
        except Exception as exc:  # keep going: one config failing must not hide the others
            print(json.dumps(dict(config=name, error=repr(exc))), flush=True)
        print(f"# {name}: {time.time() - t0:.1f} s wall", flush=True)
