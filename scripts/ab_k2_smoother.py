"""A/B of two builds of the library on the smoother paths of the group kernel (K2), bit for bit.

    PDEQ_B200_LIB=<old .so> python scripts/ab_k2_smoother.py dump old.npz
    python scripts/ab_k2_smoother.py dump new.npz
    python scripts/ab_k2_smoother.py compare old.npz new.npz

`dump` solves a handful of smoother problems (config 3's fixed-point smoother with overstepped checkpoints, the same
with clipped steps, the fixed-interval smoother on a fixed grid in both terminal conventions, the isotropic model, ts1) and stores every tensor of every solution; `compare` reports, per case and tensor, whether the
two files agree bitwise and otherwise the largest relative difference. A library is loaded once per process, hence the
two `dump` runs. (The built-in instantiations have no CTA-per-instance smoother; that mode shares the code.)"""

import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")


def tensors_of(obj, prefix="", seen=None, out=None):
    """Every torch tensor reachable through the attributes of a solution object."""
    out = {} if out is None else out
    seen = set() if seen is None else seen
    if id(obj) in seen or obj is None:
        return out
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        out[prefix.strip(".")] = obj.detach().cpu().numpy()
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            tensors_of(v, f"{prefix}{i}.", seen, out)
    elif hasattr(obj, "__dict__") and type(obj).__module__.startswith("probdiffeq_b200"):
        for k, v in sorted(vars(obj).items()):
            if k in ("prior",):
                continue
            tensors_of(v, f"{prefix}{k}.", seen, out)
    return out


def cases():
    import bench_configs as bc
    from probdiffeq_b200 import ivpsolve, probdiffeq
    from probdiffeq_b200 import problems as pb

    yield "config3_fixedpoint_save_at", lambda: bc.config3(256)()[0]

    def pleiades(strategy, B=64):
        u0 = pb.pleiades_ensemble(B, seed=1)
        vf = probdiffeq.ode("pleiades")
        ssm = probdiffeq.state_space_model_blockdiag()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
        ts0 = ssm.constraint_ode_ts0(vf)
        return vf, u0, ssm.prior_wiener_integrated(tcoeffs), ts0, strategy

    def clipped():
        vf, u0, prior, ts0, st = pleiades(probdiffeq.strategy_smoother_fixedpoint())
        solver = probdiffeq.solver_dynamic(strategy=st, constraint=ts0)
        error = probdiffeq.error_residual_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, clip_dt=True, warn=False)
        return solve(prior, save_at=np.linspace(0.0, 1.0, 9), atol=1e-9, rtol=1e-6, dt0=ivpsolve.dt0(vf, (u0,), t=0.0))

    yield "pleiades_fixedpoint_clipped", clipped

    def general_kernel_mle():
        vf, u0, prior, ts0, st = pleiades(probdiffeq.strategy_smoother_fixedpoint(), B=32)
        solver = probdiffeq.solver_mle(strategy=st, constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, warn=False)
        return solve(prior, save_at=np.linspace(0.0, 1.0, 5), atol=1e-8, rtol=1e-5, dt0=ivpsolve.dt0(vf, (u0,), t=0.0))

    yield "pleiades_fixedpoint_mle_state_std", general_kernel_mle

    for terminal in ("reference", "aligned"):

        def fixed(terminal=terminal):
            vf, u0, prior, ts0, st = pleiades(probdiffeq.strategy_smoother_fixedinterval(terminal=terminal), B=32)
            solver = probdiffeq.solver(strategy=st, constraint=ts0)
            return ivpsolve.solve_fixed_grid(solver=solver)(prior, grid=np.linspace(0.0, 0.5, 41))

        yield f"pleiades_fixedinterval_{terminal}", fixed

    def lv_iso():
        params, u0 = pb.lotka_volterra_ensemble(128, seed=0)
        vf = probdiffeq.ode("lotka_volterra", params=params)
        ssm = probdiffeq.state_space_model_isotropic()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver_mle(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, warn=False)
        return solve(ssm.prior_wiener_integrated(tcoeffs), save_at=np.linspace(0.0, 10.0, 21), atol=1e-8, rtol=1e-6)

    yield "lv_isotropic_fixedpoint", lv_iso

    def lv_bd_fixedinterval():
        params, u0 = pb.lotka_volterra_ensemble(64, seed=2)
        vf = probdiffeq.ode("lotka_volterra", params=params)
        ssm = probdiffeq.state_space_model_blockdiag()
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_smoother_fixedinterval(),
                                           constraint=ssm.constraint_ode_ts1(vf))
        return ivpsolve.solve_fixed_grid(solver=solver)(ssm.prior_wiener_integrated(tcoeffs), grid=np.linspace(0.0, 2.0, 101))

    yield "lv_blockdiag_ts1_fixedinterval", lv_bd_fixedinterval


def dump(path):
    torch.cuda.set_device(0)
    out = {}
    for name, run in cases():
        sol = run()
        torch.cuda.synchronize()
        for k, v in tensors_of(sol).items():
            out[f"{name}/{k}"] = v
        print(json.dumps(dict(case=name, tensors=len([k for k in out if k.startswith(name + "/")]),
                              failed=int((sol.status != 0).sum().item()))), flush=True)
    np.savez_compressed(path, **out)


def compare(pa, pb_):
    a, b = np.load(pa), np.load(pb_)
    worst = {}
    ok = True
    for k in a.files:
        if k not in b.files:
            print(json.dumps(dict(tensor=k, missing_in=pb_)))
            ok = False
            continue
        x, y = a[k], b[k]
        case = k.split("/")[0]
        same = x.shape == y.shape and x.tobytes() == y.tobytes()
        rel = 0.0
        if not same and x.shape == y.shape and x.dtype.kind == "f":
            fin = np.isfinite(x) & np.isfinite(y)
            den = np.maximum(np.abs(x[fin]), 1e-300)
            rel = float(np.max(np.abs(x[fin] - y[fin]) / den)) if fin.any() else 0.0
            if (np.isfinite(x) != np.isfinite(y)).any():
                rel = float("inf")
        elif not same:
            rel = float("inf")
        w = worst.setdefault(case, dict(case=case, tensors=0, bitwise_equal=0, max_rel_diff=0.0, worst_tensor=None))
        w["tensors"] += 1
        w["bitwise_equal"] += int(same)
        if rel > w["max_rel_diff"]:
            w["max_rel_diff"], w["worst_tensor"] = rel, k.split("/", 1)[1]
        ok = ok and same
    for w in worst.values():
        print(json.dumps(w))
    print(json.dumps(dict(all_bitwise_equal=ok)))


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        compare(sys.argv[2], sys.argv[3])
