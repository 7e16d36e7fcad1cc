"""A/B of the tail compaction of the thread-per-instance kernel (csrc/pdeq_loop_thread.cuh, ThreadLoop::run) on one GPU.

For each ensemble size (2^20 = the whole BASELINE config-2 ensemble on one GPU, 2^17 = its per-GPU share on 8 GPUs, ...)
the headline solve is timed with the compaction off (PDEQ_K1_POOL=0: a warp runs until its last lane is done) and on
with several segment lengths (PDEQ_K1_SEG), and every output is compared byte for byte with the run without it: where
an instance runs must not change a single bit. `strong_eff_vs_2^20` is the throughput relative to the 2^20-instance
run of the same setting -- the strong-scaling efficiency a GPU of an N-GPU split of the 2^20 ensemble would see.

usage: python scripts/sweep_k1_pool.py [--sizes 1048576,131072] [--segs 16,32,64] [--steps K]
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1048576,524288,262144,131072")
    ap.add_argument("--segs", default="16,32,64")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--spec", default=None)
    args = ap.parse_args()

    import torch

    from probdiffeq_b200 import ivpsolve, probdiffeq, problems, sharding

    dev = torch.device("cuda", 0)
    if args.spec is not None:
        os.environ["PDEQ_K1_SPEC"] = args.spec
    full = 1 << 20
    params_np, u0_np = problems.lotka_volterra_ensemble(full, seed=0)
    perm = sharding.permutation(full, seed=0)
    ssm = probdiffeq.state_space_model_isotropic()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = {}
    for B in [int(x) for x in args.sizes.split(",")]:
        idx = perm[:B]  # the first rank's shard of the permuted ensemble when 2^20 / B GPUs split it
        params = torch.from_numpy(params_np[idx]).to(dev)
        u0 = torch.from_numpy(u0_np[idx]).to(dev)
        vf = probdiffeq.ode("lotka_volterra", params=params)
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
        prior = ssm.prior_wiener_integrated(tcoeffs)
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error,
                                                        control=ivpsolve.control_proportional_integral())
        ref = None
        settings = [("off", None)] + [("on", int(s)) for s in args.segs.split(",")]
        for pool, seg in settings:
            os.environ["PDEQ_K1_POOL"] = "0" if pool == "off" else "1"
            if seg is not None:
                os.environ["PDEQ_K1_SEG"] = str(seg)
            held = None
            for _ in range(3):
                cur = solve(prior, t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6)
                flush.fill_(1)
                held = cur
            del held, cur
            torch.cuda.synchronize()
            evs, sol = [], None
            for _ in range(args.steps):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sol = solve(prior, t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6)
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            outs = [o.clone() for o in (sol.t, sol.u.mean_flat, sol.u.cholesky_flat, sol.num_steps, sol.num_attempts, sol.status)]
            if ref is None:
                ref = outs
            same = all(torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8)) for a, b in zip(outs, ref))
            steps = int(sol.num_steps.sum().item())
            key = (pool, seg)
            rate = steps / (ms * 1e-3)
            if B == full:
                base[key] = rate
            line = dict(instances=B, pool=pool, seg=seg, ms=round(ms, 4), accepted_steps=steps, steps_per_s=rate,
                        failed=int((sol.status != 0).sum().item()), bitwise_equal_to_pool_off=bool(same))
            if key in base:
                line["strong_eff_vs_2^20"] = rate / base[key]
            print(json.dumps(line), flush=True)
    os.environ.pop("PDEQ_K1_POOL", None)
    os.environ.pop("PDEQ_K1_SEG", None)


if __name__ == "__main__":
    main()
