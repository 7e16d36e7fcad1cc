"""A/B of longest-first service in the thread-per-instance kernel (pdeq_problem.order, `solve(..., cost_hint=)`).

For each ensemble size (2^20 = the whole BASELINE config-2 ensemble on one GPU, 2^17 = its per-GPU share on 8 GPUs, ...)
the headline solve is timed
  plain    index order, full grid (what round 1 measured);
  hint     `cost_hint` from a quadratic cost model fitted on a 4096-instance pilot solve (sharding.QuadraticCostModel):
           longest-first service; the model evaluation and the sort are inside the timed region;
  hintall  as hint, also beyond COST_HINT_MAX_ROUNDS instances per lane (where `solve` otherwise ignores the hint);
  oracle   hintall with the TRUE attempt counts of a previous solve as the hint (what a perfect predictor would give);
and every output is compared byte for byte with the plain run: where and when an instance runs must not change a bit.
`strong_eff_vs_2^20` = throughput relative to the 2^20-instance run of the same setting.

usage: python scripts/sweep_k1_order.py [--sizes 1048576,131072] [--steps K] [--modes plain,hint,oracle]
"""

from __future__ import annotations

import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1048576,524288,262144,131072")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--modes", default="plain,hint,oracle")
    ap.add_argument("--pilot", type=int, default=4096)
    args = ap.parse_args()

    import torch

    from probdiffeq_b200 import ivpsolve, probdiffeq, problems, sharding

    dev = torch.device("cuda", 0)
    full = 1 << 20
    params_np, u0_np = problems.lotka_volterra_ensemble(full, seed=0)
    perm = sharding.permutation(full, seed=0)
    ssm = probdiffeq.state_space_model_isotropic()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = {}
    default_rounds = ivpsolve.COST_HINT_MAX_ROUNDS
    for B in [int(x) for x in args.sizes.split(",")]:
        idx = perm[:B]  # the first rank's shard of the permuted ensemble when 2^20 / B GPUs split it
        params = torch.from_numpy(params_np[idx]).to(dev)
        u0 = torch.from_numpy(u0_np[idx]).to(dev)
        inputs = torch.cat([params, u0], dim=1)
        vf = probdiffeq.ode("lotka_volterra", params=params)
        tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
        prior = ssm.prior_wiener_integrated(tcoeffs)
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error,
                                                        control=ivpsolve.control_proportional_integral())  # fmt: skip
        kw = dict(t0=0.0, t1=50.0, atol=1e-8, rtol=1e-6)
        first = solve(prior, **kw)
        true_cost = first.num_attempts.to(torch.float64)
        M = min(args.pilot, B)
        model = sharding.QuadraticCostModel.fit(inputs[:M].cpu().numpy(), true_cost[:M].cpu().numpy())
        rank_err = float((model.predict(inputs) - true_cost).std().item())

        def run(mode):
            ivpsolve.COST_HINT_MAX_ROUNDS = 64 if mode in ("hintall", "oracle") else default_rounds
            if mode == "plain":
                return solve(prior, **kw)
            if mode in ("hint", "hintall"):
                return solve(prior, cost_hint=model.predict(inputs), **kw)
            return solve(prior, cost_hint=true_cost, **kw)

        ref = None
        for mode in args.modes.split(","):
            held = None
            for _ in range(3):
                cur = run(mode)
                flush.fill_(1)
                held = cur
            del held, cur
            torch.cuda.synchronize()
            evs, sol = [], None
            for _ in range(args.steps):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sol = run(mode)
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            outs = [o.clone() for o in (sol.t, sol.u.mean_flat, sol.u.cholesky_flat, sol.num_steps, sol.num_attempts, sol.status)]
            if ref is None:
                ref = outs
            same = all(torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8)) for a, b in zip(outs, ref))
            steps = int(sol.num_steps.sum().item())
            rate = steps / (ms * 1e-3)
            if B == full:
                base[mode] = rate
            line = dict(instances=B, mode=mode, ms=round(ms, 4), accepted_steps=steps, steps_per_s=rate,
                        failed=int((sol.status != 0).sum().item()), bitwise_equal_to_first_mode=bool(same),
                        cost_model_residual_std_attempts=round(rank_err, 2))  # fmt: skip
            if mode in base:
                line["strong_eff_vs_2^20"] = rate / base[mode]
            if "plain" in base:
                line["eff_vs_plain_2^20"] = rate / base["plain"]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
