"""The parity report's comparison logic (scripts/parity_report.py) exercised on the CPU: the "kernel" side is played by
the oracle re-run with dt0 moved by a few ulp, which must come out as identical sequences with tiny terminal errors,
and by a deliberately wrong tolerance, which must come out as a divergence far from the accept threshold."""

import multiprocessing as mp
import sys

import numpy as np

sys.path.insert(0, "scripts")


def _fake_product(pr, cfg, B, dt0, rtol):
    from oracle import problems as o_problems

    tc = o_problems.taylor_coefficients_batched("lotka_volterra", cfg["params"], cfg["inits"], 0.0, cfg["nu"])
    cfg2 = dict(cfg, rtol=rtol)
    with mp.get_context("spawn").Pool(2) as pool:
        ora = pr.run_oracle(cfg, tc, np.full(B, 0.1), pool)
        fake = pr.run_oracle(cfg2, tc, np.full(B, dt0), pool)
        cap = max(len(f["trace"]) for f in fake) + 4
        trace = np.full((B, cap, 4), np.nan)
        for b, f in enumerate(fake):
            trace[b, : len(f["trace"])] = f["trace"]
        prod = dict(tcoeffs=tc, dt0=np.full(B, 0.1), trace=trace,
                    num_attempts=np.asarray([len(f["trace"]) for f in fake]),
                    num_steps=np.stack([f["num_steps"] for f in fake]), status=np.zeros(B, dtype=np.int32),
                    mean=np.stack([f["mean"] for f in fake]), t=None)  # fmt: skip
        return pr.compare(cfg, prod, ora, pool)


def test_compare_identical_and_divergent():
    import parity_report as pr

    B = 4
    cfg = pr.config_inputs("2", B)
    cfg["save_at"] = np.asarray([0.0, 6.0])
    res = _fake_product(pr, cfg, B, 0.1 * (1 + 4.4e-16), cfg["rtol"])
    assert res["kernel_vs_oracle"]["identical_sequence"] == B
    assert res["oracle_vs_perturbed_oracle"]["identical_sequence"] == B
    assert res["terminal_coeff0"]["max_rel_identical"] < 1e-9
    assert res["divergences_not_explained_by_oracle_conditioning"] == []
    res = _fake_product(pr, cfg, B, 0.1, 3e-6)  # a "defect": the wrong tolerance
    assert res["kernel_vs_oracle"]["identical_sequence"] < B
    assert res["kernel_vs_oracle"]["max_dist_to_threshold_at_divergence"] > 1e-3  # nowhere near a tie
    assert len(res["divergences_not_explained_by_oracle_conditioning"]) >= 1  # ... and the control does not excuse it
    row = res["per_instance"][res["divergences_not_explained_by_oracle_conditioning"][0]["instance"]]
    assert row["kernel_vs_oracle"]["first_differing_attempt"] is not None
    assert row["oracle_vs_perturbed_oracle"]["first_differing_attempt"] is None


def test_pair_stats_prefix_and_amplification():
    import parity_report as pr

    a = np.asarray([[0.0, 0.1, 1.2, 1.0], [0.1, 0.2, 0.9, 0.0], [0.1, 0.15, 1.1, 1.0]])
    b = a.copy()
    b[1:, 1] *= 1 + 1e-7
    st = pr.pair_stats(a, b)
    assert st["first_differing_attempt"] is None
    assert st["dt_rel_diff_first_exceeds"]["1e-09"] == 1 and st["dt_rel_diff_first_exceeds"]["1e-06"] is None
    st = pr.pair_stats(a[:2], b)  # a prefix: the sequences part where the shorter one ends
    assert st["first_differing_attempt"] == 2
    b[2, 3] = 0.0
    st = pr.pair_stats(a, b)
    assert st["first_differing_attempt"] == 2 and abs(st["dist_to_threshold"] - 0.1) < 1e-12
