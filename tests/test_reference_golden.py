"""The oracle against outputs of the REFERENCE's own code (tests/golden/reference_numpy_backend.npz).

The fixtures were produced in the build container by running the unmodified modules of /root/reference on a NumPy
implementation of their array backend (oracle/refshim; tests/golden/make_reference_golden.py, which also records how
closely the oracle agreed when the fixtures were made: reference_numpy_backend.report.json).  This is what pins the
restatement to the reference's algorithm, line for line; XLA's arithmetic itself is not reproduced."""

import json
import pathlib

import numpy as np
import pytest

import pdeq_test_helpers as H

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "reference_numpy_backend.npz"


def load_cases():
    data = np.load(GOLDEN, allow_pickle=False)
    names = sorted({k.split("/")[0] for k in data.files})
    cases = []
    for name in names:
        c = json.loads(str(data[f"{name}/meta"]))
        c["ref"] = {k: np.asarray(data[f"{name}/{k}"]) for k in ("t", "num_steps", "output_scale", "mean", "cov")}
        c["tcoeffs"] = np.asarray(data[f"{name}/tcoeffs"])
        cases.append(c)
    return cases


CASES = load_cases()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def solution_coefficient(mean, fact, i):
    """Taylor coefficient i of a mean in the reference's flat layout (leading checkpoint axis or not)."""
    if fact == "isotropic":
        return mean[..., i, :]
    if fact == "blockdiag":
        return mean[..., :, i]
    raise ValueError(fact)


def check_against_reference(c, got):
    """`got`: dict like c['ref'] in the same layouts. Accepted-step counts and checkpoint times must be the reference's
    and the ODE solution (Taylor coefficient 0) must agree to 1e-8 (or 100 x the reference's own one-ulp sensitivity of
    that coefficient where that is larger: the full-horizon stiff solves); all coefficients, covariances and output scales to
    max(stated tolerance, 100 x the reference's own sensitivity to a one-ulp change of its input) -- the fixtures carry
    that sensitivity, measured when they were made. A quantity the reference itself moves by more than a per cent under
    one ulp is not compared: that is the terminal covariance and output scale of the stiff, dynamically calibrated HIRES
    solve (a factor of three; its last whitened residual is rounding noise) -- everything else moves by 1e-16 ... 1e-5."""
    ref, s, sens = c["ref"], c["spec"], c["reference_one_ulp_sensitivity"]
    if c.get("step_count_rtol", 0.0) > 0.0:  # full-horizon solves: the reference's own count is not stable to one ulp
        diff = np.max(np.abs(np.asarray(got["num_steps"]) - ref["num_steps"]) / ref["num_steps"])
        assert diff <= c["step_count_rtol"], (c["name"], got["num_steps"], ref["num_steps"])
        k = c.get("exact_step_prefix", 0)  # ... but identical at the first checkpoints, where the reference's are
        if k:
            assert np.array_equal(np.asarray(got["num_steps"])[:k], ref["num_steps"][:k]), (c["name"], got["num_steps"])
    else:
        assert np.array_equal(np.asarray(got["num_steps"]), ref["num_steps"]), (c["name"], got["num_steps"], ref["num_steps"])
    assert rel(got["t"], ref["t"]) < 1e-13
    n = c["problem"]["nu"] + 1
    d = len(c["problem"]["u0"])
    mean_g, mean_r = np.asarray(got["mean"]), ref["mean"]
    assert mean_g.shape == mean_r.shape
    if s["fact"] == "dense":
        lead = mean_r.shape[:-1]
        c0g, c0r = mean_g.reshape(*lead, n, d)[..., 0, :], mean_r.reshape(*lead, n, d)[..., 0, :]
    else:
        c0g, c0r = solution_coefficient(mean_g, s["fact"], 0), solution_coefficient(mean_r, s["fact"], 0)
    assert rel(c0g, c0r) < max(1e-8, 100.0 * sens.get("solution", 0.0)), (c["name"], rel(c0g, c0r))
    for key, stated in (("mean", 1e-6), ("cov", 1e-5), ("output_scale", 1e-6)):
        if sens[key] > 1e-2:
            continue  # not a reproducible quantity: the reference itself moves it by more than a per cent under one ulp
        tol = max(stated, 100.0 * sens[key])
        assert rel(got[key], ref[key]) < tol, (c["name"], key, rel(got[key], ref[key]), tol)


def diffuse_std(c):
    """The diffuse start of the constraint_init cases: exact initial value, unit standard deviation on every derivative
    ((n,) for the isotropic model, (n, d) otherwise) -- tests/golden/make_reference_golden.py:diffuse_std."""
    n, d = c["problem"]["nu"] + 1, len(c["problem"]["u0"])
    std = np.ones((n,) if c["spec"]["fact"] == "isotropic" else (n, d))
    std[0] = 0.0
    return std


def oracle_run(c):
    s, prob = c["spec"], c["problem"]
    params = np.asarray(prob["params"]) if prob["params"] else None
    grid = np.asarray(c["grid"])
    init_std = diffuse_std(c) if c.get("diffuse_start") else None
    scale = None if c.get("output_scale") is None else np.asarray(c["output_scale"])
    if c.get("prior_kwargs"):  # is_exact=False: every Taylor coefficient known to inexact_eps only
        n, d = c["tcoeffs"].shape
        init_std = np.full((n,) if s["fact"] == "isotropic" else (n, d), c["prior_kwargs"]["inexact_eps"])
    if c["kind"] == "fixed":
        sol = H.oracle_solve_fixed(s, c["tcoeffs"], params, grid, init_std=init_std, output_scale=scale)
    else:
        sol, _ = H.oracle_solve_save_at(s, c["tcoeffs"], params, grid, c["atol"], c["rtol"], dt0=c["dt0"],
                                        init_std=init_std, output_scale=scale, **(c.get("solve_kwargs") or {}))  # fmt: skip
        if c["kind"] == "terminal":
            sol = sol.terminal()
    u = sol.u
    if isinstance(u, list):
        mean = np.stack([r.mean for r in u])
        cov = np.stack([r.chol @ np.swapaxes(r.chol, -1, -2) for r in u])
    else:
        mean, cov = np.asarray(u.mean), u.chol @ np.swapaxes(u.chol, -1, -2)
    return dict(t=sol.t, num_steps=sol.num_steps, output_scale=sol.output_scale, mean=mean, cov=cov)


def test_fixtures_cover_the_strategy_factorisation_grid():
    names = {c["name"] for c in CASES}
    assert len(CASES) >= 28
    for fact in ("isotropic", "blockdiag", "dense"):
        for tail in ("solver_residual_i", "dynamic_residual_i", "mle_state_pi_ts1", "fixedpoint_mle",
                     "fixedgrid_filter", "fixedgrid_fixedinterval_mle"):  # fmt: skip
            assert f"lv_{fact}_{tail}" in names
    assert {"lv_iso_ts0_terminal_t50", "hires_dense_ts1_dynamic", "pleiades_blockdiag_fixedpoint",
            "vanderpol_dense_ts1_dynamic_state_i", "burgers_blockdiag_ts0_d16",
            "lv_isotropic_state_deriv2_unitstep_rms_then_scale", "lv_blockdiag_residual_unitstep_clipped_save_at",
            "lv_isotropic_constraint_init_fixedgrid", "lv_dense_ts1_constraint_init_fixedgrid_mle",
            "lv_blockdiag_ts1_constraint_init_adaptive_dynamic"} <= names  # fmt: skip
    report = json.loads((GOLDEN.with_suffix(".report.json")).read_text())
    assert all(r["same_step_counts"] or r.get("step_count_rel_diff", 1.0) <= 0.05 for r in report)


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_the_reference(c):
    check_against_reference(c, oracle_run(c))


def test_reference_runs_here_on_the_numpy_backend():
    """Where the reference sources exist (the build container), run one case again through the shim: the fixture is
    what the reference's code returns today."""
    from oracle import refshim

    if not refshim.available():
        pytest.skip("the reference sources are not on this machine")
    import importlib.util
    import warnings

    spec = importlib.util.spec_from_file_location("make_reference_golden", GOLDEN.parent / "make_reference_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    c = next(c for c in CASES if c["name"] == "lv_blockdiag_dynamic_residual_i")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = mod.reference_arrays(mod.run_reference(dict(c, tcoeffs=c["tcoeffs"])))
    for k, v in c["ref"].items():
        assert np.array_equal(np.asarray(got[k]), v), k
