"""The oracle against outputs of the REFERENCE's own code for the functions either side of the step loop
(tests/golden/reference_numpy_backend_aux.npz, written by tests/golden/make_reference_golden_aux.py from the unmodified
reference on the NumPy backend of oracle/refshim): Taylor-mode initialisation (jet_expansion_algorithms.py:49-152,
on an exact-rational `jet`), `dt0`, `dt0_adaptive` (stepsize_initialisers.py:7-78),
`loss_lml_terminal_values`, `loss_lml_timeseries` (estimators_and_losses.py:20-105), `MarkovSequence.sample` given the
draws (estimators_and_losses.py:233-271) and `solver.offgrid_marginals`
(solvers.py:149-203) for all three factorisations."""

import json
import pathlib

import numpy as np
import pytest

import pdeq_test_helpers as H
from oracle import ivpsolve as o_ivp
from oracle import probdiffeq as o_pdq

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "reference_numpy_backend_aux.npz"


def load_cases():
    data = np.load(GOLDEN, allow_pickle=False)
    names = sorted({k.split("/")[0] for k in data.files})
    cases = []
    for name in names:
        c = json.loads(str(data[f"{name}/meta"]))
        c["arrays"] = {k.split("/", 1)[1]: np.asarray(data[k]) for k in data.files
                       if k.startswith(name + "/") and not k.endswith("/meta")}  # fmt: skip
        cases.append(c)
    return cases


CASES = load_cases()
IDS = [c["name"] for c in CASES]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def cov(L):
    L = np.asarray(L)
    return L @ np.swapaxes(L, -1, -2)


def problem_of(c):
    prob = c["base"]["problem"]
    return prob, (np.asarray(prob["params"]) if prob["params"] else None), np.asarray(prob["u0"])


def check(c, got):
    """`got`: the case's outputs, computed by the implementation under test, in the reference's layouts. Scalars (step
    sizes, log-likelihoods) to 1e-9; off-grid means to 1e-6 and covariances to 1e-5 (high Taylor coefficients carry the
    solve's own conditioning, tests/test_reference_golden.py)."""
    a = c["arrays"]
    if c["aux"] == "taylor":
        # every coefficient to 1e-13 of its own size (the exact coefficients, correctly rounded, are the fixture)
        assert np.asarray(got["tcoeffs"]).shape == a["tcoeffs"].shape
        for i in range(a["tcoeffs"].shape[0]):
            assert rel(got["tcoeffs"][i], a["tcoeffs"][i]) < 1e-13, (c["name"], i)
    elif c["aux"] in ("dt0", "dt0_adaptive"):
        assert rel(got["value"], a["value"]) < 1e-12
    elif c["aux"] == "lml_terminal":
        for idx in (0, 1):
            assert rel(got[f"lml{idx}"], a[f"lml{idx}"]) < 1e-9, (c["name"], idx, got[f"lml{idx}"], a[f"lml{idx}"])
    elif c["aux"] == "lml_timeseries":
        for key in ("lml_avg", "lml_sum"):
            assert rel(got[key], a[key]) < 1e-7, (c["name"], key, got[key], a[key])
    elif c["aux"] == "sample":
        # with all draws zero the sample is the chain of conditional means; with the recorded draws it also depends on
        # the signs of the factors' columns, which the reference leaves to LAPACK: the implementation under test was
        # given the draws times (sign of its own factor diagonals) x (sign of the reference's), `draws_for`
        assert rel(got["samples_zero_draws"], a["samples_zero_draws"]) < 1e-8, c["name"]
        assert rel(got["samples"], a["samples"]) < 1e-7, (c["name"], rel(got["samples"], a["samples"]))
    else:
        assert np.asarray(got["mean"]).shape == a["mean"].shape
        assert rel(got["mean"], a["mean"]) < 1e-6, (c["name"], rel(got["mean"], a["mean"]))
        for k in range(len(a["ts"])):
            assert rel(got["cov"][k], a["cov"][k]) < 1e-5, (c["name"], k)


def draws_for(a, own_factor_diagonals):
    """The fixture's draws with the column-sign convention of the reference's factors translated into the
    implementation's: a factor of a covariance is unique up to the signs of its columns, and flipping a column's sign
    together with its draw leaves the sample unchanged."""
    own = np.sign(np.asarray(own_factor_diagonals)).reshape(a["base"].shape)
    return a["base"] * np.where(own == 0.0, 1.0, own) * a["factor_diag_sign"]


def oracle_outputs(c):
    b = dict(c["base"])
    prob, params, u0 = problem_of(c)
    s = b["spec"]
    ovf = o_pdq.ode(prob["vf"], params)
    if c["aux"] == "taylor":
        inits = [u0] if c["du0"] is None else [u0, np.asarray(c["du0"])]
        return dict(tcoeffs=np.asarray(ovf.taylor_coefficients(inits, c["t"], c["num"])))
    if c["aux"] == "dt0":
        return dict(value=o_ivp.dt0(ovf, (u0,), t=0.0))
    if c["aux"] == "dt0_adaptive":
        kw = {k: c[k] for k in ("error_contraction_rate", "rtol", "atol")}
        return dict(value=o_ivp.dt0_adaptive(ovf, (u0,), 0.0, **kw))
    a = c["arrays"]
    grid = np.asarray(b["grid"])
    if b["kind"] == "fixed":
        sol = H.oracle_solve_fixed(s, a["tcoeffs"], params, grid)
    else:
        sol, _ = H.oracle_solve_save_at(s, a["tcoeffs"], params, grid, b["atol"], b["rtol"], dt0=b["dt0"])
        if b["kind"] == "terminal":
            sol = sol.terminal()
    assert np.array_equal(np.asarray(sol.num_steps), a["num_steps"])
    if c["aux"] == "lml_terminal":
        return {f"lml{i}": o_pdq.loss_lml_terminal_values(tcoeff_index=i)(a[f"data{i}"], marginals=sol.u, std=a[f"std{i}"])
                for i in (0, 1)}  # fmt: skip
    if c["aux"] == "lml_timeseries":
        post = sol.solution_full.posterior.remove_filtering_distributions()
        return {key: o_pdq.loss_lml_timeseries(average_pdfs=avg)(a["data"], posterior=post, std=a["std"])
                for key, avg in (("lml_avg", True), ("lml_sum", False))}  # fmt: skip
    if c["aux"] == "sample":
        n, d = prob["nu"] + 1, len(prob["u0"])

        def stacked(states):
            return np.stack([x if s["fact"] == "isotropic" else (x.T if s["fact"] == "blockdiag" else x.reshape(n, d))
                             for x in states])  # fmt: skip

        post = sol.solution_full.posterior
        seq = post.remove_filtering_distributions()
        own = [np.diagonal(cnd.alg.preconditioner_apply(cnd).noise.chol, axis1=-2, axis2=-1) for cnd in seq.conditional]
        own.append(np.diagonal(seq.marginal.chol, axis1=-2, axis2=-1))
        base = draws_for(a, np.stack(own))
        return dict(samples=stacked(post.sample(base)), samples_zero_draws=stacked(post.sample(0.0 * base)))
    slv = H._build(o_pdq, o_ivp, s, H.oracle_vf(s, params))[1]
    rvs = [slv.offgrid_marginals(float(t), solution=sol) for t in a["ts"]]
    return dict(mean=np.stack([rv.mean for rv in rvs]), cov=np.stack([cov(rv.chol) for rv in rvs]))


def test_aux_fixtures_cover_every_function_and_factorisation():
    names = set(IDS)
    assert {"dt0_lv", "dt0_hires", "dt0_pleiades", "dt0_adaptive_lv", "dt0_adaptive_hires"} <= names
    for prob in ("lv", "hires", "vanderpol", "burgers_d16"):
        assert {f"taylor_{prob}_jetexpand_ode_padded_scan", f"taylor_{prob}_jetexpand_ode_unroll"} <= names
    for fact in ("isotropic", "blockdiag", "dense"):
        assert {f"lml_terminal_{fact}", f"lml_timeseries_{fact}_fixedpoint", f"lml_timeseries_{fact}_fixedinterval",
                f"sample_{fact}_fixedpoint",
                f"offgrid_{fact}_filter_save_at", f"offgrid_{fact}_filter_dynamic_ts1",
                f"offgrid_{fact}_fixedinterval"} <= names  # fmt: skip
    # off-grid times include the last interval of a save_at solution (JAX's index clamping, solvers.py:185)
    c = next(c for c in CASES if c["name"] == "offgrid_dense_filter_save_at")
    assert c["arrays"]["ts"][-1] > c["base"]["grid"][-2]


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_oracle_reproduces_the_reference(c):
    check(c, oracle_outputs(c))
