"""Generate tests/golden/reference_numpy_backend_aux.npz: outputs of the REFERENCE's own code for the functions either
side of the step loop -- Taylor-mode initialisation (`jetexpand_ode_padded_scan` / `jetexpand_ode_unroll`,
jet_expansion_algorithms.py:49-152), `ivpsolve.dt0` / `dt0_adaptive` (stepsize_initialisers.py:7-78), `loss_lml_terminal_values` and
`loss_lml_timeseries` (estimators_and_losses.py:20-105), `MarkovSequence.sample` (estimators_and_losses.py:233-271, given
the draws), `solver.offgrid_marginals` (solvers.py:149-203).

Same mechanism as make_reference_golden.py (which holds the step-loop cases): the reference's unmodified modules run
from /root/reference on the NumPy array backend of oracle/refshim; the oracle is compared on the spot and the
reference's outputs are written as fixtures for the CPU (oracle) and GPU (CUDA path) tests:

    python tests/golden/make_reference_golden_aux.py       # needs /root/reference; a few seconds

Each case is a step-loop case in make_reference_golden.py's format (`base`) plus the inputs of the function applied to
its solution.  `offgrid_marginals` relies on JAX clamping out-of-range integer indices (solvers.py:173-185: for a time
in the LAST interval of a save_at solution it reads `output_scale[T - 1]` of T - 1 entries, and it indexes solution
leaves that have no time axis); the shim's array type reproduces that clamping (oracle/refshim/backend/_array.py), the
oracle restates it (oracle/probdiffeq.py:419-422), and the off-grid times below include the last interval.
"""

import importlib.util
import json
import pathlib
import sys
import warnings

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import pdeq_test_helpers as H  # noqa: E402
from oracle import ivpsolve as o_ivp  # noqa: E402
from oracle import probdiffeq as o_pdq  # noqa: E402
from oracle import problems as o_problems  # noqa: E402
from oracle import refshim  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_reference_golden", HERE / "make_reference_golden.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)

OUT = HERE / "reference_numpy_backend_aux.npz"
LV = mk.LV
HIRES = dict(vf="hires", nu=5, params=[], u0=list(o_problems.hires_u0()))
PLEIADES = dict(vf="pleiades", nu=5, params=[], u0=list(o_problems.pleiades_u0()))
CASES = []


def base(kind, grid, atol=None, rtol=None, dt0=0.1, problem=LV, **spec):
    return dict(name="", kind=kind, grid=list(map(float, grid)), atol=atol, rtol=rtol, dt0=dt0, problem=problem,
                diffuse_start=False, spec=H.spec(vf=problem["vf"], **spec))  # fmt: skip


def aux(name, kind, b, **extra):
    CASES.append(dict(name=name, aux=kind, base=b, **extra))


for prob_name, prob in (("lv", LV), ("hires", HIRES), ("pleiades", PLEIADES)):
    aux(f"dt0_{prob_name}", "dt0", base("terminal", [0.0, 1.0], problem=prob))
aux("dt0_adaptive_lv", "dt0_adaptive", base("terminal", [0.0, 1.0]), error_contraction_rate=5, rtol=1e-6, atol=1e-8)
aux("dt0_adaptive_hires", "dt0_adaptive", base("terminal", [0.0, 1.0], problem=HIRES), error_contraction_rate=6,
    rtol=1e-8, atol=1e-11)  # fmt: skip
# Taylor-mode initialisation (jet_expansion_algorithms.py:49-152): the reference's recursion on the shim's exact-rational
# `jet` (polynomial right-hand sides; the second-order problem starts with a non-zero velocity)
VDP = dict(vf="vanderpol", nu=4, params=[1e3], u0=[2.0])
BURGERS = dict(vf="burgers", nu=3, params=[0.01], u0=list(o_problems.burgers_u0(16)))
for prob_name, prob, num, du0 in (("lv", LV, 4, None), ("hires", HIRES, 5, None), ("vanderpol", VDP, 3, [0.5]),
                                  ("burgers_d16", BURGERS, 3, None)):  # fmt: skip
    for alg in ("jetexpand_ode_padded_scan", "jetexpand_ode_unroll"):
        aux(f"taylor_{prob_name}_{alg}", "taylor", base("terminal", [0.0, 1.0], problem=prob), alg=alg, num=num,
            du0=du0, t=0.25)  # fmt: skip
# BASELINE configs[4] (the variant bench.py runs: Burgers, block-diagonal ts0, solver + state error + PI, t in [0, 1],
# rtol 1e-4, atol 1e-7, then the log-marginal-likelihood of terminal-value data) at d = 64, full horizon (935 steps)
aux("lml_terminal_burgers_d64_config5_full_horizon", "lml_terminal",
    base("terminal", [0.0, 1.0], 1e-7, 1e-4, dt0=1.7584139942631142e-3, fact="blockdiag",
         problem=dict(vf="burgers", nu=3, params=[0.01], u0=list(o_problems.burgers_u0(64)))))  # fmt: skip
for fact in ("isotropic", "blockdiag", "dense"):
    aux(f"lml_terminal_{fact}", "lml_terminal",
        base("terminal", [0.0, 2.0], 1e-6, 1e-4, fact=fact, solver="solver_mle"))  # fmt: skip
    aux(f"lml_timeseries_{fact}_fixedpoint", "lml_timeseries",
        base("save_at", np.linspace(0.0, 4.0, 13), 1e-4, 1e-4, fact=fact, strategy="fixedpoint", solver="solver_mle",
             error="residual_std", control="i", clip_dt=False))  # fmt: skip
    aux(f"lml_timeseries_{fact}_fixedinterval", "lml_timeseries",
        base("fixed", np.linspace(0.0, 1.0, 17), fact=fact, strategy="fixedinterval", solver="solver_mle"))
    aux(f"sample_{fact}_fixedpoint", "sample",
        base("save_at", np.linspace(0.0, 3.0, 10), 1e-3, 1e-3, fact=fact, strategy="fixedpoint", solver="solver_mle",
             error="residual_std", control="i", clip_dt=False))  # fmt: skip
    aux(f"offgrid_{fact}_filter_save_at", "offgrid",
        base("save_at", np.linspace(0.0, 3.0, 9), 1e-5, 1e-4, fact=fact, solver="solver_mle", error="residual_std",
             control="i", clip_dt=False), ts=[0.01, 0.1875, 1.2345, 2.99])  # fmt: skip
    aux(f"offgrid_{fact}_filter_dynamic_ts1", "offgrid",
        base("save_at", np.linspace(0.0, 3.0, 9), 1e-5, 1e-4, fact=fact, solver="solver_dynamic", constraint="ts1",
             error="residual_std", control="i", clip_dt=False), ts=[0.01, 0.1875, 1.2345, 2.99])  # fmt: skip
    aux(f"offgrid_{fact}_fixedinterval", "offgrid",
        base("fixed", np.linspace(0.0, 1.0, 17), fact=fact, strategy="fixedinterval", solver="solver_mle"),
        ts=[0.01, 0.33, 0.46875, 0.97])  # fmt: skip


def observations(c, mean_coeff, T=None):
    """Deterministic 'data' around a solution's Taylor coefficient, and observation noise levels in the shape the model
    expects: a scalar per time for the isotropic model, one entry per dimension otherwise."""
    rng = np.random.Generator(np.random.PCG64(len(c["name"])))
    mean_coeff = np.asarray(mean_coeff)
    data = mean_coeff + 0.05 * rng.normal(size=mean_coeff.shape)
    d = mean_coeff.shape[-1]
    iso = c["base"]["spec"]["fact"] == "isotropic"
    if T is None:
        std = np.asarray(0.07) if iso else 0.07 * (1.0 + 0.5 * np.arange(d))
    else:
        sd = 0.05 + 0.01 * np.arange(T)
        std = sd if iso else np.stack([sd * (1.0 + 0.5 * j) for j in range(d)], axis=1)
    return data, std


def solution_coefficient(sol_mean_flat, fact, n, d, i):
    m = np.asarray(sol_mean_flat)
    if fact == "isotropic":
        return m[..., i, :]
    if fact == "blockdiag":
        return m[..., :, i]
    return m.reshape(*m.shape[:-1], n, d)[..., i, :]


def oracle_solver(b):
    prob = b["problem"]
    params = np.asarray(prob["params"]) if prob["params"] else None
    return H._build(o_pdq, o_ivp, b["spec"], H.oracle_vf(b["spec"], params))[1]


def rel(a, b):
    return mk.rel(a, b)


def run(c):
    """Returns (fixture arrays, report row). Fails if the oracle disagrees with the reference."""
    ivp, pdq = refshim.load()
    b = c["base"]
    prob, s = b["problem"], b["spec"]
    n, d = prob["nu"] + 1, len(prob["u0"])
    params = np.asarray(prob["params"]) if prob["params"] else None
    ovf = o_pdq.ode(prob["vf"], params)
    u0 = np.asarray(prob["u0"])
    out, row = {}, dict(case=c["name"])
    if c["aux"] in ("dt0", "dt0_adaptive"):
        vf_r = mk.reference_vf(pdq, prob)
        if c["aux"] == "dt0":
            ref, ora = ivp.dt0(vf_r, (u0,), t=0.0), o_ivp.dt0(ovf, (u0,), t=0.0)
        else:
            kw = {k: c[k] for k in ("error_contraction_rate", "rtol", "atol")}
            ref, ora = ivp.dt0_adaptive(vf_r, (u0,), 0.0, **kw), o_ivp.dt0_adaptive(ovf, (u0,), 0.0, **kw)
        out["value"] = np.asarray(float(ref))
        row["rel"] = rel(ora, ref)
        assert row["rel"] < 1e-14, row
        return out, row
    if c["aux"] == "taylor":
        vf_r = mk.reference_vf(pdq, prob)
        inits = (u0,) if c["du0"] is None else (u0, np.asarray(c["du0"]))
        ref, _ = getattr(pdq, c["alg"])(num=c["num"])(vf_r, inits, t=c["t"])
        ref = np.stack([np.asarray(x, dtype=np.float64) for x in ref])
        ora = np.asarray(ovf.taylor_coefficients(list(inits), c["t"], c["num"]))
        out["tcoeffs"] = ref
        row["rel_per_coefficient"] = [rel(ora[i], ref[i]) for i in range(ref.shape[0])]
        assert max(row["rel_per_coefficient"]) < 1e-14, row
        return out, row
    b["tcoeffs"] = np.asarray(ovf.taylor_coefficients([u0], b["grid"][0], prob["nu"]))
    out["tcoeffs"] = b["tcoeffs"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rsol = mk.run_reference(b)
        osol = mk.run_oracle(b)
    assert np.array_equal(np.asarray(rsol.num_steps), np.asarray(osol.num_steps)), c["name"]
    out["num_steps"] = np.asarray(rsol.num_steps)
    if c["aux"] == "lml_terminal":
        for idx in (0, 1):
            data, std = observations(c, solution_coefficient(rsol.u.mean_flat, s["fact"], n, d, idx))
            ref = pdq.loss_lml_terminal_values(tcoeff_index=idx)(data, marginals=rsol.u, std=std)
            ora = o_pdq.loss_lml_terminal_values(tcoeff_index=idx)(data, marginals=osol.u, std=std)
            out[f"data{idx}"], out[f"std{idx}"], out[f"lml{idx}"] = data, np.asarray(std), np.asarray(float(ref))
            row[f"rel{idx}"] = rel(ora, ref)
            assert row[f"rel{idx}"] < 1e-8, row
    elif c["aux"] == "lml_timeseries":
        T = len(b["grid"])
        data, std = observations(c, solution_coefficient(rsol.u.mean_flat, s["fact"], n, d, 0), T=T)
        out["data"], out["std"] = data, np.asarray(std)
        opost = osol.solution_full.posterior.remove_filtering_distributions()
        for key, avg in (("lml_avg", True), ("lml_sum", False)):
            ref = pdq.loss_lml_timeseries(average_pdfs=avg)(data, posterior=rsol.solution_full.posterior, std=std)
            ora = o_pdq.loss_lml_timeseries(average_pdfs=avg)(data, posterior=opost, std=std)
            out[key] = np.asarray(float(ref))
            row["rel_" + key] = rel(ora, ref)
            assert row["rel_" + key] < 1e-8, row
    elif c["aux"] == "sample":
        # MarkovSequence.sample (estimators_and_losses.py:233-271) with the shim's generator; the standard-normal
        # numbers it draws are recorded in call order (terminal marginal first, then the conditionals backwards) and
        # stored per grid point, so that the oracle and the CUDA path can be given the same draws
        from oracle.refshim.backend import random as shim_random

        draws, original = [], shim_random.normal

        def recording(key, /, shape, dtype=None):
            x = original(key, shape, dtype)
            draws.append(np.asarray(x, dtype=np.float64))
            return x

        shim_random.normal = recording
        try:
            smp = rsol.solution_full.posterior.sample(shim_random.prng_key(seed=7))
        finally:
            shim_random.normal = original
        # ... and once more with every draw set to zero: the sample is then the chain of conditional means, which does
        # not depend on the sign convention of the factors
        shim_random.normal = lambda key, /, shape, dtype=None: np.zeros(shape)
        try:
            smp0 = rsol.solution_full.posterior.sample(shim_random.prng_key(seed=7))
        finally:
            shim_random.normal = original
        T = len(b["grid"])
        assert len(draws) == T
        base_draws = np.stack(draws[::-1])  # draws[0] belongs to the last grid point
        ref = np.stack([np.asarray(x, dtype=np.float64) for x in smp], axis=1)  # (T, n, d)
        states = osol.solution_full.posterior.sample(base_draws)
        if s["fact"] == "isotropic":
            ora = np.stack(states)
        elif s["fact"] == "blockdiag":
            ora = np.stack([x.T for x in states])
        else:
            ora = np.stack([x.reshape(n, d) for x in states])
        out["base"], out["samples"] = base_draws, ref
        # signs of the diagonals of the factors the reference drew with, per grid point: an implementation whose
        # factors carry other column signs multiplies the draws by the ratio and must then reproduce the sample
        from oracle.refshim.backend import tree as shim_tree

        rpost = rsol.solution_full.posterior.remove_filtering_distributions()
        signs = np.ones_like(base_draws)
        signs[T - 1] = np.sign(np.diagonal(np.asarray(rpost.marginal.cholesky_flat), axis1=-2, axis2=-1))
        for k in range(T - 2, -1, -1):
            cond_k = shim_tree.tree_map(lambda x, k=k: x[k], rpost.conditional)
            drawn_from = cond_k.apply_flat(rpost.marginal.mean_flat)
            signs[k] = np.sign(np.diagonal(np.asarray(drawn_from.cholesky_flat), axis1=-2, axis2=-1))
        out["factor_diag_sign"] = np.where(signs == 0.0, 1.0, signs)
        out["samples_zero_draws"] = np.stack([np.asarray(x, dtype=np.float64) for x in smp0], axis=1)
        states0 = osol.solution_full.posterior.sample(np.zeros_like(base_draws))
        ora0 = np.stack([x if s["fact"] == "isotropic" else (x.T if s["fact"] == "blockdiag" else x.reshape(n, d))
                         for x in states0])  # fmt: skip
        row["rel_zero_draws"] = rel(ora0, out["samples_zero_draws"])
        assert row["rel_zero_draws"] < 1e-8, row
        row["rel"] = rel(ora, ref)
        row["draw_shape"] = list(base_draws.shape)
        # A sample is mean + factor @ draws, and the factor of a covariance is unique only up to the signs of its
        # columns: the reference takes whatever signs LAPACK's QR returns (it never normalises them), so the sample
        # for GIVEN draws is pinned only where those signs happen to agree with the oracle's Householder convention.
        # Where they do not (the conditional means and covariances still agree, see the lml / offgrid cases), the
        # fixture is kept as a record and marked; the tests then skip the value comparison.
        c["factor_signs_agree"] = bool(row["rel"] < 1e-6)
        row["factor_signs_agree"] = c["factor_signs_agree"]
    elif c["aux"] == "offgrid":
        _ssm, rslv, _e, _c = H._build(pdq, ivp, s, mk.reference_vf(pdq, prob))
        oslv = oracle_solver(b)
        means, covs, worst = [], [], 0.0
        for t in c["ts"]:
            rv = rslv.offgrid_marginals(np.asarray(t), solution=rsol)
            orv = oslv.offgrid_marginals(t, solution=osol)
            means.append(np.asarray(rv.mean_flat))
            covs.append(mk._cov(rv.cholesky_flat))
            worst = max(worst, rel(orv.mean, means[-1]), rel(mk._cov(orv.chol), covs[-1]))
        out["ts"], out["mean"], out["cov"] = np.asarray(c["ts"]), np.stack(means), np.stack(covs)
        row["rel"] = worst
        assert worst < 1e-6, row
    else:
        raise ValueError(c["aux"])
    return out, row


def main():
    if not refshim.available():
        raise SystemExit("needs the reference sources under /root/reference")
    out, report = {}, []
    for c in CASES:
        arrays, row = run(c)
        report.append(row)
        print(json.dumps(row), flush=True)
        for k, v in arrays.items():
            out[f"{c['name']}/{k}"] = v
        meta = {k: v for k, v in c.items() if k != "base"}
        meta["base"] = {k: v for k, v in c["base"].items() if k != "tcoeffs"}
        out[f"{c['name']}/meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(OUT, **out)
    OUT.with_suffix(".report.json").write_text(json.dumps(report, indent=1))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
