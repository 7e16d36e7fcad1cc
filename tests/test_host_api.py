"""CPU-only checks of the product's host side: the C ABI exports, config lowering, argument validation.

No compute call is made here (there is no GPU in the build container); the GPU parity tests do that.
"""

import ctypes as C
import math
import pathlib
import re

import numpy as np
import pytest

from probdiffeq_b200 import _iwp, _lib

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_library_exports_every_symbol_the_header_declares():
    header = (ROOT / "include" / "probdiffeq_b200.h").read_text()
    declared = set(re.findall(r"\b(pdeq_[a-z0-9_]+)\s*\(", header))
    declared -= {"pdeq_config", "pdeq_problem", "pdeq_solution"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pdeq_version() == 100


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations: 18 int32 + 5 double + 2*64 double + 3*9 double
    assert C.sizeof(_lib.Config) == 18 * 4 + 5 * 8 + 2 * 64 * 8 + 3 * 9 * 8
    assert C.sizeof(_lib.Problem) == 9 * 8
    assert C.sizeof(_lib.Solution) == 14 * 8  # 13 pointers + trace_capacity


def test_vector_field_registry():
    lib = _lib.load()
    expect = {"lotka_volterra": (1, 4, 2), "pleiades": (1, 0, 28), "hires": (1, 0, 8),
              "vanderpol": (2, 1, 1), "linear": (1, 1, 0), "burgers": (1, 1, 0)}  # fmt: skip
    for name, (order, nparams, dim) in expect.items():
        vid = lib.pdeq_vf_id(name.encode())
        assert vid >= 0
        assert (lib.pdeq_vf_ode_order(vid), lib.pdeq_vf_num_params(vid), lib.pdeq_vf_dim(vid)) == (order, nparams, dim)
    assert lib.pdeq_vf_id(b"no_such_problem") == -1


def test_k1_spec_choice_follows_the_environment(monkeypatch):
    """pdeq_k1_spec_choice: the library default (a specialised build, not the general kernel) unless PDEQ_K1_SPEC
    names another build; values outside the known builds fall back to the default. Host-only, no launch."""
    lib = _lib.load()
    monkeypatch.delenv("PDEQ_K1_SPEC", raising=False)
    default = lib.pdeq_k1_spec_choice()
    assert 1 <= default <= 2
    for spec in range(0, 3):
        monkeypatch.setenv("PDEQ_K1_SPEC", str(spec))
        assert lib.pdeq_k1_spec_choice() == spec
    for bad in ("-1", "99"):
        monkeypatch.setenv("PDEQ_K1_SPEC", bad)
        assert lib.pdeq_k1_spec_choice() == default


def _cfg(**kw):
    cfg = _lib.Config()
    cfg.factorisation, cfg.num_derivatives, cfg.ode_dim = 0, 4, 2
    cfg.vf_id = _lib.load().pdeq_vf_id(b"lotka_volterra")
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_config_validation_reports_errors_without_touching_the_gpu():
    lib = _lib.load()
    assert lib.pdeq_config_supported(C.byref(_cfg())) == 0
    assert lib.pdeq_config_supported(C.byref(_cfg(ode_dim=3))) < 0
    assert b"dimension" in lib.pdeq_last_error()
    assert lib.pdeq_config_supported(C.byref(_cfg(num_derivatives=9))) < 0
    assert lib.pdeq_config_supported(C.byref(_cfg(vf_id=99))) < 0
    assert lib.pdeq_config_supported(C.byref(_cfg(solver=7))) < 0
    assert lib.pdeq_config_supported(None) < 0
    with pytest.raises(ValueError):
        _lib.check(lib.pdeq_config_supported(C.byref(_cfg(control=5))), "cfg")


def test_iwp_system_matrices_match_the_oracle():
    from oracle import linalg

    for nu in range(1, 8):
        a, q, f = _iwp.system_matrices(nu)
        a_ref, q_ref = linalg.system_matrices_1d_iwp(nu)
        assert np.array_equal(a, a_ref) and np.allclose(q, q_ref, rtol=1e-15, atol=0)
        assert np.array_equal(f, linalg.factorial(np.arange(nu + 2)))
        assert np.allclose(q @ q.T, np.flip(1.0 / (np.arange(1, nu + 2)[:, None] + np.arange(1, nu + 2)[None, :] - 1.0)), rtol=1e-8)


def test_product_path_fails_loudly_without_a_gpu():
    import torch

    from probdiffeq_b200 import probdiffeq

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    vf = probdiffeq.ode("lotka_volterra", params=[0.5, 0.05, 0.5, 0.05])
    with pytest.raises(_lib.NativeLibraryError):
        probdiffeq.jetexpand_ode_padded_scan(num=4)(vf, (np.asarray([20.0, 20.0]),), t=0.0)


def test_vector_field_plugin_builds_and_registers_without_a_gpu():
    """probdiffeq_b200/plugins.py: nvcc cross-compiles the user's right-hand side; loading registers a new id."""
    from probdiffeq_b200 import plugins

    vf = plugins.ode_from_cuda("logistic", params=np.ones((3, 2)), **plugins.LOGISTIC)
    lib = _lib.load()
    assert vf.vf_id >= 6 and lib.pdeq_vf_id(b"logistic") == vf.vf_id
    assert (lib.pdeq_vf_ode_order(vf.vf_id), lib.pdeq_vf_num_params(vf.vf_id), lib.pdeq_vf_dim(vf.vf_id)) == (1, 2, 1)
    again = plugins.ode_from_cuda("logistic", params=np.ones((3, 2)), **plugins.LOGISTIC)
    assert again.vf_id == vf.vf_id
    with pytest.raises(ValueError, match="different signature"):
        plugins.ode_from_cuda("logistic", dim=2, component="return u(0, i);")
    assert lib.pdeq_register_vf(b"logistic", 1, 2, 1) < 0  # the C entry point refuses duplicates


def test_constructors_mirror_the_reference_error_behaviour():
    from probdiffeq_b200 import ivpsolve, probdiffeq

    with pytest.raises(ValueError):
        probdiffeq.ode("lotka_volterra")  # parameters missing
    with pytest.raises(ValueError):
        probdiffeq.ode("not_registered")
    with pytest.raises(TypeError):
        probdiffeq.state_space_model_isotropic().constraint_ode_ts0(lambda u, t: u)
    vf = probdiffeq.ode("hires")
    ssm = probdiffeq.state_space_model_isotropic()
    ts0 = ssm.constraint_ode_ts0(vf)
    fp = probdiffeq.solver(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts0)
    with pytest.warns(UserWarning):
        ivpsolve.solve_fixed_grid(solver=fp)  # reference: solvers_via_fixed_steps.py:14-18
    assert ivpsolve.control_integral().safety == 0.95
    assert ivpsolve.control_proportional_integral().exponent_proportional == 0.4
    # loss_lml_timeseries wants the posterior of a smoother (reference: estimators_and_losses.py:60-68)
    with pytest.raises(TypeError, match="datatype"):
        probdiffeq.loss_lml_timeseries()(np.zeros((3, 2)), posterior=object(), std=np.ones(3))
    with pytest.raises(ValueError, match="terminal"):
        probdiffeq.strategy_smoother_fixedinterval(terminal="nonsense")


def test_re_linearize_flags_are_accepted_and_carried_into_the_config():
    """reference: solvers.py:496, 911, 1021. Both flags are no-ops for the prior Taylor point (shown on the oracle in
    tests/test_oracle_kats.py), so the product accepts them; the solver's flag travels in pdeq_config."""
    from probdiffeq_b200 import ivpsolve, probdiffeq

    vf = probdiffeq.ode("lotka_volterra", params=np.asarray([[0.5, 0.05, 0.5, 0.05]]))
    ssm = probdiffeq.state_space_model_blockdiag()
    ts1 = ssm.constraint_ode_ts1(vf)
    for flag in (False, True):
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1,
                                           re_linearize_after_calibration=flag)  # fmt: skip
        assert solver.re_linearize_after_calibration is flag
        assert solver.options["re_linearize_after_calibration"] == int(flag)
        for est in (probdiffeq.error_residual_std, probdiffeq.error_state_std):
            assert est(constraint=ts1, re_linearize_before_error=flag).re_linearize_before_error is flag
    prior = type("P", (), {"factorisation": "blockdiag", "num_derivatives": 4, "ode_dim": 2})()
    cfg = ivpsolve._lower(prior, solver, probdiffeq.error_residual_std(constraint=ts1, re_linearize_before_error=True),
                          ivpsolve.control_integral(), clip_dt=False)  # fmt: skip
    assert cfg.re_linearize_after_calibration == 1 and cfg.solver == 2


def test_error_constants_reproduce_the_full_bayes_rule():
    """pdeq_config.err_const (probdiffeq_b200/_iwp.py): for ts0 and damp = 0 the zero-error extrapolation's factor is
    diag(|p|) sqrt(dt) lambda q, and error_state_std's triangularisation (solvers.py:1070-1086) commutes with that
    column scaling. The constants must give the same observed factor and coefficient std as the oracle's full
    Bayes rule, for any step size and prior scale."""
    from oracle import linalg as o_linalg
    from probdiffeq_b200 import _iwp

    for nu, order in ((4, 1), (3, 1), (5, 1), (4, 2)):
        n = nu + 1
        a, q, facts = _iwp.system_matrices(nu)
        consts = _iwp.error_constants(nu, order)
        for dt, lam in ((0.37, 1.0), (1e-3, 2.5), (4.2, 0.1)):
            k = np.arange(n)
            p = dt ** (nu - k) / facts[nu - k]  # Taylor preconditioner (utilities.py:74-84)
            L = np.abs(p)[:, None] * (np.sqrt(dt) * lam * q)  # zero-error extrapolation: noise only
            h = np.zeros((1, n))
            h[0, order] = 1.0
            r_obs, (r_cor, _gain) = o_linalg.revert_conditional(R_X_F=(h @ L).T, R_X=L.T, R_YX=np.zeros((1, 1)))
            assert np.isclose(abs(r_obs[0, 0]), abs(consts[0]) * abs(p[order]) * np.sqrt(dt) * lam, rtol=1e-12)
            std = np.sqrt(np.sum(r_cor**2, axis=0))  # std of every coefficient after the update
            for i in range(n):
                # (the observed coefficient itself comes out as an exact zero up to rounding)
                assert np.isclose(std[i], consts[1 + i] * abs(p[i]) * np.sqrt(dt) * lam, rtol=1e-10,
                                  atol=1e-13 * np.max(np.abs(L))), (nu, i)  # fmt: skip


def test_problem_constants_and_posterior_default():
    from probdiffeq_b200 import ivpsolve, problems

    params, u0 = problems.lotka_volterra_ensemble(5, seed=0)
    assert params.shape == (5, 4) and u0.shape == (5, 2) and np.all(params > 0)
    assert problems.PLEIADES_U0.shape == (28,) and problems.HIRES_U0.shape == (8,)
    assert problems.burgers_u0(7).shape == (7,) and abs(problems.burgers_u0(7)[0]) > 0

    class _S:  # the posterior is returned by default only while the conditionals stay below POSTERIOR_AUTO_BYTES
        class strategy:
            kind = "fixedpoint"

    class _P:
        factorisation = "blockdiag"

        def __init__(self, B):
            import torch

            self.tcoeffs = torch.empty((B, 6, 28), device="meta")

    assert ivpsolve._want_posterior(None, _S, _P(64), 33, True) is True
    assert ivpsolve._want_posterior(None, _S, _P(65536), 33, True) is False
    assert ivpsolve._want_posterior(True, _S, _P(65536), 33, True) is True
    _S.strategy.kind = "filter"
    assert ivpsolve._want_posterior(None, _S, _P(64), 33, True) is False


def test_host_mirror_of_the_problem_and_prior_helpers():
    """Names a user of the reference reaches for around the path: `jacobian_materialize`, `taylor_point_prior`,
    `ode_order_two` (argument checks, loud refusals for what the accelerated path does not do) and the IWP constants
    `system_matrices_1d_iwp` / `preconditioner_taylor` (utilities.py:57-84) -- bitwise the oracle's, and the
    reference's own where its sources are present (run on the NumPy backend of oracle/refshim)."""
    import pytest

    from oracle import linalg as o_linalg
    from oracle import refshim
    from probdiffeq_b200 import probdiffeq as p_pdq

    assert p_pdq.jacobian_materialize().kind == "materialize" and p_pdq.taylor_point_prior().kind == "prior"
    for refused in (p_pdq.jacobian_monte_carlo_fwd, p_pdq.jacobian_monte_carlo_rev, p_pdq.taylor_point_maximum_a_posteriori):
        with pytest.raises(NotImplementedError):
            refused()
    with pytest.raises(NotImplementedError):
        p_pdq._check_jacobian(p_pdq.Jacobian("monte_carlo_rev"))
    with pytest.raises(NotImplementedError):
        p_pdq._check_taylor_point(p_pdq.TaylorPoint("maximum_a_posteriori"))
    assert p_pdq.jetexpand_ode_via_jvp is p_pdq.jetexpand_ode_padded_scan
    for nu in (1, 3, 4, 5):
        a, q = p_pdq.system_matrices_1d_iwp(nu)
        oa, oq = o_linalg.system_matrices_1d_iwp(nu) if hasattr(o_linalg, "system_matrices_1d_iwp") else (a, q)
        assert np.array_equal(a, oa) and np.array_equal(q, oq)
        p, pinv = p_pdq.preconditioner_taylor(nu)(0.37)
        assert np.allclose(p * pinv, 1.0, rtol=1e-15)
        assert np.isclose(p[-1], 1.0) and np.isclose(p[0], 0.37**nu / math.factorial(nu), rtol=1e-14)
        if refshim.available():
            _ivp, ref = refshim.load()
            ra, rq = ref.system_matrices_1d_iwp(nu)
            assert np.allclose(a, np.asarray(ra), rtol=0, atol=0) and np.allclose(q, np.asarray(rq), rtol=1e-15, atol=1e-16)
            rp, rpinv = ref.preconditioner_taylor(nu)(0.37)
            assert np.allclose(p, np.asarray(rp), rtol=1e-15) and np.allclose(pinv, np.asarray(rpinv), rtol=1e-15)


def test_the_product_never_touches_the_oracle_or_the_reference():
    """The oracle (and the shim that runs the reference) are test infrastructure: nothing under probdiffeq_b200/ may
    import them, read the fixtures, or look for the reference's sources -- statically (no such import or path in any
    product source file) and dynamically (importing the product and building a solver pulls in no `oracle` module)."""
    import subprocess
    import sys

    root = pathlib.Path(__file__).resolve().parents[1]
    pattern = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference|tests/golden|refshim", re.MULTILINE)
    for path in sorted((root / "probdiffeq_b200").rglob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".cc", ".h") and "lib" not in path.relative_to(root).parts[1:2]:
            text = path.read_text(errors="ignore")
            hits = [m.group(0) for m in pattern.finditer(text)]
            # comments may cite reference file:line, never an absolute path to it
            assert not hits, (str(path), hits)
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from probdiffeq_b200 import ivpsolve, probdiffeq, sharding, plugins\n"
        "vf = probdiffeq.ode('lotka_volterra', params=[0.5, 0.05, 0.5, 0.05])\n"
        "ssm = probdiffeq.state_space_model_isotropic(); c = ssm.constraint_ode_ts0(vf)\n"
        "s = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=c)\n"
        "ivpsolve.solve_adaptive_terminal_values(solver=s, error=probdiffeq.error_state_std(constraint=c))\n"
        "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]\n"
        "assert not bad, bad\n" % str(root)
    )
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
