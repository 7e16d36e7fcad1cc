"""The plain-C restatement (oracle/c, used as CPU baseline) against the NumPy oracle."""

import numpy as np

import pdeq_test_helpers as H
from oracle import c_port
from oracle import problems as o_problems


def test_c_port_matches_numpy_oracle_on_lotka_volterra():
    B = 6
    params, u0 = H.lv_ensemble(B, seed=5)
    tcoeffs = o_problems.taylor_coefficients_batched("lotka_volterra", params, (u0,), 0.0, 4)
    res = c_port.solve_lv_terminal(tcoeffs, params, t0=0.0, t1=10.0, atol=1e-8, rtol=1e-6, num_threads=2)
    s = H.spec()
    for b in range(B):
        osol, trace = H.oracle_solve_save_at(s, tcoeffs[b], params[b], np.asarray([0.0, 10.0]), 1e-8, 1e-6)
        assert res["num_steps"][b] == osol.num_steps[-1]
        assert res["num_attempts"][b] == len(trace)
        assert np.allclose(res["mean"][b], osol.u_mean[-1], rtol=1e-7, atol=1e-9)
        cov_c = res["chol"][b] @ res["chol"][b].T
        cov_o = osol.u_chol[-1] @ osol.u_chol[-1].T
        assert np.allclose(cov_c, cov_o, rtol=1e-5, atol=1e-6 * np.abs(cov_o).max())


def test_batched_taylor_coefficients_match_per_instance():
    params, u0 = H.lv_ensemble(5, seed=7)
    batched = o_problems.taylor_coefficients_batched("lotka_volterra", params, (u0,), 0.0, 4)
    for b in range(5):
        single = o_problems.Ode("lotka_volterra", params[b]).taylor_coefficients((u0[b],), 0.0, 4)
        assert np.allclose(batched[b], single, rtol=1e-14)


def test_c_port_reproduces_the_reference_on_the_headline_configuration():
    """The CPU baseline `bench.py` times (and its `--impl reference` arm) against the REFERENCE's own output for one
    instance of BASELINE configs[1] at its full horizon (tests/golden/reference_numpy_backend.npz, produced by running
    the unmodified reference on the NumPy backend of oracle/refshim): same accepted steps, terminal value to 1e-8."""
    from test_reference_golden import CASES, rel

    c = next(c for c in CASES if c["name"] == "lv_iso_ts0_terminal_t50")
    params = np.asarray(c["problem"]["params"])[None, :]
    res = c_port.solve_lv_terminal(c["tcoeffs"][None], params, t0=c["grid"][0], t1=c["grid"][1], atol=c["atol"],
                                   rtol=c["rtol"], dt0=c["dt0"], num_threads=1)  # fmt: skip
    assert int(res["num_steps"][0]) == int(c["ref"]["num_steps"])
    assert rel(res["mean"][0][0], c["ref"]["mean"][0]) < 1e-8
    sens = c["reference_one_ulp_sensitivity"]
    assert rel(res["mean"][0], c["ref"]["mean"]) < max(1e-6, 100 * sens["mean"])
    cov = res["chol"][0] @ res["chol"][0].T
    assert rel(cov, c["ref"]["cov"]) < max(1e-5, 100 * sens["cov"])
