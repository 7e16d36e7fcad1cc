"""BASELINE configs 3 and 4a at their full ensemble sizes, through size-independent properties: no failed instance,
sub-ensemble / permutation bitwise equality (instances are independent), monotone checkpoints, and one instance
against the oracle (accepted-step sequence; values within the conditioning measured in the small-size tests).
Config 2 is covered in test_gpu_edge_cases.py, the Burgers configuration in test_gpu_group_and_smoother.py."""

import sys

import numpy as np
import pytest

import pdeq_test_helpers as H

sys.path.insert(0, "scripts")

pytestmark = pytest.mark.gpu


def test_config3_pleiades_fixedpoint_full_size(cuda):
    import torch

    import bench_configs as bc

    B = 65536
    run = bc.config3(B)
    sol, _ = run()
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    steps = sol.num_steps.cpu().numpy()
    assert np.all(np.diff(steps, axis=1) >= 1) and steps[:, 0].max() == 0
    assert np.allclose(sol.t.cpu().numpy(), np.linspace(0.0, 3.0, 33)[None, :], rtol=0, atol=2e-8)  # eps = 1e-8
    assert bool(torch.isfinite(sol.u.mean_flat).all()) and bool(torch.isfinite(sol.u.cholesky_flat).all())
    # the same instances solved as a small ensemble give bitwise the same result
    rng = np.random.Generator(np.random.PCG64(1))
    from oracle import problems as o_problems
    from probdiffeq_b200 import ivpsolve, probdiffeq

    u0 = o_problems.pleiades_u0()[None, :] + 1e-3 * rng.normal(size=(B, 28))
    idx = np.asarray([0, 777, 40000, B - 1])
    vf = probdiffeq.ode("pleiades")
    ssm = probdiffeq.state_space_model_blockdiag()
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0[idx],), t=0.0)
    ts0 = ssm.constraint_ode_ts0(vf)
    solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts0)
    error = probdiffeq.error_residual_std(constraint=ts0)
    solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, control=ivpsolve.control_integral())
    dt0 = ivpsolve.dt0(vf, (u0[idx],), t=0.0)
    save_at = np.linspace(0.0, 3.0, 33)
    sub = solve(ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-9, rtol=1e-6, dt0=dt0)
    assert np.array_equal(sub.num_steps.cpu().numpy(), steps[idx])
    assert torch.equal(sub.u.mean_flat, sol.u.mean_flat[torch.as_tensor(idx, device="cuda")])
    # one instance against the oracle: the accepted-step counts at every checkpoint
    s = H.spec(vf="pleiades", fact="blockdiag", strategy="fixedpoint", solver="solver_dynamic", error="residual_std",
               control="i", clip_dt=False)  # fmt: skip
    osol, _ = H.oracle_solve_save_at(s, tcoeffs[0].cpu().numpy(), None, save_at, 1e-9, 1e-6, dt0=float(dt0[0]))
    # Identical accepted-step sequence up to the first close encounter of the bodies (t ~ 1.4); from there on one
    # accept/reject decision within 1e-10 of the threshold falls the other way and the two step sequences part for
    # good (the oracle does the same to itself under a 1-ulp change of dt0). Both remain
    # solutions of the same IVP at rtol 1e-6.
    osteps = np.asarray(osol.num_steps)
    assert np.array_equal(steps[0, 1:14], osteps[:13])
    # How far the final counts may part: over the 256 instances of profiles/parity_r2.json the kernel's and the
    # oracle's attempt counts of this config differ by 2.7 % (std), 8.5 % (max) -- the oracle against itself under a
    # 1-ulp change of dt0 parts even earlier (median attempt 372 against 427). One instance is held to 12 %.
    assert abs(int(steps[0, -1]) - int(osteps[-1])) <= 0.12 * osteps[-1]
    got, ref = sub.u.mean_flat[0, :, 0].cpu().numpy(), np.asarray(osol.u_mean)[:, 0]  # (33, 28) positions/velocities
    assert np.max(np.abs(got[:14] - ref[:14])) < 1e-8
    assert np.max(np.abs(got - ref)) < 1e-3


def test_config4a_hires_dense_full_size(cuda):
    import torch

    import bench_configs as bc

    B = 16384
    sol, _ = bc.config4a(B)()
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0
    steps = sol.num_steps.cpu().numpy()
    assert steps.min() > 500 and steps.max() < 5000
    assert np.all(np.abs(sol.t.cpu().numpy() - 321.8122) < 1e-9)
    mean = sol.u.mean[0].cpu().numpy()  # the eight concentrations at t1
    assert np.all(np.isfinite(mean)) and np.all(mean > -1e-6) and np.all(mean < 1.2)
    # a sub-ensemble reproduces its rows bitwise
    sub, _ = bc.config4a(64)()
    assert np.array_equal(sub.num_steps.cpu().numpy(), steps[:64])
    assert np.array_equal(sub.u.mean_flat.cpu().numpy(), sol.u.mean_flat[:64].cpu().numpy())


@pytest.mark.parametrize("d,stride", [(64, 256), (128, 1024)], ids=["d64-16-instances", "d128-4-instances"])
def test_config5_variant_burgers_full_horizon_parity_and_lml(cuda, d, stride):
    """BASELINE config 5 as bench.py runs it (d = 128, the largest power of two for which the
    oracle completes the seed-3 ensemble; the specified d = 1024 diverges in the restated algorithm itself):
    blockdiag ts0 filter, solver + error_state_std + PI, t in [0, 1], rtol 1e-4, atol 1e-7. Instances spread over the
    4096-instance ensemble (sixteen at d = 64, four at d = 128, where one oracle solve takes ~10 s) against the oracle
    over the FULL horizon: accepted / attempted counts (wherever the oracle reproduces its own under a 1-ulp change of
    dt0), terminal ODE solution, and the per-instance log-marginal-likelihood of noisy observations of the
    viscosity-0.01 solution (loss_lml_terminal_values)."""
    import torch

    from oracle import ivpsolve as o_ivp
    from oracle import probdiffeq as o_pdq
    from oracle import problems as o_problems
    from probdiffeq_b200 import ivpsolve, probdiffeq

    B_all = 4096
    visc_all = 0.01 * np.random.Generator(np.random.PCG64(3)).uniform(0.5, 2.0, size=(B_all, 1))
    visc = visc_all[::stride]
    B = visc.shape[0]
    assert B == B_all // stride
    u0 = np.repeat(o_problems.burgers_u0(d)[None, :], B, axis=0)
    vf = probdiffeq.ode("burgers", params=visc)
    ssm = probdiffeq.state_space_model_blockdiag()
    ts0 = ssm.constraint_ode_ts0(vf)
    solve = ivpsolve.solve_adaptive_terminal_values(
        solver=probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0),
        error=probdiffeq.error_state_std(constraint=ts0), control=ivpsolve.control_proportional_integral())  # fmt: skip
    tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0,), t=0.0)
    dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
    sol = solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=1.0, atol=1e-7, rtol=1e-4, dt0=dt0)
    torch.cuda.synchronize()
    assert int(sol.status.abs().max()) == 0

    s = H.spec(vf="burgers", fact="blockdiag", solver="solver", error="state_std", control="pi", clip_dt=True)
    tc = tcoeffs.cpu().numpy()
    grid = np.asarray([0.0, 1.0])
    # observations: the oracle's terminal solution at viscosity 0.01 + 1e-2 N(0, 1), std 1e-2
    ovf = o_pdq.ode("burgers", np.asarray([0.01]))
    tc_ref = np.asarray(ovf.taylor_coefficients([u0[0]], 0.0, 3))
    ref, _ = H.oracle_solve_save_at(s, tc_ref, np.asarray([0.01]), grid, 1e-7, 1e-4, dt0=o_ivp.dt0(ovf, (u0[0],), t=0.0))
    data = np.asarray(ref.u_mean)[-1][0] + 1e-2 * np.random.Generator(np.random.PCG64(33)).normal(size=(d,))
    std = np.full((d,), 1e-2)
    got_lml = probdiffeq.loss_lml_terminal_values()(data[None, :], marginals=sol.u, std=std[None, :]).cpu().numpy()
    mean = sol.u.mean_flat.cpu().numpy()
    steps, attempts = sol.num_steps.cpu().numpy(), sol.num_attempts.cpu().numpy()
    dt0_h = dt0.cpu().numpy().reshape(-1)
    stable = 0
    for b in range(B):
        osol, otr = H.oracle_solve_save_at(s, tc[b], visc[b], grid, 1e-7, 1e-4, dt0=float(dt0_h[b]))
        pert, ptr = H.oracle_solve_save_at(s, tc[b], visc[b], grid, 1e-7, 1e-4, dt0=float(dt0_h[b]) * (1 + 2.3e-16))
        o_u, p_u = np.asarray(osol.u_mean)[-1][0], np.asarray(pert.u_mean)[-1][0]
        sens = np.max(np.abs(p_u - o_u)) / np.max(np.abs(o_u))
        if [r[3] for r in ptr] == [r[3] for r in otr]:
            stable += 1
            assert int(attempts[b]) == len(otr) and int(steps[b]) == int(osol.num_steps[-1]), (b, attempts[b], len(otr))
        rel = np.max(np.abs(mean[b, 0] - o_u)) / np.max(np.abs(o_u))
        assert rel < max(1e-8, 100 * sens), (b, rel, sens)
        lml_o = o_pdq.loss_lml_terminal_values()(data, marginals=osol.u[-1], std=std)
        lml_p = o_pdq.loss_lml_terminal_values()(data, marginals=pert.u[-1], std=std)
        assert abs(got_lml[b] - lml_o) <= max(1e-6 * abs(lml_o), 100 * abs(lml_p - lml_o)), (b, got_lml[b], lml_o, lml_p)
    assert stable >= B // 2, stable
