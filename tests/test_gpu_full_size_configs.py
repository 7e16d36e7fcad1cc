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
    assert abs(int(steps[0, -1]) - int(osteps[-1])) <= 0.05 * osteps[-1]
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
