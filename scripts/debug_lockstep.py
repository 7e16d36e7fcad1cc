"""Development aid: fixed-grid K2 filter (Pleiades, block-diagonal, solver_dynamic) against the oracle, point by point."""
import sys

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "scripts")
import parity_report as pr
import pdeq_test_helpers as H

cfg = pr.config_inputs("3", 2)
prod = pr.run_product(dict(cfg, save_at=np.linspace(0.0, 3.0, 33)))
tr = prod["trace"][0][: prod["num_attempts"][0]]
acc = tr[tr[:, 3] > 0.5]
full = np.concatenate([acc[:1, 0], acc[:, 0] + acc[:, 1]])
print("accepted grid", len(full), "monotone", bool(np.all(np.diff(full) > 0)), "min dt", np.diff(full).min())
s = dict(cfg["spec"], strategy="filter")
for T in (8, 64, 300, len(full)):
    grid = full[:T]
    p_pdq, p_ivp, vf, ssm, solver, _e, _c = H.product_build(s, None)
    prior = ssm.prior_wiener_integrated(torch.as_tensor(prod["tcoeffs"][:1], device="cuda"))
    sol = p_ivp.solve_fixed_grid(solver=solver)(prior, grid=grid)
    torch.cuda.synchronize()
    o = H.oracle_solve_fixed(s, prod["tcoeffs"][0], None, grid)
    got = sol.u.mean_flat[0].cpu().numpy()
    ref = np.asarray(o.u_mean)
    rel = np.max(np.abs(got - ref).reshape(T, -1), axis=1) / np.max(np.abs(ref).reshape(T, -1), axis=1)
    rel0 = np.max(np.abs(got - ref)[:, 0], axis=1) / np.max(np.abs(ref)[:, 0], axis=1)
    bad = np.nonzero(rel > 1e-6)[0]
    print(T, "status", int(sol.status[0]), "max rel", rel.max(), "coeff0", rel0.max(), "first bad point", bad[:3],
          "scale", sol.output_scale[0, -1, :3].cpu().numpy(), np.asarray(o.output_scale)[-1][:3])
