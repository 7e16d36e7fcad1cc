"""Development aid (test infrastructure): print the first attempts of the CUDA path next to the oracle's."""
import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import pdeq_test_helpers as H
s = H.spec(clip_dt=True, solver=sys.argv[1] if len(sys.argv) > 1 else "solver_mle")
params, u0 = H.lv_ensemble(8, seed=2)
p_pdq, p_ivp, vf, ssm, solver, err, ctrl = H.product_build(s, params)
tcoeffs, _ = p_pdq.jetexpand_ode_padded_scan(num=4)(vf, (u0,), t=0.0)
prior = ssm.prior_wiener_integrated(tcoeffs)
save_at = np.linspace(0.0, 5.0, 41)
sol = p_ivp.solve_adaptive_save_at(solver=solver, error=err, control=ctrl, clip_dt=True)(prior, save_at=save_at, atol=1e-7, rtol=1e-5, trace_capacity=400)
torch.cuda.synchronize()
b = 3
tr = sol.trace[b].cpu().numpy()
osol, otr = H.oracle_solve_save_at(s, tcoeffs[b].cpu().numpy(), params[b], save_at, 1e-7, 1e-5)
otr = np.asarray(otr)
for i in range(60, 85):
    print(i, "gpu t=%.17g dt=%.17g ep=%.10g %d | ora t=%.17g dt=%.17g ep=%.10g %d" % (*tr[i], *otr[i]))
print("----- first divergence")
m = min(len(otr), 400)
rel = np.abs(tr[:m,1]-otr[:m,1])/otr[:m,1]
first = int(np.argmax(rel > 1e-9))
for i in range(max(first-3,0), first+4):
    print(i, "gpu t=%.17g dt=%.17g ep=%.10g %d | ora t=%.17g dt=%.17g ep=%.10g %d" % (*tr[i], *otr[i]))
print(save_at[:12])
