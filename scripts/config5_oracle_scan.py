"""Which Burgers resolution does BASELINE config 5 (blockdiag ts0, solver + error_state_std + PI, t in [0, 1],
rtol 1e-4, atol 1e-7, viscosity 0.01 U(0.5, 2), PCG64 seed 3) complete at IN THE ORACLE?  Runs the NumPy restatement
of the reference on every `stride`-th instance of the 4096-instance ensemble for a given d and reports how many end
finite (test infrastructure; results recorded in DESIGN.md section 7 and profiles/).

usage: python scripts/config5_oracle_scan.py D [stride] [workers] [constraint]
"""
import json
import multiprocessing as mp
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")


def one(job):
    d, visc, cons = job
    np.seterr(all="ignore")
    warnings.simplefilter("ignore")
    import pdeq_test_helpers as H
    from oracle import ivpsolve as oi
    from oracle import probdiffeq as opq
    from oracle import problems as op

    s = H.spec(vf="burgers", fact="blockdiag", constraint=cons, solver="solver", error="state_std", control="pi", clip_dt=True)
    vf = opq.ode("burgers", np.asarray([visc]))
    u0 = op.burgers_u0(d)
    tc = np.asarray(vf.taylor_coefficients([u0], 0.0, 3))
    dt0 = oi.dt0(vf, (u0,), t=0.0)
    try:
        sol, tr = H.oracle_solve_save_at(s, tc, np.asarray([visc]), np.asarray([0.0, 1.0]), 1e-7, 1e-4, dt0=dt0)
        fin = bool(np.all(np.isfinite(np.asarray(sol.u_mean)[-1])))
        return visc, fin, int(np.asarray(sol.num_steps)[-1]), len(tr)
    except Exception:
        return visc, False, -1, -1


if __name__ == "__main__":
    d = int(sys.argv[1])
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    cons = sys.argv[4] if len(sys.argv) > 4 else "ts0"
    visc = 0.01 * np.random.Generator(np.random.PCG64(3)).uniform(0.5, 2.0, size=(4096, 1))[::stride, 0]
    t0 = time.time()
    with mp.get_context("spawn").Pool(workers) as pool:
        res = pool.map(one, [(d, float(v), cons) for v in visc], chunksize=4)
    bad = [r for r in res if not r[1]]
    out = dict(d=d, constraint=cons, instances=len(res), stride=stride, finite=len(res) - len(bad),
               failed_viscosities=[r[0] for r in bad][:32], steps_min=min(r[2] for r in res if r[1]),
               steps_max=max(r[2] for r in res), seconds=time.time() - t0)
    print(json.dumps(out), flush=True)
