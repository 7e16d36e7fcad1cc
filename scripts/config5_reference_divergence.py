"""BASELINE config 5 as literally specified (Burgers semi-discretisation, block-diagonal ts0, solver + error_state_std +
PI, rtol 1e-4, atol 1e-7) run through the REFERENCE's own code (the unmodified modules of /root/reference on the NumPy
array backend of oracle/refshim -- see tests/golden/make_reference_golden.py) beside the oracle: does the solve end
finite, and after how many accepted steps does it not?  Needs /root/reference (the build container); the result is
committed as profiles/r3k_config5_reference_divergence.jsonl and cited in DESIGN.md section 7.

usage: python scripts/config5_reference_divergence.py D T1 [DT0]      (viscosity 0.01, the ensemble's nominal value)
"""
import importlib.util
import json
import pathlib
import sys
import time
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import pdeq_test_helpers as H  # noqa: E402
from oracle import ivpsolve as o_ivp  # noqa: E402
from oracle import probdiffeq as o_pdq  # noqa: E402
from oracle import problems as o_problems  # noqa: E402

_spec = importlib.util.spec_from_file_location("mk", ROOT / "tests" / "golden" / "make_reference_golden.py")
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)

if __name__ == "__main__":
    warnings.simplefilter("ignore")
    np.seterr(all="ignore")
    d, t1 = int(sys.argv[1]), float(sys.argv[2])
    ovf = o_pdq.ode("burgers", np.asarray([0.01]))
    u0 = o_problems.burgers_u0(d)
    dt0 = float(sys.argv[3]) if len(sys.argv) > 3 else float(o_ivp.dt0(ovf, (u0,), t=0.0))
    c = dict(name="config5", kind="terminal", grid=[0.0, t1], atol=1e-7, rtol=1e-4, dt0=dt0, diffuse_start=False,
             problem=dict(vf="burgers", nu=3, params=[0.01], u0=list(u0)), spec=H.spec(vf="burgers", fact="blockdiag"))
    c["tcoeffs"] = np.asarray(ovf.taylor_coefficients([u0], 0.0, 3))
    out = dict(d=d, t1=t1, dt0=dt0, constraint="ts0")
    for who, run, arrays in (("reference", mk.run_reference, mk.reference_arrays),
                             ("oracle", mk.run_oracle, lambda sol: mk.oracle_arrays(sol, False))):  # fmt: skip
        t = time.time()
        r = arrays(run(c))
        out[who] = dict(accepted_steps=int(np.max(r["num_steps"])), finite=bool(np.all(np.isfinite(r["mean"]))),
                        seconds=round(time.time() - t, 1))  # fmt: skip
    print(json.dumps(out), flush=True)
