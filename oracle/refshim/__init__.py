"""Run the UNMODIFIED reference (`/root/reference/probdiffeq`) on a NumPy array backend.  TEST INFRASTRUCTURE.

The reference touches JAX only through its own `probdiffeq/backend/` package (845 lines of thin wrappers); everything
else -- the adaptive loop, the solvers, the strategies, the three state-space factorisations, the Cholesky utilities --
is written against that interface.  JAX is not installable here, so this package supplies a second implementation of
that interface on NumPy / SciPy (`oracle/refshim/backend/`: arrays, linear algebra, Python control flow, a small
pytree library, `vmap` as a loop, derivatives by the complex-step method) and `load()` imports the reference's own
modules from `/root/reference` on top of it.  What runs is then the reference's code, line for line; what is NOT
reproduced is XLA's arithmetic (operation fusion, its LAPACK calls) and JAX's autodiff -- results agree with JAX at
rounding level, not bit for bit.

The backend is validated by the reference's OWN test-suite: `python -m oracle.refshim.run_reference_tests` runs the
modules of `/root/reference/tests` that concern this path on it (224 pass, none fails; tests/test_refshim_backend.py).

Used by `tests/golden/make_reference_golden.py` (which only runs where `/root/reference` exists, i.e. in the build
container) to write `tests/golden/reference_numpy_backend.npz`: outputs of the reference's own step loop that pin the
oracle and the CUDA path.  Nothing under `probdiffeq_b200/` imports this.
"""

from __future__ import annotations

import importlib
import pathlib
import sys
import types

REFERENCE = pathlib.Path("/root/reference")
_BACKEND_MODULES = ("abc", "flow", "func", "inspect", "linalg", "np", "ode", "random", "structs", "testing", "timing",
                    "tree", "typing", "warnings")  # fmt: skip


def available() -> bool:
    return (REFERENCE / "probdiffeq" / "ivpsolve.py").exists()


def load():
    """Import the reference with its backend swapped; returns its (ivpsolve, probdiffeq) modules."""
    if not available():
        raise RuntimeError("the reference sources are not on this machine (/root/reference)")
    if "probdiffeq" in sys.modules and getattr(sys.modules["probdiffeq"], "_pdeq_refshim", False):
        return sys.modules["probdiffeq.ivpsolve"], sys.modules["probdiffeq.probdiffeq"]
    if "probdiffeq" in sys.modules:
        raise RuntimeError("another `probdiffeq` is already imported in this process")
    pkg = types.ModuleType("probdiffeq")
    pkg.__path__ = [str(REFERENCE / "probdiffeq")]
    pkg._pdeq_refshim = True
    sys.modules["probdiffeq"] = pkg
    ver = types.ModuleType("probdiffeq._version")
    ver.version = "reference-on-numpy-backend"
    sys.modules["probdiffeq._version"] = ver
    backend = types.ModuleType("probdiffeq.backend")
    backend.__path__ = []  # a package whose submodules are all pre-loaded below
    sys.modules["probdiffeq.backend"] = backend
    for name in _BACKEND_MODULES:
        mod = importlib.import_module(f"oracle.refshim.backend.{name}")
        sys.modules[f"probdiffeq.backend.{name}"] = mod
        setattr(backend, name, mod)
    ivpsolve = importlib.import_module("probdiffeq.ivpsolve")
    probdiffeq = importlib.import_module("probdiffeq.probdiffeq")
    return ivpsolve, probdiffeq
