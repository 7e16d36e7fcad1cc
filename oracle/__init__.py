"""CPU oracle: a float64 NumPy restatement of probdiffeq's adaptive probabilistic IVP step loop.

THIS PACKAGE IS TEST INFRASTRUCTURE. It is not part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker (or as the reported CPU baseline) -- never as a fallback for the
CUDA path.

Parity status: pinned against outputs of the reference itself
-------------------------------------------------------------
JAX (and ``matfree``) cannot be installed in this image, and the reference holds no golden outputs of
the step loop.  But the reference touches JAX only through its own ``probdiffeq/backend`` package, so
``oracle/refshim`` supplies that interface on NumPy / SciPy and the reference's UNMODIFIED modules
(``_ivpsolve``, ``_probdiffeq``, ``util``) are imported from ``/root/reference`` on top of it
(``tests/golden/make_reference_golden.py``, ``make_reference_golden_aux.py``).  Their outputs are
committed as fixtures (``tests/golden/reference_numpy_backend*.npz``) and this restatement is held
against them in ``tests/test_reference_golden*.py``: every strategy x factorisation on
Lotka-Volterra, the headline configuration to t = 50, HIRES dense ts1, Pleiades fixed-point, Van der
Pol, Burgers, ``constraint_init``, the error-estimate options, Taylor-mode initialisation, ``dt0`` /
``dt0_adaptive``, both log-marginal-likelihood losses, ``sample`` and ``offgrid_marginals`` --
identical accepted-step counts everywhere, most values bitwise, the rest at the reference's own
one-ulp sensitivity (recorded in the fixtures).  The CUDA path is held against the same fixtures in
``tests/test_gpu_reference_golden.py``.

What that does NOT pin: XLA's own arithmetic (operation fusion, its QR, the JAX PRNG stream) -- the
shim runs the reference's algorithm in NumPy float64 with SciPy's LAPACK; agreement with a JAX run is
expected at rounding level, not bit for bit.  ``jax.experimental.jet`` is replaced in the shim by an
exact-rational truncated-series evaluation for polynomial right-hand sides.

Beside the fixtures, the reference's analytic known-answer tests (IWP transition ``A(dt)``, ``Q(dt)``;
``revert_conditional`` identities; log-pdf vs dense MVN; PI==I controller identity) and its
cross-implementation identities (dense == isotropic, vmap(dense) == blockdiag,
fixed-grid-on-adaptive-grid == adaptive, save_at == terminal values, accuracy vs an independent
integrator) are re-run against this restatement in ``tests/test_oracle_*.py``.

Every function cites the reference file:line it restates (paths relative to ``/root/reference``).
"""

from oracle import ivpsolve, probdiffeq  # noqa: F401
