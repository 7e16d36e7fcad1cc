"""CPU oracle: a float64 NumPy restatement of probdiffeq's adaptive probabilistic IVP step loop.

THIS PACKAGE IS TEST INFRASTRUCTURE. It is not part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker (or as the reported CPU baseline) -- never as a fallback for the
CUDA path.

Parity status
-------------
The reference (``/root/reference``, pure Python on top of JAX) cannot be imported in this image:
``jax`` and ``matfree`` are absent and there is no network.  The reference holds no golden outputs of
the step loop (its only ``.npy`` fixtures are pickled float32 ``jax.Array`` objects for unrelated
problems).  The oracle is therefore pinned by

* the reference's analytic known-answer tests (IWP transition ``A(dt)``, ``Q(dt)``;
  ``revert_conditional`` identities; log-pdf vs dense MVN; PI==I controller identity), and
* the reference's cross-implementation identities (dense == isotropic, vmap(dense) == blockdiag,
  fixed-grid-on-adaptive-grid == adaptive, save_at == terminal values, accuracy vs an independent
  integrator),

all re-run against this restatement in ``tests/test_oracle_*.py``.  Bit-level parity with
JAX/XLA/LAPACK output is **unpinned** (the third-party arithmetic -- ``jax[cpu]`` unpinned in
``pyproject.toml:25-30`` -- is not on disk); mathematical parity at the 1e-10 / 1e-8 level is pinned.

Every function cites the reference file:line it restates (paths relative to ``/root/reference``).
"""

from oracle import ivpsolve, probdiffeq  # noqa: F401
