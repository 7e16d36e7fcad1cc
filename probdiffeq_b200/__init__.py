"""probdiffeq_b200: the adaptive probabilistic IVP step loop of pnkraemer/probdiffeq, B200-native.

``probdiffeq_b200.ivpsolve`` and ``probdiffeq_b200.probdiffeq`` mirror the reference's
``probdiffeq.ivpsolve`` / ``probdiffeq.probdiffeq`` namespaces for the ensemble hot path and call
hand-written sm_100a CUDA kernels through the C ABI in ``include/probdiffeq_b200.h``.
There is no CPU fallback.
"""

from probdiffeq_b200 import ivpsolve, probdiffeq  # noqa: F401

__version__ = "0.1.0"
