"""Initial values and ensembles of the benchmark problems (host-side constants; the right-hand sides themselves are
the registered device functors, see `probdiffeq.registered_vector_fields`).

Values restated from the reference's benchmark scripts: benchmarks/A0_work-precision-lotkavolterra.py:104-112,
A1_work-precision-pleiades.py:95-104, A2_work-precision-hires.py:173, A5_work-precision-burgers-pde.py:123-138.
The ensembles follow BASELINE.json's configs (seeded `numpy.random.Generator(PCG64)` draws).
"""

from __future__ import annotations

import numpy as np

LOTKA_VOLTERRA_PARAMS = np.asarray([0.5, 0.05, 0.5, 0.05])
LOTKA_VOLTERRA_U0 = np.asarray([20.0, 20.0])
HIRES_U0 = np.asarray([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0057])
# seven bodies: x positions, y positions, x velocities, y velocities
PLEIADES_U0 = np.asarray(
    [3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0]
    + [3.0, -3.0, 2.0, 0.0, 0.0, -4.0, 4.0]
    + [0.0, 0.0, 0.0, 0.0, 0.0, 1.75, -1.5]
    + [0.0, 0.0, 0.0, -1.25, 1.0, 0.0, 0.0]
)


def burgers_u0(d: int) -> np.ndarray:
    """sin^3(3 pi x) (1 - x)^1.5 on the d interior points of a uniform grid on [0, 1]."""
    x = np.linspace(0.0, 1.0, d + 2, endpoint=True)[1:-1]
    return np.sin(3.0 * np.pi * x) ** 3 * (1.0 - x) ** 1.5


def lotka_volterra_ensemble(num_instances: int, seed: int = 0):
    """BASELINE config 2: parameters = base * U(0.8, 1.2)^4 and u0 = 20 * U(0.8, 1.2)^2, drawn as one (B, 6) array."""
    draw = np.random.Generator(np.random.PCG64(seed)).uniform(0.8, 1.2, size=(num_instances, 6))
    return LOTKA_VOLTERRA_PARAMS[None, :] * draw[:, :4], 20.0 * draw[:, 4:]


def pleiades_ensemble(num_instances: int, seed: int = 1) -> np.ndarray:
    """BASELINE config 3: base + 1e-3 N(0, 1)^28."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return PLEIADES_U0[None, :] + 1e-3 * rng.normal(size=(num_instances, 28))
