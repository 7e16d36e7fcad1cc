"""Reference solutions / example problems of the reference's test-suite (reference: probdiffeq/backend/ode.py): not
needed to run the step loop; used when the reference's own tests are run on this backend
(oracle/refshim/run_reference_tests.py)."""
from collections import namedtuple

import numpy as _np

from oracle.refshim.backend import tree as _tree
from oracle.refshim.backend._array import to_arr as _to_arr


def odeint_and_save_at(vf, y0, /, save_at, *, atol, rtol):
    """The reference calls `jax.experimental.ode.odeint` (Dormand-Prince 5(4)); here SciPy's DOP853 at the same
    tolerances -- an independent integrator either way."""
    import scipy.integrate

    assert isinstance(y0, (tuple, list)) and len(y0) == 1
    save_at = _np.asarray(save_at, dtype=_np.float64)
    flat0, unravel = _tree.ravel_pytree(y0[0])

    def rhs(t, y):
        [vfx] = vf.vector_field(jet_coords=(unravel(_to_arr(_np.asarray(y))),), t=t)
        return _np.asarray(_tree.ravel_pytree(vfx)[0], dtype=_np.float64)

    sol = scipy.integrate.solve_ivp(rhs, (save_at[0], save_at[-1]), _np.asarray(flat0, dtype=_np.float64),
                                    t_eval=save_at, method="DOP853", atol=atol, rtol=rtol)  # fmt: skip
    states = [unravel(_to_arr(sol.y[:, k])) for k in range(sol.y.shape[1])]
    return _tree.tree_array_stack(states)


def ivp_lotka_volterra():
    """reference: backend/ode.py `ivp_lotka_volterra` -- the same deliberately awkward pytree state."""
    t0, t1 = (0.0, 2.0)
    PredPrey = namedtuple("PredPrey", ["predators", "prey"])
    u0 = {"U": PredPrey(predators=_to_arr(_np.asarray([[[20.0]]])), prey=_to_arr(_np.asarray(20.0)))}

    def vf(x, /, *, t):  # noqa: ARG001
        y0, y1 = f((x["U"].predators.squeeze(), x["U"].prey.squeeze()))
        return {"U": PredPrey(predators=y0.reshape((1, 1, 1)), prey=y1.reshape(()))}

    def f(y, /):
        a, b, c, d = 0.5, 0.05, 0.5, 0.05
        return [a * y[0] - b * (y[0] * y[1]), -c * y[1] + d * (y[0] * y[1])]

    return vf, (u0,), (t0, t1)


def _needs_diffeqzoo(*_args, **_kwargs):
    import pytest

    pytest.skip("the reference takes this problem from diffeqzoo, which is not installed here")


ivp_three_body_1st = ivp_van_der_pol_2nd = _needs_diffeqzoo
