"""Adaptive / fixed-grid IVP loops of the oracle (test infrastructure, see oracle/__init__.py).

Restates

* _ivpsolve/solvers_via_adaptive_steps.py:16-148 (solve_adaptive_terminal_values / save_at),
  :151-375 (TimeStepState, RejectionLoop with the three interpolation branches);
* _ivpsolve/controllers.py:24-84 (PI and I controllers);
* _ivpsolve/solvers_via_fixed_steps.py:11-34 (solve_fixed_grid);
* _ivpsolve/stepsize_initialisers.py:7-64 (dt0, dt0_adaptive);
* util/test_util.py:10-80 (solve_adaptive_save_every_step, used by cross-checks).

`jax.lax.while_loop/scan/cond/switch` become Python control flow; the order of floating-point
operations in the accept test, the controller and the time accumulation follows the reference
(SURVEY.md Appendix C).  The loops additionally record every attempt (`trace`) so that the CUDA
path's accepted-step sequence can be compared.
"""

import warnings

import numpy as np

__all__ = [
    "control_integral",
    "control_proportional_integral",
    "dt0",
    "dt0_adaptive",
    "solve_adaptive_save_at",
    "solve_adaptive_save_every_step",
    "solve_adaptive_terminal_values",
    "solve_fixed_grid",
]


class control_proportional_integral:
    """_ivpsolve/controllers.py:24-63."""

    def __init__(self, *, safety=0.95, factor_min=0.2, factor_max=10.0, exponent_integral=0.3, exponent_proportional=0.4):
        self.safety = safety
        self.factor_min = factor_min
        self.factor_max = factor_max
        self.exponent_integral = exponent_integral
        self.exponent_proportional = exponent_proportional

    def init(self, dt, /):
        return 1.0

    def apply(self, dt, error_norm_inv_prev, /, *, error_power):
        gain_integral = error_power**self.exponent_integral
        gain_proportional = (error_power / error_norm_inv_prev) ** self.exponent_proportional
        ratio = self.safety * gain_integral * gain_proportional
        scale = np.maximum(self.factor_min, np.minimum(ratio, self.factor_max))
        prev = error_power if error_power >= 1.0 else error_norm_inv_prev
        return scale * dt, prev


class control_integral:
    """_ivpsolve/controllers.py:66-84."""

    def __init__(self, *, safety=0.95, factor_min=0.2, factor_max=10.0):
        self.safety = safety
        self.factor_min = factor_min
        self.factor_max = factor_max

    def init(self, dt, /):
        return ()

    def apply(self, dt, state, /, *, error_power):
        ratio = self.safety * error_power
        scale = np.maximum(self.factor_min, np.minimum(ratio, self.factor_max))
        return scale * dt, ()


class TimeStepState:
    """_ivpsolve/solvers_via_adaptive_steps.py:151-175."""

    def __init__(self, dt, step_from, interp_from, control, error_step_from):
        self.dt = dt
        self.step_from = step_from
        self.interp_from = interp_from
        self.control = control
        self.error_step_from = error_step_from


class RejectionLoop:
    """_ivpsolve/solvers_via_adaptive_steps.py:196-375."""

    def __init__(self, solver, clip_dt, error, control, trace=None):
        self.solver = solver
        self.clip_dt = clip_dt
        self.error = error
        self.control = control
        self.trace = trace  # optional list; receives (t_from, dt, error_power, accepted)

    def init(self, state_solver, dt):
        return TimeStepState(dt, state_solver, state_solver, self.control.init(dt), self.error.init_error())

    def loop(self, state0, *, t1, atol, rtol, eps, damp):
        state = state0
        if state0.step_from.t + eps < t1:
            state = self.step(state0, t1, atol, rtol, damp)
        if state.step_from.t + eps < t1:
            return state.step_from, state
        if state.step_from.t > t1 + eps:
            return self._interp(self.solver.interpolate_fwd, state, t1)
        return self._interp(self.solver.interpolate_fwd_at_t1, state, t1)

    def step(self, s, t1, atol, rtol, damp):
        dt, control = s.dt, s.control
        acceptance = 0.9  # < 1 so the loop body runs at least once (:273-275)
        proposed, error_proposed = None, None
        while acceptance < 1.0:
            if self.clip_dt:
                dt = np.minimum(dt, t1 - s.step_from.t)
            dt_used = dt
            proposed = self.solver.step(state=s.step_from, dt=dt, damp=damp)
            acceptance, error_proposed = self.error.estimate_error_norm(
                s.error_step_from, previous=s.step_from, proposed=proposed,
                dt=dt, atol=atol, rtol=rtol, damp=damp,
            )  # fmt: skip
            dt, control = self.control.apply(dt, control, error_power=acceptance)
            if self.trace is not None:
                self.trace.append((float(s.step_from.t), float(dt_used), float(acceptance), bool(acceptance >= 1.0)))
        return TimeStepState(dt, proposed, s.step_from, control, error_proposed)

    def _interp(self, fn, state, t1):
        solution, res = fn(t=t1, interp_from=state.interp_from, interp_to=state.step_from)
        new = TimeStepState(state.dt, res.step_from, res.interp_from, state.control, state.error_step_from)
        return solution, new


def solve_adaptive_save_at(*, solver, error, control=None, clip_dt=False, warn=True, trace=None):
    """_ivpsolve/solvers_via_adaptive_steps.py:46-148."""
    if not solver.is_suitable_for_save_at and warn:
        warnings.warn(f"Solver {solver} should not be used in solve_adaptive_save_at.", stacklevel=1)
    if control is None:
        control = control_integral()
    loop = RejectionLoop(solver=solver, clip_dt=clip_dt, control=control, error=error, trace=trace)

    def solve(u, save_at, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0):
        save_at = np.asarray(save_at, dtype=np.float64)
        solution0 = solver.init(t=save_at[0], u=u, damp=damp)
        state = loop.init(solution0, dt=dt0)
        stacked = []
        for t_next in save_at[1:]:
            do_continue = True
            solution = None
            while do_continue:  # body runs at least once (:128)
                solution, state = loop.loop(state, t1=t_next, atol=atol, rtol=rtol, eps=eps, damp=damp)
                do_continue = state.step_from.t + eps < t_next
            stacked.append(solution)
        return solver.userfriendly_output(solution0=solution0, solution=stacked, solution1=state.step_from)

    return solve


def solve_adaptive_terminal_values(solver, error, control=None, clip_dt=True, trace=None):
    """_ivpsolve/solvers_via_adaptive_steps.py:16-43."""
    save_at_solve = solve_adaptive_save_at(
        solver=solver, error=error, control=control, clip_dt=clip_dt, warn=False, trace=trace
    )

    def solve(u, /, *, t0, t1, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0):
        sol = save_at_solve(u, save_at=np.asarray([t0, t1]), atol=atol, rtol=rtol, dt0=dt0, eps=eps, damp=damp)
        return sol.terminal()

    return solve


def solve_fixed_grid(*, solver, terminal="reference"):
    """_ivpsolve/solvers_via_fixed_steps.py:11-34.

    `terminal="reference"` is the literal restatement: the last grid state is passed on as `solution1`, which
    `Smoother.finalize` (estimators_and_losses.py:453-454) treats as an overstepped state and marginalises through
    its own backward conditional -- for smoothers the marginals then come out one interval late at the terminal
    grid point. `terminal="aligned"` passes the last state with an identity conditional instead (what the
    overstepping save-every-step flow of util/test_util.py:10-66 amounts to), giving the Rauch-Tung-Striebel pass.
    """
    if not solver.is_suitable_for_save_every_step:
        warnings.warn(f"Solver {solver} should not be used in solve_fixed_grid.", stacklevel=1)

    def solve(u, /, *, grid, damp=0.0):
        grid = np.asarray(grid, dtype=np.float64)
        state = solver.init(t=grid[0], u=u, damp=damp)
        state0 = state
        result = []
        for dt in np.diff(grid):
            state = solver.step(state=state, dt=dt, damp=damp)
            result.append(state)
        last = state
        if terminal == "aligned" and hasattr(state.solution_full, "conditional"):
            import copy

            full = state.solution_full
            ident = full.marginal.alg.identity_conditional(full.marginal)
            last = copy.copy(state)
            last.solution_full = type(full)(full.marginal, ident, full.reverse)
        return solver.userfriendly_output(solution0=state0, solution=result, solution1=last)

    return solve


def solve_adaptive_save_every_step(*, solver, error, control=None, clip_dt=False):
    """util/test_util.py:10-80: record every accepted step."""
    if control is None:
        control = control_integral()
    loop = RejectionLoop(solver=solver, clip_dt=clip_dt, control=control, error=error)

    def solve(u, /, *, t0, t1, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0):
        solution0 = solver.init(t=t0, u=u, damp=damp)
        state = loop.init(solution0, dt=dt0)
        result = []
        while state.step_from.t < t1:
            solution, state = loop.loop(state, t1=t1, atol=atol, rtol=rtol, eps=eps, damp=damp)
            result.append(solution)
        return solver.userfriendly_output(solution0=solution0, solution=result, solution1=state.step_from)

    return solve


def dt0(vf, initial_values, /, *, t=0.0, scale=0.01, nugget=1e-5):
    """_ivpsolve/stepsize_initialisers.py:7-21."""
    f0 = vf.vector_field(initial_values, t)
    u0 = np.asarray(initial_values[0], dtype=np.float64).reshape(-1)
    return scale * np.linalg.norm(u0) / (np.linalg.norm(np.asarray(f0).reshape(-1)) + nugget)


def dt0_adaptive(vf, initial_values, /, t0, *, error_contraction_rate, rtol, atol):
    """_ivpsolve/stepsize_initialisers.py:24-64 (Hairer et al., Sec. II.4)."""
    if len(initial_values) > 1:
        raise ValueError
    y0 = np.asarray(initial_values[0], dtype=np.float64).reshape(-1)
    f0 = np.asarray(vf.vector_field((y0,), t0)).reshape(-1)
    scale = atol + np.abs(y0) * rtol
    d0, d1 = np.linalg.norm(y0), np.linalg.norm(f0)
    h0 = 1e-6 if (d0 < 1e-5) or (d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = np.asarray(vf.vector_field((y1,), t0 + h0)).reshape(-1)
    d2 = np.linalg.norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = np.maximum(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / np.maximum(d1, d2)) ** (1.0 / (error_contraction_rate + 1.0))
    return np.minimum(100.0 * h0, h1)
