"""IVP loops, mirroring ``probdiffeq.ivpsolve`` of the reference (probdiffeq/ivpsolve.py:10-14).

``solve_adaptive_terminal_values`` / ``solve_adaptive_save_at`` / ``solve_fixed_grid`` keep the
reference's signatures (probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:16-22,34-36,46-54,100-102;
probdiffeq/_ivpsolve/solvers_via_fixed_steps.py:11-13,21) but run the whole loop -- for the whole
ensemble -- inside one persistent CUDA kernel.
"""

from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import torch

from probdiffeq_b200 import _lib
from probdiffeq_b200 import probdiffeq as _pdq

# Guard against runaway adaptive loops (the reference has none: a collapsing step size loops forever inside
# `lax.while_loop`). Instances that exceed it are reported with status 2 (PDEQ_STATUS_MAX_ATTEMPTS).
DEFAULT_MAX_ATTEMPTS = 5_000_000

__all__ = [
    "control_integral",
    "control_proportional_integral",
    "dt0",
    "dt0_adaptive",
    "solve_adaptive_save_at",
    "solve_adaptive_terminal_values",
    "solve_fixed_grid",
]


class control_proportional_integral:
    """reference: _ivpsolve/controllers.py:24-63."""

    kind = "proportional_integral"

    def __init__(self, *, safety=0.95, factor_min=0.2, factor_max=10.0, exponent_integral=0.3, exponent_proportional=0.4):
        self.safety = float(safety)
        self.factor_min = float(factor_min)
        self.factor_max = float(factor_max)
        self.exponent_integral = float(exponent_integral)
        self.exponent_proportional = float(exponent_proportional)


class control_integral:
    """reference: _ivpsolve/controllers.py:66-84."""

    kind = "integral"

    def __init__(self, *, safety=0.95, factor_min=0.2, factor_max=10.0):
        self.safety = float(safety)
        self.factor_min = float(factor_min)
        self.factor_max = float(factor_max)
        self.exponent_integral = 1.0
        self.exponent_proportional = 0.0


def _lower(prior, solver, error, control, clip_dt, max_attempts=0) -> _lib.Config:
    """Object graph -> pdeq_config."""
    constraint = solver.constraint
    if error is not None and error.constraint is not constraint:
        c2 = error.constraint
        if (c2.kind, c2.ode.vf_id, c2.factorisation) != (constraint.kind, constraint.ode.vf_id, constraint.factorisation):
            raise ValueError("solver and error estimator must share the constraint on the accelerated path.")
    if constraint.factorisation != prior.factorisation:
        raise ValueError(
            f"constraint was built for the {constraint.factorisation} model, prior for {prior.factorisation}."
        )
    kw = dict(
        constraint=_lib.CONSTRAINT[constraint.kind],
        solver=_lib.SOLVER[solver.kind],
        strategy=_lib.STRATEGY[solver.strategy.kind],
        clip_dt=int(bool(clip_dt)),
        max_attempts=int(max_attempts),
    )
    for k, v in solver.options.items():
        kw[k] = v
    if error is not None:
        kw.update(
            error=_lib.ERROR[error.kind],
            error_norm=_lib.NORM[error.error_norm],
            derivative_idx=error.derivative_idx,
            error_per_unit_step=int(error.error_per_unit_step),
        )
    if control is not None:
        kw.update(
            control=_lib.CONTROL[control.kind],
            safety=control.safety,
            factor_min=control.factor_min,
            factor_max=control.factor_max,
            exponent_integral=control.exponent_integral,
            exponent_proportional=control.exponent_proportional,
        )
    cfg = _pdq._make_config(
        fact=prior.factorisation, nu=prior.num_derivatives, d=prior.ode_dim, vf=constraint.ode, **kw
    )
    rc = _lib.load().pdeq_config_supported(C.byref(cfg))
    _lib.check(rc, "pdeq_config_supported")
    return cfg


def cost_hint_is_used(B, device) -> bool:
    """Whether `solve(..., cost_hint=)` reorders an ensemble of B instances on this device (callers that would have to
    COMPUTE a hint can skip that work when it is not)."""
    lanes = torch.cuda.get_device_properties(device).multi_processor_count * 384
    return 0 < B <= COST_HINT_MAX_ROUNDS * lanes


def _service_order(cost_hint, B, device):
    """`cost_hint` (B,): any positive proxy of how expensive each instance is (a previous solve's `num_attempts`, a
    model of the parameters, ...). The persistent kernels then serve the instances longest-first (pdeq_problem.order),
    which evens out the end of a solve with few instances per lane; and because a lane's first ticket is its thread
    index, the 32 lanes of a warp start on 32 neighbours of the order -- equally long instances -- and keep finishing
    (and refilling) together. Results do not depend on it.

    Measured on the thread-per-instance kernel with a fitted quadratic cost model (scripts/sweep_k1_order.py,
    profiles/r2w_sweep_k1_order.jsonl): 2^17 instances (2.3 per lane) 2.63 -> 2.31 ms, 2^18: 4.71 -> 4.41, 2^19 (9.2
    per lane): 8.82 -> 8.59; at 2^20 (18 per lane) evaluating and sorting the hint costs more than the order gains
    (16.89 -> 17.01) -- many rounds even out by themselves -- hence COST_HINT_MAX_ROUNDS. (A planned queue that reserves the cheapest instances for the lanes that
    must serve one instance more than the others looked 7 % better in a lane-level simulation and measured 7 % worse
    than longest-first; it was removed.)"""
    hint = torch.as_tensor(cost_hint, device=device).reshape(-1)
    if hint.shape[0] != B:
        raise ValueError("cost_hint must have one entry per ensemble member.")
    if not cost_hint_is_used(B, device):
        return None
    # only the coarse ranking matters: float32 keys halve the radix passes of the sort
    return torch.argsort(hint.to(torch.float32), descending=True).to(torch.int32)


def _problem(prior, vf, cost_hint=None):
    B = prior.tcoeffs.shape[0]
    dev = prior.tcoeffs.device
    if dev.type == "cuda" and dev.index != torch.cuda.current_device():
        # kernels are launched on the current stream of the CURRENT device, scratch is allocated beside the inputs
        raise ValueError(f"the ensemble lives on {dev} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         f"wrap the solve in `with torch.cuda.device({dev.index}):`")
    params, pstride = vf.params_on_device(B)
    pr = _lib.Problem()
    pr.num_instances = B
    pr.tcoeffs = _pdq._ptr(prior.tcoeffs)
    pr.init_std = _pdq._ptr(prior.init_std)
    pr.init_std_stride = 0 if prior.init_std is None or prior.init_std.shape[0] == 1 else prior.init_std[0].numel()
    pr.prior_scale = _pdq._ptr(prior.output_scale)
    pr.prior_scale_stride = (
        0 if prior.output_scale is None or prior.output_scale.shape[0] == 1 else prior.output_scale.shape[1]
    )
    pr.params = _pdq._ptr(params)
    pr.params_stride = pstride
    order = None if cost_hint is None or B == 0 else _service_order(cost_hint, B, prior.tcoeffs.device)
    pr.order = _pdq._ptr(order)
    return pr, (params, order)


COST_HINT_MAX_ROUNDS = 12  # instances per resident lane above which `cost_hint` is ignored (9.2: -2.6 %, 18.4: +0.7 %)
POSTERIOR_AUTO_BYTES = 2 << 30  # smoothers return their backward conditionals by default up to this size


def _alloc_solution(prior, T, want_chol=True, trace_capacity=0, want_posterior=False, nan_fill=True):
    """Output buffers of one solve. `nan_fill`: an instance that gives up (status MAX_ATTEMPTS) leaves its remaining
    checkpoints unwritten in the lane-per-dimension and dense kernels, so those read NaN instead of whatever the
    allocator handed out; the thread-per-instance kernel writes the NaNs itself (pdeq_loop_thread.cuh: emit_nan), and
    its callers skip the fill -- for the headline ensemble it would be a 190 MB memset per solve."""
    B, n, d = prior.tcoeffs.shape
    dev = prior.tcoeffs.device
    fact = prior.factorisation
    f64 = dict(dtype=torch.float64, device=dev)

    def fresh(shape):
        return torch.full(shape, float("nan"), **f64) if nan_fill else torch.empty(shape, **f64)

    chol_shape = {"isotropic": (B, T, n, n), "blockdiag": (B, T, d, n, n), "dense": (B, T, n * d, n * d)}[fact]
    if fact == "dense" and d == 1:
        chol_shape = (B, T, n, n)
    scale_shape = (B, T, d) if fact == "blockdiag" else (B, T)
    bufs = dict(
        t=fresh((B, T)),
        mean=fresh((B, T, n, d)),
        chol=fresh(chol_shape) if want_chol else None,
        output_scale=fresh(scale_shape),
        num_steps=(torch.zeros if nan_fill else torch.empty)((B, T), dtype=torch.int32, device=dev),
        num_attempts=torch.empty((B,), dtype=torch.int32, device=dev),
        status=torch.empty((B,), dtype=torch.int32, device=dev),
    )
    if trace_capacity > 0:
        bufs["trace"] = torch.full((B, trace_capacity, 4), float("nan"), **f64)
    if want_posterior:
        bufs["bw_gain"] = torch.zeros(chol_shape, **f64)
        bufs["bw_mean"] = torch.zeros((B, T, n, d), **f64)
        bufs["bw_chol"] = torch.zeros(chol_shape, **f64)
        if want_posterior == "with_filtering":
            bufs["filt_mean"] = torch.zeros((B, T, n, d), **f64)
            bufs["filt_chol"] = torch.zeros(chol_shape, **f64)
    so = _lib.Solution()
    for k, v in bufs.items():
        setattr(so, k, _pdq._ptr(v))
    so.trace_capacity = int(trace_capacity)
    return so, bufs


def _want_posterior(flag, solver, prior, T, want_chol):
    """Smoothers return the posterior (terminal marginal + backward conditionals) like the reference does;
    `None` = yes unless the conditionals would exceed POSTERIOR_AUTO_BYTES."""
    if solver.strategy.kind == "filter" or not want_chol:
        if flag:
            raise ValueError("want_posterior needs a smoother and want_cholesky=True")
        return False
    if flag is not None:
        return bool(flag)
    B, n, d = prior.tcoeffs.shape
    if prior.factorisation == "dense":
        return 8 * B * T * (2 * (n * d) ** 2 + n * d) <= POSTERIOR_AUTO_BYTES
    blocks = 1 if prior.factorisation == "isotropic" else d
    return 8 * B * T * (2 * blocks * n * n + n * d) <= POSTERIOR_AUTO_BYTES


def _wrap(prior, bufs, *, terminal: bool):
    fact_u = prior.factorisation if not (prior.factorisation == "dense" and prior.ode_dim == 1) else "isotropic"
    full = None
    if "bw_gain" in bufs:
        pick = (lambda x: x[0]) if prior.unbatched else (lambda x: x)
        post = _pdq.MarkovSequence(
            marginal=_pdq.Normal(fact_u, pick(bufs["mean"][:, -1]), pick(bufs["chol"][:, -1])),
            conditional=_pdq.BackwardConditional(pick(bufs["bw_gain"]), pick(bufs["bw_mean"]), pick(bufs["bw_chol"])),
        )
        filtering = None
        if "filt_mean" in bufs:
            filtering = _pdq.Normal(fact_u, pick(bufs["filt_mean"]), pick(bufs["filt_chol"]))
        full = _pdq.SmoothingSolution(posterior=post, filtering=filtering)
    sol = _pdq.ProbabilisticSolution(
        t=bufs["t"],
        u=_pdq.Normal(fact_u, bufs["mean"], bufs["chol"]),
        output_scale=bufs["output_scale"],
        num_steps=bufs["num_steps"],
        num_attempts=bufs["num_attempts"],
        status=bufs["status"],
        solution_full=full,
        prior=prior,
    )  # fmt: skip
    trace = bufs.get("trace")
    if terminal:
        sol = sol._index(lambda x: x[:, -1])
    if prior.unbatched:
        sol = sol._index(lambda x: x[0])
        sol.num_attempts = sol.num_attempts[0]
        sol.status = sol.status[0]
        trace = None if trace is None else trace[0]
    sol.trace = trace  # (B, capacity, 4): t_from, dt, error_power, accepted -- only if trace_capacity > 0
    return sol


def _workspace(cfg, B, T, device):
    nbytes = _lib.load().pdeq_workspace_bytes(C.byref(cfg), B, T)
    return torch.empty((max(int(nbytes), 8),), dtype=torch.uint8, device=device), int(nbytes)


def _run_adaptive(prior, solver, error, control, clip_dt, save_at, atol, rtol, dt0, eps, damp, *, terminal,
                  want_chol=True, max_attempts=0, trace_capacity=0, want_posterior=None, cost_hint=None):  # fmt: skip
    if control is None:
        control = control_integral()  # solvers_via_adaptive_steps.py:87-90
    cfg = _lower(prior, solver, error, control, clip_dt, max_attempts)
    B = prior.tcoeffs.shape[0]
    grid = _pdq._as_device_f64(np.asarray(save_at, dtype=np.float64) if not isinstance(save_at, torch.Tensor) else save_at)
    if grid.ndim != 1:
        raise ValueError("save_at must be one-dimensional (shared by the ensemble).")
    T = grid.shape[0]
    dt0_t = _pdq._as_device_f64(dt0).reshape(-1)
    if dt0_t.shape[0] not in (1, B):
        raise ValueError("dt0 must be a scalar or have one entry per ensemble member.")
    pr, keep = _problem(prior, solver.constraint.ode, cost_hint)
    # the thread-per-instance kernel (filters of a compile-time d <= 8, not dense) fills abandoned checkpoints itself
    _, n_, d_ = prior.tcoeffs.shape
    interp_bytes = 0 if clip_dt else (n_ * d_ + (d_ if prior.factorisation == "blockdiag" else 1) * n_ * n_ + 1) * 1024
    on_k1 = (solver.strategy.kind == "filter" and prior.factorisation != "dense" and d_ <= 8
             and interp_bytes <= 227 * 1024)  # pdeq_api.cu: select_loop
    so, bufs = _alloc_solution(prior, T, want_chol, trace_capacity,
                               _want_posterior(want_posterior, solver, prior, T, want_chol) and not terminal,
                               nan_fill=not on_k1)
    if B > 0:  # an empty ensemble returns empty arrays (there is nothing to launch)
        ws, nbytes = _workspace(cfg, B, T, prior.tcoeffs.device)
        rc = _lib.load().pdeq_solve_adaptive_save_at(
            C.byref(cfg), C.byref(pr), _pdq._ptr(grid), T, float(atol), float(rtol), _pdq._ptr(dt0_t),
            0 if dt0_t.shape[0] == 1 else 1, float(eps), float(damp), C.byref(so), _pdq._ptr(ws), nbytes,
            _pdq._stream(),
        )  # fmt: skip
        _lib.check(rc, "pdeq_solve_adaptive_save_at")
    del keep
    return _wrap(prior, bufs, terminal=terminal)


def solve_adaptive_terminal_values(solver, error, control=None, clip_dt: bool = True, *,
                                   max_attempts: int = DEFAULT_MAX_ATTEMPTS):
    """reference: _ivpsolve/solvers_via_adaptive_steps.py:16-43."""

    def solve(u, /, *, t0, t1, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0, want_cholesky=True, trace_capacity=0,
              cost_hint=None):
        save_at = np.asarray([t0, t1], dtype=np.float64)
        return _run_adaptive(u, solver, error, control, clip_dt, save_at, atol, rtol, dt0, eps, damp,
                             terminal=True, want_chol=want_cholesky, max_attempts=max_attempts,
                             trace_capacity=trace_capacity, cost_hint=cost_hint)  # fmt: skip

    return solve


def solve_adaptive_save_at(*, solver, error, control=None, clip_dt: bool = False, warn: bool = True,
                           max_attempts: int = DEFAULT_MAX_ATTEMPTS):  # fmt: skip
    """reference: _ivpsolve/solvers_via_adaptive_steps.py:46-148."""
    if not solver.is_suitable_for_save_at and warn:
        msg = f"Solver {solver} should not be used in solve_adaptive_save_at."
        msg += " This is typically caused by the wrong strategy selection."
        msg += " Try using filters or fixed-point smoothers."
        warnings.warn(msg, stacklevel=1)

    def solve(u, save_at, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0, want_cholesky=True, trace_capacity=0,
              want_posterior=None, cost_hint=None):
        return _run_adaptive(u, solver, error, control, clip_dt, save_at, atol, rtol, dt0, eps, damp,
                             terminal=False, want_chol=want_cholesky, max_attempts=max_attempts,
                             trace_capacity=trace_capacity, want_posterior=want_posterior, cost_hint=cost_hint)  # fmt: skip

    return solve


def solve_fixed_grid(*, solver):
    """reference: _ivpsolve/solvers_via_fixed_steps.py:11-34."""
    if not solver.is_suitable_for_save_every_step:
        msg = f"Solver {solver} should not be used in solve_adaptive_save_every_step/solve_fixed_grid."
        msg += " This is typically caused by using a fixed-point smoother."
        msg += " Try using filters or fixed-interval smoothers instead."
        warnings.warn(msg, stacklevel=1)

    def solve(u, /, *, grid, damp: float = 0.0, want_cholesky=True, want_posterior=None):
        prior = u
        cfg = _lower(prior, solver, None, None, False)
        B = prior.tcoeffs.shape[0]
        g = _pdq._as_device_f64(np.asarray(grid, dtype=np.float64) if not isinstance(grid, torch.Tensor) else grid)
        if g.ndim != 1:
            raise ValueError("grid must be one-dimensional (shared by the ensemble).")
        T = g.shape[0]
        pr, keep = _problem(prior, solver.constraint.ode)
        wp = _want_posterior(want_posterior, solver, prior, T, want_cholesky)
        # fixed grid + fixed-interval smoother: keep the filtering marginals too (dense output starts from them)
        so, bufs = _alloc_solution(prior, T, want_cholesky, 0, "with_filtering" if wp else False)
        if B > 0:
            ws, nbytes = _workspace(cfg, B, T, prior.tcoeffs.device)
            rc = _lib.load().pdeq_solve_fixed_grid(
                C.byref(cfg), C.byref(pr), _pdq._ptr(g), T, float(damp), C.byref(so), _pdq._ptr(ws), nbytes,
                _pdq._stream(),
            )  # fmt: skip
            _lib.check(rc, "pdeq_solve_fixed_grid")
        del keep
        return _wrap(prior, bufs, terminal=False)

    return solve


def dt0(vf, initial_values, /, *, t=0.0, scale=0.01, nugget=1e-5):
    """Initial step size per ensemble member (reference: _ivpsolve/stepsize_initialisers.py:7-21)."""
    inits = [_pdq._as_device_f64(u) for u in initial_values]
    if len(inits) != vf.order:
        raise ValueError(f"{vf.name} is an order-{vf.order} ODE; got {len(inits)} initial values.")
    unbatched = inits[0].ndim == 1
    inits = [u.reshape(1, -1) if u.ndim == 1 else u for u in inits]
    u0 = torch.stack(inits, dim=1).contiguous()
    B, _q, d = u0.shape
    cfg = _pdq._make_config(fact="isotropic", nu=vf.order, d=d, vf=vf)
    params, stride = vf.params_on_device(B)
    out = torch.empty((B,), dtype=torch.float64, device=u0.device)
    rc = _lib.load().pdeq_dt0(
        C.byref(cfg), B, _pdq._ptr(u0), _pdq._ptr(params), stride, float(t), float(scale), float(nugget),
        _pdq._ptr(out), _pdq._stream(),
    )  # fmt: skip
    _lib.check(rc, "pdeq_dt0")
    return out[0] if unbatched else out


def dt0_adaptive(vf, initial_values, /, t0, *, error_contraction_rate, rtol, atol):
    """Initial step size from the tolerances (reference: _ivpsolve/stepsize_initialisers.py:24-64)."""
    if len(initial_values) > 1:
        raise ValueError("dt0_adaptive is defined for first-order ODEs only.")
    u0 = _pdq._as_device_f64(initial_values[0])
    unbatched = u0.ndim == 1
    if unbatched:
        u0 = u0.reshape(1, -1)
    B, d = u0.shape
    cfg = _pdq._make_config(fact="isotropic", nu=vf.order, d=d, vf=vf)
    params, stride = vf.params_on_device(B)
    out = torch.empty((B,), dtype=torch.float64, device=u0.device)
    rc = _lib.load().pdeq_dt0_adaptive(
        C.byref(cfg), B, _pdq._ptr(u0), _pdq._ptr(params), stride, float(t0), float(error_contraction_rate),
        float(rtol), float(atol), _pdq._ptr(out), _pdq._stream(),
    )  # fmt: skip
    _lib.check(rc, "pdeq_dt0_adaptive")
    return out[0] if unbatched else out
