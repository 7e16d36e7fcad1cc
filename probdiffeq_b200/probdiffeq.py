"""Solver-construction API, mirroring ``probdiffeq.probdiffeq`` of the reference for the hot path.

The reference composes Python objects (prior x strategy x constraint x solver x error) and traces
them with JAX.  Here the same constructors return light-weight descriptors; ``ivpsolve.solve_*``
lowers the object graph to the POD ``pdeq_config`` and calls the CUDA kernels through the C ABI.

Differences a reference user will notice:

* vector fields are *registered device functors* -- ``ode("lotka_volterra", params=...)`` instead of a
  decorated Python function (reference: probdiffeq/_probdiffeq/problems.py:283-312);
* every array carries a leading ensemble axis ``B`` (the reference's ``jax.vmap`` axis); unbatched
  inputs are treated as ``B = 1`` and squeezed on return;
* arrays are ``torch`` CUDA tensors (float64).

Reference files mirrored: probdiffeq/probdiffeq.py:6-17 (namespace),
_probdiffeq/ssm_impl_{isotropic,blockdiag,dense}.py (`state_space_model_*`, priors, constraints),
_probdiffeq/solvers.py:318,483,636,770,794,850,999, _probdiffeq/estimators_and_losses.py:20,347,473,
_probdiffeq/jet_expansion_algorithms.py:49,110.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any

import numpy as np
import torch

from probdiffeq_b200 import _iwp, _lib

__all__ = [
    "ProbabilisticSolution",
    "error_norm_rms_then_scale",
    "error_norm_scale_then_rms",
    "error_residual_std",
    "error_state_std",
    "jetexpand_ode_padded_scan",
    "jetexpand_ode_unroll",
    "jetexpand_ode_coefficient_increment",
    "loss_lml_terminal_values",
    "loss_lml_timeseries",
    "MarkovSequence",
    "SmoothingSolution",
    "BackwardConditional",
    "ode",
    "registered_vector_fields",
    "solver",
    "solver_dynamic",
    "solver_mle",
    "state_space_model_blockdiag",
    "state_space_model_dense",
    "state_space_model_isotropic",
    "strategy_filter",
    "strategy_smoother_fixedinterval",
    "strategy_smoother_fixedpoint",
]


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.NativeLibraryError("probdiffeq_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_f64(x, *, allow_none=False):
    """Host (numpy / pinned torch) or device array -> contiguous float64 CUDA tensor."""
    if x is None:
        if allow_none:
            return None
        raise ValueError("array expected")
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
    if x.dtype != torch.float64:
        x = x.to(torch.float64)
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=True)
    return x.contiguous()


def _ptr(x) -> int | None:
    return None if x is None else x.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------------
# Vector fields
# ------------------------------------------------------------------------------------------------------

registered_vector_fields = ("lotka_volterra", "pleiades", "hires", "vanderpol", "linear", "burgers")


class VectorField:
    """A registered right-hand side u^(order) = f(u, ..., u^(order-1), t) (reference: `JetOde`)."""

    def __init__(self, name: str, params=None):
        lib = _lib.load()
        vf_id = lib.pdeq_vf_id(name.encode())
        if vf_id < 0:
            raise ValueError(f"Unknown vector field {name!r}. Registered: {registered_vector_fields}.")
        self.name = name
        self.vf_id = vf_id
        self.num_tcoeffs_in_args = lib.pdeq_vf_ode_order(vf_id)
        self.order = self.num_tcoeffs_in_args
        self.num_params = lib.pdeq_vf_num_params(vf_id)
        self.fixed_dim = lib.pdeq_vf_dim(vf_id)
        self.tcoeff_indices_output = [self.order]
        self.is_jet_lifted = False
        if self.num_params == 0:
            if params is not None and np.size(params) != 0:
                raise ValueError(f"{name} takes no parameters.")
            self.params = None
        else:
            if params is None:
                raise ValueError(f"{name} needs {self.num_params} parameter(s).")
            p = params if isinstance(params, torch.Tensor) else np.asarray(params, dtype=np.float64)
            if p.shape[-1] != self.num_params or p.ndim > 2:
                raise ValueError(f"params must have shape ({self.num_params},) or (B, {self.num_params}).")
            self.params = p

    def __repr__(self):
        return f"VectorField({self.name!r}, num_tcoeffs_in_args={self.order})"

    def params_on_device(self, B: int):
        """-> (tensor or None, stride in doubles)."""
        if self.params is None:
            return None, 0
        p = _as_device_f64(self.params)
        if p.ndim == 1:
            return p, 0
        if p.shape[0] != B:
            raise ValueError(f"params has {p.shape[0]} rows for an ensemble of {B}.")
        return p, self.num_params


class Jacobian:
    """How a ts1 linearisation obtains the Jacobian (reference: `_probdiffeq/jacobians.py`). The device kernels always
    use the EXACT Jacobian of the registered right-hand side (its analytic `jac`, or dual numbers through `component`),
    which is what `jacobian_materialize()` asks for; the stochastic estimators are outside this path."""

    def __init__(self, kind: str):
        self.kind = kind

    def __repr__(self):
        return f"Jacobian({self.kind})"


def jacobian_materialize() -> Jacobian:
    """reference: `jacobians.jacobian_materialize` -- the full Jacobian, exactly."""
    return Jacobian("materialize")


def jacobian_monte_carlo_fwd(*args, **kwargs):
    raise NotImplementedError("Hutchinson-type Jacobian estimators are not part of the accelerated path; "
                              "use jacobian_materialize() (the kernels differentiate the right-hand side exactly).")  # fmt: skip


jacobian_monte_carlo_rev = jacobian_monte_carlo_fwd


def _check_jacobian(jacobian):
    if jacobian is not None and not (isinstance(jacobian, Jacobian) and jacobian.kind == "materialize"):
        raise NotImplementedError(f"jacobian={jacobian!r}: the accelerated path materialises the exact Jacobian only.")


def ode(name: str, /, *, params=None, jacobian: Jacobian | None = None) -> VectorField:
    """Select a registered vector field (reference: `probdiffeq.ode`, problems.py:283-295; a name of a device functor
    instead of a Python callable). `jacobian` may be `jacobian_materialize()` or None -- either way ts1 constraints
    use the exact Jacobian (the reference's default is a one-sample stochastic estimate)."""
    _check_jacobian(jacobian)
    return VectorField(name, params)


def ode_order_two(name: str, /, *, params=None, jacobian: Jacobian | None = None) -> VectorField:
    """reference: `probdiffeq.ode_order_two` (problems.py:298-312): u'' = f(u, u', t). The registered functor must be a
    second-order one (`vanderpol` among the built-ins); `ode(name)` accepts those as well."""
    _check_jacobian(jacobian)
    vf = VectorField(name, params)
    if vf.order != 2:
        raise ValueError(f"{name} is an order-{vf.order} right-hand side; ode_order_two needs a second-order one.")
    return vf


class TaylorPoint:
    """Where a constraint is linearised (reference: `_probdiffeq/taylor_points.py`)."""

    def __init__(self, kind: str):
        self.kind = kind

    def __repr__(self):
        return f"TaylorPoint({self.kind})"


def taylor_point_prior() -> TaylorPoint:
    """reference: `taylor_points.taylor_point_prior` -- linearise at the extrapolated mean (what the kernels do)."""
    return TaylorPoint("prior")


def taylor_point_maximum_a_posteriori(*args, **kwargs):
    raise NotImplementedError("MAP Taylor points (iterated linearisation) are not part of the accelerated path; "
                              "use taylor_point_prior().")  # fmt: skip


def system_matrices_1d_iwp(num_derivatives: int):
    """reference: `utilities.system_matrices_1d_iwp` (utilities.py:57-71): (A, Q) of the preconditioned once-integrated
    ... num_derivatives-times integrated Wiener process -- the constants every kernel launch carries in `pdeq_config`."""
    a, q, _facts = _iwp.system_matrices(int(num_derivatives))
    return a.copy(), q.copy()


def preconditioner_taylor(num_derivatives: int):
    """reference: `utilities.preconditioner_taylor` (utilities.py:74-84): dt -> (p, 1/p), p_i = dt^(nu-i) / (nu-i)!."""
    powers = np.arange(int(num_derivatives), -1.0, step=-1.0)
    scales = _iwp.factorial(powers)

    def precon(dt):
        return np.power(dt, powers) / scales, np.power(dt, -powers) * scales

    return precon


def _make_config(*, fact: str, nu: int, d: int, vf: VectorField, **kw) -> _lib.Config:
    cfg = _lib.Config()
    cfg.factorisation = _lib.FACT[fact]
    cfg.num_derivatives = nu
    cfg.ode_dim = d
    cfg.vf_id = vf.vf_id
    cfg.safety, cfg.factor_min, cfg.factor_max = 0.95, 0.2, 10.0
    cfg.exponent_integral, cfg.exponent_proportional = 0.3, 0.4
    cfg.correct_asymptotic_underconfidence = 1
    for k, v in kw.items():
        setattr(cfg, k, v)
    a, q, facts = _iwp.system_matrices(nu)
    n = nu + 1
    for i in range(n):
        for j in range(n):
            cfg.sys_a[i][j] = a[i, j]
            cfg.sys_q[i][j] = q[i, j]
    for k in range(n + 1):
        cfg.factorials[k] = facts[k]
        cfg.inv_factorials[k] = 1.0 / facts[k]
    if vf.order <= nu:
        for k, c in enumerate(_iwp.error_constants(nu, vf.order)):
            cfg.err_const[k] = c
    return cfg


def jetexpand_ode_padded_scan(*, num: int):
    """Taylor-mode initialisation on the device (reference: jet_expansion_algorithms.py:49-103).

    ``expand(vf, inits, t=t0)`` with ``inits`` a tuple of ``order`` arrays of shape (B, d) or (d,)
    returns ``(tcoeffs, {})`` with ``tcoeffs`` of shape (B, order + num, d).
    """

    def expand(vf: VectorField, inits, /, *, t: float = 0.0):
        if not isinstance(vf, VectorField):
            raise TypeError(vf)
        inits = [_as_device_f64(u) for u in inits]
        if len(inits) != vf.order:
            raise ValueError(f"{vf.name} is an order-{vf.order} ODE; got {len(inits)} initial values.")
        unbatched = inits[0].ndim == 1
        inits = [u.reshape(1, -1) if u.ndim == 1 else u for u in inits]
        u0 = torch.stack(inits, dim=1).contiguous()  # (B, order, d)
        B, q, d = u0.shape
        if num == 0 or B == 0:
            out = u0 if num == 0 else torch.empty((0, q + num, d), dtype=torch.float64, device=u0.device)
        else:
            n = q + num
            cfg = _make_config(fact="isotropic", nu=n - 1, d=d, vf=vf)
            out = torch.empty((B, n, d), dtype=torch.float64, device=u0.device)
            params, stride = vf.params_on_device(B)
            rc = _lib.load().pdeq_taylor_init(
                C.byref(cfg), B, _ptr(u0), _ptr(params), stride, float(t), _ptr(out), _stream()
            )
            _lib.check(rc, "pdeq_taylor_init")
        return (out[0] if unbatched else out), {}

    return expand


jetexpand_ode_unroll = jetexpand_ode_padded_scan  # same output (jet_expansion_algorithms.py:110-152)
jetexpand_ode_via_jvp = jetexpand_ode_padded_scan  # same output by recursive forward mode (:178-213)


def jetexpand_ode_coefficient_increment(*, num_arguments: int):
    """reference: jet_expansion_algorithms.py:155-177 -- ``increment(vf, taylor_coeffs, t=t)`` returns the Taylor
    series with one more coefficient. The coefficients are functions of the first ``num_arguments`` entries alone,
    so the device routine recomputes the series to the new length from those (one pass per coefficient)."""

    def increment(vf: VectorField, taylor_coeffs, *, t: float = 0.0):
        if not isinstance(vf, VectorField):
            raise TypeError(vf)
        if num_arguments != vf.order:
            raise ValueError(f"{vf.name} takes {vf.order} Taylor coefficient(s) as arguments, not {num_arguments}.")
        tc = _as_device_f64(taylor_coeffs)
        k = tc.shape[-2]
        if k < num_arguments:
            raise ValueError("fewer Taylor coefficients than the vector field has arguments")
        inits = tuple(tc[..., i, :] for i in range(num_arguments))
        out, _ = jetexpand_ode_padded_scan(num=k + 1 - num_arguments)(vf, inits, t=t)
        return out

    return increment


# ------------------------------------------------------------------------------------------------------
# State-space models, priors, constraints
# ------------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class WienerIntegratedPrior:
    """Integrated Wiener process prior over an ensemble (reference: `*WienerIntegrated`)."""

    factorisation: str
    tcoeffs: torch.Tensor  # (B, n, d)
    init_std: torch.Tensor | None  # isotropic (B|1, n); else (B|1, n, d); None = exact
    output_scale: torch.Tensor | None  # isotropic (B|1, 1); else (B|1, d); None = ones
    unbatched: bool

    @property
    def num_derivatives(self) -> int:
        return self.tcoeffs.shape[1] - 1

    @property
    def ode_dim(self) -> int:
        return self.tcoeffs.shape[2]


def _check_taylor_point(taylor_point):
    if taylor_point is not None and not (isinstance(taylor_point, TaylorPoint) and taylor_point.kind == "prior"):
        raise NotImplementedError(f"taylor_point={taylor_point!r}: the accelerated path linearises at the prior mean only.")


class _Constraint:
    def __init__(self, kind: str, vf: VectorField, factorisation: str):
        if not isinstance(vf, VectorField):
            raise TypeError(vf)
        self.kind = kind
        self.ode = vf
        self.factorisation = factorisation
        self.residual_order = vf.order + 1

    def __repr__(self):
        return f"Constraint({self.kind}, {self.ode!r}, {self.factorisation})"


class _StateSpaceModel:
    def __init__(self, factorisation: str):
        self.factorisation = factorisation

    def __repr__(self):
        return f"state_space_model_{self.factorisation}()"

    def prior_wiener_integrated(
        self, tcoeffs, /, *, is_exact=True, inexact_eps: float = 1e-6, output_scale=None
    ) -> WienerIntegratedPrior:
        """reference: ssm_impl_isotropic.py:409-482, ssm_impl_blockdiag.py:463-547, ssm_impl_dense.py:458-514."""
        tc = _as_device_f64(tcoeffs)
        unbatched = tc.ndim == 2
        if unbatched:
            tc = tc[None]
        if tc.ndim != 3:
            raise ValueError("tcoeffs must have shape (B, n, d) or (n, d).")
        B, n, d = tc.shape
        std_shape = (n,) if self.factorisation == "isotropic" else (n, d)
        if isinstance(is_exact, bool):
            std = None if is_exact else torch.full((1, *std_shape), inexact_eps, dtype=torch.float64, device=tc.device)
        else:
            mask = np.broadcast_to(np.asarray(is_exact, dtype=bool), std_shape)
            std = _as_device_f64(np.where(mask, 0.0, inexact_eps)[None])
        return self._prior(tc, std, output_scale, unbatched)

    def prior_wiener_integrated_diffuse(self, tcoeffs, tcoeffs_std, /, *, output_scale=None):
        tc = _as_device_f64(tcoeffs)
        unbatched = tc.ndim == 2
        if unbatched:
            tc = tc[None]
        B, n, d = tc.shape
        std_shape = (n,) if self.factorisation == "isotropic" else (n, d)
        std = _as_device_f64(tcoeffs_std)
        if tuple(std.shape) == std_shape:
            std = std[None]
        if tuple(std.shape[1:]) != std_shape or std.shape[0] not in (1, B):
            raise ValueError(f"tcoeffs_std must have shape {std_shape} or (B, {std_shape}).")
        return self._prior(tc, std, output_scale, unbatched)

    def _prior(self, tc, std, output_scale, unbatched):
        B, n, d = tc.shape
        k = 1 if self.factorisation == "isotropic" else d
        if output_scale is not None:
            os_ = _as_device_f64(output_scale)
            if self.factorisation == "isotropic":
                if os_.ndim not in (0, 1):
                    raise ValueError("The base-scale has the wrong shape. Expected: ().")
                os_ = os_.reshape(-1, 1)
            else:
                if os_.shape[-1] != d or os_.ndim > 2:
                    raise ValueError(f"The output-scale has the wrong shape. Expected: ({d},).")
                os_ = os_.reshape(-1, d)
            if os_.shape[0] not in (1, B):
                raise ValueError("output_scale batch axis does not match the ensemble.")
            output_scale = os_.contiguous()
        del k
        return WienerIntegratedPrior(self.factorisation, tc, std, output_scale, unbatched)

    def constraint_ode_ts0(self, vf: VectorField, /) -> _Constraint:
        """reference: ssm_impl_api.py:521-534."""
        return _Constraint("ts0", vf, self.factorisation)

    def constraint_ode_ts1(self, vf: VectorField, /, taylor_point: TaylorPoint | None = None) -> _Constraint:
        """reference: ssm_impl_api.py:536-560 (`taylor_point`: None or `taylor_point_prior()`)."""
        _check_taylor_point(taylor_point)
        return _Constraint("ts1", vf, self.factorisation)


def state_space_model_isotropic() -> _StateSpaceModel:
    return _StateSpaceModel("isotropic")


def state_space_model_blockdiag() -> _StateSpaceModel:
    return _StateSpaceModel("blockdiag")


def state_space_model_dense() -> _StateSpaceModel:
    return _StateSpaceModel("dense")


# ------------------------------------------------------------------------------------------------------
# Strategies, solvers, error estimators (descriptors)
# ------------------------------------------------------------------------------------------------------


class strategy_filter:
    """reference: estimators_and_losses.py:347-421."""

    kind = "filter"
    is_suitable_for_save_at = True
    is_suitable_for_save_every_step = True

    def __repr__(self):
        return "strategy_filter()"


class strategy_smoother_fixedpoint:
    """reference: estimators_and_losses.py:473-591."""

    kind = "fixedpoint"
    is_suitable_for_save_at = True
    is_suitable_for_save_every_step = False

    def __repr__(self):
        return "strategy_smoother_fixedpoint()"


class strategy_smoother_fixedinterval:
    """reference: estimators_and_losses.py:594-717. Accelerated for `solve_fixed_grid` (its intended use).

    `terminal="reference"` (default) reproduces the reference literally: `Smoother.finalize` (:437-470) treats the
    last grid state as an overstepped state and marginalises it through its own backward conditional first, so on
    a fixed grid the returned marginals are one interval late at the terminal point. `terminal="aligned"` starts
    the backward recursion from the filtering marginal at the last grid point (the Rauch-Tung-Striebel pass).
    """

    is_suitable_for_save_at = False
    is_suitable_for_save_every_step = True

    def __init__(self, *, terminal="reference"):
        if terminal not in ("reference", "aligned"):
            raise ValueError(f"terminal must be 'reference' or 'aligned', got {terminal!r}")
        self.terminal = terminal
        self.kind = "fixedinterval" if terminal == "reference" else "fixedinterval_aligned"

    def __repr__(self):
        return f"strategy_smoother_fixedinterval(terminal={self.terminal!r})"


class _Solver:
    kind = ""

    def __init__(self, *, strategy, constraint, constraint_init=None, **options):
        if not isinstance(constraint, _Constraint):
            raise TypeError(constraint)
        if constraint_init is not None:
            # reference: solvers.py:361-372, 526-537, 670-680 -- one Bayes update with `constraint_init` at t0. The
            # kernels run it with the step constraint's own linearisation, so the two must describe the same thing.
            if not isinstance(constraint_init, _Constraint):
                raise TypeError(constraint_init)
            c0, c1 = constraint_init, constraint
            if c0 is not c1 and not (c0.ode is c1.ode and (c0.kind, c0.factorisation) == (c1.kind, c1.factorisation)):
                raise NotImplementedError("constraint_init must be the solver's own constraint on the accelerated path.")
            options = dict(options, constraint_init=1)
        self.strategy = strategy
        self.constraint = constraint
        self.constraint_init = constraint_init
        self.options = options

    @property
    def is_suitable_for_save_at(self):
        return self.strategy.is_suitable_for_save_at

    @property
    def is_suitable_for_save_every_step(self):
        return self.strategy.is_suitable_for_save_every_step

    def __repr__(self):
        return f"{self.kind}(strategy={self.strategy!r}, constraint={self.constraint!r})"

    def offgrid_marginals(self, t, *, solution):
        """Dense output (reference: solvers.py:149-203): marginals at times strictly inside the grid and not on it.

        ``t`` is a scalar or a (Q,) array shared by the ensemble; ``solution`` comes from `solve_adaptive_save_at`
        or `solve_fixed_grid` (grid shared by the ensemble). Returns a `Normal` with mean (B, Q, n, d) -- the Q axis
        is dropped for a scalar ``t``, the B axis for an unbatched solve. Filters extrapolate from the grid point to
        the left; the fixed-interval smoother also conditions on the right one. The fixed-point smoother raises
        NotImplementedError like the reference (estimators_and_losses.py:519-523)."""
        kind = self.strategy.kind
        if kind == "fixedpoint":
            raise NotImplementedError
        prior = getattr(solution, "prior", None)
        if prior is None:
            raise ValueError("solution carries no prior; pass the object a solve_* call returned")
        mean, chol, scale, grid = solution.u.mean_flat, solution.u.cholesky_flat, solution.output_scale, solution.t
        if chol is None:
            raise ValueError("offgrid_marginals needs the Cholesky factors (want_cholesky=True)")
        unbatched = mean.ndim == 3
        if unbatched:
            mean, chol, scale, grid = mean[None], chol[None], scale[None], grid[None]
        B, T, n, d = mean.shape
        fact = solution.u.factorisation
        filt_mean = filt_chol = None
        if kind != "filter":
            full = solution.solution_full
            if full is None or full.filtering is None:
                raise ValueError("the smoothing solution carries no filtering marginals (want_posterior=True)")
            filt_mean, filt_chol = full.filtering.mean_flat, full.filtering.cholesky_flat
            if unbatched:
                filt_mean, filt_chol = filt_mean[None], filt_chol[None]
        scalar_t = (t.ndim if isinstance(t, torch.Tensor) else np.ndim(t)) == 0
        tq = _as_device_f64(t).reshape(-1).contiguous()
        Q = tq.shape[0]
        g0 = grid[0].contiguous() if B > 0 else torch.zeros((T,), dtype=torch.float64, device=mean.device)
        cfg = _make_config(fact=fact, nu=n - 1, d=d, vf=self.constraint.ode, strategy=_lib.STRATEGY[kind])
        out_mean = torch.empty((B, Q, n, d), dtype=torch.float64, device=mean.device)
        out_chol = torch.empty((B, Q, *chol.shape[2:]), dtype=torch.float64, device=mean.device)
        ps = prior.output_scale
        ps_stride = 0 if ps is None or ps.shape[0] == 1 else ps.shape[1]
        if B > 0:
            # contiguous copies are bound to names: a temporary would be freed (and its block reused by the next
            # temporary) before the launch that reads it
            mean, chol, scale = mean.contiguous(), chol.contiguous(), scale.contiguous()
            filt_mean = None if filt_mean is None else filt_mean.contiguous()
            filt_chol = None if filt_chol is None else filt_chol.contiguous()
            rc = _lib.load().pdeq_offgrid_marginals(
                C.byref(cfg), B, T, _ptr(g0), Q, _ptr(tq), _ptr(mean), _ptr(chol), _ptr(filt_mean), _ptr(filt_chol),
                _ptr(scale), _ptr(ps), ps_stride, _ptr(out_mean), _ptr(out_chol), _stream(),
            )  # fmt: skip
            _lib.check(rc, "pdeq_offgrid_marginals")
        if scalar_t:
            out_mean, out_chol = out_mean[:, 0], out_chol[:, 0]
        if unbatched:
            out_mean, out_chol = out_mean[0], out_chol[0]
        return Normal(fact, out_mean, out_chol)


class solver(_Solver):
    """Uncalibrated solver (reference: solvers.py:636-767)."""

    kind = "solver"

    def __init__(self, *, strategy, constraint, constraint_init=None):
        super().__init__(strategy=strategy, constraint=constraint, constraint_init=constraint_init)


class solver_mle(_Solver):
    """Running maximum-likelihood calibration (reference: solvers.py:318-480)."""

    kind = "solver_mle"

    def __init__(self, *, strategy, constraint, constraint_init=None, correct_asymptotic_underconfidence=True):
        super().__init__(
            strategy=strategy,
            constraint=constraint,
            constraint_init=constraint_init,
            correct_asymptotic_underconfidence=int(bool(correct_asymptotic_underconfidence)),
        )


class solver_dynamic(_Solver):
    """Per-step calibration (reference: solvers.py:483-633)."""

    kind = "solver_dynamic"

    def __init__(self, *, strategy, constraint, constraint_init=None, re_linearize_after_calibration=False):
        # reference: solvers.py:578-582. The re-linearisation happens at the MEAN of the re-extrapolated state
        # (taylor_point_prior, taylor_points.py:150-156), and the calibrated output scale changes the covariance of
        # that state, not its mean: for the ts0 / ts1 constraints built here the second linearisation is the first
        # one. The kernels therefore run the same step for both values of the flag. (In the reference the isotropic
        # model's two extrapolation code paths differ by an ulp in the mean, which the adaptive loop turns into
        # ~1e-10 differences; block-diagonal and dense are bitwise identical -- tests/test_oracle_kats.py.)
        self.re_linearize_after_calibration = bool(re_linearize_after_calibration)
        super().__init__(
            strategy=strategy,
            constraint=constraint,
            constraint_init=constraint_init,
            re_linearize_after_calibration=int(bool(re_linearize_after_calibration)),  # carried in pdeq_config
        )


def error_norm_scale_then_rms(*, norm_order=None):
    if norm_order is not None:
        raise NotImplementedError("only the 2-norm is accelerated")
    return "scale_then_rms"


def error_norm_rms_then_scale(norm_order=None):
    if norm_order is not None:
        raise NotImplementedError("only the 2-norm is accelerated")
    return "rms_then_scale"


class _ErrorEstimator:
    kind = ""

    def __init__(self, *, constraint, error_norm=None, re_linearize_before_error=False, derivative_idx=0,
                 error_per_unit_step=False):  # fmt: skip
        # reference: solvers.py:946-952, 1061-1067. Re-linearising at the zero-error extrapolation of the previous mean
        # reproduces the step's cached linearisation exactly (same point, same time) for the ts0 / ts1 constraints
        # built here, in every factorisation (bitwise in the oracle, tests/test_oracle_kats.py): the flag is accepted
        # and changes nothing.
        self.re_linearize_before_error = bool(re_linearize_before_error)
        self.constraint = constraint
        self.error_norm = "scale_then_rms" if error_norm is None else error_norm
        if self.error_norm not in _lib.NORM:
            raise ValueError(f"unknown error norm {error_norm!r}")
        self.derivative_idx = int(derivative_idx)
        self.error_per_unit_step = bool(error_per_unit_step)


class error_residual_std(_ErrorEstimator):
    """reference: solvers.py:850-996."""

    kind = "residual_std"

    def __init__(self, *, constraint, error_norm=None, re_linearize_before_error=False, error_per_unit_step=False):
        super().__init__(constraint=constraint, error_norm=error_norm,
                         re_linearize_before_error=re_linearize_before_error,
                         error_per_unit_step=error_per_unit_step)  # fmt: skip


class error_state_std(_ErrorEstimator):
    """reference: solvers.py:999-1098."""

    kind = "state_std"


# ------------------------------------------------------------------------------------------------------
# Solutions
# ------------------------------------------------------------------------------------------------------


class _CoefficientList:
    """`rv.mean` / `rv.std` of the reference: a list over Taylor coefficients."""

    def __init__(self, flat: torch.Tensor, axis: int):
        self.flat = flat
        self.axis = axis

    def __len__(self):
        return self.flat.shape[self.axis]

    def __getitem__(self, k):
        return self.flat.select(self.axis, k)

    def __iter__(self):
        return (self[k] for k in range(len(self)))


@dataclasses.dataclass
class Normal:
    """Marginal normal distributions (reference: `IsotropicNormal` / `BlockDiagNormal` / `DenseNormal`)."""

    factorisation: str
    mean_flat: torch.Tensor  # (..., n, d)
    cholesky_flat: torch.Tensor | None  # isotropic (..., n, n); blockdiag (..., d, n, n)

    @property
    def mean(self):
        return _CoefficientList(self.mean_flat, self.mean_flat.ndim - 2)

    @property
    def std_flat(self):
        """isotropic (..., n); blockdiag and dense (..., n, d)."""
        if self.cholesky_flat is None:
            raise ValueError("Cholesky factors were not requested.")
        s = torch.linalg.vector_norm(self.cholesky_flat, dim=-1)
        if self.factorisation == "isotropic":
            return s
        if self.factorisation == "dense":  # rows of the (nd, nd) factor, coefficient-major
            return s.reshape(self.mean_flat.shape)
        return s.transpose(-1, -2)

    @property
    def std(self):
        s = self.std_flat
        return _CoefficientList(s, s.ndim - 1 if self.factorisation == "isotropic" else s.ndim - 2)

    def covariance(self):
        L = self.cholesky_flat
        return L @ L.transpose(-1, -2)


@dataclasses.dataclass
class ProbabilisticSolution:
    """reference: `ProbabilisticSolution` (solvers.py:33-69), with a leading ensemble axis."""

    t: torch.Tensor
    u: Normal
    output_scale: torch.Tensor
    num_steps: torch.Tensor
    num_attempts: torch.Tensor
    status: torch.Tensor
    solution_full: Any = None
    trace: Any = None
    prior: Any = None

    def _index(self, fn):
        return ProbabilisticSolution(
            t=fn(self.t),
            u=Normal(self.u.factorisation, fn(self.u.mean_flat),
                     None if self.u.cholesky_flat is None else fn(self.u.cholesky_flat)),
            output_scale=fn(self.output_scale),
            num_steps=fn(self.num_steps),
            num_attempts=self.num_attempts,
            status=self.status,
            solution_full=self.solution_full,
            prior=self.prior,
        )  # fmt: skip


@dataclasses.dataclass
class BackwardConditional:
    """The backward conditionals of a smoothing posterior, calibrated and in natural coordinates
    (reference: `LatentCond.preconditioner_apply`): entry k maps grid point k to k - 1,
    ``x[k-1] | x[k] ~ N(gain[k] x[k] + mean[k], cholesky[k] cholesky[k]^T)``; entry 0 is unused (zeros)."""

    gain: torch.Tensor  # isotropic (..., T, n, n); blockdiag (..., T, d, n, n)
    mean: torch.Tensor  # (..., T, n, d)
    cholesky: torch.Tensor  # like gain


@dataclasses.dataclass
class MarkovSequence:
    """reference: `MarkovSequence` (estimators_and_losses.py:121-231) with ``reverse=True``: the terminal marginal
    plus one backward conditional per grid interval."""

    marginal: Normal  # at the last grid point
    conditional: BackwardConditional
    reverse: bool = True

    def sample(self, key=None, *, shape: tuple = (), base=None):
        """Joint posterior samples at all grid points (reference: `MarkovSequence.sample`, :233-271).

        ``key`` seeds a `torch.Generator` on the device (an int, or a Generator) -- the reference's JAX PRNG stream
        is not reproduced, the arithmetic given the normal draws is. Alternatively pass the draws as ``base``:
        (B, *shape, T, n) for the isotropic model (one draw per coefficient shared by all dimensions, like
        `IsotropicNormal.sample_flat`), (B, *shape, T, d, n) for the block-diagonal one, (B, *shape, T, n d) for the
        dense one (coefficient-major, like the flat state). Returns a list over Taylor
        coefficients of tensors (B, *shape, T, d); the B axis is absent for an unbatched posterior."""
        mean, chol = self.marginal.mean_flat, self.marginal.cholesky_flat
        gain, cmean, cchol = self.conditional.gain, self.conditional.mean, self.conditional.cholesky
        unbatched = mean.ndim == 2
        if unbatched:
            mean, chol, gain, cmean, cchol = mean[None], chol[None], gain[None], cmean[None], cchol[None]
        B, n, d = mean.shape
        T = cmean.shape[1]
        fact = self.marginal.factorisation
        core = {"isotropic": (T, n), "blockdiag": (T, d, n)}.get(fact, (T, n * d))  # dense: one draw per state entry
        if base is None:
            gen = key if isinstance(key, torch.Generator) else torch.Generator(device=mean.device)
            if not isinstance(key, torch.Generator):
                gen.manual_seed(0 if key is None else int(key))
            base_t = torch.randn((B, *shape, *core), dtype=torch.float64, device=mean.device, generator=gen)
        else:
            base_t = _as_device_f64(base)
            if unbatched:
                base_t = base_t[None]
            shape = tuple(base_t.shape[1 : base_t.ndim - len(core)])
            if tuple(base_t.shape) != (B, *shape, *core):
                raise ValueError(f"base must have shape (B, *shape, {core}); got {tuple(base_t.shape)}")
        S = int(np.prod(shape)) if len(shape) > 0 else 1
        mean_t = torch.zeros((B, T, n, d), dtype=torch.float64, device=mean.device)
        mean_t[:, -1] = mean
        chol_t = torch.zeros((B, T, *chol.shape[1:]), dtype=torch.float64, device=mean.device)
        chol_t[:, -1] = chol
        out = torch.empty((B, S, T, n, d), dtype=torch.float64, device=mean.device)
        cfg = _make_config(fact=fact, nu=n - 1, d=d, vf=VectorField("linear", params=[1.0]))
        if B > 0:
            gain, cmean, cchol = gain.contiguous(), cmean.contiguous(), cchol.contiguous()  # named: see offgrid_marginals
            base_c = base_t.reshape(B, S, *core).contiguous()
            rc = _lib.load().pdeq_sample_posterior(
                C.byref(cfg), B, T, S, _ptr(mean_t), _ptr(chol_t), _ptr(gain), _ptr(cmean), _ptr(cchol), _ptr(base_c),
                _ptr(out), _stream(),
            )  # fmt: skip
            _lib.check(rc, "pdeq_sample_posterior")
        out = out.reshape(B, *shape, T, n, d)
        if unbatched:
            out = out[0]
        return _CoefficientList(out, out.ndim - 2)


@dataclasses.dataclass
class SmoothingSolution:
    """reference: `SmoothingSolution` (estimators_and_losses.py:107-118)."""

    posterior: MarkovSequence
    filtering: Any = None


# ------------------------------------------------------------------------------------------------------
# Log-marginal likelihoods
# ------------------------------------------------------------------------------------------------------


def loss_lml_timeseries(*, average_pdfs: bool = True, tcoeff_index: int = 0):
    """reference: estimators_and_losses.py:53-105 (+ `MarkovSequence.evaluate_lml` :180-218).

    ``loss(u, posterior=sol.solution_full.posterior, std=...)`` with data ``u`` (T, d) -- or (B, T, d), one series per
    ensemble member -- and ``std`` (T,) for the isotropic model, (T, d) for the block-diagonal one (optionally with
    a leading ensemble axis). Returns one value per ensemble member, shape (B,)."""

    def loss(u, /, *, posterior, std):
        if not isinstance(posterior, MarkovSequence):
            msg = "The datatype of the posterior is not as expected."
            msg += f" Expected: {MarkovSequence}."
            msg += f" Received: {type(posterior)}."
            msg += " Did you perhaps use a filter instead of a smoother"
            msg += ", forget to extract the posterior from the smoothing-solution"
            msg += ", or mean to use a different loss?"
            raise TypeError(msg)
        marg, cond = posterior.marginal, posterior.conditional
        fact = marg.factorisation
        mean, chol = marg.mean_flat, marg.cholesky_flat
        gain, cmean, cchol = cond.gain, cond.mean, cond.cholesky
        unbatched = mean.ndim == 2
        if unbatched:
            mean, chol, gain, cmean, cchol = mean[None], chol[None], gain[None], cmean[None], cchol[None]
        B, n, d = mean.shape
        T = cmean.shape[1]
        data = _as_device_f64(u)
        sd = _as_device_f64(std)
        data_b = data if data.ndim == 3 else data[None]
        std_core = (T,) if fact == "isotropic" else (T, d)
        sd_b = sd if sd.ndim == len(std_core) + 1 else sd[None]
        msg = "The standard deviation container differs from what was expected."
        msg += f" Expected: shape={std_core}. Received: shape={tuple(sd.shape)}."
        msg += f" For reference: data-shape={tuple(data.shape)}."
        if tuple(sd_b.shape[1:]) != std_core or tuple(data_b.shape[1:]) != (T, d):
            raise ValueError(msg)
        for name, arr in (("data", data_b), ("std", sd_b)):
            if arr.shape[0] not in (1, B):
                raise ValueError(f"{name} batch axis does not match the ensemble.")
        # the kernel reads the terminal marginal as entry T - 1 of (B, T, ...) arrays
        mean_t = torch.zeros((B, T, n, d), dtype=torch.float64, device=mean.device)
        mean_t[:, -1] = mean
        chol_t = torch.zeros((B, T, *chol.shape[1:]), dtype=torch.float64, device=mean.device)
        chol_t[:, -1] = chol
        cfg = _make_config(fact=fact, nu=n - 1, d=d, vf=VectorField("linear", params=[1.0]))
        out = torch.empty((B,), dtype=torch.float64, device=mean.device)
        ws = torch.empty((max(B * d, 1),), dtype=torch.float64, device=mean.device)
        data_b, sd_b = data_b.contiguous(), sd_b.contiguous()
        gain, cmean, cchol = gain.contiguous(), cmean.contiguous(), cchol.contiguous()  # named: see offgrid_marginals
        rc = _lib.load().pdeq_lml_timeseries(
            C.byref(cfg), B, T, int(tcoeff_index), int(bool(average_pdfs)), _ptr(mean_t), _ptr(chol_t),
            _ptr(gain), _ptr(cmean), _ptr(cchol),
            _ptr(data_b), 0 if data_b.shape[0] == 1 else T * d, _ptr(sd_b),
            0 if sd_b.shape[0] == 1 else int(np.prod(std_core)), _ptr(out), _ptr(ws), ws.numel() * 8, _stream(),
        )  # fmt: skip
        _lib.check(rc, "pdeq_lml_timeseries")
        return out[0] if unbatched else out

    return loss



def loss_lml_terminal_values(*, tcoeff_index: int = 0):
    """reference: estimators_and_losses.py:20-50. Returns one log-pdf per ensemble member, shape (B,)."""

    def loss(u, /, *, marginals: Normal, std, vf: VectorField | None = None):
        mean = marginals.mean_flat
        chol = marginals.cholesky_flat
        if chol is None:
            raise ValueError("marginals carry no Cholesky factors")
        unbatched = mean.ndim == 2
        if unbatched:
            mean, chol = mean[None], chol[None]
        B, n, d = mean.shape
        data = _as_device_f64(u).reshape(-1, d)
        k = 1 if marginals.factorisation == "isotropic" else d
        sd = _as_device_f64(std).reshape(-1, k)
        for name, arr in (("data", data), ("std", sd)):
            if arr.shape[0] not in (1, B):
                raise ValueError(f"{name} batch axis does not match the ensemble.")
        vf_ = vf if vf is not None else VectorField("linear", params=[1.0])
        cfg = _make_config(fact=marginals.factorisation, nu=n - 1, d=d, vf=vf_)
        out = torch.empty((B,), dtype=torch.float64, device=mean.device)
        mean, chol = mean.contiguous(), chol.contiguous()  # named: a temporary's block would be reused before the launch
        rc = _lib.load().pdeq_lml_terminal_values(
            C.byref(cfg), B, int(tcoeff_index), _ptr(mean), _ptr(chol),
            _ptr(data), 0 if data.shape[0] == 1 else d, _ptr(sd), 0 if sd.shape[0] == 1 else k,
            _ptr(out), _stream(),
        )  # fmt: skip
        _lib.check(rc, "pdeq_lml_terminal_values")
        return out[0] if unbatched else out

    return loss
