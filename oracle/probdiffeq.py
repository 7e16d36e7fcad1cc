"""Probabilistic solvers, strategies, calibration and error estimators of the oracle.

Test infrastructure (see oracle/__init__.py).  Restates

* _probdiffeq/solvers.py:33-69 (ProbabilisticSolution), :205-315 (interpolation glue),
  :318-480 (solver_mle), :483-633 (solver_dynamic), :636-767 (solver),
  :770-811 (error norms), :850-996 (error_residual_std), :999-1098 (error_state_std);
* _probdiffeq/estimators_and_losses.py:20-50 (loss_lml_terminal_values), :123-178 (MarkovSequence),
  :347-421 (strategy_filter), :424-470 (Smoother), :473-591 (fixed-point), :594-717 (fixed-interval).

The public names follow the reference's `probdiffeq.probdiffeq` namespace so that tests read alike.
"""

import numpy as np

from oracle import linalg, problems, ssm

__all__ = [
    "error_norm_rms_then_scale",
    "error_norm_scale_then_rms",
    "error_residual_std",
    "error_state_std",
    "jetexpand_ode_padded_scan",
    "jetexpand_ode_unroll",
    "loss_lml_terminal_values",
    "ode",
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

ode = problems.ode


def jetexpand_ode_padded_scan(*, num):
    """jet_expansion_algorithms.py:49-103. Returns (tcoeffs (q+num, d), {})."""

    def expand(vf, inits, /, *, t):
        return vf.taylor_coefficients(inits, t, num), {}

    return expand


jetexpand_ode_unroll = jetexpand_ode_padded_scan  # :110-152, same output


# ------------------------------------------------------------------------------------------------
# State-space model factories
# ------------------------------------------------------------------------------------------------


class Constraint:
    """ts0 / ts1 linearisation. ssm_impl_api.py:138-179."""

    def __init__(self, kind, vf):
        self.kind = kind
        self.ode = vf
        self.residual_order = vf.order + 1

    def init_linearization(self):
        return None

    def linearize(self, rv, state, *, damp, t):
        fn = rv.alg.linearize_ts0 if self.kind == "ts0" else rv.alg.linearize_ts1
        return fn(self.ode, rv, damp, t), state


class _StateSpaceModel:
    def __init__(self, kind):
        self.kind = kind

    def _alg(self, d):
        if self.kind == "isotropic":
            return ssm.Isotropic()
        if self.kind == "blockdiag":
            return ssm.BlockDiag()
        return ssm.Dense(d)

    def prior_wiener_integrated(self, tcoeffs, /, *, is_exact=True, inexact_eps=1e-6, output_scale=None):
        """ssm_impl_{isotropic:409-482, blockdiag:463-547, dense:458-514}.py."""
        tcoeffs = np.asarray(tcoeffs, dtype=np.float64)
        if tcoeffs.ndim == 1:
            tcoeffs = tcoeffs[:, None]
        n, d = tcoeffs.shape
        alg = self._alg(d)
        shape = (n,) if self.kind == "isotropic" else (n, d)
        if isinstance(is_exact, bool):
            std = np.zeros(shape) if is_exact else inexact_eps * np.ones(shape)
        else:
            std = np.where(np.broadcast_to(np.asarray(is_exact, dtype=bool), shape), 0.0, inexact_eps)
        return alg.prior_wiener_integrated(tcoeffs, std, output_scale)

    def prior_wiener_integrated_diffuse(self, tcoeffs, tcoeffs_std, /, *, output_scale=None):
        tcoeffs = np.asarray(tcoeffs, dtype=np.float64)
        if tcoeffs.ndim == 1:
            tcoeffs = tcoeffs[:, None]
        return self._alg(tcoeffs.shape[1]).prior_wiener_integrated(tcoeffs, tcoeffs_std, output_scale)

    def constraint_ode_ts0(self, vf, /):
        return Constraint("ts0", vf)

    def constraint_ode_ts1(self, vf, /):
        return Constraint("ts1", vf)


def state_space_model_isotropic():
    return _StateSpaceModel("isotropic")


def state_space_model_blockdiag():
    return _StateSpaceModel("blockdiag")


def state_space_model_dense():
    return _StateSpaceModel("dense")


# ------------------------------------------------------------------------------------------------
# Markov sequences and strategies
# ------------------------------------------------------------------------------------------------


class MarkovSequence:
    """estimators_and_losses.py:121-231. marginal: Normal or list; conditional: Cond or list."""

    def __init__(self, marginal, conditional, reverse=True):
        self.marginal = marginal
        self.conditional = conditional
        self.reverse = reverse

    def rescale_cholesky(self, factor):
        marg = _map(lambda m: m.rescale_cholesky(factor), self.marginal)
        cond = _map(lambda c: c.rescale_noise(factor), self.conditional)
        return MarkovSequence(marg, cond, self.reverse)

    def evaluate_marginals(self):
        """Backward (reverse=True) marginalisation over the stacked conditionals. :156-178."""
        assert self.reverse
        marginal = self.marginal[-1] if isinstance(self.marginal, list) else self.marginal
        out = [marginal]
        x = marginal
        for cond in reversed(self.conditional):
            x = cond.marginalise(x)
            out.insert(0, x)
        return out


    def remove_filtering_distributions(self):
        """:220-231: keep the terminal marginal only."""
        if isinstance(self.marginal, list):
            return MarkovSequence(self.marginal[-1], self.conditional, self.reverse)
        return self

    def sample(self, base):
        """:233-271 with the random draws made explicit: `base[k]` are the standard-normal numbers `sample_flat`
        would draw at grid point k -- shape (n,) isotropic (ssm_impl_isotropic.py:255-258: ONE draw per coefficient,
        shared by all dimensions), (d, n) block-diagonal (ssm_impl_blockdiag.py:343-351), (n d,) dense.
        Returns the sampled states, one per grid point, in the layout of `Normal.mean`."""
        seq = self.remove_filtering_distributions()
        assert seq.reverse

        def draw(rv, eps):
            if rv.alg.name == "isotropic":
                return rv.mean + (rv.chol @ eps)[:, None]
            if rv.alg.name == "blockdiag":
                return rv.mean + np.einsum("djk,dk->dj", rv.chol, eps)
            return rv.mean + rv.chol @ eps

        x = draw(seq.marginal, base[-1])
        out = [x]
        for k in range(len(seq.conditional) - 1, -1, -1):
            x = draw(seq.conditional[k].apply_flat(x), base[k])
            out.insert(0, x)
        return out

    def evaluate_lml(self, u, *, model, average_pdfs, solve=None):
        """:180-218: backward scan -- observe the terminal state, then alternately step back through a
        conditional and observe. `u[k]`, `model[k]` belong to grid point k; conditional[k-1] maps k -> k-1."""
        assert self.reverse and not isinstance(self.marginal, list)
        pdf, rv = model[-1].bayes_rule_and_logpdf(u[-1], self.marginal, solve=solve)
        num = 1
        for k in range(len(self.conditional) - 1, -1, -1):
            predicted = self.conditional[k].marginalise(rv)
            pdf_n, rv = model[k].bayes_rule_and_logpdf(u[k], predicted, solve=solve)
            pdf = (pdf * num + pdf_n) / (num + 1) if average_pdfs else pdf + pdf_n
            num += 1
        return pdf


def _map(fn, x):
    return [fn(s) for s in x] if isinstance(x, list) else fn(x)


class InterpResult:
    """_probdiffeq/utilities.py:21-54."""

    def __init__(self, step_from, interp_from):
        self.step_from = step_from
        self.interp_from = interp_from


class strategy_filter:
    """estimators_and_losses.py:347-421."""

    is_suitable_for_save_at = True
    is_suitable_for_save_every_step = True
    is_suitable_for_offgrid_marginals = True

    def init_posterior(self, *, u):
        return u, u

    def predict(self, posterior, *, transition):
        m = transition.marginalise(posterior)
        return m, m

    def apply_updates(self, prediction, *, updates):
        return updates, updates

    def finalize(self, *, posterior0, posterior, posterior1, output_scale):
        marginals = [posterior0.rescale_cholesky(output_scale)]
        marginals += [p.rescale_cholesky(output_scale) for p in posterior]
        return marginals, marginals

    def interpolate_fwd(self, *, posterior_t0, posterior_t1, transition_t0_t, transition_t_t1):
        _, interpolated = self.predict(posterior_t0, transition=transition_t0_t)
        return (interpolated, interpolated), InterpResult(posterior_t1, interpolated)

    def interpolate_fwd_at_t1(self, *, posterior_t1):
        return (posterior_t1, posterior_t1), InterpResult(posterior_t1, posterior_t1)

    def interpolate_offgrid_marginals(self, *, posterior_t0, posterior_t1, transition_t0_t, transition_t_t1):
        """:403-414: a filter extrapolates from the left grid point."""
        _, interpolated = self.predict(posterior_t0, transition=transition_t0_t)
        return (interpolated, interpolated), InterpResult(posterior_t1, interpolated)


class _Smoother:
    """estimators_and_losses.py:424-470."""

    def init_posterior(self, *, u):
        return u, MarkovSequence(u, u.alg.identity_conditional(u))

    def apply_updates(self, prediction, *, updates):
        return updates, MarkovSequence(updates, prediction.conditional, prediction.reverse)

    def finalize(self, *, posterior0, posterior, posterior1, output_scale):
        posterior0 = posterior0.rescale_cholesky(output_scale)
        posterior = [p.rescale_cholesky(output_scale) for p in posterior]
        posterior1 = posterior1.rescale_cholesky(output_scale)
        rv_at_t1 = posterior1.conditional.marginalise(posterior1.marginal)
        full = MarkovSequence(rv_at_t1, [p.conditional for p in posterior], True)
        marginals = full.evaluate_marginals()
        filtering = [posterior0.marginal] + [p.marginal for p in posterior]
        return marginals, SmoothingSolution(full, filtering)


class SmoothingSolution:
    def __init__(self, posterior, filtering):
        self.posterior = posterior
        self.filtering = filtering


class strategy_smoother_fixedpoint(_Smoother):
    """estimators_and_losses.py:473-591."""

    is_suitable_for_save_at = True
    is_suitable_for_save_every_step = False
    is_suitable_for_offgrid_marginals = False

    def predict(self, posterior, *, transition):
        marginals, cond = transition.revert(posterior.marginal)
        cond = posterior.conditional.merge(cond)
        return marginals, MarkovSequence(marginals, cond, posterior.reverse)

    def interpolate_fwd_at_t1(self, *, posterior_t1):
        ident = posterior_t1.marginal.alg.identity_conditional(posterior_t1.marginal)
        resume = MarkovSequence(posterior_t1.marginal, ident, posterior_t1.reverse)
        return (posterior_t1.marginal, posterior_t1), InterpResult(resume, resume)

    def interpolate_fwd(self, *, posterior_t0, posterior_t1, transition_t0_t, transition_t_t1):
        _, extrapolated_t = self.predict(posterior_t0, transition=transition_t0_t)
        ident = posterior_t0.marginal.alg.identity_conditional(posterior_t0.marginal)
        previous_new = MarkovSequence(extrapolated_t.marginal, ident, extrapolated_t.reverse)
        _, extrapolated_t1 = self.predict(previous_new, transition=transition_t_t1)
        interpolated = MarkovSequence(extrapolated_t.marginal, extrapolated_t.conditional, True)
        step_from = MarkovSequence(posterior_t1.marginal, extrapolated_t1.conditional, True)
        return (interpolated.marginal, interpolated), InterpResult(step_from, previous_new)


class strategy_smoother_fixedinterval(_Smoother):
    """estimators_and_losses.py:594-717 (forward pass only; used for cross-checks)."""

    is_suitable_for_save_at = False
    is_suitable_for_save_every_step = True
    is_suitable_for_offgrid_marginals = True

    def predict(self, posterior, *, transition):
        marginals, cond = transition.revert(posterior.marginal)
        return marginals, MarkovSequence(marginals, cond, posterior.reverse)

    def interpolate_fwd(self, *, posterior_t0, posterior_t1, transition_t0_t, transition_t_t1):
        _, sol_t = self.predict(posterior_t0, transition=transition_t0_t)
        _, ext_t1 = self.predict(sol_t, transition=transition_t_t1)
        sol_t1 = MarkovSequence(posterior_t1.marginal, ext_t1.conditional, True)
        return (sol_t.marginal, sol_t), InterpResult(sol_t1, sol_t)

    def interpolate_fwd_at_t1(self, *, posterior_t1):
        return (posterior_t1.marginal, posterior_t1), InterpResult(posterior_t1, posterior_t1)

    def interpolate_offgrid_marginals(self, *, posterior_t0, posterior_t1, transition_t0_t, transition_t_t1):
        """:677-709: extrapolate the FILTERING distribution t0 -> t -> t1, then pull the (smoothed) marginal at t1
        back through the new t1 -> t conditional. `posterior_t0` is the filtering marginal at t0."""
        _, post = self.init_posterior(u=posterior_t0)
        _, ext_t = self.predict(post, transition=transition_t0_t)
        _, ext_t1 = self.predict(ext_t, transition=transition_t_t1)
        rv_at_t = ext_t1.conditional.marginalise(posterior_t1.marginal)
        sol_t = MarkovSequence(rv_at_t, ext_t.conditional, True)
        sol_t1 = MarkovSequence(posterior_t1.marginal, ext_t1.conditional, True)
        return (rv_at_t, sol_t), InterpResult(sol_t1, sol_t)


# ------------------------------------------------------------------------------------------------
# Solutions and solvers
# ------------------------------------------------------------------------------------------------


class ProbabilisticSolution:
    """_probdiffeq/solvers.py:33-69."""

    def __init__(self, *, t, u, solution_full, output_scale, num_steps, auxiliary, fun_evals, prior):
        self.t = t
        self.u = u
        self.solution_full = solution_full
        self.output_scale = output_scale
        self.num_steps = num_steps
        self.auxiliary = auxiliary
        self.fun_evals = fun_evals
        self.prior = prior

    def replace(self, **kw):
        d = dict(self.__dict__)
        d.update(kw)
        return ProbabilisticSolution(**d)

    # Conveniences on stacked outputs (u is a list of Normals after userfriendly_output)
    @property
    def u_mean(self):
        return np.stack([r.tcoeffs for r in self.u]) if isinstance(self.u, list) else self.u.tcoeffs

    @property
    def u_std(self):
        return np.stack([r.std for r in self.u]) if isinstance(self.u, list) else self.u.std

    @property
    def u_chol(self):
        return np.stack([r.chol for r in self.u]) if isinstance(self.u, list) else self.u.chol

    def terminal(self):
        """tree_map(lambda s: s[-1]). _ivpsolve/solvers_via_adaptive_steps.py:41."""
        full = self.solution_full
        if isinstance(full, list):
            full = full[-1]
        return self.replace(
            t=self.t[-1],
            u=self.u[-1],
            solution_full=full,
            output_scale=self.output_scale[-1],
            num_steps=self.num_steps[-1],
        )


class _ProbabilisticSolver:
    """_probdiffeq/solvers.py:72-315."""

    def __init__(self, *, strategy, constraint, constraint_init=None):
        self.strategy = strategy
        self.constraint = constraint
        self.constraint_init = constraint_init

    def _init_update(self, u_pred, prediction, *, t, damp):
        """The optional Bayes update at t0 (solvers.py:361-372, 526-537, 670-680): linearise `constraint_init` at the
        initial state and condition on a zero residual, the gain through the minimum-norm least-squares solve
        (`linalg.lstsq_svd`) because the initial observation factor may be singular (exact Taylor coefficients)."""
        u0, posterior, _rms = self._init_update_and_rms(u_pred, prediction, t=t, damp=damp)
        return u0, posterior

    def _init_update_and_rms(self, u_pred, prediction, *, t, damp):
        """... and the whitened RMS residual of that update, which `solver_mle` takes as the first term of its running
        calibration (solvers.py:361-374: `bayes_rule_and_residual_whitened_rms_tree`, `num_data = 1.0`). The residual
        is whitened by a plain triangular solve with the observed factor (ssm_impl_isotropic.py:203-207), not by the
        least-squares solve of the gain: a singular observed factor gives NaN there, as in the reference."""
        if self.constraint_init is None:
            return u_pred, prediction, None
        fx_init, _ = self.constraint_init.linearize(u_pred, self.constraint_init.init_linearization(), damp=damp, t=t)
        observed, reverted = fx_init.revert(u_pred, solve=linalg.lstsq_triu)
        zeros = np.zeros_like(fx_init.noise.mean)
        with np.errstate(divide="ignore", invalid="ignore"):
            rms = fx_init.alg.residual_whitened_rms(observed, zeros)
        updates = reverted.apply_flat(zeros)
        return (*self.strategy.apply_updates(prediction, updates=updates), rms)

    @property
    def is_suitable_for_save_at(self):
        return self.strategy.is_suitable_for_save_at

    @property
    def is_suitable_for_save_every_step(self):
        return self.strategy.is_suitable_for_save_every_step

    def _zeros_like_fx(self, rv, t, damp):
        # eval_shape + zeros_like (solvers.py:357-359): a zero linearisation of the right shape.
        fx, _ = self.constraint.linearize(rv, None, damp=damp, t=t)
        noise = ssm.Normal(0.0 * fx.noise.mean, 0.0 * fx.noise.chol, fx.alg)
        return ssm.Cond(0.0 * fx.A, noise, 0.0 * fx.to_latent, 0.0 * fx.to_observed)

    def offgrid_marginals(self, t, *, solution):
        """solvers.py:149-203: the marginal at one time strictly inside the grid and not on it."""
        if not self.strategy.is_suitable_for_offgrid_marginals:
            raise NotImplementedError
        index = int(np.searchsorted(solution.t, t))
        full = solution.solution_full
        posterior_t0 = full.filtering[index - 1] if isinstance(full, SmoothingSolution) else full[index - 1]
        t0, t1 = solution.t[index - 1], solution.t[index]
        # solver / solver_mle return T - 1 (identical) scales for T grid points (solvers.py:468-469, 744-745); JAX
        # clamps the out-of-range index of the last interval, which is restated here
        scales = np.asarray(solution.output_scale)
        output_scale = scales[min(index, scales.shape[0] - 1)]
        _, posterior_t1 = self.strategy.init_posterior(u=solution.u[index])
        tr0 = solution.prior.transition(dt=t - t0, output_scale=output_scale)
        tr1 = solution.prior.transition(dt=t1 - t, output_scale=output_scale)
        (estimate, _), _ = self.strategy.interpolate_offgrid_marginals(
            posterior_t0=posterior_t0, posterior_t1=posterior_t1, transition_t0_t=tr0, transition_t_t1=tr1
        )
        return estimate

    def interpolate_fwd(self, *, t, interp_from, interp_to):
        """solvers.py:205-269."""
        output_scale = interp_to.output_scale
        tr0 = interp_from.prior.transition(dt=t - interp_from.t, output_scale=output_scale)
        tr1 = interp_from.prior.transition(dt=interp_to.t - t, output_scale=output_scale)
        (estimate, interpolated), res = self.strategy.interpolate_fwd(
            posterior_t0=interp_from.solution_full,
            posterior_t1=interp_to.solution_full,
            transition_t0_t=tr0,
            transition_t_t1=tr1,
        )
        step_from = interp_to.replace(solution_full=res.step_from)
        solution = interp_to.replace(t=t, solution_full=interpolated, u=estimate)
        new_from = interp_from.replace(t=t, solution_full=res.interp_from)
        return solution, InterpResult(step_from, new_from)

    def interpolate_fwd_at_t1(self, *, t, interp_from, interp_to):
        """solvers.py:271-315."""
        (estimate, interpolated), res = self.strategy.interpolate_fwd_at_t1(
            posterior_t1=interp_to.solution_full
        )
        prev = interp_from.replace(t=interp_to.t, solution_full=res.interp_from)
        sol = interp_to.replace(solution_full=interpolated, u=estimate)
        acc = interp_to.replace(solution_full=res.step_from)
        return sol, InterpResult(acc, prev)

    def _stack(self, solution0, solution, estimate, posterior, output_scale):
        ts = np.asarray([solution0.t] + [s.t for s in solution])
        return ProbabilisticSolution(
            t=ts,
            u=estimate,
            solution_full=posterior,
            output_scale=output_scale,
            num_steps=np.asarray([s.num_steps for s in solution]),
            auxiliary=[s.auxiliary for s in solution],
            fun_evals=[s.fun_evals for s in solution],
            prior=solution0.prior,
        )


class solver(_ProbabilisticSolver):
    """Uncalibrated solver. solvers.py:636-767."""

    def init(self, t, u, *, damp):
        prior = u
        u_pred, prediction = self.strategy.init_posterior(u=prior.init)
        u0, posterior = self._init_update(u_pred, prediction, t=t, damp=damp)
        fx = self._zeros_like_fx(u_pred, t, damp)
        output_scale = np.ones_like(prior.alg.prototype_output_scale(u_pred))
        return ProbabilisticSolution(
            t=t, u=u0, solution_full=posterior, num_steps=0, auxiliary=None,
            output_scale=output_scale, fun_evals=fx, prior=prior,
        )  # fmt: skip

    def step(self, state, *, dt, damp):
        output_scale = np.ones_like(state.output_scale)
        transition = state.prior.transition(dt=dt, output_scale=output_scale)
        u_pred, prediction = self.strategy.predict(state.solution_full, transition=transition)
        fx, aux = self.constraint.linearize(u_pred, state.auxiliary, damp=damp, t=state.t + dt)
        updates = fx.bayes_rule(np.zeros_like(fx.noise.mean), u_pred)
        u, posterior = self.strategy.apply_updates(prediction, updates=updates)
        return ProbabilisticSolution(
            t=state.t + dt, u=u, solution_full=posterior, output_scale=output_scale,
            auxiliary=aux, num_steps=state.num_steps + 1, fun_evals=fx, prior=state.prior,
        )  # fmt: skip

    def userfriendly_output(self, *, solution0, solution, solution1):
        output_scale = np.ones_like(solution[-1].output_scale)
        u, posterior = self.strategy.finalize(
            posterior0=solution0.solution_full,
            posterior=[s.solution_full for s in solution],
            posterior1=solution1.solution_full,
            output_scale=output_scale,
        )
        scales = np.stack([np.ones_like(s.output_scale) * output_scale for s in solution])
        return self._stack(solution0, solution, u, posterior, scales)


class solver_mle(_ProbabilisticSolver):
    """Maximum-likelihood (running RMS) calibration. solvers.py:318-480."""

    def __init__(self, *, strategy, constraint, constraint_init=None, correct_asymptotic_underconfidence=True):
        super().__init__(strategy=strategy, constraint=constraint, constraint_init=constraint_init)
        self.correct_asymptotic_underconfidence = correct_asymptotic_underconfidence

    def init(self, t, u, *, damp):
        prior = u
        u_pred, prediction = self.strategy.init_posterior(u=prior.init)
        u0, posterior, rms0 = self._init_update_and_rms(u_pred, prediction, t=t, damp=damp)
        output_scale_prior = np.ones_like(prior.alg.prototype_output_scale(u_pred))
        fx = self._zeros_like_fx(u_pred, t, damp)
        if rms0 is None:
            auxiliary = (None, np.zeros_like(output_scale_prior), 0.0)
        else:  # solvers.py:366-374: the update at t0 is the first datum of the calibration
            auxiliary = (None, rms0 * np.ones_like(output_scale_prior), 1.0)
        return ProbabilisticSolution(
            t=t, u=u0, solution_full=posterior, auxiliary=auxiliary,
            output_scale=output_scale_prior, num_steps=0, fun_evals=fx, prior=prior,
        )  # fmt: skip

    def step(self, state, *, dt, damp):
        output_scale = np.ones_like(state.prior.alg.prototype_output_scale(state.u))
        transition = state.prior.transition(dt=dt, output_scale=output_scale)
        u, prediction = self.strategy.predict(state.solution_full, transition=transition)
        lin_state, running, num_data = state.auxiliary
        fx, cstate = self.constraint.linearize(u, lin_state, damp=damp, t=state.t + dt)
        new_term, updates = fx.bayes_rule_and_residual_whitened_rms(np.zeros_like(fx.noise.mean), u)
        u, posterior = self.strategy.apply_updates(prediction, updates=updates)
        x1 = np.sqrt(num_data / (num_data + 1)) * running
        x2 = np.sqrt(1 / (num_data + 1)) * new_term
        running = np.hypot(x1, x2)
        return ProbabilisticSolution(
            t=state.t + dt, u=u, solution_full=posterior, output_scale=state.output_scale,
            auxiliary=(cstate, running, num_data + 1), num_steps=state.num_steps + 1,
            fun_evals=fx, prior=state.prior,
        )  # fmt: skip

    def userfriendly_output(self, *, solution0, solution, solution1):
        _, output_scale, _ = solution1.auxiliary
        if self.correct_asymptotic_underconfidence:
            output_scale = output_scale / np.sqrt(solution[-1].num_steps)
        estimate, posterior = self.strategy.finalize(
            posterior0=solution0.solution_full,
            posterior=[s.solution_full for s in solution],
            posterior1=solution1.solution_full,
            output_scale=output_scale,
        )
        scales = np.stack([np.ones_like(s.auxiliary[1]) * output_scale for s in solution])
        return self._stack(solution0, solution, estimate, posterior, scales)


class solver_dynamic(_ProbabilisticSolver):
    """Per-step (dynamic) calibration. solvers.py:483-633."""

    def __init__(self, *, strategy, constraint, constraint_init=None, re_linearize_after_calibration=False):
        super().__init__(strategy=strategy, constraint=constraint, constraint_init=constraint_init)
        self.re_linearize_after_calibration = re_linearize_after_calibration

    def init(self, t, u, *, damp):
        prior = u
        u_pred, prediction = self.strategy.init_posterior(u=prior.init)
        u0, posterior = self._init_update(u_pred, prediction, t=t, damp=damp)
        output_scale = np.ones_like(prior.alg.prototype_output_scale(u_pred))
        fx = self._zeros_like_fx(u_pred, t, damp)
        return ProbabilisticSolution(
            t=t, u=u0, solution_full=posterior, auxiliary=None,
            output_scale=output_scale, num_steps=0, fun_evals=fx, prior=prior,
        )  # fmt: skip

    def step(self, state, *, dt, damp):
        lin_state = state.auxiliary
        ones = np.ones_like(state.prior.alg.prototype_output_scale(state.u))
        transition = state.prior.transition(dt=dt, output_scale=ones)
        u = transition.apply_flat(state.u.mean)
        fx, lin_state = self.constraint.linearize(u, lin_state, damp=damp, t=state.t + dt)
        observed = fx.marginalise(u)
        output_scale = observed.alg.residual_whitened_rms(observed, np.zeros_like(fx.noise.mean))
        transition = state.prior.transition(dt=dt, output_scale=output_scale)
        u, prediction = self.strategy.predict(state.solution_full, transition=transition)
        if self.re_linearize_after_calibration:
            fx, lin_state = self.constraint.linearize(u, lin_state, damp=damp, t=state.t + dt)
        updates = fx.bayes_rule(np.zeros_like(fx.noise.mean), u)
        u, posterior = self.strategy.apply_updates(prediction, updates=updates)
        return ProbabilisticSolution(
            t=state.t + dt, u=u, solution_full=posterior, num_steps=state.num_steps + 1,
            auxiliary=lin_state, output_scale=output_scale, fun_evals=fx, prior=state.prior,
        )  # fmt: skip

    def userfriendly_output(self, *, solution0, solution, solution1):
        output_scale = np.ones_like(solution[-1].output_scale)
        estimate, posterior = self.strategy.finalize(
            posterior0=solution0.solution_full,
            posterior=[s.solution_full for s in solution],
            posterior1=solution1.solution_full,
            output_scale=output_scale,
        )
        scales = np.stack([solution0.output_scale] + [s.output_scale for s in solution])
        return self._stack(solution0, solution, estimate, posterior, scales)


# ------------------------------------------------------------------------------------------------
# Error norms and estimators
# ------------------------------------------------------------------------------------------------


def _rms(s):
    return np.linalg.norm(s) / np.sqrt(s.size)


def error_norm_scale_then_rms():
    """solvers.py:770-791."""

    def normalize(error_abs, reference, atol, rtol):
        return _rms(error_abs / (atol + rtol * np.abs(reference)))

    return normalize


def error_norm_rms_then_scale():
    """solvers.py:794-811."""

    def normalize(error_abs, reference, atol, rtol):
        return _rms(error_abs) / (atol + rtol * _rms(reference))

    return normalize


class error_residual_std:
    """solvers.py:850-996."""

    def __init__(self, *, constraint, error_norm=None, re_linearize_before_error=False, error_per_unit_step=False):
        self.constraint = constraint
        self.error_norm = error_norm_scale_then_rms() if error_norm is None else error_norm
        self.re_linearize_before_error = re_linearize_before_error
        self.error_per_unit_step = error_per_unit_step

    def init_error(self):
        return self.constraint.init_linearization()

    def estimate_error_norm(self, state, previous, proposed, *, dt, atol, rtol, damp):
        alg = previous.prior.alg
        ones = np.ones_like(alg.prototype_output_scale(proposed.u))
        transition = previous.prior.transition(dt=dt, output_scale=ones)
        rv = transition.apply_flat(previous.u.mean)
        if self.re_linearize_before_error:
            linearized, state = self.constraint.linearize(rv, state, damp=damp, t=proposed.t)
        else:
            linearized = proposed.fun_evals
        observed = linearized.marginalise(rv)
        output_scale = alg.residual_whitened_rms(observed, np.zeros_like(linearized.noise.mean))
        observed = observed.rescale_cholesky(output_scale)
        error = np.asarray(observed.std).reshape(-1)

        prev_nd, prop_nd = previous.u.tcoeffs, proposed.u.tcoeffs
        error_contraction_rate = prev_nd.shape[0]
        reference = np.maximum(np.abs(prev_nd[0]), np.abs(prop_nd[0]))
        n = self.constraint.residual_order - 1
        if self.error_per_unit_step:
            n += 1
        if error.shape not in [(1,), reference.shape]:
            raise ValueError("The error-estimate and reference have different shapes.")
        error_abs = error * dt**n / linalg.factorial(n)
        error_norm = self.error_norm(error_abs, reference, atol=atol, rtol=rtol)
        return error_norm ** (-1.0 / error_contraction_rate), state


class error_state_std:
    """solvers.py:999-1098."""

    def __init__(self, *, constraint, error_norm=None, re_linearize_before_error=False, derivative_idx=0, error_per_unit_step=False):
        self.constraint = constraint
        self.error_norm = error_norm_scale_then_rms() if error_norm is None else error_norm
        self.re_linearize_before_error = re_linearize_before_error
        self.derivative_idx = derivative_idx
        self.error_per_unit_step = error_per_unit_step

    def init_error(self):
        return self.constraint.init_linearization()

    def estimate_error_norm(self, state, previous, proposed, *, dt, atol, rtol, damp):
        alg = previous.prior.alg
        ones = np.ones_like(alg.prototype_output_scale(proposed.u))
        transition = previous.prior.transition(dt=dt, output_scale=ones)
        rv = transition.apply_flat(previous.u.mean)
        error_contraction_rate = previous.u.tcoeffs.shape[0]
        if self.re_linearize_before_error:
            linearized, state = self.constraint.linearize(rv, state, damp=damp, t=proposed.t)
        else:
            linearized = proposed.fun_evals
        output_scale, conditional = linearized.bayes_rule_and_residual_whitened_rms(
            np.zeros_like(linearized.noise.mean), rv
        )
        n = self.derivative_idx
        std = np.asarray(conditional.std)[n]
        error = np.asarray(output_scale * std).reshape(-1)
        reference = np.maximum(np.abs(previous.u.tcoeffs[n]), np.abs(proposed.u.tcoeffs[n]))
        if self.error_per_unit_step:
            n += 1
        error_abs = error * dt**n / linalg.factorial(n)
        error_norm = self.error_norm(error_abs, reference, atol=atol, rtol=rtol)
        return error_norm ** (-1.0 / error_contraction_rate), state


# ------------------------------------------------------------------------------------------------
# Log-marginal-likelihood of terminal-value data
# ------------------------------------------------------------------------------------------------


def loss_lml_timeseries(*, average_pdfs=True, tcoeff_index=0, solve=None):
    """estimators_and_losses.py:53-105. `solve` defaults to the least-squares triangular solve the reference
    passes (backend/linalg.py:60-61), which equals the exact solve whenever the observed factor is non-singular.
    `u`: (T, d) data; `std`: (T,) isotropic, (T, d) block-diagonal / dense."""
    if solve is None:
        solve = linalg.lstsq_triu

    def loss(u, /, *, posterior, std):
        if not isinstance(posterior, MarkovSequence):
            raise TypeError("The datatype of the posterior is not as expected. Did you perhaps use a filter?")
        u = np.asarray(u, dtype=np.float64)
        std = np.asarray(std, dtype=np.float64)
        posterior = posterior.remove_filtering_distributions()
        alg = posterior.marginal.alg
        std_expected = np.stack([np.asarray(posterior.marginal.std)[tcoeff_index]] * u.shape[0])
        if std.shape != std_expected.shape:
            raise ValueError("The standard deviation container differs from what was expected.")
        model = [alg.to_derivative(posterior.marginal, tcoeff_index, s) for s in std]
        data = [alg.from_nd(x[None, :]) for x in u]
        return posterior.evaluate_lml(data, model=model, average_pdfs=average_pdfs, solve=solve)

    return loss


def loss_lml_terminal_values(*, tcoeff_index=0):
    """estimators_and_losses.py:20-50."""

    def loss(u, /, *, marginals, std):
        alg = marginals.alg
        std_expected = np.asarray(marginals.std)[tcoeff_index]
        std = np.asarray(std, dtype=np.float64)
        if std.shape != std_expected.shape:
            raise ValueError("The standard deviation container differs from what was expected.")
        model = alg.to_derivative(marginals, tcoeff_index, std)
        marg = model.marginalise(marginals)
        u = np.asarray(u, dtype=np.float64)
        data = alg.from_nd(u[None, :])
        return alg.logpdf(marg, data)

    return loss
