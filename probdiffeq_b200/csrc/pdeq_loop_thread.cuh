// K1: the whole adaptive step loop of one IVP instance in one thread, state resident in registers.
//
// Restates, flattened into a single persistent loop whose body is ONE step attempt,
//   probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:100-146 (scan over checkpoints + advance while)
//   probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:227-338 (RejectionLoop.loop/step/step_attempt)
//   probdiffeq/_ivpsolve/solvers_via_fixed_steps.py:21-32     (fixed grid)
//   probdiffeq/_probdiffeq/solvers.py:395-437, 549-599, 702-733 (solver_mle / solver_dynamic / solver .step)
//   probdiffeq/_probdiffeq/solvers.py:925-996, 1037-1098       (error_residual_std / error_state_std)
//   probdiffeq/_ivpsolve/controllers.py:46-63, 78-84           (PI / I control)
//   probdiffeq/_probdiffeq/estimators_and_losses.py:371-421    (strategy_filter predict / interpolate)
// for the isotropic and block-diagonal factorisations with a scalar observation per block (ts0, ts1).
//
// Why one flat loop: under `jax.vmap` the reference's nested while-loops run until the slowest lane is
// done.  Here a lane that finishes its instance immediately pulls the next instance index from a global
// counter, so all 32 lanes of a warp keep executing the same attempt body on different instances.
//
// Scalar arithmetic: divisions and square roots use branch-free MUFU + Newton sequences (pdeq_blockops.cuh),
// and the controller works on log2(error_power) -- error_power = norm^(-1/n) is never formed: accept iff
// log2(norm) <= 0, and the PI gain e^0.3 (e/e_prev)^0.4 is one exp2 of a linear combination of logs. These
// differ from the reference's pow() chain by rounding only (1e-16 relative in dt).
#pragma once

#include "pdeq_blockops.cuh"
#include "pdeq_vf.cuh"

namespace pdeq {

struct LoopArgs {
  pdeq_config cfg;
  pdeq_problem prob;
  pdeq_solution sol;
  const double* grid;  // save_at (adaptive) or the fixed grid, [T]
  int32_t T;
  int32_t fixed_grid;
  double atol, rtol, eps, damp;
  const double* dt0;
  int64_t dt0_stride;
  unsigned long long* work_counter;  // zero-initialised by the host wrapper
};

constexpr int K1_THREADS = 128;
// Resident CTAs per SM the register allocation is sized for. Measured on the headline kernel (2^20 Lotka-Volterra
// instances): 2 CTAs (254 registers, no spills) 29.9 ms; 3 CTAs (168 registers, 284 B of spills) 26.3 ms;
// 4 CTAs (128 registers, 892 B of spills) 29.2 ms. One shared factor (isotropic) fits the 168-register budget;
// per-dimension factors (block-diagonal) keep 255 registers.
#ifndef PDEQ_K1_MIN_BLOCKS
#define PDEQ_K1_MIN_BLOCKS(FACT) ((FACT) == PDEQ_FACT_ISOTROPIC ? 3 : 2)
#endif

// Resident CTAs per SM of the specialised builds (SPEC = 1..4, see ThreadLoop). Registers are per SM sub-partition
// (16 K each), so 13..16 resident warps all mean <= 128 registers and 9..12 mean <= 168: there is no build "between"
// 2 and 1, whatever the CTA size.
#define PDEQ_K1_SPEC_MIN_BLOCKS(SPEC) ((SPEC) >= 2 ? 4 : 3)

template <int n>
PDEQ_DI double ipow_small(double x, int k) {
  double r = 1.0;
#pragma unroll
  for (int e = 0; e < n + 1; ++e) {
    if (e < k) r *= x;
  }
  return r;
}

// sqrt(x) for x > 0, else 0 -- in select form (no branch: the attempt body stays one basic block, which lets the
// scheduler overlap the error estimate with the triangularisations); same value as `x > 0 ? fast_sqrt(x) : 0`.
PDEQ_DI double safe_sqrt(double x) {
  const bool pos = x > 0.0;
  const double r = fast_sqrt(pos ? x : 1.0);
  return pos ? r : 0.0;
}

// SPEC = 0: every solver / error / control option is a run-time (warp-uniform) branch on pdeq_config.
// SPEC = 1: the headline combination is fixed at compile time -- adaptive with clip_dt, `solver` (no calibration),
//           error_state_std on coefficient 0 through the constant-matrix shortcut (ts0, damp = 0), scale-then-rms
//           norm, unit prior scale. The arithmetic that remains is the SPEC = 0 arithmetic
//           operation for operation (bitwise the same results); what goes away are the values the run-time
//           branches keep alive (the noise-only factor Lq, calibration state), i.e. registers and spills.
//           The launcher checks the conditions on the host (k1_spec_matches).
// SPEC = 2: the same loop compiled for four resident CTAs per SM (128 registers) instead of three (168).
// (Round 1 also measured builds with the accepted state in shared memory, five resident CTAs, and parameters re-read
// per attempt -- all slower, profiles/r1e_sweep_k1_spec.jsonl -- they are gone.)
template <class VF, int NU, int FACT, int D, bool TS0, int SPEC = 0>
struct ThreadLoop {
  static constexpr bool SP = SPEC != 0;
  static constexpr int THREADS = K1_THREADS;
  static_assert(!SP || TS0, "the specialised loop is ts0 only");
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int NB = (FACT == PDEQ_FACT_BLOCKDIAG) ? D : 1;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  static constexpr int IF_SLOTS = n * D + NB * n * n + 1;  // interp_from: mean, chol, t
  static constexpr int NLOW = NB * (n * (n + 1)) / 2;
  static_assert(q < n, "need more Taylor coefficients than the ODE order");

  PDEQ_DI static constexpr int blk(int j) { return FACT == PDEQ_FACT_BLOCKDIAG ? j : 0; }

  struct RegAcc {
    const double (&m)[n][D];
    PDEQ_DI double operator()(int k, int i) const {
      double v = 0.0;
#pragma unroll
      for (int kk = 0; kk < q; ++kk) {
#pragma unroll
        for (int ii = 0; ii < D; ++ii) {
          if (kk == k && ii == i) v = m[kk][ii];
        }
      }
      return v;
    }
  };

  // Write one checkpoint of the solution.
  PDEQ_DI static void emit(const LoopArgs& a, long b, int ck, double t, const double (&m)[n][D],
                           const double (&L)[NB][n][n], const double (&scale)[NB], int nsteps) {
    const long bt = b * a.T + ck;
    a.sol.t[bt] = t;
    a.sol.num_steps[bt] = nsteps;
    double* mo = a.sol.mean + bt * (n * D);
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) mo[i * D + j] = m[i][j];
    }
    if (a.sol.chol != nullptr) {
      double* co = a.sol.chol + bt * (NB * n * n);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
          for (int j = 0; j < n; ++j) co[(k * n + i) * n + j] = (j <= i) ? L[k][i][j] : 0.0;
        }
      }
    }
    if (a.sol.output_scale != nullptr) {
#pragma unroll
      for (int k = 0; k < NB; ++k) a.sol.output_scale[bt * NB + k] = scale[k];
    }
  }

  // A checkpoint the instance never reached (max_attempts bail-out).
  PDEQ_DI static void emit_nan(const LoopArgs& a, long b, int ck, int nsteps) {
    const long bt = b * a.T + ck;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    a.sol.t[bt] = nan;
    a.sol.num_steps[bt] = nsteps;
    for (int e = 0; e < n * D; ++e) a.sol.mean[bt * (n * D) + e] = nan;
    if (a.sol.chol != nullptr) {
      for (int e = 0; e < NB * n * n; ++e) a.sol.chol[bt * (NB * n * n) + e] = nan;
    }
    if (a.sol.output_scale != nullptr) {
      for (int k = 0; k < NB; ++k) a.sol.output_scale[bt * NB + k] = nan;
    }
  }

  // interp_from lives in shared memory (one column per thread): it is written on every accepted step but read
  // only when a checkpoint is overstepped, so it should not occupy registers.
  PDEQ_DI static void if_store(double* __restrict__ sm, int nthreads, int tid, const double (&m)[n][D],
                               const double (&L)[NB][n][n], double t) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) sm[(i * D + j) * nthreads + tid] = m[i][j];
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) sm[(n * D + (k * n + i) * n + j) * nthreads + tid] = L[k][i][j];
      }
    }
    sm[(IF_SLOTS - 1) * nthreads + tid] = t;
  }
  PDEQ_DI static void if_load(const double* __restrict__ sm, int nthreads, int tid, double (&m)[n][D],
                              double (&L)[NB][n][n], double& t) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) m[i][j] = sm[(i * D + j) * nthreads + tid];
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) L[k][i][j] = sm[(n * D + (k * n + i) * n + j) * nthreads + tid];
      }
    }
    t = sm[(IF_SLOTS - 1) * nthreads + tid];
  }

  // Whitened RMS of the observation residual per block (IsotropicNormal.residual_whitened_rms_flat,
  // ssm_impl_isotropic.py:203-207; BlockDiagNormal..., ssm_impl_blockdiag.py:285-292) for 1x1 factors r[k].
  PDEQ_DI static void whitened_rms(const double (&mobs)[D], const double (&r)[NB], double (&out)[NB]) {
    if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
      for (int j = 0; j < D; ++j) out[blk(j)] = fabs(mobs[j] * fast_rcp(r[blk(j)]));
    } else {
      const double inv = fast_rcp(r[0]);
      double ss = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double w = mobs[j] * inv;
        ss = fma(w, w, ss);
      }
      out[0] = safe_sqrt(ss) * rsqrt((double)D);
    }
  }

  // Linearise the constraint at a mean (ts0: ssm_impl_isotropic.py:304-317 / ssm_impl_blockdiag.py:129-144;
  // ts1: ssm_impl_isotropic.py:326-355 / ssm_impl_blockdiag.py:153-183): observation rows h (coefficients 0..q of each
  // block) and the observed mean mobs = h m + bias.
  PDEQ_DI static void linearise(const double (&mp)[n][D], const double (&params)[P], double t_new, double (&h)[NB][q + 1],
                                double (&mobs)[D]) {
    RegAcc acc{mp};
    double f[D];
#pragma unroll
    for (int j = 0; j < D; ++j) f[j] = VF::template component<double>(j, D, acc, params, t_new);
    if (TS0) {
#pragma unroll
      for (int k = 0; k < NB; ++k) {
#pragma unroll
        for (int c = 0; c <= q; ++c) h[k][c] = (c == q) ? 1.0 : 0.0;
      }
#pragma unroll
      for (int j = 0; j < D; ++j) mobs[j] = mp[q][j] + (-f[j]);
    } else {
      if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
#pragma unroll
          for (int c = 0; c < q; ++c) h[blk(j)][c] = -VF::jac(j, c, j, D, acc, params, t_new);
          h[blk(j)][q] = 1.0;
        }
      } else {
#pragma unroll
        for (int c = 0; c < q; ++c) {
          double tr = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) tr += -VF::jac(j, c, j, D, acc, params, t_new);
          h[0][c] = tr / (double)D;
        }
        h[0][q] = 1.0;  // trace(I_d) / d
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double r = mp[q][j] - f[j];
        double hm = 0.0;
#pragma unroll
        for (int c = 0; c <= q; ++c) hm = fma(h[blk(j)][c], mp[c][j], hm);
        const double bias = r - hm;
        mobs[j] = hm + bias;
      }
    }
  }

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem_if) {
    const pdeq_config& cfg = a.cfg;
    const double(*__restrict__ A)[PDEQ_MAX_COEFFS] = cfg.sys_a;
    const double(*__restrict__ Q)[PDEQ_MAX_COEFFS] = cfg.sys_q;
    const double* __restrict__ fact = cfg.factorials;
    const double* __restrict__ ifact = cfg.inv_factorials;
    const bool adaptive = SP ? true : a.fixed_grid == 0;
    const bool clip = SP ? true : cfg.clip_dt != 0;
    const int cfg_solver = SP ? (int)PDEQ_SOLVER_PLAIN : cfg.solver;
    const int cfg_error = SP ? (int)PDEQ_ERROR_STATE_STD : cfg.error;
    const int cfg_norm = SP ? (int)PDEQ_NORM_SCALE_THEN_RMS : cfg.error_norm;
    const double damp = SP ? 0.0 : a.damp;
    const bool needs_interp = adaptive && !clip;
    const int T = a.T;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const double inv_sqrt_d = rsqrt((double)D);
    const double neg_inv_n = -1.0 / (double)n;

    // ---- per-instance state (registers) ----
    double m[n][D], L[NB][n][n], prior[NB], sig[NB], run_scale[NB], params[P];
    double t = 0.0, dt = 0.0, ctrl_lprev = 0.0, ndata = 0.0, t_next = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;

    bool need_load = true, first_ticket = true;
    // A lane that finishes its instance pulls the next ticket from the global counter, so all 32 lanes of a warp keep
    // executing the same attempt body on different instances; a lane leaves when the tickets run out. The lanes of a
    // warp finish at different times, and the ~450 instructions of a switch-over (last checkpoint, status, ticket,
    // inputs, first checkpoint) are issued for the whole warp each time: 32 times per instance length, ~5 % of the
    // run (measured: with the instances served in exact order of their attempt counts the lanes of a warp switch
    // together and a 2^20-instance pass takes 16.4 instead of 17.1 ms). Hiding the memory latency of the switch-over
    // (ticket drawn ahead, inputs staged with cp.async) changed nothing: it is the instruction issue that costs.
    while (true) {
      // ------------------------------------------------------------------ fetch the next instance
      if (need_load) {
        // The first ticket of a lane is its global thread index, later ones come from the counter (offset by the grid
        // size): which lane starts on which entry of the service order is then fixed -- the 32 lanes of a warp start
        // on 32 NEIGHBOURS of the order (equally long instances, if the host sorted it) instead of on whatever the
        // race for the counter hands them (2^17 instances served longest-first: 2.47 -> 2.34 ms). Tried on top of
        // it and removed: finished lanes sitting out a few attempts so that the lanes of a warp switch over together
        // (2.34 -> 2.39 ms or worse for every patience tried -- the idle lane-attempts cost more than the batched
        // switch-overs save).
        if (first_ticket) {
          b = (long)blockIdx.x * nthreads + tid;
          first_ticket = false;
        } else {
          b = (long)atomicAdd(a.work_counter, 1ULL) + (long)gridDim.x * nthreads;
        }
        if (b >= B) break;
        if (a.prob.order != nullptr) b = (long)a.prob.order[b];  // service order (pdeq_problem.order)
        need_load = false;
        const double* tc = a.prob.tcoeffs + b * (n * D);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
          for (int j = 0; j < D; ++j) m[i][j] = tc[i * D + j];
        }
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        dt = adaptive ? a.dt0[b * a.dt0_stride] : 0.0;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) L[k][i][j] = 0.0;
          }
        }
        if (a.prob.init_std != nullptr) {
          const double* sd = a.prob.init_std + b * a.prob.init_std_stride;
#pragma unroll
          for (int k = 0; k < NB; ++k) {
#pragma unroll
            for (int i = 0; i < n; ++i) L[k][i][i] = (FACT == PDEQ_FACT_BLOCKDIAG) ? sd[i * D + k] : sd[i];
          }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          prior[k] = (!SP && a.prob.prior_scale != nullptr) ? a.prob.prior_scale[b * a.prob.prior_scale_stride + k] : 1.0;
          sig[k] = 1.0;
          run_scale[k] = 0.0;
        }
        t = a.grid[0];
        ctrl_lprev = 0.0;  // log2 of the PI controller's initial state 1.0 (controllers.py:42-44)
        ndata = 0.0;
        nsteps = 0;
        nattempts = 0;
        status = 0;
        if (!SP && cfg.constraint_init != 0) {
          // solver.init with constraint_init (solvers.py:361-372, 526-537, 670-680): condition the initial state on a
          // zero residual of the constraint linearised at it; a zero observed factor gives a zero gain (lstsq_svd)
          double h0[NB][q + 1], mobs0[D], gain0[NB][n], ry0[NB];
          linearise(m, params, t, h0, mobs0);
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            double Ln0[n][n];
            revert_obs<n, q, TS0>(L[k], h0[k], damp, ry0[k], gain0[k], Ln0);
#pragma unroll
            for (int i = 0; i < n; ++i) {
              gain0[k][i] = (ry0[k] == 0.0) ? 0.0 : gain0[k][i];
#pragma unroll
              for (int j = 0; j <= i; ++j) L[k][i][j] = Ln0[i][j];
            }
          }
          if (cfg_solver == PDEQ_SOLVER_MLE) {
            // solver_mle.init (solvers.py:361-374): the update at t0 is the first datum of the running calibration;
            // the residual is whitened by a plain division (a singular observed factor reads NaN, as there)
            whitened_rms(mobs0, ry0, run_scale);
            ndata = 1.0;
          }
#pragma unroll
          for (int j = 0; j < D; ++j) {
#pragma unroll
            for (int i = 0; i < n; ++i) m[i][j] = fma(-gain0[blk(j)][i], mobs0[j], m[i][j]);
          }
        }
        emit(a, b, 0, t, m, L, sig, 0);
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        if (needs_interp) if_store(smem_if, nthreads, tid, m, L, t);
      }

      // ------------------------------------------------------------------ checkpoint reached?
      // adaptive: RejectionLoop.loop's interpolation switch (solvers_via_adaptive_steps.py:241-247)
      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          if (needs_interp && t > t_next + a.eps) {
            // interp_beyond_t1 -> strategy_filter.interpolate_fwd: predict from interp_from to t_next
            double mi[n][D], Li[NB][n][n], mo[n][D], Lo[NB][n][n], p[n], pinv[n], t_if;
            if_load(smem_if, nthreads, tid, mi, Li, t_if);
            const double dti = t_next - t_if;
            preconditioner<n>(dti, ifact, fact, p, pinv);
            const double sq = safe_sqrt(fabs(dti));
#pragma unroll
            for (int j = 0; j < D; ++j) {
              double col[n], out[n];
#pragma unroll
              for (int i = 0; i < n; ++i) col[i] = mi[i][j];
              predict_mean<n>(col, p, pinv, A, out);
#pragma unroll
              for (int i = 0; i < n; ++i) mo[i][j] = out[i];
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) predict_chol<n>(Li[k], p, pinv, sq * prior[k] * sig[k], A, Q, Lo[k]);
            emit(a, b, ck, t_next, mo, Lo, sig, nsteps);
            if_store(smem_if, nthreads, tid, mo, Lo, t_next);
          } else {
            // interp_at_t1: the state itself is the solution; interpolation restarts from it
            emit(a, b, ck, t, m, L, sig, nsteps);
            if (needs_interp) if_store(smem_if, nthreads, tid, m, L, t);
          }
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
        }
        if (ck >= T) {
          // -------------------------------------------------------------- finish the instance
          if (cfg_solver == PDEQ_SOLVER_MLE) {
            // solver_mle.userfriendly_output (solvers.py:439-480): rescale everything by the calibrated scale
            double fin[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              fin[k] = run_scale[k];
              if (cfg.correct_asymptotic_underconfidence) fin[k] = fin[k] / sqrt((double)nsteps);
            }
            for (int c = 0; c < T; ++c) {
              const long bt = b * T + c;
              if (a.sol.chol != nullptr) {
                double* co = a.sol.chol + bt * (NB * n * n);
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                  for (int e = 0; e < n * n; ++e) co[k * n * n + e] = fin[k] * co[k * n * n + e];
                }
              }
              if (a.sol.output_scale != nullptr) {
#pragma unroll
                for (int k = 0; k < NB; ++k) a.sol.output_scale[bt * NB + k] = fin[k];
              }
            }
          }
          bool finite = true;
#pragma unroll
          for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) finite = finite && isfinite(m[i][j]);
          }
          if (status == 0 && !finite) status = PDEQ_STATUS_NONFINITE;
          a.sol.status[b] = status;
          if (a.sol.num_attempts != nullptr) a.sol.num_attempts[b] = nattempts;
          need_load = true;
        }
        continue;
      }

      // ------------------------------------------------------------------ one step attempt
      nattempts += 1;
      double dtc;
      if (adaptive) {
        dtc = clip ? fmin(dt, t_next - t) : dt;  // solvers_via_adaptive_steps.py:300-301
      } else {
        dtc = a.grid[ck] - a.grid[ck - 1];  // np.diff(grid), solvers_via_fixed_steps.py:28
      }
      double p[n], pinv[n];
      preconditioner<n>(dtc, ifact, fact, p, pinv);
      const double sq = safe_sqrt(fabs(dtc));

      // mean extrapolation (identical for transition.apply_flat and transition.marginalise)
      double mp[n][D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double col[n], out[n];
#pragma unroll
        for (int i = 0; i < n; ++i) col[i] = m[i][j];
        predict_mean<n>(col, p, pinv, A, out);
#pragma unroll
        for (int i = 0; i < n; ++i) mp[i][j] = out[i];
      }

      // linearise at the extrapolated mean
      const double t_new = t + dtc;
      double h[NB][q + 1], mobs[D];
      linearise(mp, params, t_new, h, mobs);

      // Cholesky factor of the zero-error extrapolation (process noise only), shared by solver_dynamic and
      // the error estimators
      double Lq[NB][n][n], robs[NB];
      const bool need_robs = adaptive ? (cfg_solver == PDEQ_SOLVER_DYNAMIC || cfg_error == PDEQ_ERROR_RESIDUAL_STD)
                                      : (cfg_solver == PDEQ_SOLVER_DYNAMIC);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        noise_chol<n>(p, sq * prior[k], Q, Lq[k]);
        robs[k] = need_robs ? obs_marginal_chol<n, q, TS0>(Lq[k], h[k], damp) : 1.0;
      }

      // solver_dynamic: calibrate the output scale before extrapolating (solvers.py:552-573)
      double sig_new[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) sig_new[k] = 1.0;
      if (cfg_solver == PDEQ_SOLVER_DYNAMIC) whitened_rms(mobs, robs, sig_new);

      // extrapolate the Cholesky factor and correct (strategy_filter.predict + bayes_rule)
      double Ln[NB][n][n], gain[NB][n], ry[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double Lp[n][n];
        predict_chol<n>(L[k], p, pinv, sq * prior[k] * sig_new[k], A, Q, Lp);
        revert_obs<n, q, TS0>(Lp, h[k], damp, ry[k], gain[k], Ln[k]);
      }
      double mn[n][D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int i = 0; i < n; ++i) mn[i][j] = fma(-gain[blk(j)][i], mobs[j], mp[i][j]);
      }

      // solver_mle: running RMS of the whitened residuals (solvers.py:412-424)
      double run_new[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) run_new[k] = run_scale[k];
      if (cfg_solver == PDEQ_SOLVER_MLE) {
        const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
        double term[NB];
        whitened_rms(mobs, ry, term);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const double x1 = w1 * run_scale[k], x2 = w2 * term[k];
          run_new[k] = safe_sqrt(fma(x1, x1, x2 * x2));  // hypot
        }
      }

      // ------------------------------------------------------------------ error estimate + control
      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        double err[D], ref[D];
        int kpow;
        if (cfg_error == PDEQ_ERROR_RESIDUAL_STD) {
          // solvers.py:955-968,978-991
          double se[NB];
          whitened_rms(mobs, robs, se);
#pragma unroll
          for (int j = 0; j < D; ++j) err[j] = se[blk(j)] * fabs(robs[blk(j)]);
#pragma unroll
          for (int j = 0; j < D; ++j) ref[j] = fmax(fabs(m[0][j]), fabs(mn[0][j]));
          kpow = q;
        } else {
          // solvers.py:1070-1086: Bayes rule on the zero-error extrapolation, std of one coefficient
          const int idx = SP ? 0 : cfg.derivative_idx;
          double rye[NB], sd[NB];
          if (SP || (TS0 && damp == 0.0 && cfg.err_const[0] != 0.0)) {
            // The zero-error extrapolation's factor is a constant matrix with scaled columns, and column scalings
            // commute with the triangularisation (pdeq_config.err_const): no reflector is needed at all.
            double pq = 0.0, pi = 0.0, ci = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              pq = (i == q) ? fabs(p[i]) : pq;
              pi = (i == idx) ? fabs(p[i]) : pi;
              ci = (i == idx) ? cfg.err_const[1 + i] : ci;
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              const double s = sq * prior[k];
              rye[k] = cfg.err_const[0] * (pq * s);
              sd[k] = ci * (pi * s);
            }
          } else if (idx == 0) {
            // common case: only R_Y and the first row of the corrected factor are needed, which lets the
            // compiler drop all but the first two reflectors of the (n+1)x(n+1) triangularisation
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              double Lc[n][n], g_unused[n];
              Lc[0][0] = 0.0;
              revert_obs<n, q, TS0, 0>(Lq[k], h[k], damp, rye[k], g_unused, Lc);
              sd[k] = fabs(Lc[0][0]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              double Lc[n][n], g_unused[n];
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) Lc[i][j] = 0.0;
              }
              revert_obs<n, q, TS0>(Lq[k], h[k], damp, rye[k], g_unused, Lc);
              sd[k] = row_norm<n>(Lc, idx);
            }
          }
          double se[NB];
          whitened_rms(mobs, rye, se);
#pragma unroll
          for (int j = 0; j < D; ++j) err[j] = se[blk(j)] * sd[blk(j)];
#pragma unroll
          for (int j = 0; j < D; ++j) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              if (i == idx) {
                a0 = m[i][j];
                a1 = mn[i][j];
              }
            }
            ref[j] = fmax(fabs(a0), fabs(a1));
          }
          kpow = idx;
        }
        if (!SP && cfg.error_per_unit_step) kpow += 1;
        double escale = ipow_small<n>(dtc, kpow);  // dt^k / k!
#pragma unroll
        for (int e = 0; e <= n; ++e) {
          if (e == kpow) escale *= ifact[e];
        }
        // log2 of the error norm. scale-then-rms: norm = sqrt(sum w^2 / d), so log2(norm) = log2(sum w^2 / d) / 2 and
        // the square root is never taken.
        double l2norm;
        if (cfg_norm == PDEQ_NORM_SCALE_THEN_RMS) {
          double ss = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double w = (err[j] * escale) * fast_rcp(fma(a.rtol, ref[j], a.atol));
            ss = fma(w, w, ss);
          }
          const double ms = ss * (1.0 / (double)D);
          l2norm = 0.5 * (SP ? log2_select(ms) : log2(ms));
        } else {
          // rms(error_abs) / (atol + rtol * rms(reference)); the isotropic error has size 1
          double se2 = 0.0, sr2 = 0.0;
          constexpr int ne = (FACT == PDEQ_FACT_BLOCKDIAG) ? D : 1;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double ea = err[j] * escale;
            if (j < ne) se2 = fma(ea, ea, se2);
            sr2 = fma(ref[j], ref[j], sr2);
          }
          const double norm = (safe_sqrt(se2) * rsqrt((double)ne)) * fast_rcp(fma(a.rtol, safe_sqrt(sr2) * inv_sqrt_d, a.atol));
          l2norm = SP ? log2_select(norm) : log2(norm);
        }
        // error_power = norm^(-1/n) (solvers.py:995); accept iff !(error_power < 1) (solvers_via_adaptive_steps.py:256-258)
        const double lep = neg_inv_n * l2norm;
        accept = !(lep < 0.0);

        double lratio;  // log2 of the unclipped step ratio / safety
        if (cfg.control == PDEQ_CONTROL_PI) {
          lratio = fma(cfg.exponent_integral, lep, cfg.exponent_proportional * (lep - ctrl_lprev));
          if (lep >= 0.0) ctrl_lprev = lep;
        } else {
          lratio = lep;
        }
        const double ratio = cfg.safety * (SP ? exp2_select(lratio) : exp2(lratio));
        const double sc = fmax(cfg.factor_min, fmin(ratio, cfg.factor_max));
        dt_next = sc * dtc;
        if (a.sol.trace != nullptr && nattempts <= a.sol.trace_capacity) {
          double* tr = a.sol.trace + (b * a.sol.trace_capacity + (nattempts - 1)) * 4;
          tr[0] = t;
          tr[1] = dtc;
          tr[2] = exp2(lep);
          tr[3] = accept ? 1.0 : 0.0;
        }
        if (nattempts >= max_attempts) {
          // give up on this instance: the checkpoints it never reached are NaN, not whatever the buffers held
          status = PDEQ_STATUS_MAX_ATTEMPTS;
          accept = true;
          for (int c = ck; c < T; ++c) emit_nan(a, b, c, nsteps);
          ck = T;
        }
      }

      // ------------------------------------------------------------------ commit
      dt = dt_next;
      if (accept) {
        if (needs_interp) if_store(smem_if, nthreads, tid, m, L, t);  // interp_from <- step_from (:330-338)
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
          for (int j = 0; j < D; ++j) m[i][j] = mn[i][j];
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
              L[k][i][j] = Ln[k][i][j];
            }
          }
          if (cfg_solver == PDEQ_SOLVER_DYNAMIC) sig[k] = sig_new[k];
          run_scale[k] = run_new[k];
        }
        ndata += 1.0;
        t = t_new;
        nsteps += 1;
        if (!adaptive) {
          emit(a, b, ck, t, m, L, sig, nsteps);
          ck += 1;
        }
      }
    }
  }
};

template <class VF, int NU, int FACT, int D, bool TS0, int SPEC = 0>
__global__ void __launch_bounds__(K1_THREADS, SPEC != 0 ? PDEQ_K1_SPEC_MIN_BLOCKS(SPEC) : PDEQ_K1_MIN_BLOCKS(FACT)) k1_loop_kernel(const __grid_constant__ LoopArgs a) {
  extern __shared__ double smem_if[];
  ThreadLoop<VF, NU, FACT, D, TS0, SPEC>::run(a, smem_if);
}

// Host-side test for SPEC = 1 (see ThreadLoop).
inline bool k1_spec_matches(const LoopArgs& a, bool ts0) {
  const pdeq_config& c = a.cfg;
  return ts0 && a.fixed_grid == 0 && c.clip_dt != 0 && c.solver == PDEQ_SOLVER_PLAIN &&
         c.error == PDEQ_ERROR_STATE_STD && c.derivative_idx == 0 && c.error_per_unit_step == 0 &&
         c.error_norm == PDEQ_NORM_SCALE_THEN_RMS && a.damp == 0.0 && c.err_const[0] != 0.0 &&
         a.prob.prior_scale == nullptr && c.constraint_init == 0;
}

}  // namespace pdeq
