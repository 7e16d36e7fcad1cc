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

template <int n>
PDEQ_DI double ipow_small(double x, int k) {
  double r = 1.0;
#pragma unroll
  for (int e = 0; e < n + 1; ++e) {
    if (e < k) r *= x;
  }
  return r;
}

template <class VF, int NU, int FACT, int D, bool TS0>
struct ThreadLoop {
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int NB = (FACT == PDEQ_FACT_BLOCKDIAG) ? D : 1;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  static constexpr int IF_SLOTS = n * D + NB * n * n + 1;  // interp_from: mean, chol, t
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

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem_if) {
    const pdeq_config& cfg = a.cfg;
    const double(*__restrict__ A)[PDEQ_MAX_COEFFS] = cfg.sys_a;
    const double(*__restrict__ Q)[PDEQ_MAX_COEFFS] = cfg.sys_q;
    const double* __restrict__ fact = cfg.factorials;
    const bool adaptive = a.fixed_grid == 0;
    const bool clip = cfg.clip_dt != 0;
    const bool needs_interp = adaptive && !clip;
    const int T = a.T;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
#define PDEQ_IF(slot) smem_if[(slot) * nthreads + tid]

    // ---- per-instance state (registers) ----
    double m[n][D], L[NB][n][n], prior[NB], sig[NB], run_scale[NB], params[P];
    double t = 0.0, dt = 0.0, ctrl_prev = 1.0, ndata = 0.0, t_next = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;
    bool need_load = true;

    while (true) {
      // ------------------------------------------------------------------ fetch the next instance
      if (need_load) {
        b = (long)atomicAdd(a.work_counter, 1ULL);
        if (b >= B) break;
        need_load = false;
        const double* tc = a.prob.tcoeffs + b * (n * D);
#pragma unroll
        for (int i = 0; i < n; ++i) {
#pragma unroll
          for (int j = 0; j < D; ++j) m[i][j] = tc[i * D + j];
        }
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
          prior[k] = a.prob.prior_scale != nullptr ? a.prob.prior_scale[b * a.prob.prior_scale_stride + k] : 1.0;
          sig[k] = 1.0;
          run_scale[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        t = a.grid[0];
        dt = adaptive ? a.dt0[b * a.dt0_stride] : 0.0;
        ctrl_prev = 1.0;
        ndata = 0.0;
        nsteps = 0;
        nattempts = 0;
        status = 0;
        emit(a, b, 0, t, m, L, sig, 0);
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        if (needs_interp) {
#pragma unroll
          for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) PDEQ_IF(i * D + j) = m[i][j];
          }
#pragma unroll
          for (int k = 0; k < NB; ++k) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
              for (int j = 0; j <= i; ++j) PDEQ_IF(n * D + (k * n + i) * n + j) = L[k][i][j];
            }
          }
          PDEQ_IF(IF_SLOTS - 1) = t;
        }
      }

      // ------------------------------------------------------------------ checkpoint reached?
      // adaptive: RejectionLoop.loop's interpolation switch (solvers_via_adaptive_steps.py:241-247)
      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          if (needs_interp && t > t_next + a.eps) {
            // interp_beyond_t1 -> strategy_filter.interpolate_fwd: predict from interp_from to t_next
            double mi[n][D], Li[NB][n][n], mo[n][D], Lo[NB][n][n], p[n], pinv[n];
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
              for (int j = 0; j < D; ++j) mi[i][j] = PDEQ_IF(i * D + j);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) Li[k][i][j] = PDEQ_IF(n * D + (k * n + i) * n + j);
              }
            }
            const double t_if = PDEQ_IF(IF_SLOTS - 1);
            const double dti = t_next - t_if;
            preconditioner<n>(dti, fact, p, pinv);
            const double sq = sqrt(fabs(dti));
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
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
              for (int j = 0; j < D; ++j) PDEQ_IF(i * D + j) = mo[i][j];
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) PDEQ_IF(n * D + (k * n + i) * n + j) = Lo[k][i][j];
              }
            }
            PDEQ_IF(IF_SLOTS - 1) = t_next;
          } else {
            // interp_at_t1: the state itself is the solution; interpolation restarts from it
            emit(a, b, ck, t, m, L, sig, nsteps);
            if (needs_interp) {
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j) PDEQ_IF(i * D + j) = m[i][j];
              }
#pragma unroll
              for (int k = 0; k < NB; ++k) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
#pragma unroll
                  for (int j = 0; j <= i; ++j) PDEQ_IF(n * D + (k * n + i) * n + j) = L[k][i][j];
                }
              }
              PDEQ_IF(IF_SLOTS - 1) = t;
            }
          }
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
        }
        if (ck >= T) {
          // -------------------------------------------------------------- finish the instance
          if (cfg.solver == PDEQ_SOLVER_MLE) {
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
      preconditioner<n>(dtc, fact, p, pinv);
      const double sq = sqrt(fabs(dtc));

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

      // linearise at the extrapolated mean (ts0: ssm_impl_isotropic.py:304-317 / ssm_impl_blockdiag.py:129-144;
      // ts1: ssm_impl_isotropic.py:326-355 / ssm_impl_blockdiag.py:153-183)
      const double t_new = t + dtc;
      double h[NB][q + 1], mobs[D];
      {
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

      // Cholesky factor of the zero-error extrapolation (process noise only), shared by solver_dynamic and
      // the error estimators
      double Lq[NB][n][n], robs[NB];
      const bool need_robs = adaptive ? (cfg.solver == PDEQ_SOLVER_DYNAMIC || cfg.error == PDEQ_ERROR_RESIDUAL_STD)
                                      : (cfg.solver == PDEQ_SOLVER_DYNAMIC);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        noise_chol<n>(p, sq * prior[k], Q, Lq[k]);
        robs[k] = need_robs ? obs_marginal_chol<n, q, TS0>(Lq[k], h[k], a.damp) : 1.0;
      }

      // solver_dynamic: calibrate the output scale before extrapolating (solvers.py:552-573)
      double sig_new[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) sig_new[k] = 1.0;
      if (cfg.solver == PDEQ_SOLVER_DYNAMIC) {
        if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
          for (int j = 0; j < D; ++j) sig_new[blk(j)] = fabs(mobs[j] / robs[blk(j)]);
        } else {
          double ss = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double w = mobs[j] / robs[0];
            ss = fma(w, w, ss);
          }
          sig_new[0] = sqrt(ss) / sqrt((double)D);
        }
      }

      // extrapolate the Cholesky factor and correct (strategy_filter.predict + bayes_rule)
      double Ln[NB][n][n], gain[NB][n], ry[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double Lp[n][n];
        predict_chol<n>(L[k], p, pinv, sq * prior[k] * sig_new[k], A, Q, Lp);
        revert_obs<n, q, TS0>(Lp, h[k], a.damp, ry[k], gain[k], Ln[k]);
      }
      double mn[n][D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int i = 0; i < n; ++i) mn[i][j] = mp[i][j] - gain[blk(j)][i] * mobs[j];
      }

      // solver_mle: running RMS of the whitened residuals (solvers.py:412-424)
      double run_new[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) run_new[k] = run_scale[k];
      if (cfg.solver == PDEQ_SOLVER_MLE) {
        const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
        if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double term = fabs(mobs[j] / ry[blk(j)]);
            run_new[blk(j)] = hypot(w1 * run_scale[blk(j)], w2 * term);
          }
        } else {
          double ss = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double w = mobs[j] / ry[0];
            ss = fma(w, w, ss);
          }
          run_new[0] = hypot(w1 * run_scale[0], w2 * (sqrt(ss) / sqrt((double)D)));
        }
      }

      // ------------------------------------------------------------------ error estimate + control
      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        double err[D], ref[D];
        int kpow;
        if (cfg.error == PDEQ_ERROR_RESIDUAL_STD) {
          // solvers.py:955-968,978-991
          if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
            for (int j = 0; j < D; ++j) err[j] = fabs(mobs[j] / robs[blk(j)]) * fabs(robs[blk(j)]);
          } else {
            double ss = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
              const double w = mobs[j] / robs[0];
              ss = fma(w, w, ss);
            }
            const double e = (sqrt(ss) / sqrt((double)D)) * fabs(robs[0]);
#pragma unroll
            for (int j = 0; j < D; ++j) err[j] = e;
          }
#pragma unroll
          for (int j = 0; j < D; ++j) ref[j] = fmax(fabs(m[0][j]), fabs(mn[0][j]));
          kpow = q;
        } else {
          // solvers.py:1070-1086: Bayes rule on the zero-error extrapolation, std of one coefficient
          const int idx = cfg.derivative_idx;
          double se[NB], sd[NB];
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            double Lc[n][n], g_unused[n], rye;
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
              for (int j = 0; j <= i; ++j) Lc[i][j] = 0.0;
            }
            revert_obs<n, q, TS0>(Lq[k], h[k], a.damp, rye, g_unused, Lc);
            sd[k] = row_norm<n>(Lc, idx);
            se[k] = rye;
          }
          if (FACT == PDEQ_FACT_BLOCKDIAG) {
#pragma unroll
            for (int j = 0; j < D; ++j) err[j] = fabs(mobs[j] / se[blk(j)]) * sd[blk(j)];
          } else {
            double ss = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
              const double w = mobs[j] / se[0];
              ss = fma(w, w, ss);
            }
            const double e = (sqrt(ss) / sqrt((double)D)) * sd[0];
#pragma unroll
            for (int j = 0; j < D; ++j) err[j] = e;
          }
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
        if (cfg.error_per_unit_step) kpow += 1;
        const double dtk = ipow_small<n>(dtc, kpow);
        double fk = 1.0;
#pragma unroll
        for (int e = 0; e <= n; ++e) {
          if (e == kpow) fk = fact[e];
        }
        double norm;
        if (cfg.error_norm == PDEQ_NORM_SCALE_THEN_RMS) {
          double ss = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double w = (err[j] * dtk / fk) / (a.atol + a.rtol * ref[j]);
            ss = fma(w, w, ss);
          }
          norm = sqrt(ss) / sqrt((double)D);
        } else {
          // rms(error_abs) / (atol + rtol * rms(reference)); the isotropic error has size 1
          double se2 = 0.0, sr2 = 0.0;
          const int ne = (FACT == PDEQ_FACT_BLOCKDIAG) ? D : 1;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double ea = err[j] * dtk / fk;
            if (j < ne) se2 = fma(ea, ea, se2);
            sr2 = fma(ref[j], ref[j], sr2);
          }
          norm = (sqrt(se2) / sqrt((double)ne)) / (a.atol + a.rtol * (sqrt(sr2) / sqrt((double)D)));
        }
        const double ep = pow(norm, -1.0 / (double)n);  // solvers.py:995
        accept = !(ep < 1.0);                            // solvers_via_adaptive_steps.py:256-258

        double ratio;
        if (cfg.control == PDEQ_CONTROL_PI) {
          const double gi = pow(ep, cfg.exponent_integral);
          const double gp = pow(ep / ctrl_prev, cfg.exponent_proportional);
          ratio = cfg.safety * gi * gp;
          if (ep >= 1.0) ctrl_prev = ep;
        } else {
          ratio = cfg.safety * ep;
        }
        const double sc = fmax(cfg.factor_min, fmin(ratio, cfg.factor_max));
        dt_next = sc * dtc;
        if (nattempts >= max_attempts) {
          status = PDEQ_STATUS_MAX_ATTEMPTS;
          accept = true;
          // give up on this instance: jump beyond every remaining checkpoint
          ck = T;
        }
      }

      // ------------------------------------------------------------------ commit
      dt = dt_next;
      if (accept) {
        if (needs_interp) {  // interp_from <- step_from (solvers_via_adaptive_steps.py:330-338)
#pragma unroll
          for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) PDEQ_IF(i * D + j) = m[i][j];
          }
#pragma unroll
          for (int k = 0; k < NB; ++k) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
#pragma unroll
              for (int j = 0; j <= i; ++j) PDEQ_IF(n * D + (k * n + i) * n + j) = L[k][i][j];
            }
          }
          PDEQ_IF(IF_SLOTS - 1) = t;
        }
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
            for (int j = 0; j <= i; ++j) L[k][i][j] = Ln[k][i][j];
          }
          if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig[k] = sig_new[k];
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
#undef PDEQ_IF
  }
};

template <class VF, int NU, int FACT, int D, bool TS0>
__global__ void __launch_bounds__(K1_THREADS) k1_loop_kernel(const __grid_constant__ LoopArgs a) {
  extern __shared__ double smem_if[];
  ThreadLoop<VF, NU, FACT, D, TS0>::run(a, smem_if);
}

}  // namespace pdeq
