// K2: one IVP instance per warp (d <= 32) or per CTA (d <= 1024), ONE ODE DIMENSION PER LANE (a lane of a CTA
// serves up to 4 dimensions in turn when d > 256).
//
// The block-diagonal factorisation (probdiffeq/_probdiffeq/ssm_impl_blockdiag.py) is d independent (n x n)
// filters/smoothers that are coupled only through (i) the vector field, (ii) the error norm and (iii) the shared
// step size.  Each lane therefore runs the register-resident block algebra of pdeq_blockops.cuh for its own
// dimension; the extrapolated state u is exchanged through shared memory for the vector-field evaluation, and
// the error norm is a warp-shuffle / shared-memory reduction.  The isotropic factorisation
// (ssm_impl_isotropic.py) runs on the same kernel: every lane carries an identical copy of the shared factor,
// and calibration / error scales become rms-reductions over the lanes.
//
// Besides the filter this kernel implements the fixed-point smoother
// (probdiffeq/_probdiffeq/estimators_and_losses.py:473-591; Smoother.finalize :437-470; MarkovSequence
// .evaluate_marginals :156-178): the backward conditional is carried per lane, merged at every step, stored per
// checkpoint in a per-group scratch ring in global memory, and marginalised backwards when the instance ends.
//
// Loop structure, accept/reject, interpolation branches, calibration and error estimation restate the same
// reference lines as K1 (see pdeq_loop_thread.cuh).  The accepted state lives in shared memory in a coalesced
// [field][dim] layout (Burgers d = 1024, nu = 3: 17 fields x 1024 x 8 B = 136 KB per instance, one CTA per SM);
// a proposal lives in registers (one dimension per lane) or thread-local scratch (several).
#pragma once

#include "pdeq_blockops_smoother.cuh"
#include "pdeq_limits.cuh"
#include "pdeq_loop_thread.cuh"

namespace pdeq {

#ifndef PDEQ_K2_WARP_FILTER_BLOCKS
#define PDEQ_K2_WARP_FILTER_BLOCKS 3  // resident 128-thread CTAs per SM of the warp-mode filter kernels (<= 168 registers)
#endif

// MODE 0: a warp per instance (d <= 32). MODE 1: a CTA per instance, one dimension per lane (d <= 256).
// MODE 2: a CTA per instance, up to K2_MAX_DPL dimensions per lane (d <= 1024); its proposal arrays cost ~160
// registers, which is why the one-dimension-per-lane case has its own instantiation (two CTAs per SM instead of one).
//
// SPEC = 0: solver / error estimator are run-time (group-uniform) branches on pdeq_config.
// SPEC = 1: `solver_dynamic` + `error_residual_std` (BASELINE configs 3 and 4a's combination) fixed at compile time.
//           Same arithmetic, operation for operation; what goes away is what the run-time branches keep alive across
//           the extrapolation -- above all the full noise-only factor Lq (n (n+1) / 2 doubles per lane) that only
//           error_state_std's general path reads -- and the code of the unused estimators (instruction cache).
template <class VF, int NU, int FACT, bool TS0, bool FP, int MODE, int SPEC = 0>
struct GroupLoop {
  static constexpr bool SPD = SPEC >= 1;
  static constexpr bool CTA = MODE != 0;
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int TRI = n * (n + 1) / 2;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  static constexpr bool ISO = FACT == PDEQ_FACT_ISOTROPIC;
  // fields of one stored state, each a [d] row in shared memory
  static constexpr int F_M = 0, F_L = n, F_G = n + TRI, F_XI = F_G + n * n, F_XIC = F_XI + n, F_TL = F_XIC + TRI,
                       F_TO = F_TL + n;
  static constexpr int F_SIG = FP ? F_TO + n : n + TRI;  // per-dimension scalars: output scale, running mle
  static constexpr int F_RUN = F_SIG + 1, F_PRIOR = F_RUN + 1;  // scale, prior scale
  static constexpr int NF = F_PRIOR + 1;
  static constexpr int NFC = n * n + n + TRI + 2 * n;  // fields of a stored conditional
  static constexpr int MAXR = (MODE == 2 && !FP && !ISO) ? K2_MAX_DPL : 1;
  static_assert(q < n, "need more Taylor coefficients than the ODE order");

  struct Group {
    int lane, size;
    double* red;  // CTA mode: 32 doubles of scratch
    PDEQ_DI void sync() const {
      if (CTA) __syncthreads();
      else __syncwarp();
    }
    // Sum over the group; every lane receives the bitwise-identical result.
    PDEQ_DI double sum(double v) const {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (CTA) {
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[w] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < nw; ++i) t += red[i];
        v = t;
      }
      return v;
    }
    PDEQ_DI long bcast_long(long v) const {
      if (CTA) {
        __syncthreads();
        if (threadIdx.x == 0) reinterpret_cast<long*>(red)[0] = v;
        __syncthreads();
        v = reinterpret_cast<long*>(red)[0];
        __syncthreads();
        return v;
      }
      return __shfl_sync(0xffffffffu, v, 0);
    }
  };

  struct ExchAcc {  // jet coordinates of the extrapolated state, one row per coordinate
    const double* u;
    int d;
    PDEQ_DI double operator()(int k, int i) const { return u[k * d + i]; }
  };

  PDEQ_DI static void st_store(double* st, int d, int j, const double (&m)[n], const double (&L)[n][n]) {
#pragma unroll
    for (int i = 0; i < n; ++i) st[(F_M + i) * d + j] = m[i];
    int e = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int c = 0; c <= i; ++c) st[(F_L + e++) * d + j] = L[i][c];
    }
  }
  PDEQ_DI static void st_load(const double* st, int d, int j, double (&m)[n], double (&L)[n][n]) {
#pragma unroll
    for (int i = 0; i < n; ++i) m[i] = st[(F_M + i) * d + j];
    int e = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int c = 0; c < n; ++c) L[i][c] = (c <= i) ? st[(F_L + e++) * d + j] : 0.0;
    }
  }
  // conditionals: `base` points at field 0 of the conditional, fields strided by d
  PDEQ_DI static void cond_store(double* base, int d, int j, const BlockCond<n>& c) {
    int e = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) base[(e++) * d + j] = c.G[i][k];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) base[(e++) * d + j] = c.xi[i];
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k <= i; ++k) base[(e++) * d + j] = c.Xi[i][k];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) base[(e++) * d + j] = c.tl[i];
#pragma unroll
    for (int i = 0; i < n; ++i) base[(e++) * d + j] = c.to[i];
  }
  PDEQ_DI static void cond_load(const double* base, int d, int j, BlockCond<n>& c) {
    int e = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) c.G[i][k] = base[(e++) * d + j];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) c.xi[i] = base[(e++) * d + j];
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) c.Xi[i][k] = (k <= i) ? base[(e++) * d + j] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) c.tl[i] = base[(e++) * d + j];
#pragma unroll
    for (int i = 0; i < n; ++i) c.to[i] = base[(e++) * d + j];
  }

  // One backward conditional of the posterior (MarkovSequence.conditional after rescale_cholesky,
  // estimators_and_losses.py:151-154), written in natural coordinates (LatentCond.preconditioner_apply,
  // ssm_impl_blockdiag.py:86-96): gain = T_o G T_l, mean = T_o xi, chol = scale |T_o| Xi.
  PDEQ_DI static void emit_conditional(const LoopArgs& a, long bt, int d, int j, const BlockCond<n>& c, double scale) {
#pragma unroll
    for (int i = 0; i < n; ++i) a.sol.bw_mean[(bt * n + i) * (long)d + j] = c.to[i] * c.xi[i];
    if (ISO && j != 0) return;
    const long off = ISO ? bt * (long)(n * n) : (bt * d + j) * (long)(n * n);
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < n; ++k) {
        a.sol.bw_gain[off + i * n + k] = c.to[i] * c.G[i][k] * c.tl[k];
        a.sol.bw_chol[off + i * n + k] = (k <= i) ? scale * fabs(c.to[i]) * c.Xi[i][k] : 0.0;
      }
    }
  }

  // SmoothingSolution.filtering (estimators_and_losses.py:464-467): before the smoothing marginal replaces it, the
  // forward pass's marginal of this lane's dimension moves to filt_mean / filt_chol, calibrated.
  PDEQ_DI static void keep_filtering(const LoopArgs& a, long bt, int d, int j, double scale) {
#pragma unroll
    for (int i = 0; i < n; ++i) a.sol.filt_mean[(bt * n + i) * (long)d + j] = a.sol.mean[(bt * n + i) * (long)d + j];
    if (a.sol.filt_chol == nullptr || a.sol.chol == nullptr || (ISO && j != 0)) return;
    const long off = ISO ? bt * (long)(n * n) : (bt * d + j) * (long)(n * n);
    for (int e = 0; e < n * n; ++e) a.sol.filt_chol[off + e] = scale * a.sol.chol[off + e];
  }

  PDEQ_DI static void write_chol(const LoopArgs& a, long bt, int d, int j, const double (&L)[n][n], double scale) {
    if (a.sol.chol == nullptr || (ISO && j != 0)) return;
    double* co = ISO ? a.sol.chol + bt * (long)(n * n) : a.sol.chol + (bt * d + j) * (long)(n * n);
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int c = 0; c < n; ++c) co[i * n + c] = (c <= i) ? scale * L[i][c] : 0.0;
    }
  }

  // one checkpoint of the solution for dimension j
  PDEQ_DI static void emit(const LoopArgs& a, long b, int ck, int d, int j, double t, const double (&m)[n],
                           const double (&L)[n][n], double scale, int nsteps) {
    const long bt = b * a.T + ck;
    if (j == 0) {
      a.sol.t[bt] = t;
      a.sol.num_steps[bt] = nsteps;
    }
    double* mo = a.sol.mean + bt * (long)(n * d);
#pragma unroll
    for (int i = 0; i < n; ++i) mo[i * d + j] = m[i];
    write_chol(a, bt, d, j, L, 1.0);
    if (a.sol.output_scale != nullptr) {
      if (ISO) {
        if (j == 0) a.sol.output_scale[bt] = scale;
      } else {
        a.sol.output_scale[bt * d + j] = scale;
      }
    }
  }

  // rms over the group of w_j (isotropic) or |w_j| per lane (block-diagonal)
  PDEQ_DI static double whitened(const Group& g, double w, bool active, double inv_sqrt_d) {
    if (ISO) {
      const double ss = g.sum(active ? w * w : 0.0);
      return safe_sqrt(ss) * inv_sqrt_d;
    }
    return fabs(w);
  }

  PDEQ_DI static void cond_copy(double* dst, const double* src, int d, int j) {
#pragma unroll
    for (int e = 0; e < NFC; ++e) dst[e * d + j] = src[e * d + j];
  }
  PDEQ_DI static void cond_store_identity(double* dst, int d, int j) {
    BlockCond<n> ident;
    cond_identity<n>(ident);
    cond_store(dst, d, j, ident);
  }

  PDEQ_DI static void extrapolate(const double (&L)[n][n], const double (&m)[n], const double (&p)[n],
                                  const double (&pinv)[n], double s, const double (*A)[PDEQ_MAX_COEFFS],
                                  const double (*Q)[PDEQ_MAX_COEFFS], double (&Lp)[n][n]) {
    predict_chol<n>(L, p, pinv, s, A, Q, Lp);
  }
  // Smoother (strategy_smoother_fixedpoint.predict, estimators_and_losses.py:526-532): the transition is reverted and
  // the new backward conditional composed with the carried one. Both halves live in pdeq_blockops_smoother.cuh:
  // revert_push (with every attempt; parks the new conditional in the lane's working column) and remainder_merge (for
  // accepted steps; overwrites the carried conditional in place).
  static constexpr int NFW = FP ? SmootherScratch<n>::NFW : 0;  // fields of a working column

  // For the smoother the interp_from copy of the state (written on every accepted step, read only when a checkpoint
  // is overstepped) lives in an L2-resident global scratch slot instead of shared memory: it halves the shared
  // memory per instance and doubles the number of resident warps (4 -> 8 per SM for Pleiades).
  static constexpr bool IF_GLOBAL = FP;
  // wk_smem: the smoother's working columns live in shared memory (warp mode: per PDEQ_K2_WARP_WK_GLOBAL; CTA mode:
  // when they fit), otherwise in an L2-resident global slot per resident group.
  PDEQ_HDI static constexpr size_t smem_doubles_per_group(int d, bool needs_interp, bool wk_smem = true) {
    return (size_t)((needs_interp && !IF_GLOBAL) ? 2 : 1) * NF * d + (size_t)q * d + 32 +
           (wk_smem ? (size_t)NFW * d : 0);
  }

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem, double* __restrict__ cond_ring,
                          double* __restrict__ if_scratch, double* __restrict__ wk_scratch, int groups_per_cta) {
    const pdeq_config& cfg = a.cfg;
    const double(*__restrict__ A)[PDEQ_MAX_COEFFS] = cfg.sys_a;
    const double(*__restrict__ Q)[PDEQ_MAX_COEFFS] = cfg.sys_q;
    const double* __restrict__ fact = cfg.factorials;
    const double* __restrict__ ifact = cfg.inv_factorials;
    const bool adaptive = a.fixed_grid == 0;
    const bool clip = cfg.clip_dt != 0;
    const bool needs_interp = adaptive && !clip;
    const int cfg_solver = SPD ? (int)PDEQ_SOLVER_DYNAMIC : cfg.solver;
    const int cfg_error = SPD ? (int)PDEQ_ERROR_RESIDUAL_STD : cfg.error;
    const bool per_unit_step = SPD ? false : cfg.error_per_unit_step != 0;
    const int T = a.T;
    // a vector field of fixed dimension (validated by the host) makes every [field][dim] offset a compile-time constant
    const int d = VF::fixed_dim > 0 ? VF::fixed_dim : cfg.ode_dim;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const double inv_sqrt_d = rsqrt((double)d);
    const double neg_inv_n = -1.0 / (double)n;

    const int gidx = CTA ? 0 : (threadIdx.x >> 5);
    Group g;
    g.lane = CTA ? threadIdx.x : (threadIdx.x & 31);
    g.size = CTA ? blockDim.x : 32;
    const int nrounds = (d + g.size - 1) / g.size;  // <= MAXR, enforced by the launcher

    // shared memory of this group: [state_from | interp_from | exchange u | reduction scratch]
    const size_t per_group = smem_doubles_per_group(d, needs_interp, wk_scratch == nullptr);
    const size_t group_slot = (size_t)blockIdx.x * groups_per_cta + gidx;
    double* base = smem + (size_t)gidx * per_group;
    double* st_from = base;
    double* st_if = IF_GLOBAL ? if_scratch + group_slot * (size_t)NF * d : base + NF * d;
    double* exch = base + ((needs_interp && !IF_GLOBAL) ? 2 : 1) * NF * d;
    g.red = exch + q * d;
    // The smoother's working columns: shared memory, or an L2-resident global slot per resident group. Which one is
    // said at compile time for the warp mode (pdeq_limits.cuh: PDEQ_K2_WARP_WK_GLOBAL), so that its accesses are
    // LDS / STS or LDG / STG, never generic; the CTA mode goes global only when shared memory cannot hold them.
#if PDEQ_K2_WARP_WK_GLOBAL
    double* wk = !FP ? nullptr
                     : ((!CTA || wk_scratch != nullptr) ? wk_scratch + group_slot * (size_t)NFW * d : g.red + 32);
#else
    double* wk = !FP ? nullptr
                     : ((CTA && wk_scratch != nullptr) ? wk_scratch + group_slot * (size_t)NFW * d : g.red + 32);
#endif
    // global scratch ring for the per-checkpoint conditionals of the instance this group is working on
    double* ring = FP ? cond_ring + group_slot * (size_t)T * NFC * d : nullptr;

    double params[P];
    double t = 0.0, dt = 0.0, ctrl_lprev = 0.0, ndata = 0.0, t_next = 0.0, t_if = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;
    bool need_load = true;

    // thread-local proposal scratch (registers when a lane serves one dimension)
    double pm[MAXR][n], pL[MAXR][n][n], psig[MAXR], prun[MAXR];

    while (true) {
      // ------------------------------------------------------------------ fetch the next instance
      if (need_load) {
        long nb = 0;
        if (g.lane == 0) nb = (long)atomicAdd(a.work_counter, 1ULL);
        b = g.bcast_long(nb);
        if (b >= B) break;
        if (a.prob.order != nullptr) b = (long)a.prob.order[b];  // service order (pdeq_problem.order)
        need_load = false;
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        t = a.grid[0];
        for (int r = 0; r < nrounds; ++r) {
          const int j = g.lane + r * g.size;
          if (j >= d) continue;
          double m[n], L[n][n];
#pragma unroll
          for (int i = 0; i < n; ++i) {
            m[i] = a.prob.tcoeffs[(b * n + i) * d + j];
#pragma unroll
            for (int c = 0; c < n; ++c) L[i][c] = 0.0;
          }
          if (a.prob.init_std != nullptr) {
            const double* sd = a.prob.init_std + b * a.prob.init_std_stride;
#pragma unroll
            for (int i = 0; i < n; ++i) L[i][i] = ISO ? sd[i] : sd[i * d + j];
          }
          double prior = 1.0;
          if (a.prob.prior_scale != nullptr)
            prior = a.prob.prior_scale[b * a.prob.prior_scale_stride + (ISO ? 0 : j)];
          st_store(st_from, d, j, m, L);
          st_from[F_SIG * d + j] = 1.0;
          st_from[F_RUN * d + j] = 0.0;
          st_from[F_PRIOR * d + j] = prior;
          if (FP) {
            BlockCond<n> c;
            cond_identity<n>(c);
            cond_store(st_from + F_G * d, d, j, c);
            if (needs_interp) cond_store(st_if + F_G * d, d, j, c);
          }
        }
        if (cfg.constraint_init != 0) {
          // solver.init with constraint_init (solvers.py:361-372, 526-537, 670-680): condition the initial state on a
          // zero residual of the constraint linearised at it; a zero observed factor gives a zero gain (lstsq_svd)
          g.sync();
          for (int r = 0; r < nrounds; ++r) {
            const int j = g.lane + r * g.size;
            if (j >= d) continue;
#pragma unroll
            for (int c = 0; c < q; ++c) exch[c * d + j] = st_from[(F_M + c) * d + j];
          }
          g.sync();
          for (int r = 0; r < MAXR; ++r) {
            if (r >= nrounds) break;
            const int j = g.lane + r * g.size;
            const bool active = j < d;
            if (!active && !ISO) continue;
            const int jj = active ? j : 0;
            double m[n], L[n][n], h[q + 1], mobs;
            st_load(st_from, d, jj, m, L);
            ExchAcc acc{exch, d};
            const double f = VF::template component<double>(jj, d, acc, params, t);
            if (TS0) {
#pragma unroll
              for (int c = 0; c <= q; ++c) h[c] = (c == q) ? 1.0 : 0.0;
              mobs = m[q] + (-f);
            } else {
#pragma unroll
              for (int c = 0; c < q; ++c) {
                const double jd = -VF::jac(jj, c, jj, d, acc, params, t);
                h[c] = ISO ? g.sum(active ? jd : 0.0) / (double)d : jd;
              }
              h[q] = 1.0;
              const double rres = m[q] - f;
              double hm = 0.0;
#pragma unroll
              for (int c = 0; c <= q; ++c) hm = fma(h[c], m[c], hm);
              mobs = hm + (rres - hm);
            }
            double ry0, gain0[n], Ln0[n][n];
            revert_obs<n, q, TS0>(L, h, a.damp, ry0, gain0, Ln0);
#pragma unroll
            for (int i = 0; i < n; ++i) m[i] = (ry0 == 0.0) ? m[i] : fma(-gain0[i], mobs, m[i]);
            if (active) st_store(st_from, d, j, m, Ln0);
            if (cfg_solver == PDEQ_SOLVER_MLE) {
              // solver_mle.init (solvers.py:361-374): the update at t0 is the first datum of the running calibration
              const double term0 = whitened(g, mobs * fast_rcp(ry0), active, inv_sqrt_d);
              if (active) st_from[F_RUN * d + j] = term0;
            }
          }
          g.sync();
        }
        for (int r = 0; r < nrounds; ++r) {
          const int j = g.lane + r * g.size;
          if (j >= d) continue;
          double m[n], L[n][n];
          st_load(st_from, d, j, m, L);
          emit(a, b, 0, d, j, t, m, L, 1.0, 0);
          if (needs_interp) st_store(st_if, d, j, m, L);
        }
        dt = adaptive ? a.dt0[b * a.dt0_stride] : 0.0;
        ctrl_lprev = 0.0;
        ndata = (cfg.constraint_init != 0 && cfg_solver == PDEQ_SOLVER_MLE) ? 1.0 : 0.0;
        nsteps = 0;
        nattempts = 0;
        status = 0;
        t_if = t;
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        g.sync();
      }

      // ------------------------------------------------------------------ checkpoint reached?
      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          const bool beyond = needs_interp && t > t_next + a.eps;
          for (int r = 0; r < nrounds; ++r) {
            const int j = g.lane + r * g.size;
            if (j >= d) continue;
            const double sig = st_from[F_SIG * d + j], prior = st_from[F_PRIOR * d + j];
            if (beyond) {
              // interp_beyond_t1 (solvers_via_adaptive_steps.py:346-360 -> solvers.py:205-269)
              double mi[n], Li[n][n], mo[n], Lo[n][n], p[n], pinv[n];
              st_load(st_if, d, j, mi, Li);
              const double dt0_ = t_next - t_if;
              preconditioner<n>(dt0_, ifact, fact, p, pinv);
              predict_mean<n>(mi, p, pinv, A, mo);
              if (FP) {  // the conditional checkpoint ck -> ck - 1: merged in place in its ring slot
                double* slot = ring + (size_t)ck * NFC * d;
                revert_push<n>(st_if + j, d, p, pinv, safe_sqrt(fabs(dt0_)) * prior * sig, A, Q, wk + j, Lo);
                cond_copy(slot, st_if + F_G * d, d, j);
                remainder_merge<n>(slot + j, d, p, pinv, wk + j);
              } else {
                extrapolate(Li, mi, p, pinv, safe_sqrt(fabs(dt0_)) * prior * sig, A, Q, Lo);
              }
              emit(a, b, ck, d, j, t_next, mo, Lo, sig, nsteps);
              if (FP) {
                // second half: from the interpolated point to the overstepped state, with a fresh backward
                // model (estimators_and_losses.py:549-591)
                double p1[n], pinv1[n], Ltmp[n][n];
                const double dt1_ = t - t_next;
                preconditioner<n>(dt1_, ifact, fact, p1, pinv1);
                st_store(st_if, d, j, mo, Lo);
                revert_push<n>(st_if + j, d, p1, pinv1, safe_sqrt(fabs(dt1_)) * prior * sig, A, Q, wk + j, Ltmp);
                cond_store_identity(st_from + F_G * d, d, j);
                remainder_merge<n>(st_from + F_G * d + j, d, p1, pinv1, wk + j);
                cond_store_identity(st_if + F_G * d, d, j);
              } else {
                st_store(st_if, d, j, mo, Lo);
              }
            } else {
              // interp_at_t1 (solvers_via_adaptive_steps.py:362-375 -> solvers.py:271-315)
              double m[n], L[n][n];
              st_load(st_from, d, j, m, L);
              emit(a, b, ck, d, j, t, m, L, sig, nsteps);
              if (FP) {
                cond_copy(ring + (size_t)ck * NFC * d, st_from + F_G * d, d, j);
                cond_store_identity(st_from + F_G * d, d, j);
                if (needs_interp) cond_store_identity(st_if + F_G * d, d, j);
              }
              if (needs_interp) st_store(st_if, d, j, m, L);
            }
          }
          t_if = beyond ? t_next : t;
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
        }
        if (ck >= T) {
          // -------------------------------------------------------------- finish the instance
          double bad = 0.0;
          for (int r = 0; r < nrounds; ++r) {
            const int j = g.lane + r * g.size;
            if (j >= d) continue;
            double m[n], L[n][n];
            st_load(st_from, d, j, m, L);
#pragma unroll
            for (int i = 0; i < n; ++i) bad += isfinite(m[i]) ? 0.0 : 1.0;
            double fin = 1.0;
            if (cfg_solver == PDEQ_SOLVER_MLE) {
              // solver_mle.userfriendly_output (solvers.py:439-480)
              fin = st_from[F_RUN * d + j];
              if (cfg.correct_asymptotic_underconfidence) fin = fin / sqrt((double)nsteps);
            }
            if (FP && status == 0) {
              // Smoother.finalize (estimators_and_losses.py:437-470): marginalise the overstepped state back to
              // the last checkpoint, then run the backward recursion over the stored conditionals.
              double mo[n], Lo[n][n];
              BlockCond<n> c;
              // On a fixed grid the carried conditional is the identity (the aligned pass); the reference hands
              // finalize the last grid state itself, whose conditional is the one stored for the last interval.
              if (!adaptive && cfg.strategy == PDEQ_STRATEGY_FIXEDINTERVAL && T > 1)
                cond_load(ring + (size_t)(T - 1) * NFC * d, d, j, c);
              else
                cond_load(st_from + F_G * d, d, j, c);
              cond_marginalise<n>(c, m, L, mo, Lo);
              for (int k = T - 1; k >= 0; --k) {
                const long bt = b * T + k;
                double* mout = a.sol.mean + bt * (long)(n * d);
                if (a.sol.filt_mean != nullptr) keep_filtering(a, bt, d, j, fin);
#pragma unroll
                for (int i = 0; i < n; ++i) mout[i * d + j] = mo[i];
                write_chol(a, bt, d, j, Lo, fin);
                if (k > 0) {
                  cond_load(ring + (size_t)k * NFC * d, d, j, c);
                  if (a.sol.bw_gain != nullptr) emit_conditional(a, bt, d, j, c, fin);
#pragma unroll
                  for (int i = 0; i < n; ++i) {
                    m[i] = mo[i];
#pragma unroll
                    for (int cc = 0; cc < n; ++cc) L[i][cc] = Lo[i][cc];
                  }
                  cond_marginalise<n>(c, m, L, mo, Lo);
                }
              }
            } else if (cfg_solver == PDEQ_SOLVER_MLE && a.sol.chol != nullptr && (!ISO || j == 0)) {
              for (int k = 0; k < T; ++k) {
                const long bt = b * T + k;
                double* co = ISO ? a.sol.chol + bt * (long)(n * n) : a.sol.chol + (bt * d + j) * (long)(n * n);
                for (int e = 0; e < n * n; ++e) co[e] = fin * co[e];
              }
            }
            if (cfg_solver == PDEQ_SOLVER_MLE && a.sol.output_scale != nullptr) {
              for (int k = 0; k < T; ++k) {
                if (ISO) {
                  if (j == 0) a.sol.output_scale[b * T + k] = fin;
                } else {
                  a.sol.output_scale[(b * T + k) * d + j] = fin;
                }
              }
            }
          }
          bad = g.sum(bad);
          if (status == 0 && bad > 0.0) status = PDEQ_STATUS_NONFINITE;
          if (g.lane == 0) {
            a.sol.status[b] = status;
            if (a.sol.num_attempts != nullptr) a.sol.num_attempts[b] = nattempts;
          }
          need_load = true;
        }
        g.sync();
        continue;
      }

      // ------------------------------------------------------------------ one step attempt
      nattempts += 1;
      double dtc;
      if (adaptive) {
        dtc = clip ? fmin(dt, t_next - t) : dt;
      } else {
        dtc = a.grid[ck] - a.grid[ck - 1];
      }
      double p[n], pinv[n];
      preconditioner<n>(dtc, ifact, fact, p, pinv);
      const double sq = safe_sqrt(fabs(dtc));
      const double t_new = t + dtc;

      // phase A: extrapolate the means and publish the jet coordinates the vector field reads
      g.sync();  // earlier readers of the exchange buffer are done
      for (int r = 0; r < nrounds; ++r) {
        const int j = g.lane + r * g.size;
        if (j >= d) continue;
        double m[n], mp[n];
#pragma unroll
        for (int i = 0; i < n; ++i) m[i] = st_from[(F_M + i) * d + j];
        predict_mean<n>(m, p, pinv, A, mp);
#pragma unroll
        for (int c = 0; c < q; ++c) exch[c * d + j] = mp[c];
      }
      g.sync();

      // phase B: per dimension -- linearise, calibrate, extrapolate, correct, local error
      const bool need_robs = adaptive ? (cfg_solver == PDEQ_SOLVER_DYNAMIC || cfg_error == PDEQ_ERROR_RESIDUAL_STD)
                                      : (cfg_solver == PDEQ_SOLVER_DYNAMIC);
      int kpow = (cfg_error == PDEQ_ERROR_RESIDUAL_STD) ? q : cfg.derivative_idx;
      if (per_unit_step) kpow += 1;
      double escale = ipow_small<n>(dtc, kpow);
#pragma unroll
      for (int e = 0; e <= n; ++e) {
        if (e == kpow) escale *= ifact[e];
      }
      double acc_w2 = 0.0, acc_e2 = 0.0, acc_r2 = 0.0;  // error-norm partial sums of this lane
      for (int r = 0; r < MAXR; ++r) {
        if (r >= nrounds) break;
        const int j = g.lane + r * g.size;
        const bool active = j < d;
        if (!active && !ISO) continue;  // isotropic reductions need every lane (one round only)
        const int jj = active ? j : 0;
        double m[n], L[n][n], mp[n];
        st_load(st_from, d, jj, m, L);
        const double prior = st_from[F_PRIOR * d + jj], run_scale = st_from[F_RUN * d + jj];
        predict_mean<n>(m, p, pinv, A, mp);

        double h[q + 1], mobs;
        {
          ExchAcc acc{exch, d};
          const double f = VF::template component<double>(jj, d, acc, params, t_new);
          if (TS0) {
#pragma unroll
            for (int c = 0; c <= q; ++c) h[c] = (c == q) ? 1.0 : 0.0;
            mobs = mp[q] + (-f);
          } else {
#pragma unroll
            for (int c = 0; c < q; ++c) {
              const double jd = -VF::jac(jj, c, jj, d, acc, params, t_new);
              h[c] = ISO ? g.sum(active ? jd : 0.0) / (double)d : jd;
            }
            h[q] = 1.0;
            const double rres = mp[q] - f;
            double hm = 0.0;
#pragma unroll
            for (int c = 0; c <= q; ++c) hm = fma(h[c], mp[c], hm);
            mobs = hm + (rres - hm);
          }
        }

        double Lq[n][n];
        noise_chol<n>(p, sq * prior, Q, Lq);
        const double robs = need_robs ? obs_marginal_chol<n, q, TS0>(Lq, h, a.damp) : 1.0;
        double sig_new = 1.0;
        if (cfg_solver == PDEQ_SOLVER_DYNAMIC) sig_new = whitened(g, mobs * fast_rcp(robs), active, inv_sqrt_d);

        double Lp[n][n], Ln[n][n], gain[n], ry, mn[n];
        if (FP) {
          revert_push<n>(st_from + jj, d, p, pinv, sq * prior * sig_new, A, Q, wk + jj, Lp);
        } else {
          extrapolate(L, m, p, pinv, sq * prior * sig_new, A, Q, Lp);
        }
        revert_obs<n, q, TS0>(Lp, h, a.damp, ry, gain, Ln);
#pragma unroll
        for (int i = 0; i < n; ++i) mn[i] = fma(-gain[i], mobs, mp[i]);

        double run_new = run_scale;
        if (cfg_solver == PDEQ_SOLVER_MLE) {
          const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
          const double term = whitened(g, mobs * fast_rcp(ry), active, inv_sqrt_d);
          const double x1 = w1 * run_scale, x2 = w2 * term;
          run_new = safe_sqrt(fma(x1, x1, x2 * x2));
        }

        if (adaptive) {
          double err, ref;
          if (cfg_error == PDEQ_ERROR_RESIDUAL_STD) {
            // (SPEC = 1: the calibration above already holds this very quantity)
            err = (SPD ? sig_new : whitened(g, mobs * fast_rcp(robs), active, inv_sqrt_d)) * fabs(robs);
            ref = fmax(fabs(m[0]), fabs(mn[0]));
          } else {
            const int idx = cfg.derivative_idx;
            double rye, sd;
            if (TS0 && a.damp == 0.0 && cfg.err_const[0] != 0.0) {
              // constant matrix with scaled columns: no triangularisation needed (pdeq_config.err_const)
              double pq = 0.0, pi = 0.0, ci = 0.0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                pq = (i == q) ? fabs(p[i]) : pq;
                pi = (i == idx) ? fabs(p[i]) : pi;
                ci = (i == idx) ? cfg.err_const[1 + i] : ci;
              }
              const double s = sq * prior;
              rye = cfg.err_const[0] * (pq * s);
              sd = ci * (pi * s);
            } else {
              double Lc[n][n], g_unused[n];
#pragma unroll
              for (int i = 0; i < n; ++i) {
#pragma unroll
                for (int c = 0; c <= i; ++c) Lc[i][c] = 0.0;
              }
              revert_obs<n, q, TS0>(Lq, h, a.damp, rye, g_unused, Lc);
              sd = row_norm<n>(Lc, idx);
            }
            err = whitened(g, mobs * fast_rcp(rye), active, inv_sqrt_d) * sd;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              if (i == idx) {
                a0 = m[i];
                a1 = mn[i];
              }
            }
            ref = fmax(fabs(a0), fabs(a1));
          }
          if (active) {
            const double ea = err * escale;
            const double w = ea * fast_rcp(fma(a.rtol, ref, a.atol));
            acc_w2 = fma(w, w, acc_w2);
            acc_e2 = fma(ea, ea, acc_e2);
            acc_r2 = fma(ref, ref, acc_r2);
          }
        }
        // stash the proposal
#pragma unroll
        for (int i = 0; i < n; ++i) {
          pm[r][i] = mn[i];
#pragma unroll
          for (int c = 0; c < n; ++c) pL[r][i][c] = Ln[i][c];
        }
        psig[r] = sig_new;
        prun[r] = run_new;
      }

      // ------------------------------------------------------------------ error norm + control (replicated)
      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        double norm;
        if (cfg.error_norm == PDEQ_NORM_SCALE_THEN_RMS) {
          norm = safe_sqrt(g.sum(acc_w2)) * inv_sqrt_d;
        } else {
          // rms(error_abs) / (atol + rtol rms(reference)); the isotropic error has size 1
          const double se2 = g.sum(acc_e2), sr2 = g.sum(acc_r2);
          const double rms_e = safe_sqrt(se2) * inv_sqrt_d;  // isotropic: every lane holds the same error
          norm = rms_e * fast_rcp(fma(a.rtol, safe_sqrt(sr2) * inv_sqrt_d, a.atol));
        }
        const double lep = neg_inv_n * log2(norm);
        accept = !(lep < 0.0);
        double lratio;
        if (cfg.control == PDEQ_CONTROL_PI) {
          lratio = fma(cfg.exponent_integral, lep, cfg.exponent_proportional * (lep - ctrl_lprev));
          if (lep >= 0.0) ctrl_lprev = lep;
        } else {
          lratio = lep;
        }
        const double ratio = cfg.safety * exp2(lratio);
        dt_next = fmax(cfg.factor_min, fmin(ratio, cfg.factor_max)) * dtc;
        if (a.sol.trace != nullptr && nattempts <= a.sol.trace_capacity && g.lane == 0) {
          double* tr = a.sol.trace + (b * a.sol.trace_capacity + (nattempts - 1)) * 4;
          tr[0] = t;
          tr[1] = dtc;
          tr[2] = exp2(lep);
          tr[3] = accept ? 1.0 : 0.0;
        }
        if (nattempts >= max_attempts) {
          status = PDEQ_STATUS_MAX_ATTEMPTS;
          accept = true;
          ck = T;
        }
      }

      // ------------------------------------------------------------------ commit
      dt = dt_next;
      if (accept) {
        // every lane has finished reading the accepted state (isotropic: idle lanes read lane 0's copy)
        g.sync();
        // interp_from (solvers_via_adaptive_steps.py:268-273: the state the accepted step started from) is read only by
        // the checkpoint branch that follows a step which reaches or oversteps t_next -- known here. Every other accepted
        // step's copy would be overwritten unread by the next one, so it is not made (config 3: 32 copies of the
        // 102-field state and conditional per instance instead of ~1500).
        const bool reaches = needs_interp && !(t_new + a.eps < t_next);
        for (int r = 0; r < MAXR; ++r) {
          if (r >= nrounds) break;
          const int j = g.lane + r * g.size;
          if (j >= d) continue;
          if (reaches) {  // interp_from <- step_from
            double m[n], L[n][n];
            st_load(st_from, d, j, m, L);
            st_store(st_if, d, j, m, L);
            if (FP) cond_copy(st_if + F_G * d, st_from + F_G * d, d, j);
          }
          st_store(st_from, d, j, pm[r], pL[r]);
          if (FP) {
            // the step's backward conditional (parked in the working column by this attempt's revert_push) is composed
            // with the carried one, in place
            remainder_merge<n>(st_from + F_G * d + j, d, p, pinv, wk + j);
            if (!adaptive) {
              // fixed grid: every grid point is a checkpoint. Store the backward conditional of this step and
              // restart from the identity -- the fixed-interval smoother (estimators_and_losses.py:612-620).
              cond_copy(ring + (size_t)ck * NFC * d, st_from + F_G * d, d, j);
              cond_store_identity(st_from + F_G * d, d, j);
            }
          }
          if (cfg_solver == PDEQ_SOLVER_DYNAMIC) st_from[F_SIG * d + j] = psig[r];
          st_from[F_RUN * d + j] = prun[r];
          if (!adaptive) emit(a, b, ck, d, j, t_new, pm[r], pL[r], st_from[F_SIG * d + j], nsteps + 1);
        }
        if (reaches) t_if = t;
        ndata += 1.0;
        t = t_new;
        nsteps += 1;
        if (!adaptive) ck += 1;
      }
    }
  }
};

struct GroupLaunchInfo {
  double* cond_ring;
  double* if_scratch;
  double* wk_scratch;  // smoother working columns when they do not fit in shared memory (CTA mode), else nullptr
  int groups_per_cta;
};

template <class VF, int NU, int FACT, bool TS0, bool FP, int MODE, int SPEC = 0>
__global__ void __launch_bounds__(MODE != 0 ? K2_CTA_THREADS : 128, MODE == 1 ? 2 : ((MODE == 0 && !FP) ? PDEQ_K2_WARP_FILTER_BLOCKS : 1))
    k2_loop_kernel(const __grid_constant__ LoopArgs a, const __grid_constant__ GroupLaunchInfo info) {
  extern __shared__ double smem_k2[];
  GroupLoop<VF, NU, FACT, TS0, FP, MODE, SPEC>::run(a, smem_k2, info.cond_ring, info.if_scratch, info.wk_scratch,
                                                    info.groups_per_cta);
}

// Host-side test for GroupLoop SPEC = 1.
inline bool k2_spec_matches(const LoopArgs& a) {
  return a.cfg.solver == PDEQ_SOLVER_DYNAMIC && a.cfg.error == PDEQ_ERROR_RESIDUAL_STD && a.cfg.error_per_unit_step == 0;
}

}  // namespace pdeq
