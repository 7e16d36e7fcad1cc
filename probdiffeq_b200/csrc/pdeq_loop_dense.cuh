// K3: the dense factorisation (probdiffeq/_probdiffeq/ssm_impl_dense.py), one IVP instance per CTA of 64 threads.
//
// The state is a mean of length N = n d (coefficient-major, index k d + i) and a full N x N left square root.
// Every step triangularises a 2N x N stack (extrapolation, DenseLatentCond.marginalise :24-33) and an
// (N + d) x (N + d) stack (correction, DenseLatentCond.revert :51-77 with util/cholesky_util.py:27-82); for
// HIRES (d = 8, nu = 5) these are 96 x 48 and 56 x 56.  Both live in ONE shared-memory work buffer W of
// 2N x (N + d) doubles and are triangularised by a cooperative Householder: the 64 threads own the trailing
// columns (row-major W => conflict-free, the reflector is a shared-memory broadcast), the column norm is a
// two-warp reduction.  Reflectors follow LAPACK dlarfg (see pdeq_blockops.cuh).  Stack rows are ordered
// [(H L)^T, L^T ; damp I, 0] -- a row permutation of the reference's block matrix that leaves R^T R, hence all
// covariances and the gain, unchanged -- so that the extrapolated factor is reused in place.
//
// Restates for the dense model: DenseWienerIntegrated.transition (:347-363), DenseOdeTs0.linearize (:243-259),
// DenseResidual.linearize (:290-334, exact Jacobian, jacobians.py:93-98), DenseNormal.std (:144-150),
// residual_whitened_rms_flat (:156-160); the loop / solver / error / control lines are those listed in
// pdeq_loop_thread.cuh.  Filter strategy only.
#pragma once

#include "pdeq_loop_thread.cuh"

namespace pdeq {

constexpr int K3_THREADS = 256;

struct DenseSmemLayout {
  int N, d, ld;  // ld = N + d
  size_t off_W, off_Lfrom, off_Lif, off_vec, total;
  __host__ __device__ static DenseSmemLayout make(int n, int d, int order, bool needs_interp) {
    DenseSmemLayout s;
    s.N = n * d;
    s.d = d;
    s.ld = s.N + d;
    const size_t tri = (size_t)s.N * (s.N + 1) / 2;
    size_t o = 0;
    s.off_W = o;
    o += (size_t)2 * s.N * s.ld;
    s.off_Lfrom = o;
    o += tri;
    s.off_Lif = o;
    o += needs_interp ? tri : 0;
    s.off_vec = o;
    // m_from, mp, m_new, m_if (4N) | Hs d x (order+1) d | mobs, wht, std, ref, lam (5d) | p, pinv (2 * 8) | red 8 | bc 8
    // | rdiag N + d
    o += (size_t)4 * s.N + (size_t)d * (order + 1) * d + 5 * d + 16 + 16 + s.ld;
    s.total = o;
    return s;
  }
};

template <class VF, int NU, bool TS0>
struct DenseLoop {
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  static constexpr int G = K3_THREADS;
  // All extents are compile-time: the dense kernels exist for fixed-dimension vector fields only (the host checks
  // cfg.ode_dim == VF::fixed_dim), so every index split (e / N, e % d, ...) is a multiply-shift, not a division.
  static constexpr int D = VF::fixed_dim;
  static_assert(D > 0, "the dense kernels need a compile-time ODE dimension");
  static constexpr int N = n * D, LD = N + D, HW = (q + 1) * D, TRI_N = N * (N + 1) / 2;

  struct VecAcc {
    const double* u;  // coefficient-major mean
    int d;
    PDEQ_DI double operator()(int k, int i) const { return u[k * d + i]; }
  };

  PDEQ_DI static int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower, j <= i

  PDEQ_DI static double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) tot += red[w];
    return tot;
  }

  // Sum of squares of W[j+1..h][j], by warp 0, into *dst; ends with a block barrier.
  PDEQ_DI static void column_ss(const double* W, int j, int h, double* dst) {
    if (threadIdx.x < 32) {
      double part = 0.0;
      for (int r = j + 1 + (int)threadIdx.x; r <= h; r += 32) part = fma(W[r * LD + j], W[r * LD + j], part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (threadIdx.x == 0) *dst = part;
    }
    __syncthreads();
  }

  // Cooperative in-place Householder triangularisation of the M x ncols matrix at W (leading dimension LD).
  // hi(c) = min(hi_a + c, M - 1) for c < csplit, M - 1 otherwise: the last non-zero row of column c.
  //
  // 256 threads = 64 column slots x 4 row groups. A warp holds 8 adjacent columns x 4 row groups (rows r, r+1 of
  // LD = 56 doubles are 16 banks apart: each half-warp is conflict-free), so the four partial dot products of one
  // column meet by two shuffles and there is ONE block barrier per column: the sum of squares of the next column
  // is accumulated by the lanes that update it (always slot 0, i.e. warp 0), and the new diagonal goes to `rdiag`
  // so that nobody races with the readers of alpha. Reflectors follow LAPACK dlarfg (H = I for a zero sub-column).
  template <int M, int ncols, int csplit, int hi_a>
  PDEQ_DI static void coop_qr(double* W, double* red, double* rdiag) {
    constexpr int ld = LD;
    const int tid = threadIdx.x;
    const int cslot = (tid >> 5) * 8 + (tid & 7);
    const int rg = (tid >> 3) & 3;
    const bool warp0 = tid < 32;
    auto hi = [](int c) { return (c < csplit) ? min(hi_a + c, M - 1) : M - 1; };
    column_ss(W, 0, hi(0), red);
    for (int j = 0; j < ncols; ++j) {
      const int hj = hi(j);
      const double alpha = W[j * ld + j];
      const double ss = red[j & 1];
      const bool last = j + 1 >= ncols;
      if (hj <= j || ss == 0.0) {  // uniform: nothing to annihilate
        if (tid == 0) rdiag[j] = alpha;
        if (!last) column_ss(W, j + 1, hi(j + 1), red + ((j + 1) & 1));
        continue;
      }
      const double tt = fma(alpha, alpha, ss);
      const double y = fast_rsqrt(tt);
      const double nrm = tt * y;
      const double sgn_nrm = copysign(nrm, alpha);
      const double v0 = alpha + sgn_nrm;
      const double tp = y * fast_rcp(nrm + fabs(alpha));
      if (tid == 0) rdiag[j] = -sgn_nrm;
      const int h1 = last ? hj : hi(j + 1);
      for (int c0 = j + 1; c0 < ncols; c0 += G / 4) {
        const int c = c0 + cslot;
        const bool active = c < ncols;
        double w0 = 0.0, w1 = 0.0;
        if (active) {
          int r = j + 1 + rg;
          for (; r + 4 <= hj; r += 8) {
            w0 = fma(W[r * ld + j], W[r * ld + c], w0);
            w1 = fma(W[(r + 4) * ld + j], W[(r + 4) * ld + c], w1);
          }
          if (r <= hj) w0 = fma(W[r * ld + j], W[r * ld + c], w0);
          if (rg == 0) w1 = fma(v0, W[j * ld + c], w1);
        }
        double w = w0 + w1;
        w += __shfl_xor_sync(0xffffffffu, w, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 16);
        w *= tp;
        const bool gather = warp0 && c0 == j + 1;  // warp-uniform: this warp owns column j + 1
        double sq = 0.0, sq1 = 0.0;
        if (active) {
          if (rg == 0) W[j * ld + c] = fma(-w, v0, W[j * ld + c]);
          if (gather) {
            // ... and gathers that column's sum of squares below its diagonal for the next reflector
            int r = j + 1 + rg;
            for (; r + 4 <= hj; r += 8) {
              const double x0 = fma(-w, W[r * ld + j], W[r * ld + c]);
              const double x1 = fma(-w, W[(r + 4) * ld + j], W[(r + 4) * ld + c]);
              W[r * ld + c] = x0;
              W[(r + 4) * ld + c] = x1;
              sq = (cslot == 0 && r > j + 1) ? fma(x0, x0, sq) : sq;
              sq1 = (cslot == 0) ? fma(x1, x1, sq1) : sq1;
            }
            if (r <= hj) {
              const double x0 = fma(-w, W[r * ld + j], W[r * ld + c]);
              W[r * ld + c] = x0;
              sq = (cslot == 0 && r > j + 1) ? fma(x0, x0, sq) : sq;
            }
            sq += sq1;
          } else {
#pragma unroll 2
            for (int r = j + 1 + rg; r <= hj; r += 4) W[r * ld + c] = fma(-w, W[r * ld + j], W[r * ld + c]);
          }
        }
        if (gather) {
          if (tid == 0)  // rows of column j + 1 below this reflector's reach
            for (int r = hj + 1; r <= h1; ++r) sq = fma(W[r * ld + c], W[r * ld + c], sq);
          sq += __shfl_xor_sync(0xffffffffu, sq, 8);
          sq += __shfl_xor_sync(0xffffffffu, sq, 16);
          if (tid == 0) red[(j + 1) & 1] = sq;
        }
      }
      __syncthreads();
    }
    // the last column may have taken the barrier-free skip path: make its rdiag entry (and everyone's last read
    // of the old diagonal) visible before the diagonal is restored
    __syncthreads();
    for (int j = tid; j < ncols; j += G) W[j * ld + j] = rdiag[j];
    __syncthreads();
  }

  // Write one checkpoint (mean [n][d], chol [N][N]) from shared memory.
  PDEQ_DI static void emit(const LoopArgs& a, long b, int ck, const DenseSmemLayout& lay, double t, const double* m,
                           const double* Lpk, double chol_scale, double scale, int nsteps) {
    const long bt = b * a.T + ck;
    (void)lay;
    if (threadIdx.x == 0) {
      a.sol.t[bt] = t;
      a.sol.num_steps[bt] = nsteps;
      if (a.sol.output_scale != nullptr) a.sol.output_scale[bt] = scale;
    }
    for (int e = threadIdx.x; e < N; e += G) a.sol.mean[bt * N + e] = m[e];
    if (a.sol.chol != nullptr) {
      double* co = a.sol.chol + bt * (long)N * N;
      for (int e = threadIdx.x; e < N * N; e += G) {
        const int i = e / N, j = e % N;
        co[e] = (j <= i) ? chol_scale * Lpk[tri(i, j)] : 0.0;
      }
    }
  }

  // Fill W[0..2N) x [d..d+N) with the extrapolation stack [(A (pinv L))^T ; (s Q)^T], A = kron(a, I_d),
  // Q = kron(q, diag(lam)); triangularise; scale the columns of R by |p| so that rows 0..N-1 hold L_pred^T.
  PDEQ_DI static void extrapolate_chol(double* W, const DenseSmemLayout& lay, const double* Lpk, const double* p,
                                       const double* pinv, const double* lam, double s,
                                       const double (*A)[PDEQ_MAX_COEFFS], const double (*Qm)[PDEQ_MAX_COEFFS],
                                       double* red) {
    constexpr int d = D, ld = LD;
    (void)lay;
    for (int e = threadIdx.x; e < N * N; e += G) {
      const int r = e / N, c = e % N;
      const int ci = c / d, cj = c % d;
      // (A L~)[c][r] = sum_k a[ci][k] pinv_k L[k d + cj][r]   (k >= ci, and k d + cj >= r)
      double acc = 0.0;
      for (int k = ci; k < n; ++k) {
        const int row = k * d + cj;
        if (row >= r) acc = fma(A[ci][k], pinv[k] * Lpk[tri(row, r)], acc);
      }
      W[r * ld + d + c] = acc;
      const int ri = r / d, rj = r % d;
      W[(N + r) * ld + d + c] = (cj == rj && ci >= ri) ? s * Qm[ci][ri] * lam[cj] : 0.0;
    }
    __syncthreads();
    coop_qr<2 * N, N, N, N>(W + d, red, red + 16);
    for (int e = threadIdx.x; e < N * N; e += G) {
      const int r = e / N, c = e % N;
      if (r <= c) W[r * ld + d + c] *= fabs(p[c / d]);
      else W[r * ld + d + c] = 0.0;  // drop the stored reflectors
    }
    __syncthreads();
  }

  // With rows 0..N-1, cols d..d+N-1 of W holding an upper-triangular factor U = L^T, build and triangularise
  // [(H L)^T | L^T ; damp I | 0]. Afterwards R_Y = W[0..d)[0..d), R12 = W[0..d)[d..), R_XY = W[d..d+N)[d..).
  PDEQ_DI static void revert_stack(double* W, const DenseSmemLayout& lay, const double* Hs, double damp, double* red) {
    constexpr int d = D, ld = LD, hw = HW;
    (void)lay;
    for (int e = threadIdx.x; e < N * d; e += G) {
      const int r = e / d, a_ = e % d;
      double acc = 0.0;  // (H L)[a][r] = sum_e' H[a][e'] L[e'][r], L[e'][r] = U[r][e'] (e' >= r)
      for (int ee = r; ee < hw; ++ee) acc = fma(Hs[a_ * hw + ee], W[r * ld + d + ee], acc);
      W[r * ld + a_] = acc;
    }
    for (int e = threadIdx.x; e < d * ld; e += G) {
      const int r = e / ld, c = e % ld;
      W[(N + r) * ld + c] = (c == r) ? damp : 0.0;
    }
    __syncthreads();
    coop_qr<N + d, N + d, d, N>(W, red, red + 16);
  }

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem) {
    const pdeq_config& cfg = a.cfg;
    const double(*__restrict__ A)[PDEQ_MAX_COEFFS] = cfg.sys_a;
    const double(*__restrict__ Qm)[PDEQ_MAX_COEFFS] = cfg.sys_q;
    const double* __restrict__ fact = cfg.factorials;
    const double* __restrict__ ifact = cfg.inv_factorials;
    const bool adaptive = a.fixed_grid == 0;
    const bool clip = cfg.clip_dt != 0;
    const bool needs_interp = adaptive && !clip;
    const int T = a.T;
    constexpr int d = D;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const double inv_sqrt_d = rsqrt((double)d);
    const double neg_inv_n = -1.0 / (double)n;
    const int tid = threadIdx.x;
    const DenseSmemLayout lay = DenseSmemLayout::make(n, d, q, needs_interp);
    constexpr int ld = LD, hw = HW;

    double* W = smem + lay.off_W;
    double* Lfrom = smem + lay.off_Lfrom;
    double* Lif = smem + lay.off_Lif;
    double* v = smem + lay.off_vec;
    double* m_from = v;
    double* mp = v + N;
    double* m_new = v + 2 * N;
    double* m_if = v + 3 * N;
    double* Hs = v + 4 * N;            // d x (q+1) d: the non-zero columns of the linearisation
    double* mobs = Hs + d * hw;        // d
    double* wht = mobs + d;            // d: whitened residual
    double* stdv = wht + d;            // d: error estimate per dimension
    double* refv = stdv + d;           // d
    double* lam = refv + d;            // d: prior output scale (diagonal of Lambda)
    double* p = lam + d;               // 8
    double* pinv = p + 8;              // 8
    double* red = pinv + 8;            // 8
    double* bc = red + 8;              // 8: broadcast slots; red + 16: N + d new diagonal entries (coop_qr)

#ifdef PDEQ_K3_PROFILE
    // phase cycle counters (development aid): totals per instance go to the first 8 trace slots
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = clock64();
#define PDEQ_K3_TICK(i) { const long long now_ = clock64(); prof_acc[i] += now_ - prof_last; prof_last = now_; }
#define PDEQ_K3_DUMP if (tid == 0 && a.sol.trace != nullptr && a.sol.trace_capacity >= 2) { \
      for (int i_ = 0; i_ < 8; ++i_) { a.sol.trace[b * a.sol.trace_capacity * 4 + i_] = (double)prof_acc[i_]; prof_acc[i_] = 0; } }
#else
#define PDEQ_K3_TICK(i)
#define PDEQ_K3_DUMP
#endif
    double params[P];
    double t = 0.0, dt = 0.0, ctrl_lprev = 0.0, ndata = 0.0, t_next = 0.0, t_if = 0.0, sig = 1.0, run_scale = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;
    bool need_load = true;

    while (true) {
      if (need_load) {
        __syncthreads();
        if (tid == 0) reinterpret_cast<long*>(bc)[0] = (long)atomicAdd(a.work_counter, 1ULL);
        __syncthreads();
        b = reinterpret_cast<long*>(bc)[0];
        if (b >= B) break;
        need_load = false;
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        for (int e = tid; e < N; e += G) m_from[e] = a.prob.tcoeffs[b * N + e];
        for (int e = tid; e < TRI_N; e += G) Lfrom[e] = 0.0;
        __syncthreads();
        if (a.prob.init_std != nullptr) {
          const double* sd = a.prob.init_std + b * a.prob.init_std_stride;
          for (int e = tid; e < N; e += G) Lfrom[tri(e, e)] = sd[e];
        }
        for (int e = tid; e < d; e += G)
          lam[e] = a.prob.prior_scale != nullptr ? a.prob.prior_scale[b * a.prob.prior_scale_stride + e] : 1.0;
        t = a.grid[0];
        dt = adaptive ? a.dt0[b * a.dt0_stride] : 0.0;
        ctrl_lprev = 0.0;
        ndata = 0.0;
        nsteps = 0;
        nattempts = 0;
        status = 0;
        sig = 1.0;
        run_scale = 0.0;
        __syncthreads();
        emit(a, b, 0, lay, t, m_from, Lfrom, 1.0, 1.0, 0);
        if (needs_interp) {
          for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
          for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
        }
        t_if = t;
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        __syncthreads();
      }

      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          if (needs_interp && t > t_next + a.eps) {
            // interp_beyond_t1: extrapolate interp_from to t_next with the scale of the overstepped state
            const double dti = t_next - t_if;
            if (tid == 0) {
              double pp[n], pi[n];
              preconditioner<n>(dti, ifact, fact, pp, pi);
#pragma unroll
              for (int i = 0; i < n; ++i) {
                p[i] = pp[i];
                pinv[i] = pi[i];
              }
            }
            __syncthreads();
            for (int e = tid; e < N; e += G) {
              const int i = e / d, jd = e % d;
              double acc = 0.0;
              for (int k = i; k < n; ++k) acc = fma(A[i][k], pinv[k] * m_if[k * d + jd], acc);
              mp[e] = p[i] * acc;
            }
            extrapolate_chol(W, lay, Lif, p, pinv, lam, safe_sqrt(fabs(dti)) * sig, A, Qm, red);
            for (int e = tid; e < N * N; e += G) {
              const int jj = e / N, i = e % N;  // L[i][jj] = W[jj][d + i], lower triangle only
              if (jj <= i) Lif[tri(i, jj)] = W[jj * ld + d + i];
            }
            for (int e = tid; e < N; e += G) m_if[e] = mp[e];
            __syncthreads();
            emit(a, b, ck, lay, t_next, m_if, Lif, 1.0, sig, nsteps);
            t_if = t_next;
          } else {
            emit(a, b, ck, lay, t, m_from, Lfrom, 1.0, sig, nsteps);
            if (needs_interp) {
              for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
              for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
            }
            t_if = t;
          }
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
        }
        if (ck >= T) {
          if (cfg.solver == PDEQ_SOLVER_MLE) {
            double fin = run_scale;
            if (cfg.correct_asymptotic_underconfidence) fin = fin / sqrt((double)nsteps);
            __syncthreads();
            for (int c = 0; c < T; ++c) {
              const long bt = b * T + c;
              if (a.sol.chol != nullptr) {
                double* co = a.sol.chol + bt * (long)N * N;
                for (int e = tid; e < N * N; e += G) co[e] = fin * co[e];
              }
              if (tid == 0 && a.sol.output_scale != nullptr) a.sol.output_scale[bt] = fin;
            }
          }
          double bad = 0.0;
          for (int e = tid; e < N; e += G) bad += isfinite(m_from[e]) ? 0.0 : 1.0;
          bad = block_sum(bad, red);
          if (status == 0 && bad > 0.0) status = PDEQ_STATUS_NONFINITE;
          PDEQ_K3_DUMP
          if (tid == 0) {
            a.sol.status[b] = status;
            if (a.sol.num_attempts != nullptr) a.sol.num_attempts[b] = nattempts;
          }
          need_load = true;
        }
        __syncthreads();
        continue;
      }

      // ------------------------------------------------------------------ one step attempt
      nattempts += 1;
      PDEQ_K3_TICK(0)
      double dtc;
      if (adaptive) {
        dtc = clip ? fmin(dt, t_next - t) : dt;
      } else {
        dtc = a.grid[ck] - a.grid[ck - 1];
      }
      const double sq = safe_sqrt(fabs(dtc));
      const double t_new = t + dtc;
      __syncthreads();
      if (tid == 0) {
        double pp[n], pi[n];
        preconditioner<n>(dtc, ifact, fact, pp, pi);
#pragma unroll
        for (int i = 0; i < n; ++i) {
          p[i] = pp[i];
          pinv[i] = pi[i];
        }
      }
      __syncthreads();
      for (int e = tid; e < N; e += G) {
        const int i = e / d, jd = e % d;
        double acc = 0.0;
        for (int k = i; k < n; ++k) acc = fma(A[i][k], pinv[k] * m_from[k * d + jd], acc);
        mp[e] = p[i] * acc;
      }
      __syncthreads();

      // linearise at the extrapolated mean
      for (int e = tid; e < d * hw; e += G) Hs[e] = 0.0;
      __syncthreads();
      {
        VecAcc acc{mp, d};
        for (int jd = tid; jd < d; jd += G) {
          const double f = VF::template component<double>(jd, d, acc, params, t_new);
          Hs[jd * hw + q * d + jd] = 1.0;
          if (TS0) {
            mobs[jd] = mp[q * d + jd] + (-f);
          } else {
            const double rres = mp[q * d + jd] - f;
            double hm = mp[q * d + jd];
            for (int c = 0; c < q; ++c) {
              for (int l = 0; l < d; ++l) {
                const double hv = -VF::jac(jd, c, l, d, acc, params, t_new);
                Hs[jd * hw + c * d + l] = hv;
                hm = fma(hv, mp[c * d + l], hm);
              }
            }
            mobs[jd] = hm + (rres - hm);
          }
        }
      }
      __syncthreads();

      PDEQ_K3_TICK(1)
      // observation of the zero-error extrapolation: R_obs = qr_r([(H L_u)^T ; damp I]) (solver_dynamic, residual error)
      const bool need_obs = adaptive ? (cfg.solver == PDEQ_SOLVER_DYNAMIC || cfg.error == PDEQ_ERROR_RESIDUAL_STD)
                                     : (cfg.solver == PDEQ_SOLVER_DYNAMIC);
      double sig_new = 1.0, whitened_obs = 0.0;
      if (need_obs) {
        // (H L_u)[a][k d + l] = sum_{i >= k} H[a][i d + l] |p_i| sq q[i][k] lam_l
        for (int e = tid; e < N * d; e += G) {
          const int r = e / d, a_ = e % d;
          const int k = r / d, l = r % d;
          double acc = 0.0;
          for (int i = k; i <= q && i < n; ++i) acc = fma(Hs[a_ * hw + i * d + l], fabs(p[i]) * sq * Qm[i][k], acc);
          W[r * ld + a_] = acc * lam[l];
        }
        for (int e = tid; e < d * d; e += G) W[(N + e / d) * ld + (e % d)] = (e / d == e % d) ? a.damp : 0.0;
        __syncthreads();
        coop_qr<N + d, d, 0, 0>(W, red, red + 16);
        if (tid == 0) {
          // whitened residual: solve R_obs^T w = mobs (forward substitution), rms; and the row norms of R_obs^T
          double ss = 0.0;
          for (int i = 0; i < d; ++i) {
            double acc = mobs[i];
            for (int l = 0; l < i; ++l) acc = fma(-W[l * ld + i], wht[l], acc);
            wht[i] = acc * fast_rcp(W[i * ld + i]);
            ss = fma(wht[i], wht[i], ss);
            double rn = 0.0;
            for (int l = 0; l <= i; ++l) rn = fma(W[l * ld + i], W[l * ld + i], rn);
            stdv[i] = safe_sqrt(rn);
          }
          bc[1] = safe_sqrt(ss) * inv_sqrt_d;
        }
        __syncthreads();
        whitened_obs = bc[1];
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig_new = whitened_obs;
      }

      if (adaptive && cfg.error != PDEQ_ERROR_RESIDUAL_STD) {
        // error_state_std (solvers.py:1070-1086): Bayes rule on the zero-error extrapolation. Done first because
        // it needs the work buffer, which afterwards holds the proposal until the step is accepted or rejected.
        const int idx = cfg.derivative_idx;
        for (int e = tid; e < N * N; e += G) {
          const int r = e / N, c = e % N;  // U = L_u^T: U[r][c] = L_u[c][r], block lower in (ci >= ri), same l
          const int ci = c / d, cj = c % d, ri = r / d, rj = r % d;
          W[r * ld + d + c] = (cj == rj && ci >= ri) ? fabs(p[ci]) * sq * Qm[ci][ri] * lam[cj] : 0.0;
        }
        __syncthreads();
        revert_stack(W, lay, Hs, a.damp, red);
        if (tid == 0) {
          double ss = 0.0;
          for (int i = 0; i < d; ++i) {
            double acc = mobs[i];
            for (int l = 0; l < i; ++l) acc = fma(-W[l * ld + i], wht[l], acc);
            wht[i] = acc * fast_rcp(W[i * ld + i]);
            ss = fma(wht[i], wht[i], ss);
          }
          bc[3] = safe_sqrt(ss) * inv_sqrt_d;
        }
        __syncthreads();
        for (int e = tid; e < d; e += G) {
          const int col = idx * d + e;  // std of coefficient idx, dimension e: column norm of R_XY
          double rn = 0.0;
          for (int r = 0; r <= col; ++r) rn = fma(W[(d + r) * ld + d + col], W[(d + r) * ld + d + col], rn);
          stdv[e] = bc[3] * safe_sqrt(rn);
        }
        __syncthreads();
      }

      // extrapolate the factor, correct
      PDEQ_K3_TICK(2)
      extrapolate_chol(W, lay, Lfrom, p, pinv, lam, sq * sig_new, A, Qm, red);
      PDEQ_K3_TICK(3)
      revert_stack(W, lay, Hs, a.damp, red);
      PDEQ_K3_TICK(4)
      // gain^T = R_Y^-1 R12 (d x N), in place over R12; one column per thread
      for (int c = tid; c < N; c += G) {
        for (int i = d - 1; i >= 0; --i) {
          double acc = W[i * ld + d + c];
          for (int l = i + 1; l < d; ++l) acc = fma(-W[i * ld + l], W[l * ld + d + c], acc);
          W[i * ld + d + c] = acc * fast_rcp(W[i * ld + i]);
        }
      }
      __syncthreads();
      for (int e = tid; e < N; e += G) {
        double acc = mp[e];
        for (int a_ = 0; a_ < d; ++a_) acc = fma(-W[a_ * ld + d + e], mobs[a_], acc);
        m_new[e] = acc;
      }
      // the proposal's factor stays in W (rows / columns d .. d+N) until the step is accepted
      double run_new = run_scale;
      if (cfg.solver == PDEQ_SOLVER_MLE) {
        if (tid == 0) {
          double ss = 0.0;
          for (int i = 0; i < d; ++i) {
            double acc = mobs[i];
            for (int l = 0; l < i; ++l) acc = fma(-W[l * ld + i], wht[l], acc);
            wht[i] = acc * fast_rcp(W[i * ld + i]);
            ss = fma(wht[i], wht[i], ss);
          }
          bc[2] = safe_sqrt(ss) * inv_sqrt_d;
        }
        __syncthreads();
        const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
        const double x1 = w1 * run_scale, x2 = w2 * bc[2];
        run_new = safe_sqrt(fma(x1, x1, x2 * x2));
      }
      __syncthreads();

      PDEQ_K3_TICK(5)
      // ------------------------------------------------------------------ error estimate + control
      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        int kpow, idx = 0;
        if (cfg.error == PDEQ_ERROR_RESIDUAL_STD) {
          // error_d = sigma * std_d of the observed marginal (solvers.py:955-960); stdv holds the row norms
          for (int e = tid; e < d; e += G) {
            stdv[e] = whitened_obs * stdv[e];
            refv[e] = fmax(fabs(m_from[e]), fabs(m_new[e]));
          }
          kpow = q;
        } else {
          // error_state_std (solvers.py:1070-1086): Bayes rule on the zero-error extrapolation
          idx = cfg.derivative_idx;  // stdv was computed ahead of the extrapolation (W is the proposal's home)
          for (int e = tid; e < d; e += G) refv[e] = fmax(fabs(m_from[idx * d + e]), fabs(m_new[idx * d + e]));
          kpow = idx;
        }
        if (cfg.error_per_unit_step) kpow += 1;
        __syncthreads();
        if (tid == 0) {
          double escale = ipow_small<n>(dtc, kpow);
          for (int e = 0; e <= n; ++e) {
            if (e == kpow) escale *= ifact[e];
          }
          double norm;
          if (cfg.error_norm == PDEQ_NORM_SCALE_THEN_RMS) {
            double ss = 0.0;
            for (int e = 0; e < d; ++e) {
              const double w = (stdv[e] * escale) * fast_rcp(fma(a.rtol, refv[e], a.atol));
              ss = fma(w, w, ss);
            }
            norm = safe_sqrt(ss) * inv_sqrt_d;
          } else {
            double se2 = 0.0, sr2 = 0.0;
            for (int e = 0; e < d; ++e) {
              const double ea = stdv[e] * escale;
              se2 = fma(ea, ea, se2);
              sr2 = fma(refv[e], refv[e], sr2);
            }
            norm = (safe_sqrt(se2) * inv_sqrt_d) * fast_rcp(fma(a.rtol, safe_sqrt(sr2) * inv_sqrt_d, a.atol));
          }
          bc[4] = neg_inv_n * log2(norm);
        }
        __syncthreads();
        const double lep = bc[4];
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
        if (a.sol.trace != nullptr && nattempts <= a.sol.trace_capacity && tid == 0) {
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
      PDEQ_K3_TICK(6)
      dt = dt_next;
      if (accept) {
        if (needs_interp) {
          for (int e = tid; e < N; e += G) m_if[e] = m_from[e];
          for (int e = tid; e < TRI_N; e += G) Lif[e] = Lfrom[e];
          t_if = t;
        }
        __syncthreads();
        for (int e = tid; e < N; e += G) m_from[e] = m_new[e];
        for (int e = tid; e < N * N; e += G) {
          const int jj = e / N, i = e % N;
          if (jj <= i) Lfrom[tri(i, jj)] = W[(d + jj) * ld + d + i];
        }
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig = sig_new;
        run_scale = run_new;
        ndata += 1.0;
        t = t_new;
        nsteps += 1;
        __syncthreads();
        if (!adaptive) {
          emit(a, b, ck, lay, t, m_from, Lfrom, 1.0, sig, nsteps);
          ck += 1;
        }
      }
    }
  }
};

template <class VF, int NU, bool TS0>
__global__ void __launch_bounds__(K3_THREADS, 4) k3_loop_kernel(const __grid_constant__ LoopArgs a) {
  extern __shared__ double smem_k3[];
  DenseLoop<VF, NU, TS0>::run(a, smem_k3);
}

}  // namespace pdeq
