// Auxiliary operations on solutions of the DENSE factorisation (probdiffeq/_probdiffeq/ssm_impl_dense.py), run-time
// dimensions, one CTA per unit of work, every matrix in shared memory:
//   sample_dense_kernel          MarkovSequence.sample            (estimators_and_losses.py:233-271)
//   lml_timeseries_dense_kernel  MarkovSequence.evaluate_lml      (:180-218, under loss_lml_timeseries :53-105)
//   offgrid_dense_kernel         ProbabilisticSolver.offgrid_marginals (solvers.py:149-203) for the filter
//                                (strategy_filter.interpolate_offgrid_marginals :403-414) and the fixed-interval
//                                smoother (:677-709)
// The dense algebra restates DenseLatentCond.marginalise (:24-33) and .revert (:51-77 with
// util/cholesky_util.py:27-82), DenseWienerIntegrated.transition (:347-363) and DenseNormal.logpdf (:196-208) the way
// the dense smoother kernel does (pdeq_smooth_dense.cuh), and shares its CTA-cooperative Householder routine
// (pdeq_dense_qr.cuh). None of this is on a hot path: it runs once per solve, after the step loop.
#pragma once

#include "pdeq_dense_qr.cuh"
#include "pdeq_loop_thread.cuh"

namespace pdeq {

constexpr int AUXD_THREADS = 256;
constexpr int AUXD_MAX_COEFFS = 8;  // n <= 8 (PDEQ_MAX_COEFFS)

// shared memory (doubles) of the three kernels, for the host-side launchers
inline size_t auxd_lml_smem_doubles(int N, int d) {
  return (size_t)2 * N * (N + d) + 2 * (size_t)N * N + 3 * N + (size_t)N * d + (size_t)d * d + 2 * d + 48;
}
inline size_t auxd_offgrid_smem_doubles(int N, bool smooth) {
  return (smooth ? (size_t)8 * N * N : (size_t)4 * N * N) + 7 * (size_t)N + 4 * AUXD_MAX_COEFFS + 64 + 48;
}

// ---------------------------------------------------------------------------------------------------
// x_{T-1} = m + L eps_{T-1}; x_{k-1} = G_k x_k + xi_k + Xi_k eps_{k-1}. CTA (b, sample); base / out are [B][S][T][N].
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
    sample_dense_kernel(int S, int T, int N, const double* __restrict__ mean, const double* __restrict__ chol,
                        const double* __restrict__ bw_gain, const double* __restrict__ bw_mean,
                        const double* __restrict__ bw_chol, const double* __restrict__ base,
                        double* __restrict__ out) {
  extern __shared__ double sm_sample[];
  double* x = sm_sample;       // current state
  double* e = sm_sample + N;   // this grid point's draws
  const int64_t bs = blockIdx.x, b = bs / S;
  const int tid = threadIdx.x;
  for (int k = T - 1; k >= 0; --k) {
    for (int i = tid; i < N; i += blockDim.x) e[i] = base[(bs * T + k) * (int64_t)N + i];
    __syncthreads();
    double xn[2] = {0.0, 0.0};  // rows tid and tid + 64 (N <= 128)
    for (int rr = 0, i = tid; i < N; i += blockDim.x, ++rr) {
      const int64_t bt = b * T + (k == T - 1 ? T - 1 : k + 1);  // marginal at T-1, conditional k+1 maps k+1 -> k
      double acc;
      if (k == T - 1) {
        const double* Lr = chol + (bt * N + i) * (int64_t)N;
        acc = mean[bt * N + i];
        for (int c = 0; c <= i; ++c) acc = fma(Lr[c], e[c], acc);
      } else {
        const double* Gr = bw_gain + (bt * N + i) * (int64_t)N;
        const double* Xr = bw_chol + (bt * N + i) * (int64_t)N;
        acc = bw_mean[bt * N + i];
        for (int c = 0; c < N; ++c) acc = fma(Gr[c], x[c], acc);
        for (int c = 0; c <= i; ++c) acc = fma(Xr[c], e[c], acc);
      }
      xn[rr] = acc;
    }
    __syncthreads();
    for (int rr = 0, i = tid; i < N; i += blockDim.x, ++rr) {
      x[i] = xn[rr];
      out[(bs * T + k) * (int64_t)N + i] = xn[rr];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Shared pieces of the two CTA-per-instance kernels below (run-time n, d; N = n d; coefficient-major state).
// ---------------------------------------------------------------------------------------------------
struct AuxDense {
  int n, d, N;
  double* red;  // block_sum scratch (>= 40 doubles)

  // cond.marginalise(N(m, L L^T)) (DenseLatentCond.marginalise :24-33) for a conditional with gain G [N][N], noise
  // mean xi [N] and factor Xi [N][N] (global or shared memory) and per-coefficient scalings tl (to_latent) / to
  // (to_observed), nullptr = ones (a conditional in natural coordinates):
  //   m <- to (G (tl m) + xi),  L <- |to| tria([(G (|tl| L))^T ; Xi^T])^T.     W: 2N x ld stack, tmp: 2 N doubles.
  __device__ void marginalise(const double* G, const double* xi, const double* Xi, const double* tl, const double* to,
                              double* m, double* L, double* W, int ld, double* tmp) const {
    const int tid = threadIdx.x, NT = AUXD_THREADS;
    double* ml = tmp + N;
    for (int e = tid; e < N; e += NT) ml[e] = (tl != nullptr ? tl[e / d] : 1.0) * m[e];
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
      double acc = 0.0;
      for (int k = 0; k < N; ++k) acc = fma(G[i * N + k], ml[k], acc);
      tmp[i] = (to != nullptr ? to[i / d] : 1.0) * (acc + xi[i]);
    }
    for (int e = tid; e < N * N; e += NT) {
      const int i = e / N, j = e % N;
      double g = 0.0;
      for (int k = j; k < N; ++k) g = fma(G[i * N + k], (tl != nullptr ? fabs(tl[k / d]) : 1.0) * L[k * N + j], g);
      W[j * ld + i] = g;
      W[(N + j) * ld + i] = (i >= j) ? Xi[i * N + j] : 0.0;
    }
    __syncthreads();
    dense_qr_smem<AUXD_THREADS>(W, ld, 2 * N, N, N, red);
    for (int e = tid; e < N * N; e += NT) {
      const int i = e / N, j = e % N;
      L[e] = (j <= i) ? (to != nullptr ? fabs(to[i / d]) : 1.0) * W[j * ld + i] : 0.0;
    }
    for (int i = tid; i < N; i += NT) m[i] = tmp[i];
    __syncthreads();
  }

  // IWP transition over dt on N(m, L L^T), output scale s = sqrt|dt| sigma, prior scale lam[d]:
  // mp = p (A (pinv m)), Lp = |p| tria([(A (|pinv| L))^T ; (s Q)^T])^T -- DenseLatentCond.marginalise of
  // DenseWienerIntegrated.transition. With `bw` != nullptr the transition is REVERTED as well (DenseLatentCond.revert):
  // bw = [gain N N | Xi N N | mean N] of the backward conditional in preconditioned coordinates; W then needs
  // 2N x 2N. mt / mobt: N doubles of scratch each.
  __device__ void transition(const double* m, const double* L, const double* p, const double* pinv, double s,
                             const double (*A)[PDEQ_MAX_COEFFS], const double (*Qm)[PDEQ_MAX_COEFFS],
                             const double* lam, double* mp, double* Lp, double* W, int ld, double* mt, double* mobt,
                             double* bw) const {
    const int tid = threadIdx.x, NT = AUXD_THREADS;
    const bool smooth = bw != nullptr;
    for (int e = tid; e < N; e += NT) mt[e] = pinv[e / d] * m[e];
    __syncthreads();
    for (int e = tid; e < N; e += NT) {
      const int i = e / d, jd = e % d;
      double acc = 0.0;
      for (int k = i; k < n; ++k) acc = fma(A[i][k], mt[k * d + jd], acc);
      mobt[e] = acc;
      mp[e] = p[i] * acc;
    }
    const int ncol = smooth ? 2 * N : N;
    for (int e = tid; e < 2 * N * ncol; e += NT) {
      const int r = e / ncol, col = e % ncol;
      double val = 0.0;
      if (col < N) {
        const int ci = col / d, cj = col % d;
        if (r < N) {
          for (int k = ci; k < n; ++k) {
            const int row = k * d + cj;
            if (row >= r) val = fma(A[ci][k], fabs(pinv[k]) * L[row * N + r], val);
          }
        } else {
          const int rr = r - N, ri = rr / d, rj = rr % d;
          val = (cj == rj && ci >= ri) ? s * Qm[ci][ri] * lam[cj] : 0.0;
        }
      } else if (r < N) {
        const int k = col - N;
        val = (r <= k) ? fabs(pinv[k / d]) * L[k * N + r] : 0.0;
      }
      W[r * ld + col] = val;
    }
    __syncthreads();
    dense_qr_smem<AUXD_THREADS>(W, ld, 2 * N, ncol, N, red);
    for (int e = tid; e < N * N; e += NT) {
      const int row = e / N, col = e % N;
      Lp[e] = (col <= row) ? fabs(p[row / d]) * W[col * ld + row] : 0.0;
    }
    __syncthreads();
    if (!smooth) return;
    // gain: row k solves R_Y x = R12[:, k] in place. The conditional stays in preconditioned coordinates; its
    // scalings are to_latent = 1 / p, to_observed = 1 / pinv (the caller's business).
    for (int k = tid; k < N; k += NT) {
      for (int i = N - 1; i >= 0; --i) {
        double acc = W[i * ld + N + k];
        for (int l = i + 1; l < N; ++l) acc = fma(-W[i * ld + l], W[l * ld + N + k], acc);
        W[i * ld + N + k] = acc * fast_rcp(W[i * ld + i]);
      }
    }
    __syncthreads();
    double* G = bw;
    double* Xi = bw + (size_t)N * N;
    double* xi = bw + 2 * (size_t)N * N;
    for (int e = tid; e < N * N; e += NT) G[e] = W[(e % N) * ld + N + e / N];
    __syncthreads();
    for (int k = tid; k < N; k += NT) {
      double acc = mt[k];
      for (int i = 0; i < N; ++i) acc = fma(-G[k * N + i], mobt[i], acc);
      xi[k] = acc;
    }
    dense_qr_smem<AUXD_THREADS>(W + (size_t)N * ld + N, ld, N, N, N, red);
    for (int e = tid; e < N * N; e += NT) {
      const int row = e / N, col = e % N;
      Xi[e] = (col <= row) ? W[(N + col) * ld + N + row] : 0.0;
    }
    __syncthreads();
  }
};

// Taylor preconditioner for run-time n (pdeq_blockops.cuh: preconditioner<n>), one thread.
__device__ inline void auxd_preconditioner(int n, double dt, const double* ifact, const double* fact, double* p,
                                           double* pinv) {
  const double idt = fast_rcp(dt);
  double pw = 1.0, ipw = 1.0;
  for (int e = 0; e < n; ++e) {
    p[n - 1 - e] = pw * ifact[e];
    pinv[n - 1 - e] = ipw * fact[e];
    pw *= dt;
    ipw *= idt;
  }
}

// ---------------------------------------------------------------------------------------------------
// Backward scan of evaluate_lml: observe the terminal marginal, then alternately step back through a stored
// conditional and observe. The observation model is to_derivative(idx, std_k): H = e_idx (x) I_d, noise diag(std_k)
// (ssm_impl_dense.py:210-222); its revert triangularises [(H L)^T, L^T ; diag(std), 0], observation columns first.
// The gain uses the least-squares triangular solve the reference passes (a zero pivot gives a zero component), the
// log-pdf the exact one (DenseNormal.logpdf). CTA per instance.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AUXD_THREADS)
    lml_timeseries_dense_kernel(int T, int n, int d, int idx, int average, const double* __restrict__ mean,
                                const double* __restrict__ chol, const double* __restrict__ bw_gain,
                                const double* __restrict__ bw_mean, const double* __restrict__ bw_chol,
                                const double* __restrict__ data, int64_t data_stride, const double* __restrict__ std_,
                                int64_t std_stride, double* __restrict__ out) {
  extern __shared__ double sm_lml[];
  const int N = n * d, LD = N + d, NT = AUXD_THREADS, tid = threadIdx.x;
  double* W = sm_lml;                       // 2N x LD
  double* L = W + (size_t)2 * N * LD;       // N x N
  double* Gs = L + (size_t)N * N;           // N x N (a conditional's gain, then its factor)
  double* m = Gs + (size_t)N * N;           // N
  double* tmp = m + N;                      // 2 N
  double* Gt = tmp + 2 * N;                 // N x d  (gain^T rows)
  double* RY = Gt + (size_t)N * d;          // d x d
  double* res = RY + (size_t)d * d;         // d
  double* wv = res + d;                     // d
  double* red = wv + d;                     // 48
  const AuxDense ax{n, d, N, red};
  const int64_t b = blockIdx.x;
  const double log2pi = 1.8378770664093454835606594728112;
  {
    const int64_t bt = b * T + (T - 1);
    for (int e = tid; e < N * N; e += NT) L[e] = (e % N <= e / N) ? chol[bt * N * N + e] : 0.0;
    for (int e = tid; e < N; e += NT) m[e] = mean[bt * N + e];
  }
  __syncthreads();
  double acc = 0.0;
  int num = 0;
  for (int k = T - 1; k >= 0; --k) {
    if (k < T - 1) {
      const int64_t bt = b * T + (k + 1);
      // the gain goes through shared memory (read N times); the factor is read once, straight from global memory
      for (int e = tid; e < N * N; e += NT) Gs[e] = bw_gain[bt * N * N + e];
      __syncthreads();
      ax.marginalise(Gs, bw_mean + bt * N, bw_chol + bt * (int64_t)N * N, nullptr, nullptr, m, L, W, LD, tmp);
    }
    const double* sd = std_ + b * std_stride + (int64_t)k * d;
    const double* y = data + b * data_stride + (int64_t)k * d;
    for (int e = tid; e < LD * LD; e += NT) {
      const int r = e / LD, col = e % LD;
      double val = 0.0;
      if (col < d) {
        if (r < N) val = L[(idx * d + col) * N + r];
        else if (r - N == col) val = sd[col];
      } else if (r < N && r <= col - d) {
        val = L[(col - d) * N + r];
      }
      W[r * LD + col] = val;
    }
    for (int a = tid; a < d; a += NT) res[a] = y[a] - m[idx * d + a];
    __syncthreads();
    dense_qr_smem<AUXD_THREADS>(W, LD, LD, LD, LD, red);
    for (int e = tid; e < d * d; e += NT) RY[e] = (e / d <= e % d) ? W[(e / d) * LD + e % d] : 0.0;
    __syncthreads();
    for (int kk = tid; kk < N; kk += NT) {  // gain^T[:, kk] = lstsq_triu(R_Y, R12[:, kk])
      for (int i = d - 1; i >= 0; --i) {
        double a2 = W[i * LD + d + kk];
        for (int l = i + 1; l < d; ++l) a2 = fma(-RY[i * d + l], Gt[kk * d + l], a2);
        const double piv = RY[i * d + i];
        Gt[kk * d + i] = (piv == 0.0) ? 0.0 : a2 / piv;
      }
    }
    if (tid == 0) {  // log N(y; H m, R_Y^T R_Y): w = solve_tril(R_Y^T, res)
      double ww = 0.0, slogdet = 0.0;
      for (int i = 0; i < d; ++i) {
        double a2 = res[i];
        for (int l = 0; l < i; ++l) a2 = fma(-RY[l * d + i], wv[l], a2);
        wv[i] = a2 / RY[i * d + i];
        ww = fma(wv[i], wv[i], ww);
        slogdet += log(fabs(RY[i * d + i]));
      }
      const double pdf = -0.5 * ww - 0.5 * (double)d * log2pi - slogdet;
      red[40] = pdf;
    }
    __syncthreads();
    const double pdf = red[40];
    acc = average ? (acc * (double)num + pdf) / (double)(num + 1) : acc + pdf;
    num += 1;
    for (int kk = tid; kk < N; kk += NT) {
      double corr = 0.0;
      for (int a = 0; a < d; ++a) corr = fma(Gt[kk * d + a], res[a], corr);
      m[kk] += corr;
    }
    for (int e = tid; e < N * N; e += NT) {
      const int row = e / N, col = e % N;
      L[e] = (col <= row) ? W[(d + col) * LD + d + row] : 0.0;
    }
    __syncthreads();
  }
  if (tid == 0) out[b] = acc;
}

// ---------------------------------------------------------------------------------------------------
// Dense output. CTA (b, query). A filter extrapolates the marginal of the grid point to the left; the fixed-interval
// smoother extrapolates the FILTERING marginal t0 -> t, reverts the transition t -> t1 and pulls the smoothing marginal
// at t1 back through the new conditional. Both transitions use the output scale of the right grid point.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AUXD_THREADS)
    offgrid_dense_kernel(const __grid_constant__ pdeq_config cfg, int T, int Q, int smooth,
                         const double* __restrict__ grid, const double* __restrict__ queries,
                         const double* __restrict__ mean, const double* __restrict__ chol,
                         const double* __restrict__ filt_mean, const double* __restrict__ filt_chol,
                         const double* __restrict__ output_scale, const double* __restrict__ prior_scale,
                         int64_t prior_scale_stride, double* __restrict__ out_mean, double* __restrict__ out_chol) {
  extern __shared__ double sm_off[];
  const int n = cfg.num_derivatives + 1, d = cfg.ode_dim, N = n * d, NT = AUXD_THREADS, tid = threadIdx.x;
  const int ld = smooth ? 2 * N : N;
  double* W = sm_off;                                        // 2N x ld
  double* L0 = W + (size_t)2 * N * ld;                       // N x N
  double* Lt = L0 + (size_t)N * N;                           // N x N
  double* bw = Lt + (size_t)N * N;                           // smooth: 2 N N + N  (gain | Xi | mean)
  double* vec = bw + (smooth ? 2 * (size_t)N * N : 0);       // N (bw mean) + m0, mt, sa, sb: N each; tmp: 2 N
  double* m0 = vec + N;
  double* mt = m0 + N;
  double* sa = mt + N;
  double* sb = sa + N;
  double* tmp = sb + N;
  double* p = tmp + 2 * N;
  double* pinv = p + AUXD_MAX_COEFFS;
  double* tl = pinv + AUXD_MAX_COEFFS;
  double* to = tl + AUXD_MAX_COEFFS;
  double* lam = to + AUXD_MAX_COEFFS;                        // d <= 64
  double* red = lam + 64;
  const AuxDense ax{n, d, N, red};
  const int64_t b = blockIdx.x / Q;
  const int qi = blockIdx.x % Q;
  const double t = queries[qi];
  int lo = 0, hi = T;  // searchsorted(grid, t): first index with grid[index] >= t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (grid[mid] < t) lo = mid + 1;
    else hi = mid;
  }
  const int i1 = min(max(lo, 1), T - 1), i0 = i1 - 1;
  const double sig = output_scale[b * T + i1];
  const double dt0 = t - grid[i0], dt1 = grid[i1] - t;
  {
    const double* mu = smooth ? filt_mean : mean;
    const double* ch = smooth ? filt_chol : chol;
    const int64_t bt = b * T + i0;
    for (int e = tid; e < N * N; e += NT) L0[e] = (e % N <= e / N) ? ch[bt * N * N + e] : 0.0;
    for (int e = tid; e < N; e += NT) m0[e] = mu[bt * N + e];
    for (int e = tid; e < d; e += NT) lam[e] = prior_scale == nullptr ? 1.0 : prior_scale[b * prior_scale_stride + e];
    if (tid == 0) auxd_preconditioner(n, dt0, cfg.inv_factorials, cfg.factorials, p, pinv);
  }
  __syncthreads();
  ax.transition(m0, L0, p, pinv, safe_sqrt(fabs(dt0)) * sig, cfg.sys_a, cfg.sys_q, lam, mt, Lt, W, ld, sa, sb, nullptr);
  if (smooth) {
    if (tid == 0) auxd_preconditioner(n, dt1, cfg.inv_factorials, cfg.factorials, p, pinv);
    __syncthreads();
    // revert t -> t1 on N(mt, Lt Lt^T); the predicted marginal at t1 (m0 / L0 are free by now) is not needed
    ax.transition(mt, Lt, p, pinv, safe_sqrt(fabs(dt1)) * sig, cfg.sys_a, cfg.sys_q, lam, m0, L0, W, ld, sa, sb, bw);
    const int64_t bt = b * T + i1;
    for (int e = tid; e < N * N; e += NT) Lt[e] = (e % N <= e / N) ? chol[bt * N * N + e] : 0.0;
    for (int e = tid; e < N; e += NT) mt[e] = mean[bt * N + e];
    __syncthreads();
    if (tid < n) {  // scalings of the reverted transition: to_latent = 1 / p, to_observed = 1 / pinv
      tl[tid] = fast_rcp(p[tid]);
      to[tid] = fast_rcp(pinv[tid]);
    }
    __syncthreads();
    ax.marginalise(bw, bw + 2 * (size_t)N * N, bw + (size_t)N * N, tl, to, mt, Lt, W, ld, tmp);
  }
  const int64_t bq = b * Q + qi;
  for (int e = tid; e < N; e += NT) out_mean[bq * N + e] = mt[e];
  if (out_chol != nullptr)
    for (int e = tid; e < N * N; e += NT) out_chol[bq * N * N + e] = Lt[e];
}

}  // namespace pdeq
