// K3S: smoothers for the DENSE factorisation (probdiffeq/_probdiffeq/ssm_impl_dense.py), one IVP instance per CTA.
//
// The dense model's smoother carries, besides the N x N factor (N = n d), a backward conditional with a full N x N gain
// and an N x N noise factor, reverts the transition through a 2N x 2N triangularisation every attempt
// (DenseLatentCond.revert :51-77 with util/cholesky_util.py:27-82) and composes conditionals with N^3 products
// (DenseLatentCond.merge :35-49). No BASELINE configuration uses it; it is built for coverage of the reference's
// strategy x factorisation grid (strategy_smoother_fixedpoint / strategy_smoother_fixedinterval,
// estimators_and_losses.py:473-717; Smoother.finalize :437-470; MarkovSequence.evaluate_marginals :156-178), not for
// speed: every matrix lives in shared memory, row-major, and ONE generic CTA-cooperative Householder routine (a thread
// per trailing column, dlarfg reflectors like everywhere else in this library) serves all five triangularisations of
// an attempt. The filter's register-column kernel (pdeq_loop_dense.cuh) remains the fast path for dense filters.
//
// Control flow, checkpoint handling, calibration, error estimate and controller restate the same reference lines as
// the lane-per-dimension kernel's smoother (pdeq_loop_group.cuh), operation for operation in the dense algebra.
#pragma once

#include "pdeq_dense_qr.cuh"
#include "pdeq_limits.cuh"
#include "pdeq_loop_thread.cuh"

namespace pdeq {

constexpr int K3S_THREADS = 256;

template <class VF, int NU, bool TS0>
struct DenseSmootherLoop {
  static constexpr int n = NU + 1;
  static constexpr int q = VF::order;
  static constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  static constexpr int D = VF::fixed_dim;
  static_assert(D > 0, "the dense kernels need a compile-time ODE dimension");
  static constexpr int N = n * D, HW = (q + 1) * D, LD = 2 * N, NT = K3S_THREADS;
  static constexpr int NN = N * N;
  // a stored conditional: gain | noise factor (full lower) | mean | tl | to    (preconditioned coordinates)
  static constexpr int C_G = 0, C_XI = NN, C_MEAN = 2 * NN, C_TL = 2 * NN + N, C_TO = C_TL + n, NFC = C_TO + n;
  // a stored state: mean | factor (full lower)
  static constexpr int NFS = N + NN;

  // shared memory (doubles)
  static constexpr int O_W = 0, O_L = O_W + 4 * NN, O_LP = O_L + NN, O_BW = O_LP + NN,  // inner conditional (NFC)
                       O_CC = O_BW + NFC,                                               // carried conditional (NFC)
                       O_M = O_CC + NFC, O_MP = O_M + N, O_MNEW = O_MP + N, O_MT = O_MNEW + N, O_MOBT = O_MT + N,
                       O_VTMP = O_MOBT + N, O_INVD = O_VTMP + N, O_HS = O_INVD + N, O_RY = O_HS + D * HW,
                       O_MOBS = O_RY + D * D, O_WHT = O_MOBS + D, O_STD = O_WHT + D, O_REF = O_STD + D,
                       O_LAM = O_REF + D, O_P = O_LAM + D, O_PINV = O_P + 8, O_RED = O_PINV + 8, O_BC = O_RED + 40,
                       O_END = O_BC + 8;
  PDEQ_HDI static constexpr size_t smem_doubles() { return (size_t)((O_END + 1) / 2 * 2); }
  PDEQ_HDI static constexpr size_t ring_doubles_per_cta(int T) { return (size_t)T * NFC + NFS + NFC; }

  struct VecAcc {
    const double* u;
    PDEQ_DI double operator()(int k, int i) const { return u[k * D + i]; }
  };

  PDEQ_DI static double block_sum(double v, double* red) { return dense_block_sum<NT>(v, red); }
  PDEQ_DI static void qr_smem(double* W, int ld, int M, int NC, int NP, double* red) {
    dense_qr_smem<NT>(W, ld, M, NC, NP, red);
  }

  struct Ctx {
    double *W, *L, *Lp, *bw, *cc, *m, *mp, *mnew, *mt, *mobt, *vtmp, *invd, *Hs, *RY, *mobs, *wht, *stdv, *refv, *lam, *p,
        *pinv, *red, *bc;
    const double (*A)[PDEQ_MAX_COEFFS];
    const double (*Qm)[PDEQ_MAX_COEFFS];
    double inv_sqrt_d;
  };

  PDEQ_DI static void cond_identity(double* c) {
    for (int e = threadIdx.x; e < NFC; e += NT) {
      double v = 0.0;
      if (e < NN) v = (e / N == e % N) ? 1.0 : 0.0;
      if (e >= C_TL) v = 1.0;
      c[e] = v;
    }
  }
  PDEQ_DI static void copy(double* dst, const double* src, int count) {
    for (int e = threadIdx.x; e < count; e += NT) dst[e] = src[e];
  }

  // mp = p (A (pinv m)); also mt = pinv m and mobt = A (pinv m) for the reverted transition
  PDEQ_DI static void predict_mean(const Ctx& c, const double* msrc) {
    for (int e = threadIdx.x; e < N; e += NT) c.mt[e] = c.pinv[e / D] * msrc[e];
    __syncthreads();
    for (int e = threadIdx.x; e < N; e += NT) {
      const int i = e / D, jd = e % D;
      double acc = 0.0;
      for (int k = i; k < n; ++k) acc = fma(c.A[i][k], c.mt[k * D + jd], acc);
      c.mobt[e] = acc;
      c.mp[e] = c.p[i] * acc;
    }
    __syncthreads();
  }

  // DenseOdeTs0.linearize :243-259 / DenseResidual.linearize :290-334 at `mean`
  PDEQ_DI static void linearise_at(const Ctx& c, const double* mean, double tt, const double (&params)[P]) {
    for (int e = threadIdx.x; e < D * HW; e += NT) c.Hs[e] = 0.0;
    __syncthreads();
    VecAcc acc{mean};
    for (int jd = threadIdx.x; jd < D; jd += NT) {
      const double f = VF::template component<double>(jd, D, acc, params, tt);
      c.Hs[jd * HW + q * D + jd] = 1.0;
      if (TS0) {
        c.mobs[jd] = mean[q * D + jd] + (-f);
      } else {
        const double rres = mean[q * D + jd] - f;
        double hm = mean[q * D + jd];
        for (int cc = 0; cc < q; ++cc) {
          for (int l = 0; l < D; ++l) {
            const double hv = -VF::jac(jd, cc, l, D, acc, params, tt);
            c.Hs[jd * HW + cc * D + l] = hv;
            hm = fma(hv, mean[cc * D + l], hm);
          }
        }
        c.mobs[jd] = hm + (rres - hm);
      }
    }
    __syncthreads();
  }

  // whitened residual of an observation with upper-triangular factor RY: solve R^T w = mobs, rms; optionally the row
  // norms of R^T. One thread.
  PDEQ_DI static double whiten(const Ctx& c, bool want_std) {
    double ss = 0.0;
    for (int i = 0; i < D; ++i) {
      double acc = c.mobs[i];
      for (int l = 0; l < i; ++l) acc = fma(-c.RY[l * D + i], c.wht[l], acc);
      c.wht[i] = acc * fast_rcp(c.RY[i * D + i]);
      ss = fma(c.wht[i], c.wht[i], ss);
      if (want_std) {
        double rn = 0.0;
        for (int l = 0; l <= i; ++l) rn = fma(c.RY[l * D + i], c.RY[l * D + i], rn);
        c.stdv[i] = safe_sqrt(rn);
      }
    }
    return safe_sqrt(ss) * c.inv_sqrt_d;
  }

  // Lp <- the noise-only factor |p| (s q (x) diag(lam)) of the transition (apply_flat, ssm_impl_dense.py:15-22)
  PDEQ_DI static void noise_factor(const Ctx& c, double s) {
    for (int e = threadIdx.x; e < NN; e += NT) {
      const int row = e / N, col = e % N, ci = row / D, cj = row % D, ri = col / D, rj = col % D;
      c.Lp[e] = (cj == rj && ci >= ri) ? fabs(c.p[ci]) * s * c.Qm[ci][ri] * c.lam[cj] : 0.0;
    }
    __syncthreads();
  }

  // R_obs = qr_r([(H Lp)^T ; damp I]) -> RY (DenseLatentCond.marginalise of the observation model)
  PDEQ_DI static void observe_marginal(const Ctx& c, double damp) {
    for (int e = threadIdx.x; e < (N + D) * D; e += NT) {
      const int r = e / D, a = e % D;
      double val = 0.0;
      if (r < N) {
        for (int k = r; k < HW; ++k) val = fma(c.Hs[a * HW + k], c.Lp[k * N + r], val);
      } else if (r - N == a) {
        val = damp;
      }
      c.W[r * LD + a] = val;
    }
    __syncthreads();
    qr_smem(c.W, LD, N + D, D, D, c.red);
    for (int e = threadIdx.x; e < D * D; e += NT) c.RY[e] = (e / D <= e % D) ? c.W[(e / D) * LD + e % D] : 0.0;
    __syncthreads();
  }

  // DenseLatentCond.revert of the observation model on N(mean_in, Lp Lp^T) (:51-77): triangularise
  // [(H Lp)^T, Lp^T ; damp I, 0], observation columns first. Afterwards R_Y is in RY, the corrected factor's transpose
  // in rows D.. / columns D.. of W, and (with_mean) mean_out = mean_in - gain mobs.
  PDEQ_DI static void correct(const Ctx& c, double damp, bool with_mean, const double* mean_in, double* mean_out,
                              bool lstsq) {
    for (int e = threadIdx.x; e < (N + D) * (N + D); e += NT) {
      const int r = e / (N + D), col = e % (N + D);
      double val = 0.0;
      if (col < D) {
        if (r < N) {
          for (int k = r; k < HW; ++k) val = fma(c.Hs[col * HW + k], c.Lp[k * N + r], val);
        } else if (r - N == col) {
          val = damp;
        }
      } else if (r < N && r <= col - D) {
        val = c.Lp[(col - D) * N + r];
      }
      c.W[r * LD + col] = val;
    }
    __syncthreads();
    qr_smem(c.W, LD, N + D, N + D, N + D, c.red);
    for (int e = threadIdx.x; e < D * D; e += NT) c.RY[e] = (e / D <= e % D) ? c.W[(e / D) * LD + e % D] : 0.0;
    __syncthreads();
    if (with_mean) {
      for (int k = threadIdx.x; k < N; k += NT) {  // gain^T[:, k] = R_Y^-1 R12[:, k]
        double x[D];
#pragma unroll
        for (int i = D - 1; i >= 0; --i) {
          double acc = c.W[i * LD + D + k];
#pragma unroll
          for (int l = i + 1; l < D; ++l) acc = fma(-c.RY[i * D + l], x[l], acc);
          const double piv = c.RY[i * D + i];
          x[i] = (lstsq && piv == 0.0) ? 0.0 : acc * fast_rcp(piv);
        }
        double corr = 0.0;
#pragma unroll
        for (int a_ = 0; a_ < D; ++a_) corr = fma(x[a_], c.mobs[a_], corr);
        mean_out[k] = mean_in[k] - corr;
      }
      __syncthreads();
    }
  }
  // factor <- the corrected factor left in W by `correct`
  PDEQ_DI static void take_corrected_factor(const Ctx& c, double* Ldst) {
    for (int e = threadIdx.x; e < NN; e += NT) {
      const int row = e / N, col = e % N;
      Ldst[e] = (col <= row) ? c.W[(D + col) * LD + D + row] : 0.0;
    }
    __syncthreads();
  }

  // DenseLatentCond.revert of the IWP transition on N(msrc, Lsrc Lsrc^T) (ssm_impl_dense.py:51-77, 347-363): the
  // predicted factor -> Lp, the backward conditional (preconditioned coordinates) -> c.bw. mt / mobt must hold
  // pinv m and A (pinv m) (predict_mean). `smooth` = false: prediction only (the 2N x N stack).
  PDEQ_DI static void revert_transition(const Ctx& c, const double* Lsrc, double s, bool smooth) {
    const int ncol = smooth ? 2 * N : N;
    for (int e = threadIdx.x; e < 2 * N * ncol; e += NT) {
      const int r = e / ncol, col = e % ncol;
      double val = 0.0;
      if (col < N) {
        const int ci = col / D, cj = col % D;
        if (r < N) {
          for (int k = ci; k < n; ++k) {
            const int row = k * D + cj;
            if (row >= r) val = fma(c.A[ci][k], fabs(c.pinv[k]) * Lsrc[row * N + r], val);
          }
        } else {
          const int rr = r - N, ri = rr / D, rj = rr % D;
          val = (cj == rj && ci >= ri) ? s * c.Qm[ci][ri] * c.lam[cj] : 0.0;
        }
      } else if (r < N) {
        const int k = col - N;
        val = (r <= k) ? fabs(c.pinv[k / D]) * Lsrc[k * N + r] : 0.0;
      }
      c.W[r * LD + col] = val;
    }
    __syncthreads();
    qr_smem(c.W, LD, 2 * N, ncol, N, c.red);
    for (int e = threadIdx.x; e < NN; e += NT) {
      const int row = e / N, col = e % N;
      c.Lp[e] = (col <= row) ? fabs(c.p[row / D]) * c.W[col * LD + row] : 0.0;
    }
    if (!smooth) {
      __syncthreads();
      return;
    }
    for (int i = threadIdx.x; i < N; i += NT) c.invd[i] = fast_rcp(c.W[i * LD + i]);
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += NT) {  // row k of the gain: solve R_Y x = R12[:, k], in place
      for (int i = N - 1; i >= 0; --i) {
        double acc = c.W[i * LD + N + k];
        for (int l = i + 1; l < N; ++l) acc = fma(-c.W[i * LD + l], c.W[l * LD + N + k], acc);
        c.W[i * LD + N + k] = acc * c.invd[i];
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NN; e += NT) c.bw[C_G + e] = c.W[(e % N) * LD + N + e / N];
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += NT) {
      double acc = c.mt[k];
      for (int i = 0; i < N; ++i) acc = fma(-c.bw[C_G + k * N + i], c.mobt[i], acc);
      c.bw[C_MEAN + k] = acc;
    }
    for (int i = threadIdx.x; i < n; i += NT) {
      c.bw[C_TL + i] = fast_rcp(c.p[i]);
      c.bw[C_TO + i] = fast_rcp(c.pinv[i]);
    }
    // backward noise: triangularise the N x N remainder, Xi = R^T
    qr_smem(c.W + N * LD + N, LD, N, N, N, c.red);
    for (int e = threadIdx.x; e < NN; e += NT) {
      const int row = e / N, col = e % N;
      c.bw[C_XI + e] = (col <= row) ? c.W[(N + col) * LD + N + row] : 0.0;
    }
    __syncthreads();
  }

  // outer.merge(inner) (DenseLatentCond.merge, ssm_impl_dense.py:35-49); `dst` may be `outer`.
  PDEQ_DI static void merge(const Ctx& c, const double* outer, const double* inner, double* dst) {
    const int tid = threadIdx.x;
    for (int e = tid; e < NN; e += NT) {
      const int i = e / N, j = e % N;
      double g = 0.0, s = 0.0;
      for (int k = 0; k < N; ++k) {
        const double T = outer[C_TL + k / D] * inner[C_TO + k / D];
        g = fma(outer[C_G + i * N + k], T * inner[C_G + k * N + j], g);
        if (k >= j) s = fma(outer[C_G + i * N + k], fabs(T) * inner[C_XI + k * N + j], s);
      }
      c.W[i * LD + N + j] = g;                                        // merged gain (staged)
      c.W[j * LD + i] = s;                                            // (A_o (|T| Xi_i))^T
      c.W[(N + j) * LD + i] = (i >= j) ? outer[C_XI + i * N + j] : 0.0;  // Xi_o^T
    }
    for (int i = tid; i < N; i += NT) {
      double acc = 0.0;
      for (int k = 0; k < N; ++k) {
        const double T = outer[C_TL + k / D] * inner[C_TO + k / D];
        acc = fma(outer[C_G + i * N + k], T * inner[C_MEAN + k], acc);
      }
      c.vtmp[i] = acc + outer[C_MEAN + i];
    }
    __syncthreads();
    qr_smem(c.W, LD, 2 * N, N, N, c.red);
    for (int e = tid; e < NN; e += NT) {
      const int i = e / N, j = e % N;
      dst[C_G + e] = c.W[i * LD + N + j];
      dst[C_XI + e] = (j <= i) ? c.W[j * LD + i] : 0.0;
    }
    for (int i = tid; i < N; i += NT) dst[C_MEAN + i] = c.vtmp[i];
    for (int i = tid; i < n; i += NT) {
      const double keep = outer[C_TO + i];
      dst[C_TL + i] = inner[C_TL + i];
      dst[C_TO + i] = keep;
    }
    __syncthreads();
  }

  // cond.marginalise(N(m, L L^T)) (DenseLatentCond.marginalise :24-33), m and L updated in place
  PDEQ_DI static void marginalise(const Ctx& c, const double* cond, double* m, double* L) {
    const int tid = threadIdx.x;
    for (int e = tid; e < N; e += NT) c.mt[e] = cond[C_TL + e / D] * m[e];
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
      double acc = 0.0;
      for (int k = 0; k < N; ++k) acc = fma(cond[C_G + i * N + k], c.mt[k], acc);
      c.vtmp[i] = cond[C_TO + i / D] * (acc + cond[C_MEAN + i]);
    }
    for (int e = tid; e < NN; e += NT) {
      const int i = e / N, j = e % N;
      double g = 0.0;
      for (int k = j; k < N; ++k) g = fma(cond[C_G + i * N + k], fabs(cond[C_TL + k / D]) * L[k * N + j], g);
      c.W[j * LD + i] = g;
      c.W[(N + j) * LD + i] = (i >= j) ? cond[C_XI + i * N + j] : 0.0;
    }
    __syncthreads();
    qr_smem(c.W, LD, 2 * N, N, N, c.red);
    for (int e = tid; e < NN; e += NT) {
      const int i = e / N, j = e % N;
      L[e] = (j <= i) ? fabs(cond[C_TO + i / D]) * c.W[j * LD + i] : 0.0;
    }
    for (int i = tid; i < N; i += NT) m[i] = c.vtmp[i];
    __syncthreads();
  }

  PDEQ_DI static void emit(const LoopArgs& a, long b, int ck, double t, const double* m, const double* L, double scale,
                           int nsteps) {
    const long bt = b * a.T + ck;
    if (threadIdx.x == 0) {
      a.sol.t[bt] = t;
      a.sol.num_steps[bt] = nsteps;
      if (a.sol.output_scale != nullptr) a.sol.output_scale[bt] = scale;
    }
    for (int e = threadIdx.x; e < N; e += NT) a.sol.mean[bt * N + e] = m[e];
    if (a.sol.chol != nullptr) {
      double* co = a.sol.chol + bt * (long)NN;
      for (int e = threadIdx.x; e < NN; e += NT) co[e] = L[e];
    }
  }

  // MarkovSequence.conditional after rescale_cholesky, in natural coordinates (preconditioner_apply :79-84)
  PDEQ_DI static void emit_conditional(const LoopArgs& a, long bt, const double* cond, double scale) {
    for (int e = threadIdx.x; e < N; e += NT) a.sol.bw_mean[bt * N + e] = cond[C_TO + e / D] * cond[C_MEAN + e];
    for (int e = threadIdx.x; e < NN; e += NT) {
      const int i = e / N, k = e % N;
      const double to = cond[C_TO + i / D];
      a.sol.bw_gain[bt * (long)NN + e] = to * cond[C_G + e] * cond[C_TL + k / D];
      a.sol.bw_chol[bt * (long)NN + e] = (k <= i) ? scale * fabs(to) * cond[C_XI + e] : 0.0;
    }
  }

  PDEQ_DI static void run(const LoopArgs& a, double* __restrict__ smem, double* __restrict__ ring_all) {
    const pdeq_config& cfg = a.cfg;
    const double* __restrict__ fact = cfg.factorials;
    const double* __restrict__ ifact = cfg.inv_factorials;
    const bool adaptive = a.fixed_grid == 0;
    const bool clip = cfg.clip_dt != 0;
    const bool needs_interp = adaptive && !clip;
    const int T = a.T;
    const long B = a.prob.num_instances;
    const int max_attempts = cfg.max_attempts > 0 ? cfg.max_attempts : 0x7fffffff;
    const double neg_inv_n = -1.0 / (double)n;
    const int tid = threadIdx.x;

    Ctx c;
    c.W = smem + O_W; c.L = smem + O_L; c.Lp = smem + O_LP; c.bw = smem + O_BW; c.cc = smem + O_CC;
    c.m = smem + O_M; c.mp = smem + O_MP; c.mnew = smem + O_MNEW; c.mt = smem + O_MT; c.mobt = smem + O_MOBT;
    c.vtmp = smem + O_VTMP; c.invd = smem + O_INVD; c.Hs = smem + O_HS; c.RY = smem + O_RY; c.mobs = smem + O_MOBS;
    c.wht = smem + O_WHT; c.stdv = smem + O_STD; c.refv = smem + O_REF; c.lam = smem + O_LAM; c.p = smem + O_P;
    c.pinv = smem + O_PINV; c.red = smem + O_RED; c.bc = smem + O_BC;
    c.A = cfg.sys_a; c.Qm = cfg.sys_q;
    c.inv_sqrt_d = rsqrt((double)D);

    // global scratch of this CTA: the conditionals of checkpoints 1..T-1 (slot 0 unused), then interp_from
    double* ring = ring_all + (size_t)blockIdx.x * ring_doubles_per_cta(T);
    double* st_if = ring + (size_t)T * NFC;  // state (NFS) | conditional (NFC)

    double params[P];
    double t = 0.0, dt = 0.0, ctrl_lprev = 0.0, ndata = 0.0, t_next = 0.0, t_if = 0.0, sig = 1.0, run_scale = 0.0;
    int nsteps = 0, nattempts = 0, ck = 0, status = 0;
    long b = -1;
    bool need_load = true;

    auto set_preconditioner = [&](double h) {
      __syncthreads();
      if (tid == 0) {
        double pp[n], pi[n];
        preconditioner<n>(h, ifact, fact, pp, pi);
#pragma unroll
        for (int i = 0; i < n; ++i) {
          c.p[i] = pp[i];
          c.pinv[i] = pi[i];
        }
      }
      __syncthreads();
    };

    while (true) {
      if (need_load) {
        __syncthreads();
        if (tid == 0) reinterpret_cast<long*>(c.bc)[0] = (long)atomicAdd(a.work_counter, 1ULL);
        __syncthreads();
        b = reinterpret_cast<long*>(c.bc)[0];
        if (b >= B) break;
        if (a.prob.order != nullptr) b = (long)a.prob.order[b];
        need_load = false;
#pragma unroll
        for (int k = 0; k < P; ++k)
          params[k] = (VF::num_params > 0) ? a.prob.params[b * a.prob.params_stride + k] : 0.0;
        for (int e = tid; e < N; e += NT) c.m[e] = a.prob.tcoeffs[b * N + e];
        for (int e = tid; e < NN; e += NT) {
          double v = 0.0;
          if (a.prob.init_std != nullptr && e / N == e % N) v = a.prob.init_std[b * a.prob.init_std_stride + e / N];
          c.L[e] = v;
        }
        for (int e = tid; e < D; e += NT)
          c.lam[e] = a.prob.prior_scale != nullptr ? a.prob.prior_scale[b * a.prob.prior_scale_stride + e] : 1.0;
        cond_identity(c.cc);
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
        if (cfg.constraint_init != 0) {
          // solver.init with constraint_init (solvers.py:361-372, 526-537, 670-680)
          linearise_at(c, c.m, t, params);
          copy(c.Lp, c.L, NN);
          __syncthreads();
          correct(c, a.damp, true, c.m, c.mnew, true);
          if (cfg.solver == PDEQ_SOLVER_MLE) {
            // solver_mle.init (solvers.py:361-374): the update at t0 is the first datum of the running calibration
            if (tid == 0) c.bc[2] = whiten(c, false);
            __syncthreads();
            run_scale = c.bc[2];
            ndata = 1.0;
          }
          copy(c.m, c.mnew, N);
          take_corrected_factor(c, c.L);
        }
        emit(a, b, 0, t, c.m, c.L, 1.0, 0);
        if (needs_interp) {
          copy(st_if, c.m, N);
          copy(st_if + N, c.L, NN);
          copy(st_if + NFS, c.cc, NFC);
        }
        t_if = t;
        ck = 1;
        t_next = (T > 1) ? a.grid[1] : t;
        __syncthreads();
      }

      const bool at_checkpoint = (ck >= T) || (adaptive && !(t + a.eps < t_next));
      if (at_checkpoint) {
        if (ck < T) {
          double* slot = ring + (size_t)ck * NFC;
          if (needs_interp && t > t_next + a.eps) {
            // interp_beyond_t1 (solvers_via_adaptive_steps.py:346-360 -> estimators_and_losses.py:549-591): from
            // interp_from to the checkpoint, merged with interp_from's conditional -> the checkpoint's conditional;
            // then from the interpolated point to the overstepped state with a fresh backward model
            const double dta = t_next - t_if;
            set_preconditioner(dta);
            predict_mean(c, st_if);
            revert_transition(c, st_if + N, safe_sqrt(fabs(dta)) * sig, true);
            merge(c, st_if + NFS, c.bw, slot);
            emit(a, b, ck, t_next, c.mp, c.Lp, sig, nsteps);
            __syncthreads();
            copy(st_if, c.mp, N);
            copy(st_if + N, c.Lp, NN);
            __syncthreads();
            const double dtb = t - t_next;
            set_preconditioner(dtb);
            predict_mean(c, st_if);
            revert_transition(c, st_if + N, safe_sqrt(fabs(dtb)) * sig, true);
            cond_identity(c.cc);
            __syncthreads();
            merge(c, c.cc, c.bw, c.cc);
            cond_identity(st_if + NFS);
            t_if = t_next;
          } else {
            // interp_at_t1 (solvers_via_adaptive_steps.py:362-375 -> estimators_and_losses.py:534-547)
            emit(a, b, ck, t, c.m, c.L, sig, nsteps);
            copy(slot, c.cc, NFC);
            __syncthreads();
            cond_identity(c.cc);
            if (needs_interp) {
              copy(st_if, c.m, N);
              copy(st_if + N, c.L, NN);
              cond_identity(st_if + NFS);
            }
            t_if = t;
          }
          ck += 1;
          if (ck < T) t_next = a.grid[ck];
          __syncthreads();
        }
        if (ck >= T) {
          // -------------------------------------------------------------- finish the instance
          double bad = 0.0;
          for (int e = tid; e < N; e += NT) bad += isfinite(c.m[e]) ? 0.0 : 1.0;
          bad = block_sum(bad, c.red);
          if (status == 0 && bad > 0.0) status = PDEQ_STATUS_NONFINITE;
          double fin = 1.0;
          if (cfg.solver == PDEQ_SOLVER_MLE) {
            fin = run_scale;
            if (cfg.correct_asymptotic_underconfidence) fin = fin / sqrt((double)nsteps);
          }
          if (status == 0) {
            // Smoother.finalize (estimators_and_losses.py:437-470): marginalise the last state through its own
            // conditional (the identity for the fixed-point smoother; on a fixed grid the reference hands finalize the
            // last grid state, whose conditional is the last interval's -- PDEQ_STRATEGY_FIXEDINTERVAL reproduces
            // that, ..._ALIGNED starts from the filtering marginal), then the backward recursion.
            const double* last = (!adaptive && cfg.strategy == PDEQ_STRATEGY_FIXEDINTERVAL && T > 1)
                                     ? ring + (size_t)(T - 1) * NFC
                                     : c.cc;
            marginalise(c, last, c.m, c.L);
            for (int k = T - 1; k >= 0; --k) {
              const long bt = b * T + k;
              if (a.sol.filt_mean != nullptr) {
                for (int e = tid; e < N; e += NT) a.sol.filt_mean[bt * N + e] = a.sol.mean[bt * N + e];
                if (a.sol.filt_chol != nullptr && a.sol.chol != nullptr)
                  for (int e = tid; e < NN; e += NT) a.sol.filt_chol[bt * (long)NN + e] = fin * a.sol.chol[bt * (long)NN + e];
              }
              for (int e = tid; e < N; e += NT) a.sol.mean[bt * N + e] = c.m[e];
              if (a.sol.chol != nullptr)
                for (int e = tid; e < NN; e += NT) a.sol.chol[bt * (long)NN + e] = fin * c.L[e];
              if (k > 0) {
                const double* cond = ring + (size_t)k * NFC;
                if (a.sol.bw_gain != nullptr) emit_conditional(a, bt, cond, fin);
                __syncthreads();
                marginalise(c, cond, c.m, c.L);
              }
            }
          } else if (cfg.solver == PDEQ_SOLVER_MLE && a.sol.chol != nullptr) {
            for (int k = 0; k < T; ++k) {
              double* co = a.sol.chol + (b * T + k) * (long)NN;
              for (int e = tid; e < NN; e += NT) co[e] = fin * co[e];
            }
          }
          if (cfg.solver == PDEQ_SOLVER_MLE && a.sol.output_scale != nullptr && tid == 0)
            for (int k = 0; k < T; ++k) a.sol.output_scale[b * T + k] = fin;
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
      double dtc;
      if (adaptive) {
        dtc = clip ? fmin(dt, t_next - t) : dt;
      } else {
        dtc = a.grid[ck] - a.grid[ck - 1];
      }
      const double sq = safe_sqrt(fabs(dtc));
      const double t_new = t + dtc;
      set_preconditioner(dtc);
      predict_mean(c, c.m);
      linearise_at(c, c.mp, t_new, params);

      const bool need_obs = adaptive ? (cfg.solver == PDEQ_SOLVER_DYNAMIC || cfg.error == PDEQ_ERROR_RESIDUAL_STD)
                                     : (cfg.solver == PDEQ_SOLVER_DYNAMIC);
      double sig_new = 1.0, whitened_obs = 0.0;
      if (need_obs) {
        noise_factor(c, sq);
        observe_marginal(c, a.damp);
        if (tid == 0) c.bc[1] = whiten(c, true);
        __syncthreads();
        whitened_obs = c.bc[1];
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig_new = whitened_obs;
      }
      if (adaptive && cfg.error != PDEQ_ERROR_RESIDUAL_STD) {
        // error_state_std (solvers.py:1070-1086): Bayes rule on the zero-error extrapolation
        const int idx = cfg.derivative_idx;
        noise_factor(c, sq);
        correct(c, a.damp, false, nullptr, nullptr, false);
        if (tid == 0) c.bc[3] = whiten(c, false);
        __syncthreads();
        for (int cj = tid; cj < D; cj += NT) {
          const int col = idx * D + cj;
          double rn = 0.0;
          for (int r = 0; r <= col; ++r) rn = fma(c.W[(D + r) * LD + D + col], c.W[(D + r) * LD + D + col], rn);
          c.stdv[cj] = c.bc[3] * safe_sqrt(rn);
        }
        __syncthreads();
      }

      // extrapolate (reverting the transition), correct; the proposal's factor stays in W until the step is decided
      revert_transition(c, c.L, sq * sig_new, true);
      correct(c, a.damp, true, c.mp, c.mnew, false);
      double run_new = run_scale;
      if (cfg.solver == PDEQ_SOLVER_MLE) {
        if (tid == 0) c.bc[2] = whiten(c, false);
        __syncthreads();
        const double w1 = sqrt(ndata / (ndata + 1.0)), w2 = sqrt(1.0 / (ndata + 1.0));
        const double x1 = w1 * run_scale, x2 = w2 * c.bc[2];
        run_new = safe_sqrt(fma(x1, x1, x2 * x2));
      }
      __syncthreads();

      bool accept = true;
      double dt_next = dt;
      if (adaptive) {
        int kpow, idx = 0;
        if (cfg.error == PDEQ_ERROR_RESIDUAL_STD) {
          for (int e = tid; e < D; e += NT) {
            c.stdv[e] = whitened_obs * c.stdv[e];
            c.refv[e] = fmax(fabs(c.m[e]), fabs(c.mnew[e]));
          }
          kpow = q;
        } else {
          idx = cfg.derivative_idx;
          for (int e = tid; e < D; e += NT) c.refv[e] = fmax(fabs(c.m[idx * D + e]), fabs(c.mnew[idx * D + e]));
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
            for (int e = 0; e < D; ++e) {
              const double w = (c.stdv[e] * escale) * fast_rcp(fma(a.rtol, c.refv[e], a.atol));
              ss = fma(w, w, ss);
            }
            norm = safe_sqrt(ss) * c.inv_sqrt_d;
          } else {
            double se2 = 0.0, sr2 = 0.0;
            for (int e = 0; e < D; ++e) {
              const double ea = c.stdv[e] * escale;
              se2 = fma(ea, ea, se2);
              sr2 = fma(c.refv[e], c.refv[e], sr2);
            }
            norm = (safe_sqrt(se2) * c.inv_sqrt_d) * fast_rcp(fma(a.rtol, safe_sqrt(sr2) * c.inv_sqrt_d, a.atol));
          }
          c.bc[4] = neg_inv_n * log2(norm);
        }
        __syncthreads();
        const double lep = c.bc[4];
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
      dt = dt_next;
      if (accept) {
        // interp_from <- step_from, only if the checkpoint branch that follows will read it (see pdeq_loop_group.cuh)
        const bool reaches = needs_interp && !(t_new + a.eps < t_next);
        if (reaches) {
          copy(st_if, c.m, N);
          copy(st_if + N, c.L, NN);
          copy(st_if + NFS, c.cc, NFC);
          t_if = t;
        }
        __syncthreads();
        copy(c.m, c.mnew, N);
        take_corrected_factor(c, c.L);
        merge(c, c.cc, c.bw, c.cc);
        if (cfg.solver == PDEQ_SOLVER_DYNAMIC) sig = sig_new;
        run_scale = run_new;
        ndata += 1.0;
        t = t_new;
        nsteps += 1;
        if (!adaptive) {
          // fixed grid: every grid point is a checkpoint; the fixed-interval smoother (estimators_and_losses.py:612-620)
          copy(ring + (size_t)ck * NFC, c.cc, NFC);
          __syncthreads();
          cond_identity(c.cc);
          emit(a, b, ck, t, c.m, c.L, sig, nsteps);
          ck += 1;
        }
        __syncthreads();
      }
    }
  }
};

template <class VF, int NU, bool TS0>
__global__ void __launch_bounds__(K3S_THREADS, 1) k3s_loop_kernel(const __grid_constant__ LoopArgs a, double* ring) {
  extern __shared__ double smem_k3s[];
  DenseSmootherLoop<VF, NU, TS0>::run(a, smem_k3s, ring);
}

}  // namespace pdeq
