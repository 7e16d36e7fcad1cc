// Fixed-point / fixed-interval smoother algebra for one (n x n) block, staged through a per-lane WORKING COLUMN.
//
// What is computed is what revert_transition + merge_cond (pdeq_blockops.cuh) compute -- LatentCond.revert for the IWP
// transition (ssm_impl_blockdiag.py:69-102, util/cholesky_util.py:27-82) and outer.merge(inner)
// (ssm_impl_blockdiag.py:45-67) under strategy_smoother_fixedpoint.predict (estimators_and_losses.py:526-532) -- with
// the same operations in the same order, hence bitwise the same numbers. What differs is where the operands live.
//
// The register-only formulation needs ~250 live doubles per lane at nu = 5 (prediction stack + reflectors, the new
// backward conditional, the carried one, the merge stack): ptxas gave it 255 registers plus a 3.5 KB local frame, the
// fully unrolled attempt body was 22 k instructions (360 KB), and the kernel ran at one instruction per ~15 cycles per
// warp: instruction-cache misses (ncu: no_instruction 6.2 stall cycles per issue) and local-memory round trips through
// DRAM (long_scoreboard 3.4). Here
//   * the new backward conditional (gain, mean, remainder) is produced ONE COLUMN PER ITERATION OF A RUN-TIME LOOP
//     and parked in the lane's working column (shared memory, [field][dim] like every other stored state);
//   * the merge reads the carried conditional one gain row per iteration of a run-time loop and overwrites it IN PLACE
//     (row i of the merged gain, mean and stack column depends on row i of the carried gain / noise only), so there is
//     neither a `merged` temporary nor a copy on acceptance;
//   * the remainder triangularisation and the merge run only once the step is ACCEPTED (the accept decision is uniform
//     across the lanes of an instance) -- without the re-triangularisation the register-only `DEFER` build paid.
// Only the prediction stack with its reflectors (in the first loop) and the scaled inner conditional (in the second)
// occupy registers, every index into a register array is a compile-time constant, and the loops keep the code small.
#pragma once

#include "pdeq_blockops.cuh"

namespace pdeq {

// Field offsets within a lane's working column (fields strided by the group's dimension count).
template <int n>
struct SmootherScratch {
  static constexpr int G = 0;               // n x n: gain of the new backward conditional, row-major
  static constexpr int XI = n * n;          // n: its mean
  static constexpr int Z = XI + n;          // n x n: remainder block (row-major); later the merge stack's top block
  static constexpr int APINV = Z + n * n;   // n: |pinv|
  static constexpr int MT = APINV + n;      // n: pinv * m
  static constexpr int NFW = MT + n;
};

// Field offsets of a stored conditional (the order of GroupLoop::cond_store).
template <int n>
struct CondLayout {
  static constexpr int TRI = n * (n + 1) / 2;
  static constexpr int G = 0, XI = n * n, XIC = XI + n, TL = XIC + TRI, TO = TL + n, NFC = TO + n;
};

// First half of LatentCond.revert for the IWP transition: triangularise the prediction stack [(A L~)^T ; (sQ)^T]
// keeping its reflectors, then push the columns of [L~^T ; 0] through them, a few per loop iteration; each gives a row of
// the gain (back substitution), an entry of the backward mean and a column of the remainder -> working column.
// `st` points at the lane's column of a stored state (fields m[n], then L packed by rows), `wk` at its working column.
template <int n>
PDEQ_DI void revert_push(const double* st, int d, const double (&p)[n], const double (&pinv)[n], double s,
                         const double (*__restrict__ A)[PDEQ_MAX_COEFFS],
                         const double (*__restrict__ Q)[PDEQ_MAX_COEFFS], double* wk, double (&Lpred)[n][n]) {
  using W = SmootherScratch<n>;
  double S[2 * n][n], v0[n], tp[n];
  {
    double L[n][n];
    int e = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int c = 0; c < n; ++c) L[i][c] = (c <= i) ? st[(n + e++) * d] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < n; ++r) {
#pragma unroll
      for (int c = 0; c < n; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int k = imax(c, r); k < n; ++k) acc = fma(A[c][k], fabs(pinv[k]) * L[k][r], acc);
        S[r][c] = acc;                               // (A L~)^T
        S[n + r][c] = (c >= r) ? s * Q[c][r] : 0.0;  // (sQ)^T
      }
    }
  }
  qr_r_inplace_keep<2 * n, n, ExtPredict<n>>(S, v0, tp);
  double inv_diag[n], mobs[n];
#pragma unroll
  for (int i = 0; i < n; ++i) inv_diag[i] = fast_rcp(S[i][i]);
  {
    double mt[n];
#pragma unroll
    for (int k = 0; k < n; ++k) mt[k] = pinv[k] * st[k * d];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = i; k < n; ++k) acc = fma(A[i][k], mt[k], acc);
      mobs[i] = acc;
    }
#pragma unroll
    for (int k = 0; k < n; ++k) {
      wk[(W::MT + k) * d] = mt[k];
      wk[(W::APINV + k) * d] = fabs(pinv[k]);
    }
  }
  // KB columns per iteration: their reflector applications and back substitutions are independent dependency chains
  // (a dot product over n + 1 rows, a triangular solve), interleaved by the compiler -- one column at a time the loop
  // ran at the FMA latency (ncu: 81 % of its stall samples were fixed-latency waits).
  constexpr int KB = (n % 3 == 0) ? 3 : ((n % 2 == 0) ? 2 : n);
#pragma unroll 1
  for (int k0 = 0; k0 < n; k0 += KB) {
    double col[KB][2 * n];  // columns k0 .. k0 + KB - 1 of [L~^T ; 0]
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int k = k0 + kk;
      const double apk = wk[(W::APINV + k) * d];
      const double* Lk = st + (size_t)(n + k * (k + 1) / 2) * d;  // row k of L; entries beyond k are discarded below
#pragma unroll
      for (int r = 0; r < n; ++r) {
        const double v = Lk[r * d];
        col[kk][r] = (r <= k) ? apk * v : 0.0;
      }
#pragma unroll
      for (int r = n; r < 2 * n; ++r) col[kk][r] = 0.0;
    }
    static_for<0, n>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr int hj = n + j;
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) {
        double w = v0[j] * col[kk][j];
#pragma unroll
        for (int r = j + 1; r <= hj; ++r) w = fma(S[r][j], col[kk][r], w);
        w *= tp[j];
        col[kk][j] = fma(-w, v0[j], col[kk][j]);
#pragma unroll
        for (int r = j + 1; r <= hj; ++r) col[kk][r] = fma(-w, S[r][j], col[kk][r]);
      }
    });
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int k = k0 + kk;
      // row k of the gain: solve R_Y x = R12[:, k]
      double Gk[n];
#pragma unroll
      for (int i = n - 1; i >= 0; --i) {
        double acc = col[kk][i];
#pragma unroll
        for (int l = i + 1; l < n; ++l) acc = fma(-S[i][l], Gk[l], acc);
        Gk[i] = acc * inv_diag[i];
      }
      double xi_acc = wk[(W::MT + k) * d];
#pragma unroll
      for (int i = 0; i < n; ++i) xi_acc = fma(-Gk[i], mobs[i], xi_acc);
#pragma unroll
      for (int i = 0; i < n; ++i) wk[(W::G + k * n + i) * d] = Gk[i];
      wk[(W::XI + k) * d] = xi_acc;
#pragma unroll
      for (int r = 0; r < n; ++r) wk[(W::Z + r * n + k) * d] = col[kk][n + r];
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < n; ++j) Lpred[i][j] = (j <= i) ? fabs(p[i]) * S[j][i] : 0.0;
  }
}

// Second half, for an accepted step: the backward noise factor (triangularisation of the remainder) and
// outer.merge(inner) with the outer conditional at `cond` (the lane's column of a stored conditional) overwritten by
// the merged one. p / pinv are the transition's preconditioner (inner.tl = 1 / p, inner.to = 1 / pinv).
template <int n>
PDEQ_DI void remainder_merge(double* cond, int d, const double (&p)[n], const double (&pinv)[n], double* wk) {
  using W = SmootherScratch<n>;
  using CL = CondLayout<n>;
  double TG[n][n], TXi[n][n], Txi[n];
  {
    double T[n];
#pragma unroll
    for (int k = 0; k < n; ++k) T[k] = cond[(CL::TL + k) * d] * fast_rcp(pinv[k]);
    {
      double Z[n][n];
#pragma unroll
      for (int r = 0; r < n; ++r) {
#pragma unroll
        for (int c = 0; c < n; ++c) Z[r][c] = wk[(W::Z + r * n + c) * d];
      }
      qr_r_inplace<n, n, ExtFull<n>>(Z);  // Xi_inner = R^T
#pragma unroll
      for (int k = 0; k < n; ++k) {
#pragma unroll
        for (int j = 0; j < n; ++j) TXi[k][j] = (j <= k) ? fabs(T[k]) * Z[j][k] : 0.0;
      }
    }
#pragma unroll
    for (int k = 0; k < n; ++k) {
      Txi[k] = T[k] * wk[(W::XI + k) * d];
#pragma unroll
      for (int j = 0; j < n; ++j) TG[k][j] = T[k] * wk[(W::G + k * n + j) * d];
    }
  }
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    double g0[n];
#pragma unroll
    for (int k = 0; k < n; ++k) g0[k] = cond[(CL::G + i * n + k) * d];
    double xacc = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) xacc = fma(g0[k], Txi[k], xacc);
    cond[(CL::XI + i) * d] = xacc + cond[(CL::XI + i) * d];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      double g = 0.0, c = 0.0;
#pragma unroll
      for (int k = 0; k < n; ++k) {
        g = fma(g0[k], TG[k][j], g);
        if (k >= j) c = fma(g0[k], TXi[k][j], c);
      }
      cond[(CL::G + i * n + j) * d] = g;
      wk[(W::Z + j * n + i) * d] = c;  // (A_o (|T| Xi_i))^T, the merge stack's top block
    }
  }
  double S[2 * n][n];
#pragma unroll
  for (int j = 0; j < n; ++j) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      S[j][i] = wk[(W::Z + j * n + i) * d];
      S[n + j][i] = (i >= j) ? cond[(CL::XIC + i * (i + 1) / 2 + j) * d] : 0.0;  // Xi_o^T
    }
  }
  qr_r_inplace<2 * n, n, ExtPredict<n>>(S);
#pragma unroll
  for (int i = 0; i < n; ++i) {
    cond[(CL::TL + i) * d] = fast_rcp(p[i]);
#pragma unroll
    for (int j = 0; j <= i; ++j) cond[(CL::XIC + i * (i + 1) / 2 + j) * d] = S[j][i];
  }
}

}  // namespace pdeq
