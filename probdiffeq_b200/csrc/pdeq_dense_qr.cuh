// CTA-cooperative Householder triangularisation of a matrix in shared memory: the one routine behind every
// triangularisation of the dense smoother kernel (pdeq_smooth_dense.cuh) and of the dense auxiliary kernels
// (pdeq_aux_dense.cuh). Restates util/cholesky_util.py:85-103 (triu_via_qr / sum_of_sqrtm_factors: R of a QR
// decomposition; R^T R is what the callers use) with LAPACK dlarfg reflectors, like pdeq_blockops.cuh.
#pragma once

#include "pdeq_blockops.cuh"

namespace pdeq {

// Sum over a CTA of NT threads; every thread receives the bitwise-identical result. red: >= NT / 32 doubles.
template <int NT>
PDEQ_DI double dense_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  return tot;
}

// Householder triangularisation of the first NP columns of the M x NC matrix at W (leading dimension ld), the
// reflectors applied to all NC columns. LAPACK dlarfg convention (beta = -sign(alpha) ||x||, H = I for a zero
// sub-column) in the unnormalised form of pdeq_blockops.cuh. On return R sits on and above the diagonal of the pivot
// columns, exact zeros below. Ends with a block barrier.
template <int NT>
PDEQ_DI void dense_qr_smem(double* W, int ld, int M, int NC, int NP, double* red) {
  const int tid = threadIdx.x;
  for (int j = 0; j < NP; ++j) {
    double part = 0.0;
    for (int r = j + 1 + tid; r < M; r += NT) part = fma(W[r * ld + j], W[r * ld + j], part);
    const double ss = dense_block_sum<NT>(part, red);
    const double alpha = W[j * ld + j];
    const bool live = ss != 0.0;
    const double tt = fma(alpha, alpha, ss);
    const double y = fast_rsqrt(live ? tt : 1.0);
    const double nrm = tt * y;
    const double sgn_nrm = copysign(nrm, alpha);
    const double v0 = alpha + sgn_nrm;
    const double tp = live ? fast_rcp(fma(nrm, fabs(alpha), tt)) : 0.0;
    const double beta = live ? -sgn_nrm : alpha;
    for (int c = j + 1 + tid; c < NC; c += NT) {
      double w0 = v0 * W[j * ld + c], w1 = 0.0;
      int r = j + 1;
      for (; r + 1 < M; r += 2) {
        w0 = fma(W[r * ld + j], W[r * ld + c], w0);
        w1 = fma(W[(r + 1) * ld + j], W[(r + 1) * ld + c], w1);
      }
      if (r < M) w0 = fma(W[r * ld + j], W[r * ld + c], w0);
      const double w = (w0 + w1) * tp;
      W[j * ld + c] = fma(-w, v0, W[j * ld + c]);
      for (r = j + 1; r < M; ++r) W[r * ld + c] = fma(-w, W[r * ld + j], W[r * ld + c]);
    }
    __syncthreads();
    if (live) {
      if (tid == 0) W[j * ld + j] = beta;
      for (int r = j + 1 + tid; r < M; r += NT) W[r * ld + j] = 0.0;
    }
    __syncthreads();
  }
}

}  // namespace pdeq
