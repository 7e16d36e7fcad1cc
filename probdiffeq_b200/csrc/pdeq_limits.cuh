// Capacity limits the host-side kernel selection (pdeq_api.cu: select_loop) shares with the kernels: how many ODE
// dimensions the lane-per-dimension kernels take, and how much shared memory an instance of the dense kernel needs.
#pragma once

#include <cstddef>

#ifndef PDEQ_K3_TPC2_MAX_N
#define PDEQ_K3_TPC2_MAX_N 48
#endif
// Where the warp-per-instance smoother keeps its working columns: 1 = an L2-resident global slot per resident warp
// (24 KB of shared memory per instance, 8 warps per SM -- the register limit), 0 = shared memory (44 KB, 5 warps per
// SM). Config 3 on B200: 2.19 s -> 1.73 s with 1.
#ifndef PDEQ_K2_WARP_WK_GLOBAL
#define PDEQ_K2_WARP_WK_GLOBAL 1
#endif
#ifndef PDEQ_K3_MIN_BLOCKS
#define PDEQ_K3_MIN_BLOCKS 3
#endif

namespace pdeq {

constexpr int K2_MAX_DPL = 4;        // dimensions per lane in CTA mode (pdeq_loop_group.cuh)
constexpr int K2_CTA_THREADS = 256;  // upper bound of a CTA-mode block

// Shared memory of one instance (doubles). Host and device agree through this one function.
struct DenseSmemLayout {
  int N, d;
  size_t off_Lfrom, off_Lif, off_vec, total;
  // lanes per column of the dense kernel: 2 up to PDEQ_K3_TPC2_MAX_N state coefficients, else 4
  __host__ __device__ static constexpr int tpc(int N) { return N <= PDEQ_K3_TPC2_MAX_N ? 2 : 4; }
  __host__ __device__ static constexpr int rows_per_thread(int N) { return (2 * N + tpc(N) - 1) / tpc(N); }
  __host__ __device__ static constexpr int rpad(int N) { return (rows_per_thread(N) + 1) / 2 * 2; }
  __host__ __device__ static constexpr int threads(int N, int d) { return (tpc(N) * (N + d) + 31) / 32 * 32; }
  __host__ __device__ static DenseSmemLayout make(int n, int d, int order, bool needs_interp) {
    DenseSmemLayout s;
    s.N = n * d;
    s.d = d;
    const size_t tri = (size_t)s.N * (s.N + 1) / 2;
    const int hw = (order + 1) * d;
    size_t o = 0;
    s.off_Lfrom = o;
    o += (tri + 1) / 2 * 2;
    s.off_Lif = o;
    o += needs_interp ? (tri + 1) / 2 * 2 : 0;
    s.off_vec = o;
    // reflector buffers 2 x (TPC rpad + 2) | m_from, mp, m_new, m_if (4N) | Hs d x hw | U hw x hw | RY d x d |
    // mobs, wht, std, ref, lam (5d) | p, pinv (16) | red 8 | bc 8
    o += 2 * (size_t)(tpc(s.N) * rpad(s.N) + 2) + (size_t)4 * s.N + (size_t)d * hw + (size_t)hw * hw + (size_t)d * d +
         5 * d + 16 + 16;
    s.total = (o + 1) / 2 * 2;
    return s;
  }
};

}  // namespace pdeq
