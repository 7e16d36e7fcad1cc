// Launcher and registrar of the dense-factorisation kernels (K3, pdeq_loop_dense.cuh).
#pragma once

#include "pdeq_dispatch.cuh"
#include "pdeq_loop_dense.cuh"
#include "pdeq_smooth_dense.cuh"

namespace pdeq {

// ---------------------------------------------------------------------------------------------------
// K3 launcher: dense factorisation, CTA per instance.
// ---------------------------------------------------------------------------------------------------
template <class VF, int NU, bool TS0>
cudaError_t k3_launch(const LoopArgs& a, void*, size_t, cudaStream_t stream) {
  const bool needs_interp = a.fixed_grid == 0 && a.cfg.clip_dt == 0;
  const DenseSmemLayout lay = DenseSmemLayout::make(NU + 1, a.cfg.ode_dim, VF::order, needs_interp);
  const size_t smem = lay.total * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = k3_loop_kernel<VF, NU, TS0>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  int per_sm = 0;
  const int threads = DenseSmemLayout::threads((NU + 1) * VF::fixed_dim, VF::fixed_dim);
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const long cap = (long)per_sm * device_sm_count();
  const int grid = (int)std::max(1L, std::min((long)a.prob.num_instances, cap));
  kern<<<grid, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// K3S launcher: dense factorisation with a smoother (pdeq_smooth_dense.cuh), CTA per instance, one per SM.
// ---------------------------------------------------------------------------------------------------
template <class VF, int NU, bool TS0>
int k3s_grid(int64_t B) {
  using SL = DenseSmootherLoop<VF, NU, TS0>;
  auto kern = k3s_loop_kernel<VF, NU, TS0>;
  const size_t smem = SL::smem_doubles() * sizeof(double);
  if (smem > 227 * 1024) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K3S_THREADS, smem) != cudaSuccess) return 0;
  if (per_sm < 1) per_sm = 1;
  return (int)std::max(1L, std::min((long)B, (long)per_sm * device_sm_count()));
}

template <class VF, int NU, bool TS0>
size_t k3s_workspace(const pdeq_config&, int64_t B, int32_t T) {
  using SL = DenseSmootherLoop<VF, NU, TS0>;
  const int grid = k3s_grid<VF, NU, TS0>(B);
  return 256 + (size_t)std::max(grid, 1) * SL::ring_doubles_per_cta(T) * sizeof(double);
}

template <class VF, int NU, bool TS0>
cudaError_t k3s_launch(const LoopArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  using SL = DenseSmootherLoop<VF, NU, TS0>;
  int grid = k3s_grid<VF, NU, TS0>(a.prob.num_instances);
  if (grid < 1) return cudaErrorInvalidValue;
  const size_t per_cta = SL::ring_doubles_per_cta(a.T) * sizeof(double);
  const size_t room = (workspace_bytes - 256) / per_cta;  // never launch more CTAs than the scratch has room for
  if (room < 1) return cudaErrorMemoryAllocation;
  grid = (int)std::min((size_t)grid, room);
  double* ring = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  k3s_loop_kernel<VF, NU, TS0><<<grid, K3S_THREADS, SL::smem_doubles() * sizeof(double), stream>>>(a, ring);
  return cudaGetLastError();
}

template <class VF, int NU>
struct K3Registrar {
  static size_t ws(const pdeq_config&, int64_t, int32_t) { return 256; }
  explicit K3Registrar(int vf_id = VF::id) {
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 1, 0}, &k3_launch<VF, NU, true>, &ws, "dense"});
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 0, 0}, &k3_launch<VF, NU, false>, &ws, "dense"});
#ifndef PDEQ_K3_NO_SMOOTHER  // (SASS inspection builds of the filter kernels only)
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 1, 1}, &k3s_launch<VF, NU, true>, &k3s_workspace<VF, NU, true>,
                   "dense-smoother"});
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 0, 1}, &k3s_launch<VF, NU, false>, &k3s_workspace<VF, NU, false>,
                   "dense-smoother"});
#endif
  }
};
#define PDEQ_INSTANTIATE_K3(VF, NU) static K3Registrar<VF, NU> _k3_##VF##_##NU;

}  // namespace pdeq
