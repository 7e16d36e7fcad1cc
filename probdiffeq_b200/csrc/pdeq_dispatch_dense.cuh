// Launcher and registrar of the dense-factorisation kernels (K3, pdeq_loop_dense.cuh).
#pragma once

#include "pdeq_dispatch.cuh"
#include "pdeq_loop_dense.cuh"

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

template <class VF, int NU>
struct K3Registrar {
  static size_t ws(const pdeq_config&, int64_t, int32_t) { return 256; }
  explicit K3Registrar(int vf_id = VF::id) {
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 1, 0}, &k3_launch<VF, NU, true>, &ws, "dense"});
    register_loop({{vf_id, NU, PDEQ_FACT_DENSE, 0, 0, 0}, &k3_launch<VF, NU, false>, &ws, "dense"});
  }
};
#define PDEQ_INSTANTIATE_K3(VF, NU) static K3Registrar<VF, NU> _k3_##VF##_##NU;

}  // namespace pdeq
