// Kernel registry: explicit template instantiations live in their own translation units (so that nvcc
// can build them in parallel) and register a launcher under a small key at load time.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "pdeq_loop_thread.cuh"

namespace pdeq {

struct KernelKey {
  int vf, nu, fact, d, ts0;
  bool operator==(const KernelKey& o) const {
    return vf == o.vf && nu == o.nu && fact == o.fact && d == o.d && ts0 == o.ts0;
  }
};

using LoopLauncher = cudaError_t (*)(const LoopArgs&, cudaStream_t);

struct LoopEntry {
  KernelKey key;
  LoopLauncher launch;
  const char* family;  // "thread" (K1) or "group" (K2) ...
};

void register_loop(const LoopEntry& e);
const LoopEntry* find_loop(const KernelKey& key);

int device_sm_count();

// ---------------------------------------------------------------------------------------------------
// K1 launcher
// ---------------------------------------------------------------------------------------------------
template <class VF, int NU, int FACT, int D, bool TS0>
cudaError_t k1_launch(const LoopArgs& a, cudaStream_t stream) {
  using TL = ThreadLoop<VF, NU, FACT, D, TS0>;
  auto kern = k1_loop_kernel<VF, NU, FACT, D, TS0>;
  const bool needs_interp = a.fixed_grid == 0 && a.cfg.clip_dt == 0;
  const size_t smem = needs_interp ? size_t(TL::IF_SLOTS) * K1_THREADS * sizeof(double) : 0;
  cudaError_t err;
  if (smem > 48 * 1024) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_THREADS, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const long want = (a.prob.num_instances + K1_THREADS - 1) / K1_THREADS;
  const long cap = (long)per_sm * device_sm_count();
  const int grid = (int)std::max(1L, std::min(want, cap));
  kern<<<grid, K1_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

template <class VF, int NU, int FACT, int D, bool TS0>
struct K1Registrar {
  K1Registrar() { register_loop({{VF::id, NU, FACT, D, TS0 ? 1 : 0}, &k1_launch<VF, NU, FACT, D, TS0>, "thread"}); }
};

#define PDEQ_INSTANTIATE_K1(VF, NU, D)                                              \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, true> _k1_iso0_##VF##_##NU##_##D; \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, false> _k1_iso1_##VF##_##NU##_##D; \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, true> _k1_bd0_##VF##_##NU##_##D;  \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, false> _k1_bd1_##VF##_##NU##_##D;

}  // namespace pdeq
