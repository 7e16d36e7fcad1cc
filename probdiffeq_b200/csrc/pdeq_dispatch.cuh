// Kernel registry: explicit template instantiations live in their own translation units (so that nvcc
// can build them in parallel) and register a launcher under a small key at load time. This header holds the registry and
// the launcher of the thread-per-instance kernels (K1); pdeq_dispatch_group.cuh and pdeq_dispatch_dense.cuh add K2 and
// K3, so that a translation unit depends only on the kernel family it instantiates (probdiffeq_b200/build.py hashes
// exactly the headers a unit includes).
#pragma once
#include <cstdlib>

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "pdeq_loop_thread.cuh"

namespace pdeq {

struct KernelKey {
  int vf, nu, fact, d, ts0;  // d == 0: the kernel takes the ODE dimension at run time
  int fixedpoint = 0;
  bool operator==(const KernelKey& o) const {
    return vf == o.vf && nu == o.nu && fact == o.fact && d == o.d && ts0 == o.ts0 && fixedpoint == o.fixedpoint;
  }
};

// `workspace` points at the caller's scratch: the first 256 bytes hold the work counter, the rest is kernel specific.
using LoopLauncher = cudaError_t (*)(const LoopArgs&, void* workspace, size_t workspace_bytes, cudaStream_t);
using WorkspaceFn = size_t (*)(const pdeq_config&, int64_t num_instances, int32_t T);

struct LoopEntry {
  KernelKey key;
  LoopLauncher launch;
  WorkspaceFn workspace_bytes;
  const char* family;  // "thread" (K1), "group" (K2), "dense" (K3)
};

void register_loop(const LoopEntry& e);
const LoopEntry* find_loop(const KernelKey& key);

int device_sm_count();

// ---------------------------------------------------------------------------------------------------
// K1 launcher
// ---------------------------------------------------------------------------------------------------
// Grid of the thread-per-instance kernel: as many 128-thread CTAs as stay resident (or as the ensemble needs).
// (Round 2 measured three ways of shortening the end of the run, where lanes go idle one by one: an in-kernel pool
// of parked instances that re-forms full warps, a one-off repack between two launches, and a grid shaped so that
// every lane serves the same number of instances. All three were slower than this plain queue -- a warp that runs
// alone is only ~1.45x faster per attempt than one of twelve, so emptier SMs buy little, and the 128-register build
// that offers the extra lanes costs 5 %. What does help is the ORDER of service, see pdeq_problem.order.)
template <class VF, int NU, int FACT, int D, bool TS0, int SPEC>
cudaError_t k1_launch_impl(const LoopArgs& a, void*, size_t, cudaStream_t stream) {
  using TL = ThreadLoop<VF, NU, FACT, D, TS0, SPEC>;
  auto kern = k1_loop_kernel<VF, NU, FACT, D, TS0, SPEC>;
  const bool needs_interp = a.fixed_grid == 0 && a.cfg.clip_dt == 0;
  const size_t smem = needs_interp ? size_t(TL::IF_SLOTS) * TL::THREADS * sizeof(double) : 0;
  cudaError_t err;
  if (smem > 48 * 1024) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TL::THREADS, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const long want = (a.prob.num_instances + TL::THREADS - 1) / TL::THREADS;
  const long cap = (long)per_sm * device_sm_count();
  const int grid = (int)std::max(1L, std::min(want, cap));
  kern<<<grid, TL::THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

// Which specialised loop to use when the configuration matches: PDEQ_K1_SPEC=0 in the environment forces the
// general kernel (A/B measurements, bitwise comparison tests), 1..2 pick a specialised build (ThreadLoop: register
// budget / resident CTAs).
#ifndef PDEQ_K1_SPEC_DEFAULT
#define PDEQ_K1_SPEC_DEFAULT 1
#endif
inline int k1_spec_choice() {
  const char* e = std::getenv("PDEQ_K1_SPEC");
  const int c = e == nullptr ? PDEQ_K1_SPEC_DEFAULT : std::atoi(e);
  return (c >= 0 && c <= 2) ? c : PDEQ_K1_SPEC_DEFAULT;
}

template <class VF, int NU, int FACT, int D, bool TS0, bool HAS_SPEC = false>
cudaError_t k1_launch(const LoopArgs& a, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if constexpr (HAS_SPEC && TS0) {
    if (k1_spec_matches(a, TS0)) {
      const int spec = k1_spec_choice();
      if (spec == 1) return k1_launch_impl<VF, NU, FACT, D, TS0, 1>(a, ws, ws_bytes, stream);
      if (spec == 2) return k1_launch_impl<VF, NU, FACT, D, TS0, 2>(a, ws, ws_bytes, stream);
    }
  }
  return k1_launch_impl<VF, NU, FACT, D, TS0, 0>(a, ws, ws_bytes, stream);
}

template <class VF, int NU, int FACT, int D, bool TS0, bool HAS_SPEC = false>
struct K1Registrar {
  static size_t ws(const pdeq_config&, int64_t, int32_t) { return 256; }
  explicit K1Registrar(int vf_id = VF::id) {
    register_loop({{vf_id, NU, FACT, D, TS0 ? 1 : 0, 0}, &k1_launch<VF, NU, FACT, D, TS0, HAS_SPEC>, &ws, "thread"});
  }
};

// As PDEQ_INSTANTIATE_K1, with the compile-time specialised loop (ThreadLoop SPEC = 1) behind the isotropic ts0 entry.
#define PDEQ_INSTANTIATE_K1_WITH_SPEC(VF, NU, D)                                          \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, true, true> _k1_iso0_##VF##_##NU##_##D; \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, false> _k1_iso1_##VF##_##NU##_##D;      \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, true> _k1_bd0_##VF##_##NU##_##D;        \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, false> _k1_bd1_##VF##_##NU##_##D;

#define PDEQ_INSTANTIATE_K1(VF, NU, D)                                              \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, true> _k1_iso0_##VF##_##NU##_##D; \
  static K1Registrar<VF, NU, PDEQ_FACT_ISOTROPIC, D, false> _k1_iso1_##VF##_##NU##_##D; \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, true> _k1_bd0_##VF##_##NU##_##D;  \
  static K1Registrar<VF, NU, PDEQ_FACT_BLOCKDIAG, D, false> _k1_bd1_##VF##_##NU##_##D;

}  // namespace pdeq
