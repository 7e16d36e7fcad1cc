// Kernel registry: explicit template instantiations live in their own translation units (so that nvcc
// can build them in parallel) and register a launcher under a small key at load time.
#pragma once
#include <cstdlib>

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "pdeq_loop_dense.cuh"
#include "pdeq_loop_group.cuh"
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
// Tail compaction (ThreadLoop::run): PDEQ_K1_POOL=0 in the environment switches it off (A/B measurements),
// PDEQ_K1_SEG sets how many loop iterations a warp runs between two visits to the pool.
inline bool k1_pool_enabled() {
  const char* e = std::getenv("PDEQ_K1_POOL");
  return e == nullptr || std::atoi(e) != 0;
}
inline int k1_pool_seg_len() {
  const char* e = std::getenv("PDEQ_K1_SEG");
  const int c = e == nullptr ? 32 : std::atoi(e);
  return (c >= 1 && c <= 4096) ? c : 32;
}
inline int k1_pool_dissolve() {
  const char* e = std::getenv("PDEQ_K1_DISSOLVE");
  const int c = e == nullptr ? 20 : std::atoi(e);
  return (c >= 0 && c <= 32) ? c : 20;
}
inline int k1_pool_patience_ns() {
  const char* e = std::getenv("PDEQ_K1_PATIENCE_NS");
  const int c = e == nullptr ? 30000 : std::atoi(e);
  return c >= 0 ? c : 30000;
}
constexpr int K1_MAX_CTAS_PER_SM = 6;
inline size_t k1_pool_ring_entries(long lanes) {
  size_t cap = 1024;
  while ((long)cap < lanes) cap <<= 1;
  return cap;
}
inline long k1_max_lanes(int64_t num_instances) {
  const long want = (num_instances + K1_THREADS - 1) / K1_THREADS;
  return std::max(1L, std::min(want, (long)K1_MAX_CTAS_PER_SM * device_sm_count())) * K1_THREADS;
}
inline size_t k1_pool_bytes(long lanes, int park_slots) {
  return k1_pool_ring_entries(lanes) * sizeof(unsigned int) + (size_t)lanes * park_slots * sizeof(double);
}

template <class VF, int NU, int FACT, int D, bool TS0, int SPEC>
cudaError_t k1_launch_impl(const LoopArgs& a_in, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  using TL = ThreadLoop<VF, NU, FACT, D, TS0, SPEC>;
  auto kern = k1_loop_kernel<VF, NU, FACT, D, TS0, SPEC>;
  LoopArgs a = a_in;
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
  // the pool of parked instances lives behind the 256-byte header of the workspace
  const long lanes = (long)grid * TL::THREADS;
  const int park = TL::PARK_BASE + (needs_interp ? TL::IF_SLOTS : 0);
  a.pool_ring = nullptr;
  a.pool_slots = nullptr;
  a.pool_ring_mask = 0;
  a.pool_num_slots = 0;
  a.pool_seg_len = k1_pool_seg_len();
  a.pool_dissolve = k1_pool_dissolve();
  a.pool_patience_ns = k1_pool_patience_ns();
  if (k1_pool_enabled() && workspace != nullptr && workspace_bytes >= 256 + k1_pool_bytes(lanes, park)) {
    const size_t ring = k1_pool_ring_entries(lanes);
    a.pool_ring = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 256);
    a.pool_ring_mask = (unsigned int)(ring - 1);
    a.pool_slots = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256 + ring * sizeof(unsigned int));
    a.pool_num_slots = lanes;
    err = cudaMemsetAsync(a.pool_ring, 0, ring * sizeof(unsigned int), stream);
    if (err != cudaSuccess) return err;
  }
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
  static size_t ws(const pdeq_config&, int64_t num_instances, int32_t) {
    using TL = ThreadLoop<VF, NU, FACT, D, TS0, 0>;
    return 256 + k1_pool_bytes(k1_max_lanes(num_instances), TL::PARK_SLOTS_MAX);
  }
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

// ---------------------------------------------------------------------------------------------------
// K2 launcher: warp per instance for d <= 32, CTA per instance otherwise.
// ---------------------------------------------------------------------------------------------------
struct K2Plan {
  bool cta;
  int mode;  // GroupLoop MODE
  int threads, groups_per_cta, grid;
  size_t smem_bytes, ring_bytes_per_group;
};

// PDEQ_K2_SPEC=0 in the environment forces the general kernel where a specialised build exists (GroupLoop SPEC);
// 2 selects the smoother build that defers the backward conditional to accepted steps.
inline int k2_spec_choice() {
  const char* e = std::getenv("PDEQ_K2_SPEC");
  const int c = e == nullptr ? 1 : std::atoi(e);
  return (c >= 0 && c <= 2) ? c : 1;
}

template <class VF, int NU, int FACT, bool TS0, bool FP, int SPEC = 0>
cudaError_t k2_plan(const pdeq_config& cfg, int64_t B, int32_t T, bool needs_interp, K2Plan* plan) {
  using GL = GroupLoop<VF, NU, FACT, TS0, FP, 0>;
  const int d = cfg.ode_dim;
  const size_t per_group = GL::smem_doubles_per_group(d, needs_interp) * sizeof(double);
  plan->cta = d > 32;
  plan->mode = d <= 32 ? 0 : (d <= K2_CTA_THREADS ? 1 : 2);
  if (plan->cta) {
    using GC = GroupLoop<VF, NU, FACT, TS0, FP, 2>;
    plan->threads = std::min(K2_CTA_THREADS, ((d + 31) / 32) * 32);
    if ((d + plan->threads - 1) / plan->threads > GC::MAXR) return cudaErrorInvalidValue;
    plan->groups_per_cta = 1;
  } else {
    plan->groups_per_cta = 4;
    while (plan->groups_per_cta > 1 && per_group * plan->groups_per_cta > 96 * 1024) plan->groups_per_cta /= 2;
    plan->threads = 32 * plan->groups_per_cta;
  }
  plan->smem_bytes = per_group * plan->groups_per_cta;
  if (plan->smem_bytes > 227 * 1024) return cudaErrorInvalidValue;
  // per resident group: the ring of per-checkpoint conditionals plus the interp_from slot (smoother only)
  plan->ring_bytes_per_group = FP ? ((size_t)T * GL::NFC + GL::NF) * d * sizeof(double) : 0;
  int per_sm = 0;
  cudaError_t err;
  if (plan->mode == 2) {
    auto kern = k2_loop_kernel<VF, NU, FACT, TS0, FP, 2>;
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
    if (err != cudaSuccess) return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan->threads, plan->smem_bytes);
  } else if (plan->mode == 1) {
    auto kern = k2_loop_kernel<VF, NU, FACT, TS0, FP, 1>;
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
    if (err != cudaSuccess) return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan->threads, plan->smem_bytes);
  } else {
    auto kern = k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, SPEC>;
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
    if (err != cudaSuccess) return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan->threads, plan->smem_bytes);
  }
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  if (const char* cap = std::getenv("PDEQ_K2_CTAS_PER_SM")) {  // tuning knob: fewer resident CTAs per SM
    const int c = std::atoi(cap);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  const long want = (B + plan->groups_per_cta - 1) / plan->groups_per_cta;
  plan->grid = (int)std::max(1L, std::min(want, (long)per_sm * device_sm_count()));
  return cudaSuccess;
}

template <class VF, int NU, int FACT, bool TS0, bool FP>
size_t k2_workspace(const pdeq_config& cfg, int64_t B, int32_t T) {
  K2Plan plan;
  if (k2_plan<VF, NU, FACT, TS0, FP>(cfg, B, T, /*needs_interp=*/true, &plan) != cudaSuccess) return 256;
  K2Plan plan2;
  size_t groups = (size_t)plan.grid * plan.groups_per_cta;
  if (k2_plan<VF, NU, FACT, TS0, FP>(cfg, B, T, /*needs_interp=*/false, &plan2) == cudaSuccess)
    groups = std::max(groups, (size_t)plan2.grid * plan2.groups_per_cta);
  return 256 + groups * plan.ring_bytes_per_group;
}

template <class VF, int NU, int FACT, bool TS0, bool FP, bool HAS_SPEC = false>
cudaError_t k2_launch(const LoopArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const bool needs_interp = a.fixed_grid == 0 && a.cfg.clip_dt == 0;
  K2Plan plan;
  // the specialised build exists for the warp-per-instance mode only (d <= 32)
  const int choice = k2_spec_choice();
  const bool spec = HAS_SPEC && a.cfg.ode_dim <= 32 && k2_spec_matches(a) && choice != 0;
  const bool defer = spec && FP && choice == 2 && a.fixed_grid == 0;
  cudaError_t err = cudaSuccess;
  if constexpr (HAS_SPEC) {
    err = spec ? k2_plan<VF, NU, FACT, TS0, FP, 1>(a.cfg, a.prob.num_instances, a.T, needs_interp, &plan)
               : k2_plan<VF, NU, FACT, TS0, FP, 0>(a.cfg, a.prob.num_instances, a.T, needs_interp, &plan);
  } else {
    err = k2_plan<VF, NU, FACT, TS0, FP>(a.cfg, a.prob.num_instances, a.T, needs_interp, &plan);
  }
  if (err != cudaSuccess) return err;
  GroupLaunchInfo info;
  info.groups_per_cta = plan.groups_per_cta;
  info.cond_ring = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  info.if_scratch = nullptr;
  if (FP) {  // never launch more groups than the scratch has room for
    using GL = GroupLoop<VF, NU, FACT, TS0, FP, 0>;
    const size_t room = (workspace_bytes - 256) / plan.ring_bytes_per_group;
    const int max_grid = (int)(room / plan.groups_per_cta);
    if (max_grid < 1) return cudaErrorMemoryAllocation;
    plan.grid = std::min(plan.grid, max_grid);
    const size_t groups = (size_t)plan.grid * plan.groups_per_cta;
    info.if_scratch = info.cond_ring + groups * (size_t)a.T * GL::NFC * a.cfg.ode_dim;
  }
  if (plan.mode == 2)
    k2_loop_kernel<VF, NU, FACT, TS0, FP, 2><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
  else if (plan.mode == 1)
    k2_loop_kernel<VF, NU, FACT, TS0, FP, 1><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
  else if (spec) {
    if constexpr (HAS_SPEC) {
      if (defer) {
        if constexpr (FP) {
          auto kern = k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, 2>;
          err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
          if (err != cudaSuccess) return err;
          kern<<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
        }
      } else {
        k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, 1><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
      }
    }
  } else
    k2_loop_kernel<VF, NU, FACT, TS0, FP, 0><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
  return cudaGetLastError();
}

template <class VF, int NU, int FACT, bool TS0, bool FP, bool HAS_SPEC = false>
struct K2Registrar {
  explicit K2Registrar(int vf_id = VF::id) {
    register_loop({{vf_id, NU, FACT, 0, TS0 ? 1 : 0, FP ? 1 : 0}, &k2_launch<VF, NU, FACT, TS0, FP, HAS_SPEC>,
                   &k2_workspace<VF, NU, FACT, TS0, FP>, "group"});
  }
};

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
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K3_THREADS, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const long cap = (long)per_sm * device_sm_count();
  const int grid = (int)std::max(1L, std::min((long)a.prob.num_instances, cap));
  kern<<<grid, K3_THREADS, smem, stream>>>(a);
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

// as PDEQ_INSTANTIATE_K2, with the specialised warp-mode builds (GroupLoop SPEC = 1) behind the two ts0 entries
#define PDEQ_INSTANTIATE_K2_WITH_SPEC(VF, NU, FACT, TAG)                      \
  static K2Registrar<VF, NU, FACT, true, false, true> _k2_f0_##VF##_##NU##_##TAG;  \
  static K2Registrar<VF, NU, FACT, false, false> _k2_f1_##VF##_##NU##_##TAG;       \
  static K2Registrar<VF, NU, FACT, true, true, true> _k2_s0_##VF##_##NU##_##TAG;   \
  static K2Registrar<VF, NU, FACT, false, true> _k2_s1_##VF##_##NU##_##TAG;

// filter + fixed-point smoother, ts0 + ts1, for one factorisation
#define PDEQ_INSTANTIATE_K2(VF, NU, FACT, TAG)                        \
  static K2Registrar<VF, NU, FACT, true, false> _k2_f0_##VF##_##NU##_##TAG;  \
  static K2Registrar<VF, NU, FACT, false, false> _k2_f1_##VF##_##NU##_##TAG; \
  static K2Registrar<VF, NU, FACT, true, true> _k2_s0_##VF##_##NU##_##TAG;   \
  static K2Registrar<VF, NU, FACT, false, true> _k2_s1_##VF##_##NU##_##TAG;

}  // namespace pdeq
