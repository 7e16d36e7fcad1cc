// Launcher and registrars of the lane-per-dimension kernels (K2, pdeq_loop_group.cuh).
#pragma once

#include "pdeq_dispatch.cuh"
#include "pdeq_loop_group.cuh"

namespace pdeq {

// ---------------------------------------------------------------------------------------------------
// K2 launcher: warp per instance for d <= 32, CTA per instance otherwise.
// ---------------------------------------------------------------------------------------------------
struct K2Plan {
  bool cta;
  bool wk_global;  // smoother working columns in a global slot per resident group instead of shared memory
  int mode;        // GroupLoop MODE
  int threads, groups_per_cta, grid;
  size_t smem_bytes, ring_bytes_per_group;
};

// PDEQ_K2_SPEC=0 in the environment forces the general kernel where a specialised build exists (GroupLoop SPEC).
inline int k2_spec_choice() {
  const char* e = std::getenv("PDEQ_K2_SPEC");
  const int c = e == nullptr ? 1 : std::atoi(e);
  return c == 0 ? 0 : 1;
}

template <class Kern>
cudaError_t k2_occupancy(Kern kern, int threads, size_t smem_bytes, int* per_sm) {
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err != cudaSuccess) return err;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kern, threads, smem_bytes);
}

template <class VF, int NU, int FACT, bool TS0, bool FP, int SPEC = 0>
cudaError_t k2_plan(const pdeq_config& cfg, int64_t B, int32_t T, bool needs_interp, K2Plan* plan) {
  using GL = GroupLoop<VF, NU, FACT, TS0, FP, 0>;
  const int d = cfg.ode_dim;
  constexpr size_t kSmemMax = 227 * 1024;
  plan->cta = d > 32;
  plan->mode = d <= 32 ? 0 : (d <= K2_CTA_THREADS ? 1 : 2);
  plan->wk_global = false;
  int per_sm = 0;
  cudaError_t err;
  if (plan->cta) {
    using GC = GroupLoop<VF, NU, FACT, TS0, FP, 2>;
    plan->threads = std::min(K2_CTA_THREADS, ((d + 31) / 32) * 32);
    if ((d + plan->threads - 1) / plan->threads > GC::MAXR) return cudaErrorInvalidValue;
    plan->groups_per_cta = 1;
    plan->smem_bytes = GL::smem_doubles_per_group(d, needs_interp, true) * sizeof(double);
    if (FP && plan->smem_bytes > kSmemMax) {
      plan->wk_global = true;
      plan->smem_bytes = GL::smem_doubles_per_group(d, needs_interp, false) * sizeof(double);
    }
    if (plan->smem_bytes > kSmemMax) return cudaErrorInvalidValue;
    if (plan->mode == 2)
      err = k2_occupancy(k2_loop_kernel<VF, NU, FACT, TS0, FP, 2>, plan->threads, plan->smem_bytes, &per_sm);
    else
      err = k2_occupancy(k2_loop_kernel<VF, NU, FACT, TS0, FP, 1>, plan->threads, plan->smem_bytes, &per_sm);
    if (err != cudaSuccess) return err;
  } else {
    // warp per instance: 4, 2 or 1 instances per CTA -- whichever keeps the most instances resident per SM (the
    // smoother's 40+ KB per instance fit five times into an SM only as single-warp CTAs)
    plan->wk_global = FP && PDEQ_K2_WARP_WK_GLOBAL;
    const size_t per_group = GL::smem_doubles_per_group(d, needs_interp, !plan->wk_global) * sizeof(double);
    if (per_group > kSmemMax) return cudaErrorInvalidValue;
    int best_groups = 0;
    for (int gpc = 4; gpc >= 1; gpc /= 2) {
      if (per_group * gpc > kSmemMax) continue;
      int ps = 0;
      err = k2_occupancy(k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, SPEC>, 32 * gpc, per_group * gpc, &ps);
      if (err != cudaSuccess) return err;
      if (ps * gpc > best_groups) {
        best_groups = ps * gpc;
        per_sm = ps;
        plan->groups_per_cta = gpc;
      }
    }
    if (best_groups == 0) return cudaErrorInvalidValue;
    plan->threads = 32 * plan->groups_per_cta;
    plan->smem_bytes = per_group * plan->groups_per_cta;
    err = k2_occupancy(k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, SPEC>, plan->threads, plan->smem_bytes, &per_sm);
    if (err != cudaSuccess) return err;
  }
  // per resident group: the ring of per-checkpoint conditionals plus the interp_from slot (smoother only), plus the
  // working columns when they are not in shared memory
  plan->ring_bytes_per_group =
      FP ? ((size_t)T * GL::NFC + GL::NF + (plan->wk_global ? GL::NFW : 0)) * d * sizeof(double) : 0;
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
  size_t groups = (size_t)plan.grid * plan.groups_per_cta, per_group = plan.ring_bytes_per_group;
  if (k2_plan<VF, NU, FACT, TS0, FP>(cfg, B, T, /*needs_interp=*/false, &plan2) == cudaSuccess) {
    groups = std::max(groups, (size_t)plan2.grid * plan2.groups_per_cta);
    per_group = std::max(per_group, plan2.ring_bytes_per_group);
  }
  return 256 + groups * per_group;
}

template <class VF, int NU, int FACT, bool TS0, bool FP, bool HAS_SPEC = false>
cudaError_t k2_launch(const LoopArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const bool needs_interp = a.fixed_grid == 0 && a.cfg.clip_dt == 0;
  K2Plan plan;
  // the specialised build exists for the warp-per-instance mode only (d <= 32)
  const bool spec = HAS_SPEC && a.cfg.ode_dim <= 32 && k2_spec_matches(a) && k2_spec_choice() != 0;
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
  info.wk_scratch = nullptr;
  if (FP) {  // never launch more groups than the scratch has room for
    using GL = GroupLoop<VF, NU, FACT, TS0, FP, 0>;
    const size_t room = (workspace_bytes - 256) / plan.ring_bytes_per_group;
    const int max_grid = (int)(room / plan.groups_per_cta);
    if (max_grid < 1) return cudaErrorMemoryAllocation;
    plan.grid = std::min(plan.grid, max_grid);
    const size_t groups = (size_t)plan.grid * plan.groups_per_cta;
    info.if_scratch = info.cond_ring + groups * (size_t)a.T * GL::NFC * a.cfg.ode_dim;
    if (plan.wk_global) info.wk_scratch = info.if_scratch + groups * (size_t)GL::NF * a.cfg.ode_dim;
  }
  if (plan.mode == 2)
    k2_loop_kernel<VF, NU, FACT, TS0, FP, 2><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
  else if (plan.mode == 1)
    k2_loop_kernel<VF, NU, FACT, TS0, FP, 1><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
  else if (spec) {
    if constexpr (HAS_SPEC)
      k2_loop_kernel<VF, NU, FACT, TS0, FP, 0, 1><<<plan.grid, plan.threads, plan.smem_bytes, stream>>>(a, info);
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
