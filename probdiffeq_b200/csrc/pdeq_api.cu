// C ABI of probdiffeq_b200 (see include/probdiffeq_b200.h): argument validation, kernel lookup, launch.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

#include "pdeq_aux_kernels.cuh"
#include "pdeq_limits.cuh"

namespace pdeq {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e > 0 ? (int)e : 1;
}

static std::vector<LoopEntry>& loop_table() {
  static std::vector<LoopEntry> t;
  return t;
}
// A second registration under the same key REPLACES the first (a plug-in rebuilt with a changed right-hand side must
// not leave the old step-loop kernel behind the new Taylor / dt0 kernels), exactly as register_aux does.
void register_loop(const LoopEntry& e) {
  for (auto& x : loop_table())
    if (x.key == e.key) {
      x = e;
      return;
    }
  loop_table().push_back(e);
}
const LoopEntry* find_loop(const KernelKey& key) {
  for (const auto& e : loop_table())
    if (e.key == key) return &e;
  return nullptr;
}
int device_sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

// Vector-field table: the six built-in functors, then whatever plug-ins register at run time (pdeq_register_vf).
struct VfInfo {
  char name[64];
  int order, num_params, fixed_dim;
};
static std::vector<VfInfo>& vf_table() {
  static std::vector<VfInfo> table = {
      {"lotka_volterra", LotkaVolterra::order, LotkaVolterra::num_params, LotkaVolterra::fixed_dim},
      {"pleiades", Pleiades::order, Pleiades::num_params, Pleiades::fixed_dim},
      {"hires", Hires::order, Hires::num_params, Hires::fixed_dim},
      {"vanderpol", VanDerPol::order, VanDerPol::num_params, VanDerPol::fixed_dim},
      {"linear", Linear::order, Linear::num_params, Linear::fixed_dim},
      {"burgers", Burgers::order, Burgers::num_params, Burgers::fixed_dim},
  };
  return table;
}
static const VfInfo* vf_info(int id) {
  auto& t = vf_table();
  return (id >= 0 && id < (int)t.size()) ? &t[id] : nullptr;
}

static std::vector<AuxEntry>& aux_registry() {
  static std::vector<AuxEntry> r;
  return r;
}
void register_aux(const AuxEntry& e) {
  for (auto& x : aux_registry())
    if (x.vf_id == e.vf_id) {
      x = e;
      return;
    }
  aux_registry().push_back(e);
}
const AuxEntry* find_aux(int vf_id) {
  for (const auto& x : aux_registry())
    if (x.vf_id == vf_id) return &x;
  return nullptr;
}

static int validate(const pdeq_config* c) {
  if (c == nullptr) return fail(-1, "config is NULL");
  if (vf_info(c->vf_id) == nullptr) return fail(-2, "unknown vf_id %d", c->vf_id);
  const VfInfo& v = *vf_info(c->vf_id);
  if (c->num_derivatives < 1 || c->num_derivatives + 1 > PDEQ_MAX_COEFFS)
    return fail(-3, "num_derivatives=%d outside [1, %d]", c->num_derivatives, PDEQ_MAX_COEFFS - 1);
  if (c->num_derivatives + 1 <= v.order)
    return fail(-3, "need num_derivatives >= ode order (%d)", v.order);
  if (c->ode_dim < 1) return fail(-4, "ode_dim must be positive");
  if (v.fixed_dim != 0 && c->ode_dim != v.fixed_dim)
    return fail(-4, "vector field '%s' has dimension %d, got ode_dim=%d", v.name, v.fixed_dim, c->ode_dim);
  if (c->factorisation < 0 || c->factorisation > 2) return fail(-5, "bad factorisation %d", c->factorisation);
  if (c->constraint < 0 || c->constraint > 1) return fail(-5, "bad constraint %d", c->constraint);
  if (c->solver < 0 || c->solver > 2) return fail(-5, "bad solver %d", c->solver);
  if (c->strategy < 0 || c->strategy > 3) return fail(-5, "bad strategy %d", c->strategy);
  if (c->error < 0 || c->error > 1) return fail(-5, "bad error estimator %d", c->error);
  if (c->error_norm < 0 || c->error_norm > 1) return fail(-5, "bad error norm %d", c->error_norm);
  if (c->control < 0 || c->control > 1) return fail(-5, "bad control %d", c->control);
  if (c->derivative_idx < 0 || c->derivative_idx > c->num_derivatives)
    return fail(-5, "derivative_idx=%d outside [0, %d]", c->derivative_idx, c->num_derivatives);
  return 0;
}

// Map a configuration to a kernel. For d == 1 the three factorisations coincide, so a dense model of a
// scalar ODE runs on the isotropic kernel.
static const LoopEntry* select_loop(const pdeq_config* c) {
  int fact = c->factorisation;
  if (fact == PDEQ_FACT_DENSE && c->ode_dim == 1) fact = PDEQ_FACT_ISOTROPIC;
  const int ts0 = c->constraint == PDEQ_CONSTRAINT_TS0;
  const int fp = c->strategy != PDEQ_STRATEGY_FILTER;  // both smoothers run on the conditional-carrying kernels
  // K1: thread per instance, compile-time d, filter only. Without step clipping every lane parks its interp_from state
  // (mean, factor, t: ThreadLoop::IF_SLOTS doubles) in shared memory; a block-diagonal model of d >= 6 at high order
  // does not fit there, and the lane-per-dimension kernel takes it instead
  if (!fp) {
    const size_t n = (size_t)c->num_derivatives + 1, blocks = fact == PDEQ_FACT_BLOCKDIAG ? (size_t)c->ode_dim : 1;
    const size_t interp_smem = c->clip_dt ? 0 : (n * c->ode_dim + blocks * n * n + 1) * 128 * sizeof(double);
    if (interp_smem <= 227 * 1024) {
      const LoopEntry* e = find_loop({c->vf_id, c->num_derivatives, fact, c->ode_dim, ts0, 0});
      if (e != nullptr) return e;
    }
  }
  // K2: lane per dimension, run-time d <= 1024, filter and fixed-point smoother
  //     (several dimensions per lane only for the block-diagonal filter; otherwise d <= 256)
  const int k2_max_d = (fact == PDEQ_FACT_BLOCKDIAG && !fp) ? K2_CTA_THREADS * K2_MAX_DPL : K2_CTA_THREADS;
  if (fact != PDEQ_FACT_DENSE && c->ode_dim <= k2_max_d) {
    const LoopEntry* e = find_loop({c->vf_id, c->num_derivatives, fact, 0, ts0, fp});
    if (e != nullptr) return e;
  }
  // K3S: dense factorisation with a smoother, CTA per instance, all matrices in shared memory: 4 N^2 work buffer,
  // two N x N factors, two stored conditionals (pdeq_smooth_dense.cuh: DenseSmootherLoop::smem_doubles)
  if (fact == PDEQ_FACT_DENSE && fp) {
    const size_t N = (size_t)(c->num_derivatives + 1) * c->ode_dim;
    if ((10 * N * N + 12 * N) * sizeof(double) + 4096 <= 227 * 1024) {
      const LoopEntry* e = find_loop({c->vf_id, c->num_derivatives, fact, 0, ts0, 1});
      if (e != nullptr) return e;
    }
  }
  // K3: dense factorisation, CTA per instance, filter only; the work buffer must fit in shared memory
  if (fact == PDEQ_FACT_DENSE && !fp) {
    const DenseSmemLayout lay = DenseSmemLayout::make(c->num_derivatives + 1, c->ode_dim, vf_info(c->vf_id)->order, true);
    if (lay.total * sizeof(double) <= 227 * 1024) {
      const LoopEntry* e = find_loop({c->vf_id, c->num_derivatives, fact, 0, ts0, 0});
      if (e != nullptr) return e;
    }
  }
  return nullptr;
}

int api_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int api_cuda_fail(cudaError_t e, const char* where) { return cuda_fail(e, where); }
int api_validate(const pdeq_config* c) { return validate(c); }

}  // namespace pdeq

using namespace pdeq;

extern "C" {

int pdeq_version(void) { return PDEQ_VERSION; }

int pdeq_k1_spec_choice(void) { return k1_spec_choice(); }
const char* pdeq_last_error(void) { return g_err; }

int pdeq_vf_id(const char* name) {
  if (name == nullptr) return -1;
  const auto& t = vf_table();
  for (int i = 0; i < (int)t.size(); ++i)
    if (std::strcmp(name, t[i].name) == 0) return i;
  return -1;
}
int pdeq_vf_num_params(int vf_id) { return vf_info(vf_id) ? vf_info(vf_id)->num_params : -1; }
int pdeq_vf_ode_order(int vf_id) { return vf_info(vf_id) ? vf_info(vf_id)->order : -1; }
int pdeq_vf_dim(int vf_id) { return vf_info(vf_id) ? vf_info(vf_id)->fixed_dim : -1; }

int pdeq_register_vf(const char* name, int32_t ode_order, int32_t num_params, int32_t fixed_dim) {
  if (name == nullptr || name[0] == 0 || std::strlen(name) >= sizeof(VfInfo{}.name))
    return fail(-2, "vector-field name must have 1..63 characters");
  if (ode_order < 1 || ode_order >= PDEQ_MAX_COEFFS) return fail(-3, "bad ode_order %d", ode_order);
  if (num_params < 0 || fixed_dim < 0) return fail(-4, "num_params / fixed_dim must be non-negative");
  if (pdeq_vf_id(name) >= 0) return fail(-2, "vector field '%s' is already registered", name);
  VfInfo v{};
  std::strncpy(v.name, name, sizeof(v.name) - 1);
  v.order = ode_order;
  v.num_params = num_params;
  v.fixed_dim = fixed_dim;
  vf_table().push_back(v);
  return (int)vf_table().size() - 1;
}

int pdeq_vf_clear_kernels(int vf_id) {
  if (vf_info(vf_id) == nullptr) return fail(-2, "unknown vf_id %d", vf_id);
  if (vf_id < 6) return fail(-2, "the kernels of the built-in vector field '%s' cannot be dropped", vf_info(vf_id)->name);
  auto& loops = loop_table();
  int dropped = 0;
  for (size_t i = 0; i < loops.size();) {
    if (loops[i].key.vf == vf_id) {
      loops.erase(loops.begin() + (long)i);
      ++dropped;
    } else {
      ++i;
    }
  }
  auto& aux = aux_registry();
  for (size_t i = 0; i < aux.size();) {
    if (aux[i].vf_id == vf_id) aux.erase(aux.begin() + (long)i);
    else ++i;
  }
  return dropped;
}

int pdeq_config_supported(const pdeq_config* cfg) {
  int rc = validate(cfg);
  if (rc != 0) return rc;
  if (select_loop(cfg) == nullptr)
    return fail(-10,
                "no kernel for vf=%s nu=%d d=%d factorisation=%d constraint=%d strategy=%d",
                vf_info(cfg->vf_id)->name, cfg->num_derivatives, cfg->ode_dim, cfg->factorisation, cfg->constraint,
                cfg->strategy);
  return 0;
}

size_t pdeq_workspace_bytes(const pdeq_config* cfg, int64_t num_instances, int32_t num_checkpoints) {
  if (validate(cfg) != 0) return 256;
  const LoopEntry* e = select_loop(cfg);
  if (e == nullptr) return 256;
  return e->workspace_bytes(*cfg, num_instances, num_checkpoints);
}

static int check_common(const pdeq_config* cfg, const pdeq_problem* pr, const pdeq_solution* so, int32_t T,
                        void* ws, size_t ws_bytes) {
  int rc = pdeq_config_supported(cfg);
  if (rc != 0) return rc;
  if (pr == nullptr || so == nullptr) return fail(-20, "problem/solution is NULL");
  if (pr->num_instances < 0) return fail(-20, "negative num_instances");
  if (T < 1) return fail(-21, "need at least one checkpoint");
  if (pr->tcoeffs == nullptr) return fail(-22, "tcoeffs is NULL");
  if (vf_info(cfg->vf_id)->num_params > 0 && pr->params == nullptr) return fail(-22, "params is NULL");
  if (so->t == nullptr || so->mean == nullptr || so->num_steps == nullptr || so->status == nullptr)
    return fail(-23, "solution.t/mean/num_steps/status must be non-NULL");
  if (ws == nullptr || ws_bytes < pdeq_workspace_bytes(cfg, pr->num_instances, T))
    return fail(-24, "workspace too small");
  return 0;
}

static int run_loop(const pdeq_config* cfg, const pdeq_problem* pr, const pdeq_solution* so, const double* grid,
                    int32_t T, int fixed, double atol, double rtol, const double* dt0, int64_t dt0_stride,
                    double eps, double damp, void* ws, size_t ws_bytes, void* stream) {
  if (pr->num_instances == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  LoopArgs a;
  a.cfg = *cfg;
  a.prob = *pr;
  a.sol = *so;
  a.grid = grid;
  a.T = T;
  a.fixed_grid = fixed;
  a.atol = atol;
  a.rtol = rtol;
  a.eps = eps;
  a.damp = damp;
  a.dt0 = dt0;
  a.dt0_stride = dt0_stride;
  a.work_counter = (unsigned long long*)ws;
  cudaError_t e = cudaMemsetAsync(ws, 0, 256, s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
  const LoopEntry* entry = select_loop(cfg);
  e = entry->launch(a, ws, ws_bytes, s);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return 0;
}

int pdeq_solve_adaptive_save_at(const pdeq_config* cfg, const pdeq_problem* problem, const double* save_at,
                                int32_t num_checkpoints, double atol, double rtol, const double* dt0,
                                int64_t dt0_stride, double eps, double damp, const pdeq_solution* solution,
                                void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(cfg, problem, solution, num_checkpoints, workspace, workspace_bytes);
  if (rc != 0) return rc;
  if (save_at == nullptr || dt0 == nullptr) return fail(-25, "save_at/dt0 is NULL");
  if (cfg->strategy == PDEQ_STRATEGY_FIXEDINTERVAL || cfg->strategy == PDEQ_STRATEGY_FIXEDINTERVAL_ALIGNED)
    return fail(-10, "the fixed-interval smoother is for solve_fixed_grid; use the fixed-point smoother with save_at "
                     "(the reference warns likewise, solvers_via_adaptive_steps.py:81-85)");
  return run_loop(cfg, problem, solution, save_at, num_checkpoints, 0, atol, rtol, dt0, dt0_stride, eps, damp,
                  workspace, workspace_bytes, stream);
}

int pdeq_solve_fixed_grid(const pdeq_config* cfg, const pdeq_problem* problem, const double* grid,
                          int32_t num_gridpoints, double damp, const pdeq_solution* solution, void* workspace,
                          size_t workspace_bytes, void* stream) {
  int rc = check_common(cfg, problem, solution, num_gridpoints, workspace, workspace_bytes);
  if (rc != 0) return rc;
  if (grid == nullptr) return fail(-25, "grid is NULL");
  if (cfg->strategy == PDEQ_STRATEGY_FIXEDPOINT)
    return fail(-10, "solve_fixed_grid with a fixed-point smoother is not accelerated (the reference warns against it)");
  return run_loop(cfg, problem, solution, grid, num_gridpoints, 1, 0.0, 0.0, nullptr, 0, 0.0, damp, workspace,
                  workspace_bytes, stream);
}

// ---------------------------------------------------------------------------------------------------
// FP64 FMA throughput probe (roofline denominator for the register-resident kernels).
// ---------------------------------------------------------------------------------------------------
__global__ void fp64_probe_kernel(int iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
         a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double x = 1.0000001, y = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y);
    a1 = fma(a1, x, y);
    a2 = fma(a2, x, y);
    a3 = fma(a3, x, y);
    a4 = fma(a4, x, y);
    a5 = fma(a5, x, y);
    a6 = fma(a6, x, y);
    a7 = fma(a7, x, y);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) sink[0] = r;
}

int pdeq_fp64_peak_probe(int32_t iters, double* ms, double* flops, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, 8);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  const int threads = 256, blocks = device_sm_count() * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_probe_kernel<<<blocks, threads, 0, s>>>(iters / 8 + 1, sink);  // warm-up
  cudaEventRecord(e0, s);
  fp64_probe_kernel<<<blocks, threads, 0, s>>>(iters, sink);
  cudaEventRecord(e1, s);
  e = cudaEventSynchronize(e1);
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return cuda_fail(e, "fp64 probe");
  if (ms) *ms = t;
  if (flops) *flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// NCCL all-reduce of the ensemble log-marginal-likelihood. NCCL is resolved from the process (the copy
// that created the communicator) so that no second NCCL is loaded.
// ---------------------------------------------------------------------------------------------------
// NCCL entry points resolved once from the copy of libnccl.so.2 already in the process (torch's), else loaded.
struct NcclApi {
  struct UniqueId {
    char internal[128];  // NCCL_UNIQUE_ID_BYTES
  };
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*get_unique_id)(UniqueId*) = nullptr;
  int (*comm_init_rank)(void**, int, UniqueId, int) = nullptr;
  int (*comm_destroy)(void*) = nullptr;
  int (*comm_count)(void*, int*) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  bool ok = false;
};
static const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (h == nullptr) return;
    api.all_reduce = (decltype(api.all_reduce))dlsym(h, "ncclAllReduce");
    api.get_unique_id = (decltype(api.get_unique_id))dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (decltype(api.comm_init_rank))dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (decltype(api.comm_destroy))dlsym(h, "ncclCommDestroy");
    api.comm_count = (decltype(api.comm_count))dlsym(h, "ncclCommCount");
    api.get_error_string = (decltype(api.get_error_string))dlsym(h, "ncclGetErrorString");
    api.ok = api.all_reduce && api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.comm_count;
  });
  return api.ok ? &api : nullptr;
}
static int nccl_fail(const NcclApi* api, int rc, const char* where) {
  return fail(100 + rc, "%s failed: %s (ncclResult %d)", where,
              api->get_error_string ? api->get_error_string(rc) : "?", rc);
}

int pdeq_nccl_unique_id(void* id128) {
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(-30, "libnccl.so.2 not found: %s", dlerror());
  if (id128 == nullptr) return fail(-31, "id buffer is NULL");
  NcclApi::UniqueId id;
  int rc = api->get_unique_id(&id);
  if (rc != 0) return nccl_fail(api, rc, "ncclGetUniqueId");
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return 0;
}

int pdeq_nccl_comm_init_rank(void** comm, int32_t nranks, const void* id128, int32_t rank) {
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(-30, "libnccl.so.2 not found: %s", dlerror());
  if (comm == nullptr || id128 == nullptr) return fail(-31, "comm/id is NULL");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(-31, "bad rank %d of %d", rank, nranks);
  NcclApi::UniqueId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  int rc = api->comm_init_rank(comm, nranks, id, rank);
  if (rc != 0) return nccl_fail(api, rc, "ncclCommInitRank");
  return 0;
}

int pdeq_nccl_comm_count(void* nccl_comm) {
  const NcclApi* api = nccl_api();
  if (api == nullptr || nccl_comm == nullptr) return -1;
  int n = -1;
  return api->comm_count(nccl_comm, &n) == 0 ? n : -1;
}

int pdeq_nccl_comm_destroy(void* nccl_comm) {
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(-30, "libnccl.so.2 not found");
  if (nccl_comm == nullptr) return 0;
  int rc = api->comm_destroy(nccl_comm);
  if (rc != 0) return nccl_fail(api, rc, "ncclCommDestroy");
  return 0;
}

int pdeq_allreduce_sum_f64(void* nccl_comm, double* buf, int64_t n, void* stream) {
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(-30, "libnccl.so.2 not found: %s", dlerror());
  if (nccl_comm == nullptr || buf == nullptr) return fail(-31, "comm/buf is NULL");
  if (n < 0) return fail(-31, "negative count");
  // ncclFloat64 = 8, ncclSum = 0
  int rc = api->all_reduce(buf, buf, (size_t)n, 8, 0, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(api, rc, "ncclAllReduce");
  return 0;
}

}  // extern "C"
