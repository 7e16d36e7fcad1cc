// Once-per-solve device routines that depend on the vector field: Taylor-coefficient initialisation and the
// initial step size. They live in a header so that a vector field compiled later (a plug-in, see
// probdiffeq_b200/plugins.py) instantiates and registers them exactly like the built-in ones.
#pragma once

#include "pdeq_dispatch.cuh"

namespace pdeq {

// ---------------------------------------------------------------------------------------------------
// Taylor-mode initialisation (probdiffeq/_probdiffeq/jet_expansion_algorithms.py:49-177).
// One launch per new coefficient ("pass"); thread (b, i) evaluates component i of the vector field on the
// truncated series built from the coefficients known so far and writes u^(pass+q)_i.
// `out` [B][n][d] holds unnormalised derivatives and doubles as the workspace.
// ---------------------------------------------------------------------------------------------------
template <int KS>
struct GlobalSeriesAcc {
  const double* __restrict__ U;  // [n][d] of one instance, unnormalised derivatives
  int d, known;                  // coefficients 0..known-1 are valid
  PDEQ_DI Series<KS> operator()(int j, int i) const {
    Series<KS> s;
    double kfact = 1.0;  // k!
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      if (k > 0) kfact *= double(k);
      s.c[k] = (k + j < known) ? U[(k + j) * d + i] / kfact : 0.0;  // (u^(j))_k = u^(k+j) / k!
    }
    return s;
  }
};

template <class VF, int KS>
__global__ void taylor_pass_kernel(int64_t B, int n, int d, int pass, const double* __restrict__ params,
                                   int64_t params_stride, double t0, double* __restrict__ out) {
  constexpr int q = VF::order, P = VF::num_params > 0 ? VF::num_params : 1;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * d) return;
  const int64_t b = gid / d;
  const int i = (int)(gid % d);
  double par[P];
#pragma unroll
  for (int k = 0; k < P; ++k) par[k] = VF::num_params > 0 ? params[b * params_stride + k] : 0.0;
  GlobalSeriesAcc<KS> acc{out + b * (int64_t)n * d, d, pass + q};
  const Series<KS> F = VF::template component<Series<KS>>(i, d, acc, par, t0);
  double fk = 0.0, pf = 1.0;  // F.c[pass] * pass!  == u^(pass+q)
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    if (k > 0) pf *= double(k);
    if (k == pass) fk = F.c[k] * pf;
  }
  out[(b * n + pass + q) * d + i] = fk;
}

template <class VF, int KS>
inline cudaError_t taylor_run(int64_t B, int n, int d, const double* u0, const double* params,
                              int64_t params_stride, double t0, double* out, cudaStream_t s) {
  constexpr int q = VF::order;
  // copy the initial values into the first q coefficient slots
  cudaError_t e = cudaMemcpy2DAsync(out, sizeof(double) * n * d, u0, sizeof(double) * q * d,
                                    sizeof(double) * q * d, B, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  const int threads = 128;
  const int64_t total = B * d;
  const int grid = (int)((total + threads - 1) / threads);
  for (int pass = 0; pass < n - q; ++pass) {
    taylor_pass_kernel<VF, KS><<<grid, threads, 0, s>>>(B, n, d, pass, params, params_stride, t0, out);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <class VF>
inline cudaError_t taylor_dispatch_ks(int ks, int64_t B, int n, int d, const double* u0, const double* params,
                                      int64_t ps, double t0, double* out, cudaStream_t s) {
  switch (ks) {
    case 1: return taylor_run<VF, 1>(B, n, d, u0, params, ps, t0, out, s);
    case 2: return taylor_run<VF, 2>(B, n, d, u0, params, ps, t0, out, s);
    case 3: return taylor_run<VF, 3>(B, n, d, u0, params, ps, t0, out, s);
    case 4: return taylor_run<VF, 4>(B, n, d, u0, params, ps, t0, out, s);
    case 5: return taylor_run<VF, 5>(B, n, d, u0, params, ps, t0, out, s);
    case 6: return taylor_run<VF, 6>(B, n, d, u0, params, ps, t0, out, s);
    case 7: return taylor_run<VF, 7>(B, n, d, u0, params, ps, t0, out, s);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------------------------------
// ivpsolve.dt0 (probdiffeq/_ivpsolve/stepsize_initialisers.py:7-21): scale * ||u0|| / (||f(u0)|| + nugget).
// One warp per instance; lanes stride over the components.
// ---------------------------------------------------------------------------------------------------
struct GlobalAcc {
  const double* __restrict__ u;  // [order][d]
  int d;
  PDEQ_DI double operator()(int k, int i) const { return u[k * d + i]; }
};

template <class VF>
__global__ void dt0_kernel(int64_t B, int d, const double* __restrict__ u0, const double* __restrict__ params,
                           int64_t params_stride, double t0, double scale, double nugget,
                           double* __restrict__ out) {
  constexpr int q = VF::order, P = VF::num_params > 0 ? VF::num_params : 1;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (b >= B) return;
  double par[P];
#pragma unroll
  for (int k = 0; k < P; ++k) par[k] = VF::num_params > 0 ? params[b * params_stride + k] : 0.0;
  GlobalAcc acc{u0 + b * (int64_t)q * d, d};
  double su = 0.0, sf = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double ui = acc(0, i);
    const double fi = VF::template component<double>(i, d, acc, par, t0);
    su = fma(ui, ui, su);
    sf = fma(fi, fi, sf);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sf += __shfl_xor_sync(0xffffffffu, sf, o);
  }
  if (lane == 0) out[b] = scale * sqrt(su) / (sqrt(sf) + nugget);
}

// ---------------------------------------------------------------------------------------------------
// ivpsolve.dt0_adaptive (probdiffeq/_ivpsolve/stepsize_initialisers.py:24-64; Hairer et al., Sec. II.4).
// First-order ODEs only, as in the reference. The Euler point y1 = y0 + h0 f(y0) is evaluated on the fly.
// ---------------------------------------------------------------------------------------------------
template <class VF>
struct EulerAcc {
  GlobalAcc base;
  const double* par;
  double t0, h0;
  PDEQ_DI double operator()(int k, int i) const {
    return base(k, i) + h0 * VF::template component<double>(i, base.d, base, par, t0);
  }
};

template <class VF>
__global__ void dt0_adaptive_kernel(int64_t B, int d, const double* __restrict__ u0,
                                    const double* __restrict__ params, int64_t params_stride, double t0,
                                    double rate, double rtol, double atol, double* __restrict__ out) {
  constexpr int P = VF::num_params > 0 ? VF::num_params : 1;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (b >= B) return;
  double par[P];
#pragma unroll
  for (int k = 0; k < P; ++k) par[k] = VF::num_params > 0 ? params[b * params_stride + k] : 0.0;
  GlobalAcc acc{u0 + b * (int64_t)d, d};
  double s0 = 0.0, s1 = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double yi = acc(0, i);
    const double fi = VF::template component<double>(i, d, acc, par, t0);
    s0 = fma(yi, yi, s0);
    s1 = fma(fi, fi, s1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const double d0 = sqrt(s0), d1 = sqrt(s1);
  const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  EulerAcc<VF> acc1{acc, par, t0, h0};
  double s2 = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double f0 = VF::template component<double>(i, d, acc, par, t0);
    const double f1 = VF::template component<double>(i, d, acc1, par, t0 + h0);
    const double w = (f1 - f0) / (atol + fabs(acc(0, i)) * rtol);
    s2 = fma(w, w, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const double d2 = sqrt(s2) / h0;
  const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / (rate + 1.0));
  if (lane == 0) out[b] = fmin(100.0 * h0, h1);
}

// ---------------------------------------------------------------------------------------------------
// Registry of the routines above per vector-field id (filled at load time, like the loop registry).
// ---------------------------------------------------------------------------------------------------
using TaylorFn = cudaError_t (*)(int ks, int64_t B, int n, int d, const double* u0, const double* params,
                                 int64_t params_stride, double t0, double* out, cudaStream_t s);
using Dt0Fn = cudaError_t (*)(int64_t B, int d, const double* u0, const double* params, int64_t params_stride,
                              double t0, double scale, double nugget, double* out, cudaStream_t s);
using Dt0AdaptiveFn = cudaError_t (*)(int64_t B, int d, const double* u0, const double* params,
                                      int64_t params_stride, double t0, double rate, double rtol, double atol,
                                      double* out, cudaStream_t s);
struct AuxEntry {
  int vf_id;
  TaylorFn taylor;
  Dt0Fn dt0;
  Dt0AdaptiveFn dt0_adaptive;
};
void register_aux(const AuxEntry& e);
const AuxEntry* find_aux(int vf_id);

template <class VF>
cudaError_t dt0_launch(int64_t B, int d, const double* u0, const double* params, int64_t params_stride, double t0,
                       double scale, double nugget, double* out, cudaStream_t s) {
  const int threads = 128;
  const int grid = (int)((B * 32 + threads - 1) / threads);
  dt0_kernel<VF><<<grid, threads, 0, s>>>(B, d, u0, params, params_stride, t0, scale, nugget, out);
  return cudaGetLastError();
}
template <class VF>
cudaError_t dt0_adaptive_launch(int64_t B, int d, const double* u0, const double* params, int64_t params_stride,
                                double t0, double rate, double rtol, double atol, double* out, cudaStream_t s) {
  const int threads = 128;
  const int grid = (int)((B * 32 + threads - 1) / threads);
  dt0_adaptive_kernel<VF><<<grid, threads, 0, s>>>(B, d, u0, params, params_stride, t0, rate, rtol, atol, out);
  return cudaGetLastError();
}

template <class VF>
struct AuxRegistrar {
  explicit AuxRegistrar(int vf_id = VF::id) {
    register_aux({vf_id, &taylor_dispatch_ks<VF>, &dt0_launch<VF>, &dt0_adaptive_launch<VF>});
  }
};

}  // namespace pdeq
