// Once-per-solve device routines around the step loop: Taylor-coefficient initialisation, initial
// step size, terminal-value log-marginal-likelihood. All are generic in the ODE dimension d.
#include <cuda_runtime.h>

#include "pdeq_aux_dense.cuh"
#include "pdeq_aux_kernels.cuh"

namespace pdeq {

int api_fail(int code, const char* fmt, ...);       // pdeq_api.cu
int api_cuda_fail(cudaError_t e, const char* where);  // pdeq_api.cu
int api_validate(const pdeq_config* c);              // pdeq_api.cu

// ---------------------------------------------------------------------------------------------------
// loss_lml_terminal_values for the isotropic and block-diagonal factorisations
// (probdiffeq/_probdiffeq/estimators_and_losses.py:20-50 with IsotropicNormal.logpdf_scalar_flat,
// ssm_impl_isotropic.py:225-236, and BlockDiagNormal.logpdf_scalar_flat, ssm_impl_blockdiag.py:305-316).
// Observing one coefficient through scalar noise gives a 1x1 factor per dimension:
//   s_i = sqrt(sum_j L[idx][j]^2 + std_i^2),  logpdf = sum_i -0.5 (2 log s_i + ((u_i - m_i)/s_i)^2 + log 2pi).
// One warp per instance; lanes stride over dimensions.
// ---------------------------------------------------------------------------------------------------
__global__ void lml_terminal_kernel(int64_t B, int n, int d, int blockdiag, int idx, const double* __restrict__ mean,
                                    const double* __restrict__ chol, const double* __restrict__ data,
                                    int64_t data_stride, const double* __restrict__ std_, int64_t std_stride,
                                    int std_per_dim, double* __restrict__ out) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (b >= B) return;
  const double* m = mean + b * (int64_t)n * d + (int64_t)idx * d;
  const double log2pi = 1.8378770664093453;
  double acc = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double* Lrow = blockdiag ? chol + ((b * d + i) * (int64_t)n + idx) * n : chol + (b * (int64_t)n + idx) * n;
    double ss = 0.0;
    for (int j = 0; j <= idx; ++j) ss = fma(Lrow[j], Lrow[j], ss);
    const double sd = std_[b * std_stride + (std_per_dim ? i : 0)];
    const double s = sqrt(fma(sd, sd, ss));
    const double w = (data[b * data_stride + i] - m[i]) / s;
    acc += -0.5 * (2.0 * log(s) + w * w + log2pi);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[b] = acc;
}

// loss_lml_terminal_values for the dense factorisation (DenseNormal.to_derivative + marginalise + logpdf,
// ssm_impl_dense.py:108-234): the observed covariance S = L_i L_i^T + diag(std^2), with L_i the d rows of the
// (nd x nd) factor that belong to coefficient idx (coefficient-major state, row idx * d + a), is formed in shared
// memory by one warp per instance, factorised (Cholesky, d <= 32) and the Gaussian log-density evaluated.
// The reference triangularises a stacked square root instead; the value is the same up to rounding.
constexpr int LML_DENSE_MAX_D = 32;
__global__ void lml_terminal_dense_kernel(int64_t B, int n, int d, int idx, const double* __restrict__ mean,
                                          const double* __restrict__ chol, const double* __restrict__ data,
                                          int64_t data_stride, const double* __restrict__ std_, int64_t std_stride,
                                          double* __restrict__ out) {
  __shared__ double S_all[4][LML_DENSE_MAX_D][LML_DENSE_MAX_D + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 4 + w;
  if (b >= B) return;  // whole warps leave together; no block barrier below
  double(*S)[LML_DENSE_MAX_D + 1] = S_all[w];
  const int N = n * d;
  const double* L = chol + b * (int64_t)N * N;
  for (int e = lane; e < d * d; e += 32) {
    const int a = e / d, c = e % d;
    if (c > a) continue;
    const double* ra = L + (int64_t)(idx * d + a) * N;
    const double* rc = L + (int64_t)(idx * d + c) * N;
    double acc = 0.0;
    for (int k = 0; k <= idx * d + c; ++k) acc = fma(ra[k], rc[k], acc);  // lower-triangular factor
    if (a == c) {
      const double sd = std_[b * std_stride + a];
      acc = fma(sd, sd, acc);
    }
    S[a][c] = acc;
  }
  __syncwarp();
  if (lane == 0) {
    const double log2pi = 1.8378770664093453;
    double logdet = 0.0, quad = 0.0;
    double y[LML_DENSE_MAX_D];
    for (int j = 0; j < d; ++j) {  // in-place Cholesky, then forward substitution with the residual
      double s = S[j][j];
      for (int k = 0; k < j; ++k) s = fma(-S[j][k], S[j][k], s);
      const double piv = sqrt(s);
      S[j][j] = piv;
      for (int i = j + 1; i < d; ++i) {
        double t = S[i][j];
        for (int k = 0; k < j; ++k) t = fma(-S[i][k], S[j][k], t);
        S[i][j] = t / piv;
      }
      double r = data[b * data_stride + j] - mean[(b * n + idx) * (int64_t)d + j];
      for (int k = 0; k < j; ++k) r = fma(-S[j][k], y[k], r);
      y[j] = r / piv;
      quad = fma(y[j], y[j], quad);
      logdet += log(piv);
    }
    out[b] = -0.5 * (2.0 * logdet + quad + (double)d * log2pi);
  }
}

// ---------------------------------------------------------------------------------------------------
// loss_lml_timeseries for the isotropic and block-diagonal factorisations
// (probdiffeq/_probdiffeq/estimators_and_losses.py:53-105; MarkovSequence.evaluate_lml :180-218;
// AbstractLatentCond.bayes_rule_and_logpdf_tree, ssm_impl_api.py:124-132): a backward scan over the grid that
// observes the terminal marginal, then alternately steps back through a stored backward conditional
// (cond_marginalise) and observes again (revert_obs with the measurement noise in the place of damp).
// Thread (b, j) runs the scan of dimension j of instance b and leaves its share of the log-density in
// `partial` [B][d]; lml_reduce_kernel adds the d shares of an instance in a fixed order.
// The conditionals are in natural coordinates (preconditioner applied), as the step loop emits them.
// A zero observed factor gives a zero gain -- the reference's least-squares solve (backend/linalg.py:60-61).
// ---------------------------------------------------------------------------------------------------
template <int n>
__global__ void lml_timeseries_kernel(int64_t B, int T, int d, int blockdiag, int idx, int average,
                                      const double* __restrict__ mean, const double* __restrict__ chol,
                                      const double* __restrict__ bw_gain, const double* __restrict__ bw_mean,
                                      const double* __restrict__ bw_chol, const double* __restrict__ data,
                                      int64_t data_stride, const double* __restrict__ std_, int64_t std_stride,
                                      double* __restrict__ partial) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * d) return;
  const int64_t b = gid / d;
  const int j = (int)(gid % d);
  const double log2pi = 1.8378770664093453;
  auto mat = [&](const double* base, int k) {  // the n x n block of grid point k
    return blockdiag ? base + ((b * T + k) * (int64_t)d + j) * (n * n) : base + (b * T + k) * (int64_t)(n * n);
  };
  double m[n], L[n][n], h[n];
#pragma unroll
  for (int i = 0; i < n; ++i) {
    m[i] = mean[((b * T + (T - 1)) * n + i) * (int64_t)d + j];
    h[i] = (i == idx) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < n; ++c) L[i][c] = (c <= i) ? mat(chol, T - 1)[i * n + c] : 0.0;
  }
  double acc = 0.0;
  int num = 0;
  for (int k = T - 1; k >= 0; --k) {
    if (k < T - 1) {  // step back through the conditional that maps grid point k + 1 to k
      BlockCond<n> c;
      const double* G = mat(bw_gain, k + 1);
      const double* X = mat(bw_chol, k + 1);
#pragma unroll
      for (int i = 0; i < n; ++i) {
        c.xi[i] = bw_mean[((b * T + (k + 1)) * n + i) * (int64_t)d + j];
        c.tl[i] = 1.0;
        c.to[i] = 1.0;
#pragma unroll
        for (int cc = 0; cc < n; ++cc) {
          c.G[i][cc] = G[i * n + cc];
          c.Xi[i][cc] = (cc <= i) ? X[i * n + cc] : 0.0;
        }
      }
      double mo[n], Lo[n][n];
      cond_marginalise<n>(c, m, L, mo, Lo);
#pragma unroll
      for (int i = 0; i < n; ++i) {
        m[i] = mo[i];
#pragma unroll
        for (int cc = 0; cc < n; ++cc) L[i][cc] = (cc <= i) ? Lo[i][cc] : 0.0;
      }
    }
    const double sd = std_[b * std_stride + (blockdiag ? (int64_t)k * d + j : k)];
    const double y = data[b * data_stride + (int64_t)k * d + j];
    double ry, gain[n], Ln[n][n];
    revert_obs<n, n - 1, false>(L, h, sd, ry, gain, Ln);
    double mi = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) mi = (i == idx) ? m[i] : mi;
    const double res = y - mi;
    const double w = res / ry;
    const double pdf = -0.5 * (2.0 * log(fabs(ry)) + w * w + log2pi);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      m[i] = fma((ry != 0.0) ? gain[i] : 0.0, res, m[i]);
#pragma unroll
      for (int cc = 0; cc < n; ++cc) L[i][cc] = (cc <= i) ? Ln[i][cc] : 0.0;
    }
    acc = average ? (acc * (double)num + pdf) / (double)(num + 1) : acc + pdf;
    num += 1;
  }
  partial[gid] = acc;
}

// ---------------------------------------------------------------------------------------------------
// Dense output: ProbabilisticSolver.offgrid_marginals (probdiffeq/_probdiffeq/solvers.py:149-203).
// Thread (b, query, dimension). A filter extrapolates the marginal of the grid point to the left
// (strategy_filter.interpolate_offgrid_marginals, estimators_and_losses.py:403-414); the fixed-interval smoother
// extrapolates the FILTERING marginal t0 -> t -> t1 and pulls the smoothing marginal at t1 back through the
// new t1 -> t conditional (:677-709). Both transitions use the output scale of the right grid point (:189-194).
// ---------------------------------------------------------------------------------------------------
template <int n, bool SMOOTH>
__global__ void offgrid_kernel(const __grid_constant__ pdeq_config cfg, int64_t B, int T, int Q, int blockdiag,
                               const double* __restrict__ grid, const double* __restrict__ queries,
                               const double* __restrict__ mean, const double* __restrict__ chol,
                               const double* __restrict__ filt_mean, const double* __restrict__ filt_chol,
                               const double* __restrict__ output_scale, const double* __restrict__ prior_scale,
                               int64_t prior_scale_stride, double* __restrict__ out_mean,
                               double* __restrict__ out_chol) {
  const int d = cfg.ode_dim;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * Q * d) return;
  const int j = (int)(gid % d);
  const int qi = (int)((gid / d) % Q);
  const int64_t b = gid / ((int64_t)d * Q);
  const double t = queries[qi];
  int lo = 0, hi = T;  // searchsorted(grid, t): first index with grid[index] >= t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (grid[mid] < t) lo = mid + 1;
    else hi = mid;
  }
  const int i1 = min(max(lo, 1), T - 1), i0 = i1 - 1;
  auto load = [&](const double* mu, const double* ch, int k, double (&m)[n], double (&L)[n][n]) {
    const double* c = blockdiag ? ch + ((b * T + k) * (int64_t)d + j) * (n * n) : ch + (b * T + k) * (int64_t)(n * n);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      m[i] = mu[((b * T + k) * n + i) * (int64_t)d + j];
#pragma unroll
      for (int cc = 0; cc < n; ++cc) L[i][cc] = (cc <= i) ? c[i * n + cc] : 0.0;
    }
  };
  const double lam = prior_scale == nullptr ? 1.0 : prior_scale[b * prior_scale_stride + (blockdiag ? j : 0)];
  const double sig = output_scale[blockdiag ? (b * T + i1) * (int64_t)d + j : b * T + i1] * lam;
  const double dt0 = t - grid[i0], dt1 = grid[i1] - t;
  double m0[n], L0[n][n], p[n], pinv[n], mt[n], Lt[n][n];
  load(SMOOTH ? filt_mean : mean, SMOOTH ? filt_chol : chol, i0, m0, L0);
  preconditioner<n>(dt0, cfg.inv_factorials, cfg.factorials, p, pinv);
  predict_mean<n>(m0, p, pinv, cfg.sys_a, mt);
  predict_chol<n>(L0, p, pinv, safe_sqrt(fabs(dt0)) * sig, cfg.sys_a, cfg.sys_q, Lt);
  if (SMOOTH) {
    double m1[n], L1[n][n], Lu[n][n], mo[n], Lo[n][n];
    BlockCond<n> bw;
    preconditioner<n>(dt1, cfg.inv_factorials, cfg.factorials, p, pinv);
    revert_transition<n>(Lt, mt, p, pinv, safe_sqrt(fabs(dt1)) * sig, cfg.sys_a, cfg.sys_q, Lu, bw);
    load(mean, chol, i1, m1, L1);
    cond_marginalise<n>(bw, m1, L1, mo, Lo);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      mt[i] = mo[i];
#pragma unroll
      for (int cc = 0; cc < n; ++cc) Lt[i][cc] = (cc <= i) ? Lo[i][cc] : 0.0;
    }
  }
  const int64_t bq = b * Q + qi;
#pragma unroll
  for (int i = 0; i < n; ++i) out_mean[(bq * n + i) * (int64_t)d + j] = mt[i];
  if (out_chol != nullptr && (blockdiag || j == 0)) {
    double* co = blockdiag ? out_chol + (bq * d + j) * (int64_t)(n * n) : out_chol + bq * (int64_t)(n * n);
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int cc = 0; cc < n; ++cc) co[i * n + cc] = (cc <= i) ? Lt[i][cc] : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// MarkovSequence.sample (probdiffeq/_probdiffeq/estimators_and_losses.py:233-271) with the standard-normal draws
// supplied by the caller: x_{T-1} = m + L eps_{T-1}, then x_{k-1} = G_k x_k + xi_k + Xi_k eps_{k-1} backwards over
// the stored conditionals (natural coordinates). Thread (b, sample, dimension). base is [B][S][T][n] for the
// isotropic model -- one draw per coefficient shared by all dimensions, as IsotropicNormal.sample_flat does
// (ssm_impl_isotropic.py:255-258) -- and [B][S][T][d][n] for the block-diagonal one (ssm_impl_blockdiag.py:343-351).
// ---------------------------------------------------------------------------------------------------
template <int n>
__global__ void sample_kernel(int64_t B, int S, int T, int d, int blockdiag, const double* __restrict__ mean,
                              const double* __restrict__ chol, const double* __restrict__ bw_gain,
                              const double* __restrict__ bw_mean, const double* __restrict__ bw_chol,
                              const double* __restrict__ base, double* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * S * d) return;
  const int j = (int)(gid % d);
  const int si = (int)((gid / d) % S);
  const int64_t b = gid / ((int64_t)d * S);
  const int64_t bs = b * S + si;
  auto mat = [&](const double* p, int k) {
    return blockdiag ? p + ((b * T + k) * (int64_t)d + j) * (n * n) : p + (b * T + k) * (int64_t)(n * n);
  };
  auto eps = [&](int k) { return blockdiag ? base + ((bs * T + k) * (int64_t)d + j) * n : base + (bs * T + k) * (int64_t)n; };
  double x[n];
  {
    const double* L = mat(chol, T - 1);
    const double* e = eps(T - 1);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < n; ++c) acc = (c <= i) ? fma(L[i * n + c], e[c], acc) : acc;
      x[i] = mean[((b * T + (T - 1)) * n + i) * (int64_t)d + j] + acc;
      out[((bs * T + (T - 1)) * n + i) * (int64_t)d + j] = x[i];
    }
  }
  for (int k = T - 1; k >= 1; --k) {
    const double* G = mat(bw_gain, k);
    const double* X = mat(bw_chol, k);
    const double* e = eps(k - 1);
    double y[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double acc = 0.0, nz = 0.0;
#pragma unroll
      for (int c = 0; c < n; ++c) {
        acc = fma(G[i * n + c], x[c], acc);
        nz = (c <= i) ? fma(X[i * n + c], e[c], nz) : nz;
      }
      y[i] = (acc + bw_mean[((b * T + k) * n + i) * (int64_t)d + j]) + nz;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) {
      x[i] = y[i];
      out[((bs * T + (k - 1)) * n + i) * (int64_t)d + j] = x[i];
    }
  }
}

__global__ void lml_reduce_kernel(int64_t B, int d, const double* __restrict__ partial, double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double acc = 0.0;
  for (int j = 0; j < d; ++j) acc += partial[b * d + j];
  out[b] = acc;
}

}  // namespace pdeq

using namespace pdeq;

// the built-in vector fields register their Taylor / step-size routines here; plug-ins do the same from their own
// translation unit (pdeq_aux_kernels.cuh)
namespace {
AuxRegistrar<LotkaVolterra> aux_lotka_volterra;
AuxRegistrar<Pleiades> aux_pleiades;
AuxRegistrar<Hires> aux_hires;
AuxRegistrar<VanDerPol> aux_vanderpol;
AuxRegistrar<Linear> aux_linear;
AuxRegistrar<Burgers> aux_burgers;
}  // namespace

// Shared-memory opt-in of a dense auxiliary kernel (pdeq_aux_dense.cuh); -10 when the state does not fit.
template <class Kern>
static int auxd_prepare(Kern kern, size_t smem_bytes, const char* what) {
  if (smem_bytes > 227 * 1024)
    return api_fail(-10, "%s: the dense state does not fit in shared memory (%zu bytes needed)", what, smem_bytes);
  if (smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return api_cuda_fail(e, what);
  }
  return 0;
}

extern "C" {

int pdeq_taylor_init(const pdeq_config* cfg, int64_t num_instances, const double* u0, const double* params,
                     int64_t params_stride, double t0, double* tcoeffs, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (u0 == nullptr || tcoeffs == nullptr) return api_fail(-22, "u0/tcoeffs is NULL");
  if (pdeq_vf_num_params(cfg->vf_id) > 0 && params == nullptr) return api_fail(-22, "params is NULL");
  if (num_instances == 0) return 0;
  const int n = cfg->num_derivatives + 1, d = cfg->ode_dim, q = pdeq_vf_ode_order(cfg->vf_id);
  const AuxEntry* aux = find_aux(cfg->vf_id);
  if (aux == nullptr) return api_fail(-10, "no Taylor initialisation registered for vf_id %d", cfg->vf_id);
  cudaError_t e = aux->taylor(n - q, num_instances, n, d, u0, params, params_stride, t0, tcoeffs, (cudaStream_t)stream);
  if (e != cudaSuccess) return api_cuda_fail(e, "taylor_init");
  return 0;
}

int pdeq_dt0(const pdeq_config* cfg, int64_t num_instances, const double* u0, const double* params,
             int64_t params_stride, double t0, double scale, double nugget, double* out, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (u0 == nullptr || out == nullptr) return api_fail(-22, "u0/out is NULL");
  if (pdeq_vf_num_params(cfg->vf_id) > 0 && params == nullptr) return api_fail(-22, "params is NULL");
  if (num_instances == 0) return 0;
  const AuxEntry* aux = find_aux(cfg->vf_id);
  if (aux == nullptr) return api_fail(-10, "no dt0 registered for vf_id %d", cfg->vf_id);
  cudaError_t e = aux->dt0(num_instances, cfg->ode_dim, u0, params, params_stride, t0, scale, nugget, out,
                           (cudaStream_t)stream);
  if (e != cudaSuccess) return api_cuda_fail(e, "dt0");
  return 0;
}

int pdeq_dt0_adaptive(const pdeq_config* cfg, int64_t num_instances, const double* u0, const double* params,
                      int64_t params_stride, double t0, double error_contraction_rate, double rtol, double atol,
                      double* out, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (u0 == nullptr || out == nullptr) return api_fail(-22, "u0/out is NULL");
  if (pdeq_vf_ode_order(cfg->vf_id) != 1)
    return api_fail(-5, "dt0_adaptive is defined for first-order ODEs only (stepsize_initialisers.py:37-38)");
  if (pdeq_vf_num_params(cfg->vf_id) > 0 && params == nullptr) return api_fail(-22, "params is NULL");
  if (num_instances == 0) return 0;
  const AuxEntry* aux = find_aux(cfg->vf_id);
  if (aux == nullptr) return api_fail(-10, "no dt0_adaptive registered for vf_id %d", cfg->vf_id);
  cudaError_t e = aux->dt0_adaptive(num_instances, cfg->ode_dim, u0, params, params_stride, t0,
                                    error_contraction_rate, rtol, atol, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return api_cuda_fail(e, "dt0_adaptive");
  return 0;
}

int pdeq_lml_terminal_values(const pdeq_config* cfg, int64_t num_instances, int32_t tcoeff_index,
                             const double* mean, const double* chol, const double* data, int64_t data_stride,
                             const double* std, int64_t std_stride, double* out, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (mean == nullptr || chol == nullptr || data == nullptr || std == nullptr || out == nullptr)
    return api_fail(-22, "NULL argument");
  if (tcoeff_index < 0 || tcoeff_index > cfg->num_derivatives) return api_fail(-5, "bad tcoeff_index");
  int fact = cfg->factorisation;
  if (fact == PDEQ_FACT_DENSE && cfg->ode_dim == 1) fact = PDEQ_FACT_ISOTROPIC;
  if (num_instances == 0) return 0;
  const int threads = 128;
  const int grid = (int)((num_instances * 32 + threads - 1) / threads);
  if (fact == PDEQ_FACT_DENSE) {
    if (cfg->ode_dim > LML_DENSE_MAX_D)
      return api_fail(-10, "lml_terminal_values: dense factorisation supports ode_dim <= %d", LML_DENSE_MAX_D);
    lml_terminal_dense_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
        num_instances, cfg->num_derivatives + 1, cfg->ode_dim, tcoeff_index, mean, chol, data, data_stride, std,
        std_stride, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_cuda_fail(e, "lml_terminal_values (dense)");
    return 0;
  }
  lml_terminal_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
      num_instances, cfg->num_derivatives + 1, cfg->ode_dim, fact == PDEQ_FACT_BLOCKDIAG, tcoeff_index, mean, chol,
      data, data_stride, std, std_stride, fact == PDEQ_FACT_BLOCKDIAG, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_cuda_fail(e, "lml_terminal_values");
  return 0;
}

int pdeq_lml_timeseries(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints, int32_t tcoeff_index,
                        int32_t average_pdfs, const double* mean, const double* chol, const double* bw_gain,
                        const double* bw_mean, const double* bw_chol, const double* data, int64_t data_stride,
                        const double* std, int64_t std_stride, double* out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (mean == nullptr || chol == nullptr || bw_gain == nullptr || bw_mean == nullptr || bw_chol == nullptr ||
      data == nullptr || std == nullptr || out == nullptr)
    return api_fail(-22, "NULL argument");
  if (num_gridpoints < 1) return api_fail(-23, "num_gridpoints must be >= 1");
  if (tcoeff_index < 0 || tcoeff_index > cfg->num_derivatives) return api_fail(-5, "bad tcoeff_index");
  int fact = cfg->factorisation;
  if (fact == PDEQ_FACT_DENSE && cfg->ode_dim == 1) fact = PDEQ_FACT_ISOTROPIC;
  if (num_instances == 0) return 0;
  if (fact == PDEQ_FACT_DENSE) {  // CTA per instance, run-time dimensions (pdeq_aux_dense.cuh)
    const int n = cfg->num_derivatives + 1, d = cfg->ode_dim;
    const size_t smem = auxd_lml_smem_doubles(n * d, d) * sizeof(double);
    rc = auxd_prepare(lml_timeseries_dense_kernel, smem, "lml_timeseries (dense)");
    if (rc != 0) return rc;
    lml_timeseries_dense_kernel<<<(unsigned)num_instances, AUXD_THREADS, smem, (cudaStream_t)stream>>>(
        num_gridpoints, n, d, tcoeff_index, average_pdfs, mean, chol, bw_gain, bw_mean, bw_chol, data, data_stride,
        std, std_stride, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_cuda_fail(e, "lml_timeseries (dense)");
    return 0;
  }
  const int64_t total = num_instances * (int64_t)cfg->ode_dim;
  if (workspace == nullptr || workspace_bytes < (size_t)total * sizeof(double))
    return api_fail(-24, "lml_timeseries needs %zu workspace bytes (num_instances * ode_dim doubles)",
                    (size_t)total * sizeof(double));
  const int threads = 64;
  const int grid = (int)((total + threads - 1) / threads);
  const int bd = fact == PDEQ_FACT_BLOCKDIAG;
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = (double*)workspace;
#define PDEQ_LML_CASE(NN)                                                                                         \
  case NN:                                                                                                        \
    lml_timeseries_kernel<NN><<<grid, threads, 0, st>>>(num_instances, num_gridpoints, cfg->ode_dim, bd,          \
                                                        tcoeff_index, average_pdfs, mean, chol, bw_gain, bw_mean, \
                                                        bw_chol, data, data_stride, std, std_stride, partial);    \
    break;
  switch (cfg->num_derivatives + 1) {
    PDEQ_LML_CASE(2)
    PDEQ_LML_CASE(3)
    PDEQ_LML_CASE(4)
    PDEQ_LML_CASE(5)
    PDEQ_LML_CASE(6)
    PDEQ_LML_CASE(7)
    PDEQ_LML_CASE(8)
    default:
      return api_fail(-10, "lml_timeseries: num_derivatives must be in 1..7");
  }
#undef PDEQ_LML_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_cuda_fail(e, "lml_timeseries");
  lml_reduce_kernel<<<(int)((num_instances + 127) / 128), 128, 0, st>>>(num_instances, cfg->ode_dim, partial, out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return api_cuda_fail(e, "lml_timeseries (reduce)");
  return 0;
}

int pdeq_sample_posterior(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints, int32_t num_samples,
                          const double* mean, const double* chol, const double* bw_gain, const double* bw_mean,
                          const double* bw_chol, const double* base, double* out, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (mean == nullptr || chol == nullptr || bw_gain == nullptr || bw_mean == nullptr || bw_chol == nullptr ||
      base == nullptr || out == nullptr)
    return api_fail(-22, "NULL argument");
  if (num_gridpoints < 1) return api_fail(-23, "num_gridpoints must be >= 1");
  int fact = cfg->factorisation;
  if (fact == PDEQ_FACT_DENSE && cfg->ode_dim == 1) fact = PDEQ_FACT_ISOTROPIC;
  if (num_instances == 0 || num_samples == 0) return 0;
  if (fact == PDEQ_FACT_DENSE) {  // CTA per (instance, sample); base is [B][S][T][n d]
    const int N = (cfg->num_derivatives + 1) * cfg->ode_dim;
    if (N > 128) return api_fail(-10, "sample_posterior: dense factorisation supports (nu + 1) d <= 128");
    sample_dense_kernel<<<(unsigned)(num_instances * num_samples), 64, 2 * N * sizeof(double),
                          (cudaStream_t)stream>>>(num_samples, num_gridpoints, N, mean, chol, bw_gain, bw_mean,
                                                  bw_chol, base, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_cuda_fail(e, "sample_posterior (dense)");
    return 0;
  }
  const int64_t total = num_instances * (int64_t)num_samples * cfg->ode_dim;
  const int threads = 128;
  const int nblocks = (int)((total + threads - 1) / threads);
  const int bd = fact == PDEQ_FACT_BLOCKDIAG;
#define PDEQ_SAMPLE_CASE(NN)                                                                                    \
  case NN:                                                                                                      \
    sample_kernel<NN><<<nblocks, threads, 0, (cudaStream_t)stream>>>(num_instances, num_samples, num_gridpoints, \
                                                                     cfg->ode_dim, bd, mean, chol, bw_gain,     \
                                                                     bw_mean, bw_chol, base, out);              \
    break;
  switch (cfg->num_derivatives + 1) {
    PDEQ_SAMPLE_CASE(2)
    PDEQ_SAMPLE_CASE(3)
    PDEQ_SAMPLE_CASE(4)
    PDEQ_SAMPLE_CASE(5)
    PDEQ_SAMPLE_CASE(6)
    PDEQ_SAMPLE_CASE(7)
    PDEQ_SAMPLE_CASE(8)
    default:
      return api_fail(-10, "sample_posterior: num_derivatives must be in 1..7");
  }
#undef PDEQ_SAMPLE_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_cuda_fail(e, "sample_posterior");
  return 0;
}

int pdeq_offgrid_marginals(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints, const double* grid,
                           int32_t num_queries, const double* queries, const double* mean, const double* chol,
                           const double* filt_mean, const double* filt_chol, const double* output_scale,
                           const double* prior_scale, int64_t prior_scale_stride, double* out_mean,
                           double* out_chol, void* stream) {
  int rc = api_validate(cfg);
  if (rc != 0) return rc;
  if (grid == nullptr || queries == nullptr || mean == nullptr || chol == nullptr || output_scale == nullptr ||
      out_mean == nullptr)
    return api_fail(-22, "NULL argument");
  if (num_gridpoints < 2) return api_fail(-23, "offgrid_marginals needs at least two grid points");
  // strategy_smoother_fixedpoint.is_suitable_for_offgrid_marginals is False (estimators_and_losses.py:519-523)
  if (cfg->strategy == PDEQ_STRATEGY_FIXEDPOINT)
    return api_fail(-10, "offgrid_marginals is not defined for the fixed-point smoother (the reference raises "
                         "NotImplementedError); use a filter or the fixed-interval smoother");
  const bool smooth = cfg->strategy != PDEQ_STRATEGY_FILTER;
  if (smooth && (filt_mean == nullptr || filt_chol == nullptr))
    return api_fail(-22, "a smoothing solution needs its filtering marginals (filt_mean / filt_chol)");
  int fact = cfg->factorisation;
  if (fact == PDEQ_FACT_DENSE && cfg->ode_dim == 1) fact = PDEQ_FACT_ISOTROPIC;
  if (num_instances == 0 || num_queries == 0) return 0;
  if (fact == PDEQ_FACT_DENSE) {  // CTA per (instance, query), run-time dimensions (pdeq_aux_dense.cuh)
    const int N = (cfg->num_derivatives + 1) * cfg->ode_dim;
    if (cfg->ode_dim > 64) return api_fail(-10, "offgrid_marginals: dense factorisation supports ode_dim <= 64");
    const size_t smem = auxd_offgrid_smem_doubles(N, smooth) * sizeof(double);
    rc = auxd_prepare(offgrid_dense_kernel, smem, "offgrid_marginals (dense)");
    if (rc != 0) return rc;
    offgrid_dense_kernel<<<(unsigned)(num_instances * num_queries), AUXD_THREADS, smem, (cudaStream_t)stream>>>(
        *cfg, num_gridpoints, num_queries, smooth ? 1 : 0, grid, queries, mean, chol, filt_mean, filt_chol,
        output_scale, prior_scale, prior_scale_stride, out_mean, out_chol);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_cuda_fail(e, "offgrid_marginals (dense)");
    return 0;
  }
  const int64_t total = num_instances * (int64_t)num_queries * cfg->ode_dim;
  const int threads = 64;
  const int nblocks = (int)((total + threads - 1) / threads);
  const int bd = fact == PDEQ_FACT_BLOCKDIAG;
  cudaStream_t st = (cudaStream_t)stream;
#define PDEQ_OFFGRID_CASE(NN)                                                                                       \
  case NN:                                                                                                          \
    if (smooth)                                                                                                     \
      offgrid_kernel<NN, true><<<nblocks, threads, 0, st>>>(*cfg, num_instances, num_gridpoints, num_queries, bd,   \
                                                            grid, queries, mean, chol, filt_mean, filt_chol,        \
                                                            output_scale, prior_scale, prior_scale_stride,          \
                                                            out_mean, out_chol);                                    \
    else                                                                                                            \
      offgrid_kernel<NN, false><<<nblocks, threads, 0, st>>>(*cfg, num_instances, num_gridpoints, num_queries, bd,  \
                                                             grid, queries, mean, chol, filt_mean, filt_chol,       \
                                                             output_scale, prior_scale, prior_scale_stride,         \
                                                             out_mean, out_chol);                                   \
    break;
  switch (cfg->num_derivatives + 1) {
    PDEQ_OFFGRID_CASE(2)
    PDEQ_OFFGRID_CASE(3)
    PDEQ_OFFGRID_CASE(4)
    PDEQ_OFFGRID_CASE(5)
    PDEQ_OFFGRID_CASE(6)
    PDEQ_OFFGRID_CASE(7)
    PDEQ_OFFGRID_CASE(8)
    default:
      return api_fail(-10, "offgrid_marginals: num_derivatives must be in 1..7");
  }
#undef PDEQ_OFFGRID_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return api_cuda_fail(e, "offgrid_marginals");
  return 0;
}

}  // extern "C"
