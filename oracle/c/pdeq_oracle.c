/*
 * Plain-C restatement of the reference's adaptive step loop for the isotropic ts0 filter
 * (BASELINE configs 1 and 2).  TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see oracle/__init__.py.
 * It is validated against the NumPy oracle in tests/test_oracle_c_port.py and is what bench.py times
 * as the "restated reference (C, N cores)" CPU baseline.  It is never linked into the product.
 *
 * Follows, in execution order (paths relative to /root/reference):
 *   probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:100-146,227-338   loops, accept/reject, clip_dt
 *   probdiffeq/_ivpsolve/controllers.py:46-63,78-84                      PI / I controllers
 *   probdiffeq/_probdiffeq/solvers.py:702-733                            solver.step
 *   probdiffeq/_probdiffeq/solvers.py:1037-1098                          error_state_std
 *   probdiffeq/_probdiffeq/ssm_impl_isotropic.py:75-133,304-317,367-378  apply_flat/marginalise/revert/ts0/transition
 *   probdiffeq/util/cholesky_util.py:27-103                              revert_conditional, sum_of_sqrtm_factors
 *   probdiffeq/_probdiffeq/utilities.py:74-84                            Taylor preconditioner
 * The QR is an unblocked Householder (LAPACK dgeqr2/dlarfg), i.e. what jnp.linalg.qr(mode="r") runs on CPU
 * for these sizes (probdiffeq/backend/linalg.py:8-10).  Dense, no structure exploitation: this is the
 * reference's algorithm, not the CUDA kernel's.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 8

/* R-factor of an m x n row-major matrix, in place (upper triangle of the first n rows). */
static void qr_r(double* a, int m, int n, int lda) {
  int kmax = m < n ? m : n;
  for (int j = 0; j < kmax; ++j) {
    double ss = 0.0;
    for (int r = j + 1; r < m; ++r) ss += a[r * lda + j] * a[r * lda + j];
    if (ss == 0.0) continue; /* dlarfg: tau = 0 */
    double alpha = a[j * lda + j];
    double xnorm = sqrt(ss);
    double beta = -copysign(hypot(alpha, xnorm), alpha);
    double tau = (beta - alpha) / beta;
    double scal = 1.0 / (alpha - beta);
    for (int r = j + 1; r < m; ++r) a[r * lda + j] *= scal;
    a[j * lda + j] = beta;
    for (int c = j + 1; c < n; ++c) {
      double w = a[j * lda + c];
      for (int r = j + 1; r < m; ++r) w += a[r * lda + j] * a[r * lda + c];
      w *= tau;
      a[j * lda + c] -= w;
      for (int r = j + 1; r < m; ++r) a[r * lda + c] -= w * a[r * lda + j];
    }
  }
}

typedef struct {
  int n;          /* number of Taylor coefficients */
  int d;          /* ODE dimension (2 for Lotka-Volterra) */
  int control_pi; /* 1 = PI, 0 = I */
  int clip_dt;
  double safety, fmin_, fmax_, exp_i, exp_p;
  double A[NMAX][NMAX], Q[NMAX][NMAX], fact[NMAX + 1];
} oracle_cfg;

static void lotka_volterra(const double* p, const double* u, double* f) {
  f[0] = p[0] * u[0] - p[1] * u[0] * u[1];
  f[1] = -p[2] * u[1] + p[3] * u[0] * u[1];
}

/* revert of a ts0 observation (row selector e_1, noise damp) on chol L (n x n lower):
   returns r_y, gain[n], Lout (n x n lower). */
static void revert_ts0(int n, const double L[NMAX][NMAX], double damp, double* r_y, double* gain,
                       double Lout[NMAX][NMAX]) {
  double S[(NMAX + 1) * (NMAX + 1)];
  int m = n + 1;
  memset(S, 0, sizeof(S));
  S[0] = damp;
  for (int r = 0; r < n; ++r) {
    S[(1 + r) * m + 0] = L[1][r]; /* (H L)^T */
    for (int c = 0; c < n; ++c) S[(1 + r) * m + 1 + c] = L[c][r]; /* L^T */
  }
  qr_r(S, m, m, m);
  *r_y = S[0];
  for (int i = 0; i < n; ++i) gain[i] = S[1 + i] / S[0];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Lout[i][j] = (j <= i) ? S[(1 + j) * m + 1 + i] : 0.0;
}

/* One instance: terminal-value solve on [t0, t1]. Returns the number of accepted steps. */
static int solve_one(const oracle_cfg* c, const double* tcoeffs, const double* params, double t0, double t1,
                     double atol, double rtol, double dt0, double eps, double damp, double* mean_out,
                     double* chol_out, double* t_out, int* attempts_out) {
  const int n = c->n, d = c->d;
  double m[NMAX][2], L[NMAX][NMAX];
  memset(L, 0, sizeof(L));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < d; ++j) m[i][j] = tcoeffs[i * d + j];
  double t = t0, dt = dt0, prev = 1.0;
  int nsteps = 0, attempts = 0;
  while (t + eps < t1) {
    double acc_factor = 0.9;
    double mn[NMAX][2], Ln[NMAX][NMAX], dtc = dt;
    while (acc_factor < 1.0) {
      attempts += 1;
      dtc = c->clip_dt ? fmin(dt, t1 - t) : dt;
      /* preconditioner: p = dt^k / k!, p_inv = dt^-k * k!, k = nu..0 */
      double p[NMAX], pinv[NMAX];
      for (int i = 0; i < n; ++i) {
        double k = (double)(n - 1 - i);
        p[i] = pow(dtc, k) / c->fact[n - 1 - i];
        pinv[i] = pow(dtc, -k) * c->fact[n - 1 - i];
      }
      double sq = sqrt(fabs(dtc));
      /* predict: marginalise */
      double S[2 * NMAX * NMAX];
      for (int r = 0; r < n; ++r)
        for (int cc = 0; cc < n; ++cc) {
          double b = 0.0; /* B[cc][r] = sum_k A[cc][k] pinv[k] L[k][r] */
          for (int k = 0; k < n; ++k) b += c->A[cc][k] * (pinv[k] * L[k][r]);
          S[r * n + cc] = b;
          S[(n + r) * n + cc] = sq * c->Q[cc][r];
        }
      qr_r(S, 2 * n, n, n);
      double Lp[NMAX][NMAX], mp[NMAX][2];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Lp[i][j] = (j <= i) ? fabs(p[i]) * S[j * n + i] : 0.0;
      for (int j = 0; j < d; ++j)
        for (int i = 0; i < n; ++i) {
          double a = 0.0;
          for (int k = 0; k < n; ++k) a += c->A[i][k] * (pinv[k] * m[k][j]);
          mp[i][j] = p[i] * a;
        }
      /* ts0 linearisation at the predicted mean + correction */
      double u[2] = {mp[0][0], mp[0][1]}, f[2];
      lotka_volterra(params, u, f);
      double ry, gain[NMAX];
      revert_ts0(n, Lp, damp, &ry, gain, Ln);
      double mobs[2];
      for (int j = 0; j < d; ++j) {
        mobs[j] = mp[1][j] - f[j];
        for (int i = 0; i < n; ++i) mn[i][j] = mp[i][j] - gain[i] * mobs[j];
      }
      /* error_state_std: Bayes rule on the zero-error extrapolation */
      double Lq[NMAX][NMAX], Lc[NMAX][NMAX], rye, g2[NMAX];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Lq[i][j] = (j <= i) ? fabs(p[i]) * (sq * c->Q[i][j]) : 0.0;
      revert_ts0(n, Lq, damp, &rye, g2, Lc);
      double ss = 0.0;
      for (int j = 0; j < d; ++j) ss += (mobs[j] / rye) * (mobs[j] / rye);
      double sigma = sqrt(ss) / sqrt((double)d);
      double err = sigma * fabs(Lc[0][0]);
      double acc = 0.0;
      for (int j = 0; j < d; ++j) {
        double ref = fmax(fabs(m[0][j]), fabs(mn[0][j]));
        double w = (err / c->fact[0]) / (atol + rtol * ref);
        acc += w * w;
      }
      double norm = sqrt(acc) / sqrt((double)d);
      acc_factor = pow(norm, -1.0 / (double)n);
      double ratio;
      if (c->control_pi) {
        ratio = c->safety * pow(acc_factor, c->exp_i) * pow(acc_factor / prev, c->exp_p);
        if (acc_factor >= 1.0) prev = acc_factor;
      } else {
        ratio = c->safety * acc_factor;
      }
      dt = fmax(c->fmin_, fmin(ratio, c->fmax_)) * dtc;
      if (attempts > 100000000) break;
    }
    memcpy(m, mn, sizeof(m));
    memcpy(L, Ln, sizeof(L));
    t = t + dtc;
    nsteps += 1;
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < d; ++j) mean_out[i * d + j] = m[i][j];
  if (chol_out)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) chol_out[i * n + j] = L[i][j];
  *t_out = t;
  *attempts_out = attempts;
  return nsteps;
}

/* Ensemble entry point (OpenMP over instances). Returns the total number of accepted steps. */
int64_t pdeq_oracle_lv_terminal(const oracle_cfg* cfg, int64_t B, const double* tcoeffs, const double* params,
                                double t0, double t1, double atol, double rtol, double dt0, double eps,
                                double damp, double* mean_out, double* chol_out, double* t_out,
                                int32_t* num_steps, int32_t* num_attempts, int32_t num_threads) {
  int64_t total = 0;
  const int n = cfg->n, d = cfg->d;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total)
  for (int64_t b = 0; b < B; ++b) {
    int att = 0;
    double tt = 0.0;
    int ns = solve_one(cfg, tcoeffs + b * n * d, params + b * 4, t0, t1, atol, rtol, dt0, eps, damp,
                       mean_out + b * n * d, chol_out ? chol_out + b * n * n : 0, &tt, &att);
    num_steps[b] = ns;
    num_attempts[b] = att;
    t_out[b] = tt;
    total += ns;
  }
  return total;
}

int pdeq_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
