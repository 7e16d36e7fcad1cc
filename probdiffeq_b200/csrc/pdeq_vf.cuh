// Registered vector-field device functors: the right-hand sides of the reference's benchmark problems.
//
// Each functor is written once, templated on the scalar type T, so that the same source evaluates the
// vector field in float64 (T = double), and propagates truncated Taylor series (T = Series<K>) for the
// on-device Taylor-coefficient initialisation that the reference obtains from `jax.experimental.jet`
// (probdiffeq/_probdiffeq/jet_expansion_algorithms.py:49-177).  Jacobians (exact, as the reference's
// `materialize_dense`, probdiffeq/_probdiffeq/jacobians.py:93-98) are written out analytically.
//
// Sources of the right-hand sides (paths relative to the reference checkout):
//   lotka_volterra  benchmarks/A0_work-precision-lotka-volterra.py:104-109
//   pleiades        benchmarks/A1_work-precision-pleiades.py:105-119 (first-order form, i == j skipped)
//   hires           benchmarks/A2_work-precision-hires.py:157-173
//   vanderpol       benchmarks/A3_work-precision-vanderpol.py:98-101 (second order, stiffness = param)
//   linear          benchmarks/A4_work-precision-linear-ode.py:114-118
//   burgers         benchmarks/A5_work-precision-burgers-pde.py:123-134 (dx = 1/(d+1))
#pragma once

#include "pdeq_blockops.cuh"

namespace pdeq {

enum VfId { VF_LOTKA_VOLTERRA = 0, VF_PLEIADES = 1, VF_HIRES = 2, VF_VANDERPOL = 3, VF_LINEAR = 4, VF_BURGERS = 5, VF_COUNT = 6 };

// ---------------------------------------------------------------------------------------------------
// Truncated Taylor series with normalised coefficients c[k] = x^(k)(t0) / k!.
// ---------------------------------------------------------------------------------------------------
template <int K>
struct Series {
  double c[K];
};

template <int K>
PDEQ_DI Series<K> operator+(const Series<K>& a, const Series<K>& b) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) r.c[k] = a.c[k] + b.c[k];
  return r;
}
template <int K>
PDEQ_DI Series<K> operator-(const Series<K>& a, const Series<K>& b) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) r.c[k] = a.c[k] - b.c[k];
  return r;
}
template <int K>
PDEQ_DI Series<K> operator-(const Series<K>& a) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) r.c[k] = -a.c[k];
  return r;
}
template <int K>
PDEQ_DI Series<K> operator*(const Series<K>& a, const Series<K>& b) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j <= k; ++j) acc = fma(a.c[j], b.c[k - j], acc);
    r.c[k] = acc;
  }
  return r;
}
template <int K>
PDEQ_DI Series<K> operator*(double s, const Series<K>& a) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) r.c[k] = s * a.c[k];
  return r;
}
template <int K>
PDEQ_DI Series<K> operator*(const Series<K>& a, double s) { return s * a; }
template <int K>
PDEQ_DI Series<K> operator+(const Series<K>& a, double s) {
  Series<K> r = a;
  r.c[0] += s;
  return r;
}
template <int K>
PDEQ_DI Series<K> operator+(double s, const Series<K>& a) { return a + s; }
template <int K>
PDEQ_DI Series<K> operator-(const Series<K>& a, double s) { return a + (-s); }
template <int K>
PDEQ_DI Series<K> operator-(double s, const Series<K>& a) { return (-a) + s; }
template <int K>
PDEQ_DI Series<K> operator/(const Series<K>& a, const Series<K>& b) {
  Series<K> q;
  const double inv = 1.0 / b.c[0];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double acc = a.c[k];
#pragma unroll
    for (int j = 0; j < k; ++j) acc = fma(-q.c[j], b.c[k - j], acc);
    q.c[k] = acc * inv;
  }
  return q;
}
// x^(3/2) through the power-rule recurrence k x_0 y_k = sum_{j=1..k} (r j - (k - j)) x_j y_{k-j}.
template <int K>
PDEQ_DI Series<K> pow_three_halves(const Series<K>& x) {
  Series<K> y;
  const double r = 1.5;
  y.c[0] = x.c[0] * sqrt(x.c[0]);
  const double inv = 1.0 / x.c[0];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    double acc = 0.0;
#pragma unroll
    for (int j = 1; j <= k; ++j) acc = fma((r * j - (k - j)) * x.c[j], y.c[k - j], acc);
    y.c[k] = acc * inv / double(k);
  }
  return y;
}
PDEQ_DI double pow_three_halves(double x) { return x * sqrt(x); }

template <class T>
PDEQ_DI T constant_like(const T&, double v);
template <>
PDEQ_DI double constant_like<double>(const double&, double v) { return v; }
template <int K>
PDEQ_DI Series<K> constant_like(const Series<K>&, double v) {
  Series<K> r;
#pragma unroll
  for (int k = 0; k < K; ++k) r.c[k] = 0.0;
  r.c[0] = v;
  return r;
}

// ---------------------------------------------------------------------------------------------------
// Forward-mode derivative of a functor's `component`: a Series<2> is a dual number (value, derivative), so seeding
// one jet coordinate with 1 and reading coefficient 1 of the result gives d f_i / d u^(k)_j. This is what the
// reference obtains from `jacfwd` (jacobians.py:93-98); a functor can use it as its `jac` instead of stating the
// derivatives by hand.
// ---------------------------------------------------------------------------------------------------
template <class Acc>
struct SeededAcc {
  const Acc& u;
  int k, j;
  PDEQ_DI Series<2> operator()(int kk, int jj) const {
    Series<2> s;
    s.c[0] = u(kk, jj);
    s.c[1] = (kk == k && jj == j) ? 1.0 : 0.0;
    return s;
  }
};
template <class VF, class Acc>
PDEQ_DI double autodiff_jac(int i, int k, int j, int d, const Acc& u, const double* p, double t) {
  const SeededAcc<Acc> seeded{u, k, j};
  return VF::template component<Series<2>>(i, d, seeded, p, t).c[1];
}

// ---------------------------------------------------------------------------------------------------
// Functors. `Acc` is any callable (k, i) -> T that returns component i of the k-th jet coordinate
// (k < order): registers for thread-per-instance kernels, shared memory for group kernels.
// component(i, ...) returns f_i; jac(i, k, j, ...) returns d f_i / d u^(k)_j.
// ---------------------------------------------------------------------------------------------------

struct LotkaVolterra {
  static constexpr int id = VF_LOTKA_VOLTERRA, order = 1, num_params = 4, fixed_dim = 2;
  template <class T, class Acc>
  PDEQ_DI static T component(int i, int /*d*/, const Acc& u, const double* p, double /*t*/) {
    const T u0 = u(0, 0), u1 = u(0, 1);
    if (i == 0) return p[0] * u0 - p[1] * u0 * u1;
    return -p[2] * u1 + p[3] * u0 * u1;
  }
  template <class Acc>
  PDEQ_DI static double jac(int i, int /*k*/, int j, int /*d*/, const Acc& u, const double* p, double /*t*/) {
    const double u0 = u(0, 0), u1 = u(0, 1);
    if (i == 0) return j == 0 ? p[0] - p[1] * u1 : -p[1] * u0;
    return j == 0 ? p[3] * u1 : -p[2] + p[3] * u0;
  }
};

struct Pleiades {
  static constexpr int id = VF_PLEIADES, order = 1, num_params = 0, fixed_dim = 28;
  template <class T, class Acc>
  PDEQ_DI static T component(int i, int /*d*/, const Acc& u, const double* /*p*/, double /*t*/) {
    if (i < 14) return u(0, 14 + i);  // x' = vx, y' = vy
    const int body = (i - 14) % 7;
    const bool is_x = i < 21;
    const T xi = u(0, body), yi = u(0, 7 + body);
    T acc = constant_like(xi, 0.0);
    for (int j = 0; j < 7; ++j) {
      if (j == body) continue;
      const T dx = u(0, j) - xi, dy = u(0, 7 + j) - yi;
      if constexpr (std::is_same<T, double>::value) {
        // step loops: 1 / r^3 from one reciprocal square root (MUFU seed + one correction, within an ulp) instead of
        // a square root and a division -- both with slow-path branches, six times per lane and attempt
        const double ir = fast_rsqrt(dx * dx + dy * dy);
        acc = fma(double(j + 1) * (is_x ? dx : dy), (ir * ir) * ir, acc);
      } else {
        const T r3 = pow_three_halves(dx * dx + dy * dy);
        acc = acc + double(j + 1) * ((is_x ? dx : dy) / r3);
      }
    }
    return acc;
  }
  template <class Acc>
  PDEQ_DI static double jac(int i, int /*k*/, int j, int /*d*/, const Acc& u, const double* /*p*/, double /*t*/) {
    if (i < 14) return j == 14 + i ? 1.0 : 0.0;
    if (j >= 14) return 0.0;
    const int body = (i - 14) % 7;
    const bool is_x = i < 21;       // which acceleration component
    const bool wrt_x = j < 7;       // derivative w.r.t. an x or a y position
    const int other = j % 7;
    const double xi = u(0, body), yi = u(0, 7 + body);
    double acc = 0.0;
    for (int l = 0; l < 7; ++l) {
      if (l == body) continue;
      if (other != body && other != l) continue;
      const double dx = u(0, l) - xi, dy = u(0, 7 + l) - yi;
      const double r2 = dx * dx + dy * dy;
      const double r3 = r2 * sqrt(r2), r5 = r3 * r2;
      const double num = is_x ? dx : dy, den = wrt_x ? dx : dy;
      // d/d(pos_l) of m num / r^3 ; d/d(pos_body) is its negative
      double g = -3.0 * num * den / r5;
      if (is_x == wrt_x) g += 1.0 / r3;
      g *= double(l + 1);
      acc += (other == l) ? g : -g;
    }
    return acc;
  }
};

struct Hires {
  static constexpr int id = VF_HIRES, order = 1, num_params = 0, fixed_dim = 8;
  template <class T, class Acc>
  PDEQ_DI static T component(int i, int /*d*/, const Acc& u, const double* /*p*/, double /*t*/) {
    switch (i) {
      case 0: return -1.71 * u(0, 0) + 0.43 * u(0, 1) + 8.32 * u(0, 2) + 0.0007;
      case 1: return 1.71 * u(0, 0) - 8.75 * u(0, 1);
      case 2: return -10.03 * u(0, 2) + 0.43 * u(0, 3) + 0.035 * u(0, 4);
      case 3: return 8.32 * u(0, 1) + 1.71 * u(0, 2) - 1.12 * u(0, 3);
      case 4: return -1.745 * u(0, 4) + 0.43 * u(0, 5) + 0.43 * u(0, 6);
      case 5: return -280.0 * u(0, 5) * u(0, 7) + 0.69 * u(0, 3) + 1.71 * u(0, 4) - 0.43 * u(0, 5) + 0.69 * u(0, 6);
      case 6: return 280.0 * u(0, 5) * u(0, 7) - 1.81 * u(0, 6);
      default: return -280.0 * u(0, 5) * u(0, 7) + 1.81 * u(0, 6);
    }
  }
  template <class Acc>
  PDEQ_DI static double jac(int i, int /*k*/, int j, int /*d*/, const Acc& u, const double* /*p*/, double /*t*/) {
    const double u5 = u(0, 5), u7 = u(0, 7);
    switch (i * 8 + j) {
      case 0 * 8 + 0: return -1.71;
      case 0 * 8 + 1: return 0.43;
      case 0 * 8 + 2: return 8.32;
      case 1 * 8 + 0: return 1.71;
      case 1 * 8 + 1: return -8.75;
      case 2 * 8 + 2: return -10.03;
      case 2 * 8 + 3: return 0.43;
      case 2 * 8 + 4: return 0.035;
      case 3 * 8 + 1: return 8.32;
      case 3 * 8 + 2: return 1.71;
      case 3 * 8 + 3: return -1.12;
      case 4 * 8 + 4: return -1.745;
      case 4 * 8 + 5: return 0.43;
      case 4 * 8 + 6: return 0.43;
      case 5 * 8 + 3: return 0.69;
      case 5 * 8 + 4: return 1.71;
      case 5 * 8 + 5: return -280.0 * u7 - 0.43;
      case 5 * 8 + 6: return 0.69;
      case 5 * 8 + 7: return -280.0 * u5;
      case 6 * 8 + 5: return 280.0 * u7;
      case 6 * 8 + 6: return -1.81;
      case 6 * 8 + 7: return 280.0 * u5;
      case 7 * 8 + 5: return -280.0 * u7;
      case 7 * 8 + 6: return 1.81;
      case 7 * 8 + 7: return -280.0 * u5;
      default: return 0.0;
    }
  }
};

struct VanDerPol {  // u'' = s ((1 - u^2) u' - u)
  static constexpr int id = VF_VANDERPOL, order = 2, num_params = 1, fixed_dim = 1;
  template <class T, class Acc>
  PDEQ_DI static T component(int /*i*/, int /*d*/, const Acc& u, const double* p, double /*t*/) {
    const T x = u(0, 0), dx = u(1, 0);
    return p[0] * ((1.0 - x * x) * dx - x);
  }
  template <class Acc>
  PDEQ_DI static double jac(int /*i*/, int k, int /*j*/, int /*d*/, const Acc& u, const double* p, double /*t*/) {
    const double x = u(0, 0), dx = u(1, 0);
    return k == 0 ? p[0] * (-2.0 * x * dx - 1.0) : p[0] * (1.0 - x * x);
  }
};

struct Linear {  // u' = scale * u, any d
  static constexpr int id = VF_LINEAR, order = 1, num_params = 1, fixed_dim = 0;
  template <class T, class Acc>
  PDEQ_DI static T component(int i, int /*d*/, const Acc& u, const double* p, double /*t*/) {
    return p[0] * u(0, i);
  }
  template <class Acc>
  PDEQ_DI static double jac(int i, int /*k*/, int j, int /*d*/, const Acc& /*u*/, const double* p, double /*t*/) {
    return i == j ? p[0] : 0.0;
  }
};

struct Burgers {  // viscous Burgers, central differences, zero Dirichlet ghosts, dx = 1/(d+1), any d
  static constexpr int id = VF_BURGERS, order = 1, num_params = 1, fixed_dim = 0;
  template <class T, class Acc>
  PDEQ_DI static T component(int i, int d, const Acc& u, const double* p, double /*t*/) {
    const double dx = 1.0 / double(d + 1);
    const T uc = u(0, i);
    const T ul = i > 0 ? u(0, i - 1) : constant_like(uc, 0.0);
    const T ur = i < d - 1 ? u(0, i + 1) : constant_like(uc, 0.0);
    const T fluxterm = ((ur * ur) * 0.5 - (ul * ul) * 0.5) * (1.0 / (2.0 * dx));
    const T lap = (ur - 2.0 * uc + ul) * (1.0 / (dx * dx));
    return p[0] * lap - fluxterm;
  }
  template <class Acc>
  PDEQ_DI static double jac(int i, int /*k*/, int j, int d, const Acc& u, const double* p, double /*t*/) {
    const double dx = 1.0 / double(d + 1);
    if (j == i) return -2.0 * p[0] / (dx * dx);
    if (j == i - 1) return p[0] / (dx * dx) + u(0, j) / (2.0 * dx);
    if (j == i + 1) return p[0] / (dx * dx) - u(0, j) / (2.0 * dx);
    return 0.0;
  }
};

}  // namespace pdeq
