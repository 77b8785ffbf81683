// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// jet.hpp: forward-mode dual numbers, the arithmetic ceres::Jet<double, N>
// performs inside ceres::AutoDiffCostFunction (third party, Ceres Solver 2.0,
// not vendored under /root/reference; call sites lidarFactor.hpp:50,99,132 and
// ceres_cost_function.h:90,179).  Only the operations the reference's functors
// use are provided.
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct Jet {
  double a = 0.0;
  double v[N];
  Jet() { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT implicit like ceres::Jet
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h;
}
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h;
}
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h;
}
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
// ceres::Jet division: (f/g)' = (f' - (f/g) g') / g
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return Jet<N>(s) * f; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { return f * Jet<N>(s); }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { return f + Jet<N>(s); }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { return f - Jet<N>(s); }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
template <int N> inline bool operator>=(const Jet<N>& f, const Jet<N>& g) { return f.a >= g.a; }
template <int N> inline bool operator<=(const Jet<N>& f, const Jet<N>& g) { return f.a <= g.a; }

template <int N> inline Jet<N> jsqrt(const Jet<N>& f) {
  Jet<N> h; h.a = std::sqrt(f.a); const double d = 1.0 / (2.0 * h.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * d;
  return h;
}
template <int N> inline Jet<N> jsin(const Jet<N>& f) {
  Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}
template <int N> inline Jet<N> jcos(const Jet<N>& f) {
  Jet<N> h; h.a = std::cos(f.a); const double s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i];
  return h;
}
template <int N> inline Jet<N> jacos(const Jet<N>& f) {
  Jet<N> h; h.a = std::acos(f.a); const double d = -1.0 / std::sqrt(1.0 - f.a * f.a);
  for (int i = 0; i < N; ++i) h.v[i] = d * f.v[i];
  return h;
}
template <int N> inline Jet<N> jabs(const Jet<N>& f) { return f.a < 0.0 ? -f : f; }

// double overloads so functors can be instantiated with T = double too.
inline double jsqrt(double f) { return std::sqrt(f); }
inline double jsin(double f) { return std::sin(f); }
inline double jcos(double f) { return std::cos(f); }
inline double jacos(double f) { return std::acos(f); }
inline double jabs(double f) { return std::fabs(f); }

}  // namespace oracle
