// TEST INFRASTRUCTURE ONLY -- part of the CPU parity oracle (see oracle/README.md).
// Nothing under oracle/ may be imported, linked or executed by the product path.
//
// Forward-mode dual numbers, the technique ForwardDiff.jl uses for
//   gradient  (EquationsOfState.jl:169,183)
//   jacobian  (EquationsOfState.jl:200,226)   -- nested: duals of duals
//   derivative(HyperelasticityMPh.jl:212, NumFluxes.jl:41,100)
// ForwardDiff itself is a third-party dependency absent from /root/reference and
// unpinned (no Manifest.toml); its published algorithm -- value + N partials, chain
// rule per primitive -- is restated here.  Derivatives are mathematically unique, so
// any correct implementation agrees to roundoff.
#pragma once
#include <cmath>

namespace hso {

template <class T, int N>
struct Dual {
  T v;
  T d[N];
  Dual() : v(T(0.0)) { for (int i = 0; i < N; ++i) d[i] = T(0.0); }
  Dual(double x) : v(T(x)) { for (int i = 0; i < N; ++i) d[i] = T(0.0); }
  // lift an inner-level value (needed for nesting: Dual<Dual<double,9>,9>(inner))
  template <class U = T, class = std::enable_if_t<!std::is_same<U, double>::value>>
  Dual(const T& x) : v(x) { for (int i = 0; i < N; ++i) d[i] = T(0.0); }
};

// ---- scalar (double) primitives with the same spelling as the dual ones ----------
inline double d_pow(double x, double p) { return std::pow(x, p); }
inline double d_exp(double x) { return std::exp(x); }
inline double d_log(double x) { return std::log(x); }
inline double d_sqrt(double x) { return std::sqrt(x); }
inline double d_abs(double x) { return std::fabs(x); }
inline double value_of(double x) { return x; }

template <class T, int N> inline double value_of(const Dual<T, N>& a) { return value_of(a.v); }

// ---- arithmetic -------------------------------------------------------------------
template <class T, int N> inline Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v / b.v;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }

// mixed with plain double (constants are never perturbed)
template <class T, int N> inline Dual<T, N> operator+(const Dual<T, N>& a, double b) { Dual<T, N> r = a; r.v = a.v + b; return r; }
template <class T, int N> inline Dual<T, N> operator+(double b, const Dual<T, N>& a) { Dual<T, N> r = a; r.v = b + a.v; return r; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a, double b) { Dual<T, N> r = a; r.v = a.v - b; return r; }
template <class T, int N> inline Dual<T, N> operator-(double b, const Dual<T, N>& a) { Dual<T, N> r = -a; r.v = b - a.v; return r; }
template <class T, int N> inline Dual<T, N> operator*(const Dual<T, N>& a, double b) {
  Dual<T, N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <class T, int N> inline Dual<T, N> operator*(double b, const Dual<T, N>& a) { return a * b; }
template <class T, int N> inline Dual<T, N> operator/(const Dual<T, N>& a, double b) {
  Dual<T, N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <class T, int N> inline Dual<T, N> operator/(double b, const Dual<T, N>& a) { return Dual<T, N>(b) / a; }

// mixed with the inner level (T is itself a Dual): an inner value is a constant at this level
template <class T, int N, class = std::enable_if_t<!std::is_same<T, double>::value>>
inline Dual<T, N> operator*(const Dual<T, N>& a, const T& b) {
  Dual<T, N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <class T, int N, class = std::enable_if_t<!std::is_same<T, double>::value>>
inline Dual<T, N> operator*(const T& b, const Dual<T, N>& a) { return a * b; }
template <class T, int N, class = std::enable_if_t<!std::is_same<T, double>::value>>
inline Dual<T, N> operator+(const Dual<T, N>& a, const T& b) { Dual<T, N> r = a; r.v = a.v + b; return r; }
template <class T, int N, class = std::enable_if_t<!std::is_same<T, double>::value>>
inline Dual<T, N> operator-(const Dual<T, N>& a, const T& b) { Dual<T, N> r = a; r.v = a.v - b; return r; }

template <class T, int N> inline bool operator<(const Dual<T, N>& a, double b) { return value_of(a) < b; }

// ---- elementary functions (DiffRules: d x^p = p x^(p-1), d exp = exp, d log = 1/x) ----
template <class T, int N> inline Dual<T, N> d_pow(const Dual<T, N>& a, double p) {
  Dual<T, N> r; r.v = d_pow(a.v, p);
  T dp = d_pow(a.v, p - 1.0) * p;
  for (int i = 0; i < N; ++i) r.d[i] = dp * a.d[i];
  return r; }
template <class T, int N> inline Dual<T, N> d_exp(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = d_exp(a.v); for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i]; return r; }
template <class T, int N> inline Dual<T, N> d_log(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = d_log(a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / a.v; return r; }
template <class T, int N> inline Dual<T, N> d_sqrt(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = d_sqrt(a.v); T h = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = h * a.d[i]; return r; }
template <class T, int N> inline Dual<T, N> d_abs(const Dual<T, N>& a) { return value_of(a) < 0.0 ? -a : a; }

}  // namespace hso
