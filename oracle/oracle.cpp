// =====================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU parity oracle for the finite-volume hot path of
// BlackSiberian/HyperelasticSolver.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product path
// (hyperelasticsolver_b200/) never does.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and Julia is
// not installed here, so this restatement cannot be checked against the reference running.
// It is pinned instead against (i) SURVEY.md Appendix B numbers (an independent derivation),
// (ii) an independent torch-autograd restatement (oracle/pyoracle.py -> tests/golden/) and
// (iii) the physical anchors of SURVEY.md section 4.
//
// The code follows the reference function for function, quirks included; each function
// cites the reference file:line it restates.  Derivatives use forward-mode duals
// (dual.hpp) exactly where the reference calls ForwardDiff, independent of the closed
// forms the CUDA kernels use.  Third-party arithmetic that is not under /root/reference:
//   ForwardDiff (gradient/jacobian/derivative)  -> dual.hpp
//   FastGaussQuadrature gausslegendre(6)/gausslobatto(6) -> constants below (unique values)
//   LinearAlgebra.det (LAPACK getrf LU, partial pivoting) -> det_lu()
//   LinearAlgebra.eigvals (LAPACK dgeev on a symmetric-to-roundoff 3x3) -> cyclic Jacobi on
//     the symmetrised matrix (eigenvalues are unique; asymmetry is <= 1e-16 relative)
// No versions are pinned by the reference (no Project.toml / Manifest.toml).
// =====================================================================================
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "dual.hpp"

namespace hso {

// ------------------------------------------------------------------------------------
// Barton2009 parameter block: EquationsOfState.jl:71-116 (same field order).
// ------------------------------------------------------------------------------------
struct Eos {
  double rho0, c0, cv, t0, b0, alpha, beta, gamma, b0sq, k0;
};

static thread_local int g_domain_error = 0;  // Julia would throw DomainError

// ------------------------------------------------------------------------------------
// SimpleLA.jl  (3x3 matrices stored column-major in 9-vectors: m(i,j) = m[i + 3 j])
// ------------------------------------------------------------------------------------
#define M_(m, i, j) (m)[((i) - 1) + 3 * ((j) - 1)]

// SimpleLA.jl:84-89
template <class S> static S det_(const S* m) {
  return M_(m, 1, 1) * (M_(m, 2, 2) * M_(m, 3, 3) - M_(m, 3, 2) * M_(m, 2, 3)) -
         M_(m, 1, 2) * (M_(m, 2, 1) * M_(m, 3, 3) - M_(m, 2, 3) * M_(m, 3, 1)) +
         M_(m, 1, 3) * (M_(m, 2, 1) * M_(m, 3, 2) - M_(m, 2, 2) * M_(m, 3, 1));
}

// SimpleLA.jl:51-79
template <class S> static void inv_(const S* m, S* minv) {
  S dt = det_(m);
  S invdet = 1.0 / dt;
  M_(minv, 1, 1) = (M_(m, 2, 2) * M_(m, 3, 3) - M_(m, 3, 2) * M_(m, 2, 3)) * invdet;
  M_(minv, 1, 2) = (M_(m, 1, 3) * M_(m, 3, 2) - M_(m, 1, 2) * M_(m, 3, 3)) * invdet;
  M_(minv, 1, 3) = (M_(m, 1, 2) * M_(m, 2, 3) - M_(m, 1, 3) * M_(m, 2, 2)) * invdet;
  M_(minv, 2, 1) = (M_(m, 2, 3) * M_(m, 3, 1) - M_(m, 2, 1) * M_(m, 3, 3)) * invdet;
  M_(minv, 2, 2) = (M_(m, 1, 1) * M_(m, 3, 3) - M_(m, 1, 3) * M_(m, 3, 1)) * invdet;
  M_(minv, 2, 3) = (M_(m, 2, 1) * M_(m, 1, 3) - M_(m, 1, 1) * M_(m, 2, 3)) * invdet;
  M_(minv, 3, 1) = (M_(m, 2, 1) * M_(m, 3, 2) - M_(m, 3, 1) * M_(m, 2, 2)) * invdet;
  M_(minv, 3, 2) = (M_(m, 3, 1) * M_(m, 1, 2) - M_(m, 1, 1) * M_(m, 3, 2)) * invdet;
  M_(minv, 3, 3) = (M_(m, 1, 1) * M_(m, 2, 2) - M_(m, 2, 1) * M_(m, 1, 2)) * invdet;
}

// SimpleLA.jl:94-96
template <class S> static S tr_(const S* m) { return M_(m, 1, 1) + M_(m, 2, 2) + M_(m, 3, 3); }

template <class S> static void matmul3(const S* a, const S* b, S* c) {  // c = a*b
  for (int j = 1; j <= 3; ++j)
    for (int i = 1; i <= 3; ++i) {
      S s = M_(a, i, 1) * M_(b, 1, j);
      s = s + M_(a, i, 2) * M_(b, 2, j);
      s = s + M_(a, i, 3) * M_(b, 3, j);
      M_(c, i, j) = s;
    }
}

// LinearAlgebra.det == LU with partial pivoting (LAPACK getrf for Float64, generic_lufact!
// for duals); call sites HyperelasticityMPh.jl:114,152,187,412-415, EquationsOfState.jl:180,228.
template <class S> static S det_lu(const S* m_in) {
  S a[9];
  for (int i = 0; i < 9; ++i) a[i] = m_in[i];
  double sign = 1.0;
  for (int k = 1; k <= 3; ++k) {
    int p = k;
    double best = std::fabs(value_of(M_(a, k, k)));
    for (int i = k + 1; i <= 3; ++i) {
      double c = std::fabs(value_of(M_(a, i, k)));
      if (c > best) { best = c; p = i; }
    }
    if (best == 0.0) return S(0.0);
    if (p != k) {
      for (int j = 1; j <= 3; ++j) std::swap(M_(a, k, j), M_(a, p, j));
      sign = -sign;
    }
    for (int i = k + 1; i <= 3; ++i) {
      M_(a, i, k) = M_(a, i, k) / M_(a, k, k);
      for (int j = k + 1; j <= 3; ++j) M_(a, i, j) = M_(a, i, j) - M_(a, i, k) * M_(a, k, j);
    }
  }
  S d = M_(a, 1, 1) * M_(a, 2, 2) * M_(a, 3, 3);
  return d * sign;
}

// ------------------------------------------------------------------------------------
// Strains.jl
// ------------------------------------------------------------------------------------
// Strains.jl:26-32   finger(a) = inv_(a * a')[:]
template <class S> static void finger(const S* a, S* g) {
  S at[9], b[9];
  for (int i = 1; i <= 3; ++i)
    for (int j = 1; j <= 3; ++j) M_(at, i, j) = M_(a, j, i);
  matmul3(a, at, b);
  inv_(b, g);
}

// Strains.jl:46-52   [tr g, 0.5 (tr(g)^2 - tr(g^2)), det g]
template <class S> static void invariants(const S* g, S* inv3) {
  S g2[9];
  matmul3(g, g, g2);
  S i1 = tr_(g);
  S t = tr_(g);
  S i2 = 0.5 * (t * t - tr_(g2));
  S i3 = det_(g);
  inv3[0] = i1; inv3[1] = i2; inv3[2] = i3;
}

// ------------------------------------------------------------------------------------
// EquationsOfState.jl -- Barton2009
// ------------------------------------------------------------------------------------
// EquationsOfState.jl:118-137.  SS is the scalar type of the entropy argument, S of G.
template <class SS, class S> static auto energy(const Eos& eos, const SS& ent, const S* G) {
  S i[3];
  invariants(G, i);
  S a = d_pow(i[2], 0.5 * eos.alpha) - 1.0;
  auto U = (0.5 * eos.k0 / (eos.alpha * eos.alpha)) * (a * a) +
           (eos.cv * eos.t0) * d_pow(i[2], 0.5 * eos.gamma) * (d_exp(ent / eos.cv) - 1.0);
  S W = (0.5 * eos.b0sq) * d_pow(i[2], 0.5 * eos.beta) * (i[0] * i[0] / 3.0 - i[1]);
  return U + W;
}

// EquationsOfState.jl:139-156 (clamp at 1e-6 is a real branch)
template <class S> static S entropy(const Eos& eos, const S& e_int, const S* G) {
  S i[3];
  invariants(G, i);
  S a = d_pow(i[2], 0.5 * eos.alpha) - 1.0;
  S s = e_int - (0.5 * eos.b0sq) * d_pow(i[2], 0.5 * eos.beta) * (i[0] * i[0] / 3.0 - i[1]) -
        (0.5 * eos.k0 / (eos.alpha * eos.alpha)) * (a * a);
  s = (s / ((eos.cv * eos.t0) * d_pow(i[2], 0.5 * eos.gamma)) + 1.0);
  if (value_of(s) != value_of(s)) g_domain_error = 1;  // NaN: log would throw upstream
  if (s < 1e-6) s = S(1e-6);
  return d_log(s) * eos.cv;
}

// EquationsOfState.jl:179-190  stress(eos, ent, F): gradient over the 9 entries of G.
template <class S> static void stress(const Eos& eos, const S& ent, const S* F, S* sig) {
  S den = eos.rho0 / det_lu(F);
  S G[9];
  finger(F, G);
  typedef Dual<S, 9> D9;
  D9 Gd[9];
  for (int i = 0; i < 9; ++i) { Gd[i] = D9(G[i]); Gd[i].d[i] = S(1.0); }
  D9 e = energy(eos, ent, Gd);  // ent is constant at this level
  S dedG[9];
  for (int i = 0; i < 9; ++i) dedG[i] = e.d[i];
  S GdedG[9];
  matmul3(G, dedG, GdedG);
  for (int i = 0; i < 9; ++i) sig[i] = (-2.0 * den) * GdedG[i];
}

// EquationsOfState.jl:223-246  acoustic(eos, ent, F, n): jacobian of stress wrt F (nested
// duals), then the 5-deep contraction with A[m,i,j,l] = (1/den) dsigma_{mi}/dF_{jl}.
static void acoustic(const Eos& eos, double ent, const double* F, const double* n, double* ac) {
  typedef Dual<double, 9> D9;
  D9 Fd[9], sig[9];
  for (int i = 0; i < 9; ++i) { Fd[i] = D9(F[i]); Fd[i].d[i] = 1.0; }
  stress(eos, D9(ent), Fd, sig);
  double den = eos.rho0 / det_lu(F);
  // dTdF[(m,i),(j,l)] = sig[m + 3 i].d[j + 3 l]  (both column-major, 0-based here)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l)
          for (int m = 0; m < 3; ++m) {
            double A = (1.0 / den) * sig[m + 3 * i].d[j + 3 * l];
            acc += A * F[k + 3 * l] * n[m] * n[k];
          }
      ac[i + 3 * j] = acc;
    }
}

// ------------------------------------------------------------------------------------
// EquationsOfState.jl -- Hank2016 (:301-364).  Dead code in the reference (never called; `energy` and
// `stress` cannot run as written: they pass a 3x3 Matrix to the Vector-only `invariants` / `finger`,
// Strains.jl:26,46).  Restated with a 3x3 tensor == its 9 column-major entries, which is the only reading the
// arithmetic allows; `pressure` (:333-346) runs as written.  Parity unpinned like the rest.
// ------------------------------------------------------------------------------------
struct Hank { double rho0, mu, gamma, pres_inf, a; };

// EquationsOfState.jl:326-328 == :341-343
template <class S> static S hank_e_el(const Hank& eos, const S* i) {
  S j1 = i[0] / d_pow(i[2], 1.0 / 3);
  S j2 = (i[0] * i[0] - 2.0 * i[1]) / d_pow(i[2], 2.0 / 3);
  return (eos.mu / (4 * eos.rho0)) * (((1 - 2 * eos.a) / 3) * (j1 * j1) + eos.a * j2 + 3 * (eos.a - 1));
}
// EquationsOfState.jl:317-331
template <class S> static S hank_energy(const Hank& eos, double den, double pres, const S* G) {
  S i[3];
  invariants(G, i);
  if (!(value_of(i[2]) > 0.0)) g_domain_error = 1;  // negative base, fractional exponent -> DomainError
  S e_el = hank_e_el(eos, i);
  double e_h = (pres + eos.gamma * eos.pres_inf) / (den * (eos.gamma - 1));
  return e_el + e_h;
}
// EquationsOfState.jl:333-346
static double hank_pressure(const Hank& eos, double den, double e_int, const double* i) {
  if (!(i[2] > 0.0)) g_domain_error = 1;
  double e_el = hank_e_el(eos, i);
  double e_h = e_int - e_el;
  return e_h * (eos.gamma - 1) * den - eos.gamma * eos.pres_inf;
}
// EquationsOfState.jl:348-356: G = finger(inv(distortion)); gradient of the energy over the 9 entries of G
static void hank_stress(const Hank& eos, double den, double pres, const double* A, double* sig) {
  double Ai[9], G[9];
  inv_(A, Ai);        // LinearAlgebra.inv: the inverse is unique, cofactor form used here
  finger(Ai, G);
  typedef Dual<double, 9> D9;
  D9 Gd[9];
  for (int i = 0; i < 9; ++i) { Gd[i] = D9(G[i]); Gd[i].d[i] = 1.0; }
  D9 e = hank_energy(eos, den, pres, Gd);
  double dedG[9], GdedG[9];
  for (int i = 0; i < 9; ++i) dedG[i] = e.d[i];
  matmul3(G, dedG, GdedG);
  for (int i = 0; i < 9; ++i) sig[i] = (-2.0 * den) * GdedG[i];
}

// eigvals(ac) (HyperelasticityMPh.jl:263): cyclic Jacobi on (ac+ac')/2, ascending order
// (Julia sorts eigvals of a general real matrix by (real, imag)).
static void eigvals_sym3(const double* ac, double* ev) {
  double a[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (ac[i + 3 * j] + ac[j + 3 * i]);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-40 * dia || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J' A
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
      }
  }
  ev[0] = a[0][0]; ev[1] = a[1][1]; ev[2] = a[2][2];
  std::sort(ev, ev + 3);
}

// ------------------------------------------------------------------------------------
// HyperelasticityMPh.jl -- per-phase pieces (15 variables per phase, F column-major)
// ------------------------------------------------------------------------------------
struct PhaseRec {  // the quantities cons2prim/flux/noncons_flux all recompute identically
  double frac, true_den, den, vel[3], e_total, e_int, F[9];
};

// HyperelasticityMPh.jl:109-121 == :147-159 == :181-194
static void recover(const Eos& eos, const double* Q, PhaseRec& r) {
  r.frac = Q[0];
  double FQ[9];
  for (int i = 0; i < 9; ++i) FQ[i] = Q[6 + i] / r.frac;
  double x = det_lu(FQ) / eos.rho0;
  if (!(x >= 0.0)) g_domain_error = 1;  // sqrt(negative) -> DomainError in Julia
  r.true_den = std::sqrt(x);
  r.den = r.frac * r.true_den;
  for (int i = 0; i < 3; ++i) r.vel[i] = Q[2 + i] / r.den;
  r.e_total = Q[5] / r.den;
  double e_kin = ((r.vel[0] * r.vel[0] + r.vel[1] * r.vel[1]) + r.vel[2] * r.vel[2]) / 2;
  r.e_int = r.e_total - e_kin;
  for (int i = 0; i < 9; ++i) r.F[i] = Q[6 + i] / r.den;
}

// HyperelasticityMPh.jl:106-133
static void cons2prim(const Eos& eos, const double* Q, double* P) {
  PhaseRec r;
  recover(eos, Q, r);
  double G[9];
  finger(r.F, G);
  double ent = entropy(eos, r.e_int, G);
  P[0] = r.frac; P[1] = r.true_den;
  for (int i = 0; i < 3; ++i) P[2 + i] = r.vel[i];
  P[5] = ent;
  for (int i = 0; i < 9; ++i) P[6 + i] = r.F[i];
}

// HyperelasticityMPh.jl:66-87
static void prim2cons(const Eos& eos, const double* P, double* Q) {
  double frac = P[0], true_den = P[1], den = frac * true_den;
  const double* vel = P + 2;
  double ent = P[5];
  const double* F = P + 6;
  double G[9];
  finger(F, G);
  double e_int = energy(eos, ent, G);
  double e_kin = ((vel[0] * vel[0] + vel[1] * vel[1]) + vel[2] * vel[2]) / 2;
  double e_total = e_int + e_kin;
  Q[0] = frac; Q[1] = den;
  for (int i = 0; i < 3; ++i) Q[2 + i] = den * vel[i];
  Q[5] = den * e_total;
  for (int i = 0; i < 9; ++i) Q[6 + i] = den * F[i];
}

// HyperelasticityMPh.jl:146-175
static void flux(const Eos& eos, const double* Q, double* f) {
  PhaseRec r;
  recover(eos, Q, r);
  double G[9], sig[9], strs[9];
  finger(r.F, G);
  double ent = entropy(eos, r.e_int, G);
  stress(eos, ent, r.F, sig);
  for (int i = 0; i < 9; ++i) strs[i] = r.frac * sig[i];
  const double s1[3] = {strs[0], strs[3], strs[6]};  // strs[begin:3:end] = row 1
  const double F1[3] = {r.F[0], r.F[3], r.F[6]};     // def_grad[begin:3:end] = row 1
  f[0] = 0;
  f[1] = r.den * r.vel[0];
  for (int i = 0; i < 3; ++i) f[2 + i] = r.den * r.vel[0] * r.vel[i] - s1[i];
  f[5] = r.den * r.vel[0] * r.e_total - ((r.vel[0] * s1[0] + r.vel[1] * s1[1]) + r.vel[2] * s1[2]);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i)
      f[6 + i + 3 * j] = r.den * (r.vel[0] * r.F[i + 3 * j] - r.vel[i] * F1[j]);
}

// HyperelasticityMPh.jl:258-266 (single phase).  Returns [spd+c_k (k asc), spd-c_k].
static void get_eigvals_phase(const Eos& eos, const double* Q, const double* n, double* out6) {
  double P[15], ac[9], ev[3];
  cons2prim(eos, Q, P);
  acoustic(eos, P[5], P + 6, n, ac);
  eigvals_sym3(ac, ev);
  double spd = P[2] * n[0] + P[3] * n[1] + P[4] * n[2];
  for (int k = 0; k < 3; ++k) {
    double c = std::sqrt(std::fabs(ev[k]));
    out6[k] = spd + c;
    out6[3 + k] = spd - c;
  }
}

// ------------------------------------------------------------------------------------
// Multiphase wrappers (nph phases of 15)
// ------------------------------------------------------------------------------------
static void cons2prim_mph(const Eos* eos, int nph, const double* Q, double* P) {  // :99-104
  for (int p = 0; p < nph; ++p) cons2prim(eos[p], Q + 15 * p, P + 15 * p);
}
static void prim2cons_mph(const Eos* eos, int nph, const double* P, double* Q) {  // :63-90
  for (int p = 0; p < nph; ++p) prim2cons(eos[p], P + 15 * p, Q + 15 * p);
}
static void flux_mph(const Eos* eos, int nph, const double* Q, double* f) {  // :140-144
  for (int p = 0; p < nph; ++p) flux(eos[p], Q + 15 * p, f + 15 * p);
}
static void get_eigvals(const Eos* eos, int nph, const double* Q, const double* n, double* out) {  // :252-256
  for (int p = 0; p < nph; ++p) get_eigvals_phase(eos[p], Q + 15 * p, n, out + 6 * p);
}

// HyperelasticityMPh.jl:178-250.  The 30x30 result is block-diagonal and in each 15x15 block
// only column 1 is ever written (:223-230), so the oracle returns the two columns
// (col[15*p + r] = B[p][r,1]); hso_noncons_flux() expands them to the dense matrix.
static void noncons_cols(const Eos* eos, const double* Qin, double* col) {
  const int nph = 2;
  PhaseRec r[2];
  double G[2][9], ent[2], strs[2][9], temp[2];
  for (int p = 0; p < nph; ++p) {
    recover(eos[p], Qin + 15 * p, r[p]);
    finger(r[p].F, G[p]);
    ent[p] = entropy(eos[p], r[p].e_int, G[p]);
    double sig[9];
    stress(eos[p], ent[p], r[p].F, sig);
    for (int i = 0; i < 9; ++i) strs[p][i] = r[p].frac * sig[i];
    // :212 temperature = d energy / d S
    Dual<double, 1> Sd(ent[p]);
    Sd.d[0] = 1.0;
    Dual<double, 1> e = energy(eos[p], Sd, G[p]);
    temp[p] = e.d[0];
  }
  const double omega = 0;
  double k[2];
  k[0] = 1.0 / 2;
  k[1] = 1 - k[0];
  const double beta[2] = {0.0, 0.0};
  double vel_i[3];
  for (int i = 0; i < 3; ++i) vel_i[i] = k[0] * r[0].vel[i] + k[1] * r[1].vel[i];  // :213
  double K[2][9];
  for (int p = 0; p < nph; ++p) {  // :216
    double trs = (strs[p][0] + strs[p][4]) + strs[p][8];
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) {
        double eye = (i == j) ? 1.0 : 0.0;
        K[p][i + 3 * j] = 1 / r[p].frac * (omega / 3 * trs * eye + (1 - omega) * strs[p][i + 3 * j]) + beta[p] * eye;
      }
  }
  double strs_i[9];
  double denom = k[0] * temp[0] + k[1] * temp[1];
  for (int i = 0; i < 9; ++i)  // :217
    strs_i[i] = (k[1] * temp[1] * K[0][i] + k[0] * temp[0] * K[1][i]) / denom;

  for (int p = 0; p < nph; ++p) {
    double* c = col + 15 * p;
    for (int i = 0; i < 15; ++i) c[i] = 0.0;
    c[0] = vel_i[0];                                            // :223
    for (int i = 0; i < 3; ++i) c[2 + i] = strs_i[i];           // :224  strs_i[:,1]
    c[5] = (strs_i[0] * vel_i[0] + strs_i[1] * vel_i[1]) + strs_i[2] * vel_i[2];  // :225
    for (int i = 0; i <= 6; i += 3)                              // :227-229
      for (int q = 0; q < 3; ++q)
        c[6 + i + q] = omega * r[p].true_den / 3 * (vel_i[0] - r[p].vel[0]) * r[p].F[i + q] +
                       r[p].true_den * r[p].F[i] * r[p].vel[q];
    // :230   B[7:3:15,1] += (1-omega) rho F' (vel_i - vel)
    double dv[3] = {vel_i[0] - r[p].vel[0], vel_i[1] - r[p].vel[1], vel_i[2] - r[p].vel[2]};
    for (int j = 0; j < 3; ++j) {
      const double sc = (1 - omega) * r[p].true_den;  // scaled matrix first, then mat-vec
      double s = ((sc * r[p].F[0 + 3 * j]) * dv[0] + (sc * r[p].F[1 + 3 * j]) * dv[1]) + (sc * r[p].F[2 + 3 * j]) * dv[2];
      c[6 + 3 * j] += s;
    }
  }
}

// ------------------------------------------------------------------------------------
// NumFluxes.jl
// ------------------------------------------------------------------------------------
// FastGaussQuadrature.gausslegendre(6) / gausslobatto(6) on [-1,1] (correctly rounded),
// mapped to [0,1] exactly as NumFluxes.jl:95 / :37 do.
static const double GLEG_X[6] = {-0.9324695142031520278, -0.6612093864662645137, -0.2386191860831969086,
                                 0.2386191860831969086,  0.6612093864662645137,  0.9324695142031520278};
static const double GLEG_W[6] = {0.1713244923791703450, 0.3607615730481386076, 0.4679139345726910474,
                                 0.4679139345726910474, 0.3607615730481386076, 0.1713244923791703450};
static const double GLOB_X[6] = {-1.0, -0.7650553239294646929, -0.2852315164806450963,
                                 0.2852315164806450963, 0.7650553239294646929, 1.0};
static const double GLOB_W[6] = {0.06666666666666666667, 0.3784749562978469803, 0.5548583770354863530,
                                 0.5548583770354863530,  0.3784749562978469803, 0.06666666666666666667};

// sum_i w_i * B(psi(s_i)) * dpsi/ds, psi = Q_l (1-s) + Q_r s   (NumFluxes.jl:97-107, :39-47)
// B*d restricted to block p is col_p * d[15p] (all other columns are exact zeros).
static void B_int(const Eos* eos, const double* Ql, const double* Qr, const double* xs, const double* ws,
                  double* out) {
  double acc[30];
  for (int q = 0; q < 6; ++q) {
    double s = (xs[q] + 1.0) / 2.0, w = ws[q] / 2.0;
    double u[30], d[30], col[30];
    for (int i = 0; i < 30; ++i) {
      u[i] = Ql[i] * (1 - s) + Qr[i] * s;  // path, :71 / :28
      d[i] = -Ql[i] + Qr[i];               // ForwardDiff.derivative of the path, :100 / :41
    }
    noncons_cols(eos, u, col);
    for (int p = 0; p < 2; ++p)
      for (int rr = 0; rr < 15; ++rr) {
        double term = (w * col[15 * p + rr]) * d[15 * p];  // (weight*B)*dvals
        acc[15 * p + rr] = (q == 0) ? term : acc[15 * p + rr] + term;
      }
  }
  for (int i = 0; i < 30; ++i) out[i] = acc[i];
}

static double min6n(const double* e, int n) { double m = e[0]; for (int i = 1; i < n; ++i) m = std::min(m, e[i]); return m; }
static double max6n(const double* e, int n) { double m = e[0]; for (int i = 1; i < n; ++i) m = std::max(m, e[i]); return m; }

// NumFluxes.jl:85-132.  eig_l / eig_r: the 12 cached speeds of the left / right cell.
static void hll_pathcons(const Eos* eos, const double* Ql, const double* Qr, const double* eig_l,
                         const double* eig_r, double* dm, double* dp, double* sl_out, double* sr_out) {
  const double n[3] = {1, 0, 0};
  double Qm[30], em[12];
  for (int i = 0; i < 30; ++i) Qm[i] = 0.5 * (Ql[i] + Qr[i]);
  get_eigvals(eos, 2, Qm, n, em);  // reference evaluates this twice (:90,:91); it is pure
  double s_l = std::min(0.0, std::min(min6n(em, 12), min6n(eig_l, 12)));
  double s_r = std::max(0.0, std::max(max6n(em, 12), max6n(eig_r, 12)));
  double Fl[30], Fr[30], b0[30], path_int[30], Qh[30], b1[30], b2[30];
  flux_mph(eos, 2, Ql, Fl);
  flux_mph(eos, 2, Qr, Fr);
  B_int(eos, Ql, Qr, GLEG_X, GLEG_W, b0);
  for (int i = 0; i < 30; ++i) path_int[i] = (b0[i] + Fr[i]) - Fl[i];                        // :109
  for (int i = 0; i < 30; ++i) Qh[i] = ((Qr[i] * s_r - Ql[i] * s_l) - path_int[i]) / (s_r - s_l);  // :111
  B_int(eos, Ql, Qh, GLEG_X, GLEG_W, b1);
  B_int(eos, Qh, Qr, GLEG_X, GLEG_W, b2);
  for (int i = 0; i < 30; ++i) {  // :128-129
    double br = ((Fr[i] - Fl[i]) + b1[i]) + b2[i];
    dm[i] = -s_l / (s_r - s_l) * br + s_l * s_r / (s_r - s_l) * (Qr[i] - Ql[i]);
    dp[i] = s_r / (s_r - s_l) * br - s_l * s_r / (s_r - s_l) * (Qr[i] - Ql[i]);
  }
  if (sl_out) *sl_out = s_l;
  if (sr_out) *sr_out = s_r;
}

// NumFluxes.jl:25-60
static void lxf(const Eos* eos, const double* Ql, const double* Qr, double lambda, double* cons,
                double* dm, double* dp) {
  double Fl[30], Fr[30], fp[30];
  flux_mph(eos, 2, Ql, Fl);
  flux_mph(eos, 2, Qr, Fr);
  for (int i = 0; i < 30; ++i) cons[i] = 0.5 * (Fl[i] + Fr[i]) - 0.5 * lambda * (Qr[i] - Ql[i]);  // :30
  B_int(eos, Ql, Qr, GLOB_X, GLOB_W, fp);
  for (int i = 0; i < 30; ++i) { dp[i] = (1.0 / 2.0) * fp[i]; dm[i] = (1.0 / 2.0) * fp[i]; }       // :50-51
}

// ------------------------------------------------------------------------------------
// Single-phase (13-variable) configuration -- SURVEY.md A.6.  The reference module
// Hyperelasticity.jl is stale and cannot run; layout / recovery / flux follow it, the EoS
// calls follow the shipped MPh variants (stress(eos,ent,F), acoustic(eos,ent,F,n)).
//   Q = [rho u (3), rho F (9, ROW-major: Q[3i+j] = rho F_ij, 1-based), rho E]
// ------------------------------------------------------------------------------------
struct SpRec { double den, vel[3], F[9] /*column-major*/, E, e_int; };

// Hyperelasticity.jl:25-36 + EquationsOfState.jl:259-264 (density = sqrt(det(rho F)/rho0))
static void sp_recover(const Eos& eos, const double* Q, SpRec& r) {
  double FQ[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) FQ[i + 3 * j] = Q[3 + 3 * i + j];
  double x = det_lu(FQ) / eos.rho0;
  if (!(x >= 0.0)) g_domain_error = 1;
  r.den = std::sqrt(x);
  for (int i = 0; i < 3; ++i) r.vel[i] = Q[i] / r.den;
  for (int i = 0; i < 9; ++i) r.F[i] = FQ[i] / r.den;
  r.E = Q[12] / r.den;
  double e_kin = 0.5 * ((r.vel[0] * r.vel[0] + r.vel[1] * r.vel[1]) + r.vel[2] * r.vel[2]);
  r.e_int = r.E - e_kin;
}

// Hyperelasticity.jl:99-114 with sigma = stress(eos, entropy, F)
static void sp_flux(const Eos& eos, const double* Q, double* f) {
  SpRec r;
  sp_recover(eos, Q, r);
  double G[9], sig[9];
  finger(r.F, G);
  double ent = entropy(eos, r.e_int, G);
  stress(eos, ent, r.F, sig);
  for (int i = 0; i < 3; ++i) {
    f[i] = r.den * r.vel[0] * r.vel[i] - sig[0 + 3 * i];                                  // sigma[1,i]
    f[i + 3] = 0;
    f[i + 6] = r.den * (r.F[1 + 3 * i] * r.vel[0] - r.F[0 + 3 * i] * r.vel[1]);
    f[i + 9] = r.den * (r.F[2 + 3 * i] * r.vel[0] - r.F[0 + 3 * i] * r.vel[2]);
  }
  f[12] = r.den * r.vel[0] * r.E - r.vel[0] * sig[0] - r.vel[1] * sig[3] - r.vel[2] * sig[6];
}

// primitive vector P13 = [rho, u(3), S, F(9, column-major like the MPh CSV)] is not part of the
// reference; SP prims are exposed as [u(3), F(9 row-major), S] = the arguments of
// Hyperelasticity.jl:70 prim2cons(eos, vel, F, S).
static void sp_prim2cons(const Eos& eos, const double* P, double* Q) {  // Hyperelasticity.jl:70-93
  const double* vel = P;
  double Fc[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Fc[i + 3 * j] = P[3 + 3 * i + j];
  double S = P[12];
  double den = eos.rho0 / det_lu(Fc);
  for (int i = 0; i < 3; ++i) Q[i] = den * vel[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Q[3 + 3 * i + j] = den * Fc[i + 3 * j];
  double e_kin = 0.5 * ((vel[0] * vel[0] + vel[1] * vel[1]) + vel[2] * vel[2]);
  double G[9];
  finger(Fc, G);
  double E = energy(eos, S, G) + e_kin;
  Q[12] = den * E;
}

static void sp_cons2prim(const Eos& eos, const double* Q, double* P) {
  SpRec r;
  sp_recover(eos, Q, r);
  double G[9];
  finger(r.F, G);
  for (int i = 0; i < 3; ++i) P[i] = r.vel[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) P[3 + 3 * i + j] = r.F[i + 3 * j];
  P[12] = entropy(eos, r.e_int, G);
}

static void sp_get_eigvals(const Eos& eos, const double* Q, double* out6) {
  SpRec r;
  sp_recover(eos, Q, r);
  double G[9], ac[9], ev[3];
  finger(r.F, G);
  double ent = entropy(eos, r.e_int, G);
  const double n[3] = {1, 0, 0};
  acoustic(eos, ent, r.F, n, ac);
  eigvals_sym3(ac, ev);
  for (int k = 0; k < 3; ++k) {
    double c = std::sqrt(std::fabs(ev[k]));
    out6[k] = r.vel[0] + c;
    out6[3 + k] = r.vel[0] - c;
  }
}

// conservative HLL flux, NumFluxes.jl:73-78 (commented one-phase form) with the cached cell
// speeds of :90-91
static void sp_hll(const Eos& eos, const double* Ql, const double* Qr, const double* eig_l,
                   const double* eig_r, double* cons) {
  double Qm[13], em[6], Fl[13], Fr[13];
  for (int i = 0; i < 13; ++i) Qm[i] = 0.5 * (Ql[i] + Qr[i]);
  sp_get_eigvals(eos, Qm, em);
  double s_l = std::min(0.0, std::min(min6n(em, 6), min6n(eig_l, 6)));
  double s_r = std::max(0.0, std::max(max6n(em, 6), max6n(eig_r, 6)));
  sp_flux(eos, Ql, Fl);
  sp_flux(eos, Qr, Fr);
  for (int i = 0; i < 13; ++i)
    cons[i] = (s_r * Fl[i] - s_l * Fr[i]) / (s_r - s_l) + s_l * s_r / (s_r - s_l) * (Qr[i] - Ql[i]);
}

static void sp_lxf(const Eos& eos, const double* Ql, const double* Qr, double lambda, double* cons) {
  double Fl[13], Fr[13];  // NumFluxes.jl:26
  sp_flux(eos, Ql, Fl);
  sp_flux(eos, Qr, Fr);
  for (int i = 0; i < 13; ++i) cons[i] = 0.5 * (Fl[i] + Fr[i]) - 0.5 * lambda * (Qr[i] - Ql[i]);
}

// ------------------------------------------------------------------------------------
// main.jl time loop
// ------------------------------------------------------------------------------------
template <class Fn> static void parallel_for(int64_t n, int nthreads, Fn fn) {  // Threads.@threads (main.jl:206,221)
  if (nthreads <= 1 || n < 2) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
  std::vector<std::thread> th;
  std::atomic<int> err(0);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      g_domain_error = 0;
      int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;  // static chunking
      for (int64_t i = lo; i < hi; ++i) fn(i);
      if (g_domain_error) err = 1;
    });
  for (auto& x : th) x.join();
  if (err) g_domain_error = 1;
}

}  // namespace hso

// =====================================================================================
// C ABI (AoS, one cell / face per column -- Julia's Array{Float64,2}(nvar, n))
// model: 0 = SP13, 1 = MPH30.   flux kind: 0 = LXF, 1 = HLL.
// =====================================================================================
using namespace hso;

extern "C" {

int hso_cons2prim(const double* eos, int model, const double* Q, double* P, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i)
    if (model == 1) cons2prim_mph(e, 2, Q + 30 * i, P + 30 * i); else sp_cons2prim(e[0], Q + 13 * i, P + 13 * i);
  return g_domain_error;
}

int hso_prim2cons(const double* eos, int model, const double* P, double* Q, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i)
    if (model == 1) prim2cons_mph(e, 2, P + 30 * i, Q + 30 * i); else sp_prim2cons(e[0], P + 13 * i, Q + 13 * i);
  return g_domain_error;
}

int hso_flux(const double* eos, int model, const double* Q, double* F, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i)
    if (model == 1) flux_mph(e, 2, Q + 30 * i, F + 30 * i); else sp_flux(e[0], Q + 13 * i, F + 13 * i);
  return g_domain_error;
}

// columns of the non-conservative matrix: col (30, n); dense B (30,30,n) if Bdense != NULL
int hso_noncons_flux(const double* eos, const double* Q, double* col, double* Bdense, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i) {
    double c[30];
    noncons_cols(e, Q + 30 * i, c);
    if (col) std::memcpy(col + 30 * i, c, sizeof(c));
    if (Bdense) {
      double* B = Bdense + 900 * i;  // column-major 30x30
      std::memset(B, 0, 900 * sizeof(double));
      for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 15; ++r) B[(15 * p + r) + 30 * (15 * p)] = c[15 * p + r];
    }
  }
  return g_domain_error;
}

// eig: (6*nph, n), per phase [u1+c_k (ascending c^2), u1-c_k]
int hso_get_eigvals(const double* eos, int model, const double* Q, double* eig, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  const double nn[3] = {1, 0, 0};
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i)
    if (model == 1) get_eigvals(e, 2, Q + 30 * i, nn, eig + 12 * i); else sp_get_eigvals(e[0], Q + 13 * i, eig + 6 * i);
  return g_domain_error;
}

// get_eigvals for an arbitrary normal n (two-phase), HyperelasticityMPh.jl:252-266
int hso_get_eigvals_n(const double* eos, const double* Q, const double* n, double* eig, int64_t cnt) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < cnt; ++i) get_eigvals(e, 2, Q + 30 * i, n, eig + 12 * i);
  return g_domain_error;
}

// individual EoS entry points (double), for unit tests
double hso_energy(const double* eos, double S, const double* G) { return energy(*reinterpret_cast<const Eos*>(eos), S, G); }
double hso_entropy(const double* eos, double e_int, const double* G) { return entropy(*reinterpret_cast<const Eos*>(eos), e_int, G); }
void hso_finger(const double* F, double* G) { finger(F, G); }
// Hank2016: per-item scalar calls; return 1 where Julia would throw DomainError
int hso_hank_energy(const double* eos, double den, double pres, const double* G, double* e) {
  g_domain_error = 0; *e = hank_energy(*reinterpret_cast<const Hank*>(eos), den, pres, G); return g_domain_error; }
int hso_hank_pressure(const double* eos, double den, double e_int, const double* inv3, double* p) {
  g_domain_error = 0; *p = hank_pressure(*reinterpret_cast<const Hank*>(eos), den, e_int, inv3); return g_domain_error; }
int hso_hank_stress(const double* eos, double den, double pres, const double* A, double* sig) {
  g_domain_error = 0; hank_stress(*reinterpret_cast<const Hank*>(eos), den, pres, A, sig); return g_domain_error; }
void hso_invariants(const double* G, double* i3) { invariants(G, i3); }
void hso_stress(const double* eos, double S, const double* F, double* sig) { stress(*reinterpret_cast<const Eos*>(eos), S, F, sig); }
void hso_acoustic(const double* eos, double S, const double* F, const double* n, double* ac) {
  acoustic(*reinterpret_cast<const Eos*>(eos), S, F, n, ac); }
double hso_temperature(const double* eos, double S, const double* G) {
  Dual<double, 1> Sd(S); Sd.d[0] = 1.0;
  Dual<double, 1> e = energy(*reinterpret_cast<const Eos*>(eos), Sd, G);
  return e.d[0];
}
void hso_quadrature(int lobatto, double* nodes, double* weights) {
  for (int q = 0; q < 6; ++q) {
    nodes[q] = ((lobatto ? GLOB_X[q] : GLEG_X[q]) + 1.0) / 2.0;
    weights[q] = (lobatto ? GLOB_W[q] : GLEG_W[q]) / 2.0;
  }
}

// hll over faces: Ql, Qr (30, n); eig_l, eig_r (12, n) cached speeds of the adjacent cells.
// Outputs cons (zeros, NumFluxes.jl:82), dm, dp (30, n); s (2, n) = [s_l, s_r].
int hso_hll(const double* eos, const double* Ql, const double* Qr, const double* eig_l, const double* eig_r,
            double* cons, double* dm, double* dp, double* s, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i) {
    double sl, sr;
    hll_pathcons(e, Ql + 30 * i, Qr + 30 * i, eig_l + 12 * i, eig_r + 12 * i, dm + 30 * i, dp + 30 * i, &sl, &sr);
    if (cons) std::memset(cons + 30 * i, 0, 30 * sizeof(double));
    if (s) { s[2 * i] = sl; s[2 * i + 1] = sr; }
  }
  return g_domain_error;
}

int hso_lxf(const double* eos, const double* Ql, const double* Qr, double lambda, double* cons, double* dm,
            double* dp, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i) lxf(e, Ql + 30 * i, Qr + 30 * i, lambda, cons + 30 * i, dm + 30 * i, dp + 30 * i);
  return g_domain_error;
}

int hso_sp_hll(const double* eos, const double* Ql, const double* Qr, const double* eig_l, const double* eig_r,
               double* cons, int64_t n) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  g_domain_error = 0;
  for (int64_t i = 0; i < n; ++i) sp_hll(e[0], Ql + 13 * i, Qr + 13 * i, eig_l + 6 * i, eig_r + 6 * i, cons + 13 * i);
  return g_domain_error;
}

// -------------------------------------------------------------------------------------
// The time loop, main.jl:202-227.  Q (nvar, ncells, nprob) is advanced in place.
//   literal != 0 : follow main.jl's structure exactly -- update_cell per cell (main.jl:43-60),
//                  i.e. every face evaluated twice (what the reference executes; used for the
//                  CPU baseline).
//   literal == 0 : every face evaluated once (all functions are pure, so results are
//                  bit-identical to literal mode; used to keep parity tests fast).
// Each problem of an ensemble has its own dt/t.  Stops a problem when t >= t_end (no clipping:
// t overshoots, main.jl:202,214) or after max_steps.  dt_hist (max_steps, nprob) optional.
// Returns 0, or 1 if a DomainError would have been thrown.
// -------------------------------------------------------------------------------------
int hso_run(const double* eos, int model, int fluxkind, double* Q, int64_t ncells, int64_t nprob, double cfl,
            double dx, double t_end, int64_t max_steps, double* t_io, int64_t* steps_io, double* dt_hist,
            int nthreads, int literal) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  const int nvar = model == 1 ? 30 : 13, neig = model == 1 ? 12 : 6;
  const double nn[3] = {1, 0, 0};
  int status = 0;
  std::vector<double> Q1((size_t)nvar * ncells), eig((size_t)neig * ncells), lam(ncells);
  std::vector<double> DM, DP;  // per-face results in dedup mode; face f is between cells f and f+1
  if (!literal) { DM.resize((size_t)nvar * ncells); DP.resize((size_t)nvar * ncells); }
  for (int64_t pr = 0; pr < nprob; ++pr) {
    double* Q0 = Q + (size_t)nvar * ncells * pr;
    double t = t_io ? t_io[pr] : 0.0;
    int64_t step = steps_io ? steps_io[pr] : 0, done = 0;
    while (t < t_end && done < max_steps) {
      g_domain_error = 0;
      // (A) CFL sweep, main.jl:204-212
      parallel_for(ncells, nthreads, [&](int64_t i) {
        if (model == 1) get_eigvals(e, 2, Q0 + nvar * i, nn, &eig[neig * i]); else sp_get_eigvals(e[0], Q0 + nvar * i, &eig[neig * i]);
        double m = 0.0;
        for (int k = 0; k < neig; ++k) m = std::max(m, std::fabs(eig[neig * i + k]));
        lam[i] = m;
      });
      double lmax = lam[0];
      for (int64_t i = 1; i < ncells; ++i) lmax = std::max(lmax, lam[i]);
      double dt = cfl * dx / lmax;
      t += dt;
      step += 1;
      // (C) update, main.jl:218-227
      for (int v = 0; v < nvar; ++v) { Q1[v] = Q0[v]; Q1[nvar * (ncells - 1) + v] = Q0[nvar * (ncells - 1) + v]; }
      const double dtdx = dt / dx, lambda = dx / dt;
      auto face = [&](int64_t f, double* cons, double* dm, double* dp) {  // face between cells f, f+1
        const double *ql = Q0 + nvar * f, *qr = Q0 + nvar * (f + 1);
        if (model == 1) {
          if (fluxkind == 1) { hll_pathcons(e, ql, qr, &eig[neig * f], &eig[neig * (f + 1)], dm, dp, nullptr, nullptr);
                               for (int v = 0; v < nvar; ++v) cons[v] = 0.0; }
          else lxf(e, ql, qr, lambda, cons, dm, dp);
        } else {
          if (fluxkind == 1) sp_hll(e[0], ql, qr, &eig[neig * f], &eig[neig * (f + 1)], cons); else sp_lxf(e[0], ql, qr, lambda, cons);
          for (int v = 0; v < nvar; ++v) { dm[v] = 0.0; dp[v] = 0.0; }
        }
      };
      if (literal) {
        parallel_for(ncells - 2, nthreads, [&](int64_t ii) {
          int64_t i = ii + 1;
          double Fl[30], Fr[30], dml[30], dpl[30], dmr[30], dpr[30];
          face(i - 1, Fl, dml, dpl);  // F_l, _, NF_l = flux_num(Q_l, Q)   main.jl:56
          face(i, Fr, dmr, dpr);      // F_r, NF_r, _ = flux_num(Q, Q_r)   main.jl:57
          for (int v = 0; v < nvar; ++v) {
            const double q = Q0[nvar * i + v];
            if (model == 1) Q1[nvar * i + v] = (fluxkind == 1) ? q - dtdx * ((Fr[v] - Fl[v]) + (dmr[v] + dpl[v]))      // main.jl:59
                                                               : q - 1.0 / lambda * ((Fr[v] - Fl[v]) + (dmr[v] + dpl[v]));  // main.jl:40
            else Q1[nvar * i + v] = (fluxkind == 1) ? q - dtdx * (Fr[v] - Fl[v]) : q - 1.0 / lambda * (Fr[v] - Fl[v]);     // main.jl:36
          }
        });
      } else {
        std::vector<double> CONS((size_t)nvar * ncells);
        parallel_for(ncells - 1, nthreads, [&](int64_t f) { face(f, &CONS[nvar * f], &DM[nvar * f], &DP[nvar * f]); });
        parallel_for(ncells - 2, nthreads, [&](int64_t ii) {
          int64_t i = ii + 1;
          for (int v = 0; v < nvar; ++v) {
            const double q = Q0[nvar * i + v];
            const double Fl = CONS[nvar * (i - 1) + v], Fr = CONS[nvar * i + v];
            const double dmr = DM[nvar * i + v], dpl = DP[nvar * (i - 1) + v];
            if (model == 1) Q1[nvar * i + v] = (fluxkind == 1) ? q - dtdx * ((Fr - Fl) + (dmr + dpl)) : q - 1.0 / lambda * ((Fr - Fl) + (dmr + dpl));
            else Q1[nvar * i + v] = (fluxkind == 1) ? q - dtdx * (Fr - Fl) : q - 1.0 / lambda * (Fr - Fl);
          }
        });
      }
      std::memcpy(Q0, Q1.data(), sizeof(double) * nvar * ncells);  // Q0 = copy(Q1), main.jl:227
      if (dt_hist) dt_hist[(size_t)max_steps * pr + done] = dt;
      done += 1;
      if (g_domain_error) { status = 1; break; }
    }
    if (t_io) t_io[pr] = t;
    if (steps_io) steps_io[pr] = step;
  }
  return status;
}

// CFL sweep only (main.jl:204-212): eig (neig, n) optional, returns max |eig|
double hso_lambda_max(const double* eos, int model, const double* Q, int64_t n, double* eig_out) {
  const Eos* e = reinterpret_cast<const Eos*>(eos);
  const int nvar = model == 1 ? 30 : 13, neig = model == 1 ? 12 : 6;
  const double nn[3] = {1, 0, 0};
  double lmax = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double eg[12];
    if (model == 1) get_eigvals(e, 2, Q + nvar * i, nn, eg); else sp_get_eigvals(e[0], Q + nvar * i, eg);
    for (int k = 0; k < neig; ++k) lmax = std::max(lmax, std::fabs(eg[k]));
    if (eig_out) std::memcpy(eig_out + neig * i, eg, sizeof(double) * neig);
  }
  return lmax;
}

int hso_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
