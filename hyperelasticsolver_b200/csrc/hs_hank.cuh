// Hank2016 equation of state (EquationsOfState.jl:301-364), closed form, FP64 -- SURVEY.md section 8 row f4.
//
// In the reference this material law is dead code: nothing calls it, and two of its three functions cannot
// run as written (`stress` hands a 3x3 Matrix to `finger`, Strains.jl:26, which only has a Vector method, and then a
// Vector to `energy(..., G::Array{<:Any,2})`; `energy` hands that Matrix to `invariants`, Strains.jl:46, again
// Vector-only).  Only `pressure` (EquationsOfState.jl:333-346, invariants in, scalar out) runs.  What is built here is
// therefore the law the code spells out, with the container mismatch resolved the only way the arithmetic allows
// (a 3x3 tensor is its 9 column-major entries):
//   j1 = I1 / I3^(1/3),  j2 = (I1^2 - 2 I2) / I3^(2/3) = tr(G^2) / I3^(2/3)                  (:326, :341)
//   e_el = mu/(4 rho0) ((1 - 2a)/3 j1^2 + a j2 + 3(a - 1))                                    (:328, :343)
//   energy(den, pres, G)      = e_el + (pres + gamma pres_inf) / (den (gamma - 1))             (:329-331)
//   pressure(den, e_int, I)   = (e_int - e_el)(gamma - 1) den - gamma pres_inf                 (:344-346)
//   stress(den, pres, A)      = -2 den G de/dG,  G = finger(inv(A)) = (A^-1 A^-T)^-1 = A^T A   (:349-356)
// Closed form of the gradient the reference takes with ForwardDiff (:353), all nine entries of G independent:
//   dI1/dG = 1,  d tr(G^2)/dG = 2 G^T,  dI3/dG = I3 G^-T, and for the symmetric G the product G G^-T = 1, so
//   sigma = -2 den mu/(4 rho0) [ (2(1-2a)/3) j1 I3^(-1/3) (G - (I1/3) 1) + a I3^(-2/3) (2 G^2 - (2/3) tr(G^2) 1) ]
// (trace-free: the hydrodynamic part of the energy does not depend on G).  I3^(1/3) is cbrt(); Julia's i3^(1/3)
// uses the double nearest to 1/3 as exponent, a relative difference of |ln I3| * 1.9e-17.
// A non-positive I3 is where Julia's `^` with a fractional exponent throws DomainError: flagged through `bad`.
// The header also compiles as plain C++ (tests/hostmath), like hs_phase.cuh.
#pragma once
#include <math.h>

#ifndef HS_HD
#ifdef __CUDACC__
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif
#endif

namespace hs {

// Hank2016 block exactly as the C ABI passes it (EquationsOfState.jl:305-310 field order)
struct HankAbi {
  double rho0, mu, gamma, pres_inf, a;
};

// elastic energy from (I1, tr G^2, I3); *iq = I3^(-1/3)
HS_HD double hank_e_el(const HankAbi& e, double I1, double T2, double I3, double* iq_out, int* bad) {
  if (!(I3 > 0.0)) *bad = 1;
  const double iq = 1.0 / cbrt(I3);
  const double j1 = I1 * iq, j2 = T2 * (iq * iq);
  *iq_out = iq;
  return e.mu / (4.0 * e.rho0) * ((1.0 - 2.0 * e.a) / 3.0 * (j1 * j1) + e.a * j2 + 3.0 * (e.a - 1.0));
}

// invariants of a 3x3 tensor given as 9 column-major entries (Strains.jl:46-52), with tr(G^2) in place of I2
HS_HD void hank_invariants(const double* G, double* I1, double* T2, double* I3) {
  *I1 = G[0] + G[4] + G[8];
  // tr(G G) = sum_ij G_ij G_ji
  *T2 = G[0] * G[0] + G[4] * G[4] + G[8] * G[8] + 2.0 * (G[1] * G[3] + G[2] * G[6] + G[5] * G[7]);
  *I3 = G[0] * (G[4] * G[8] - G[5] * G[7]) - G[3] * (G[1] * G[8] - G[7] * G[2]) + G[6] * (G[1] * G[5] - G[4] * G[2]);
}

// energy(eos::Hank2016, den, pres, G)  EquationsOfState.jl:317-331
HS_HD double hank_energy(const HankAbi& e, double den, double pres, const double* G, int* bad) {
  double I1, T2, I3, iq;
  hank_invariants(G, &I1, &T2, &I3);
  const double e_el = hank_e_el(e, I1, T2, I3, &iq, bad);
  return e_el + (pres + e.gamma * e.pres_inf) / (den * (e.gamma - 1.0));
}

// pressure(eos::Hank2016, den, e_int, i)  EquationsOfState.jl:333-346; inv3 = [I1, I2, I3] of Strains.jl:46-52
HS_HD double hank_pressure(const HankAbi& e, double den, double e_int, const double* inv3, int* bad) {
  double iq;
  const double e_el = hank_e_el(e, inv3[0], inv3[0] * inv3[0] - 2.0 * inv3[1], inv3[2], &iq, bad);
  return (e_int - e_el) * (e.gamma - 1.0) * den - e.gamma * e.pres_inf;
}

// stress(eos::Hank2016, den, pressure, distortion)  EquationsOfState.jl:348-356; A, sigma: 9 column-major entries.
// (The pressure argument does not enter: the hydrodynamic energy has no G dependence, its gradient is zero.)
HS_HD void hank_stress(const HankAbi& e, double den, const double* A, double* sig, int* bad) {
  // G = A^T A (symmetric): G_ij = sum_k A_ki A_kj, A_ki = A[k + 3 i]
  double g[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      const double s = A[3 * i] * A[3 * j] + A[1 + 3 * i] * A[1 + 3 * j] + A[2 + 3 * i] * A[2 + 3 * j];
      g[i][j] = s; g[j][i] = s;
    }
  const double I1 = g[0][0] + g[1][1] + g[2][2];
  const double detA = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[3] * (A[1] * A[8] - A[7] * A[2]) + A[6] * (A[1] * A[5] - A[4] * A[2]);
  const double I3 = detA * detA;
  if (!(I3 > 0.0)) *bad = 1;
  double g2[3][3], T2 = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      const double s = g[i][0] * g[0][j] + g[i][1] * g[1][j] + g[i][2] * g[2][j];
      g2[i][j] = s; g2[j][i] = s;
    }
  T2 = g2[0][0] + g2[1][1] + g2[2][2];
  const double iq = 1.0 / cbrt(I3);
  const double j1 = I1 * iq;
  const double c = -2.0 * den * e.mu / (4.0 * e.rho0);
  const double k1 = c * (2.0 * (1.0 - 2.0 * e.a) / 3.0) * j1 * iq;
  const double k2 = c * e.a * (iq * iq);
  const double iso = k1 * (I1 / 3.0) + k2 * (2.0 / 3.0) * T2;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) sig[i + 3 * j] = k1 * g[i][j] + 2.0 * k2 * g2[i][j] - (i == j ? iso : 0.0);
}

}  // namespace hs
