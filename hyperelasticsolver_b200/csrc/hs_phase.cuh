// Per-phase constitutive math of the hot path, closed form, FP64, branch-light.
//
// One "phase record" is the 14 conserved numbers the physics reads
//   alpha, m = alpha*rho*u (3), E = alpha*rho*E_tot, A = alpha*rho*F (9, column-major A[i+3j])
// (Q[2] = alpha*rho is evolved but never read: HyperelasticityMPh.jl:110,148,183).  The
// single-phase model is the same record with alpha == 1.
//
// What each routine replaces in the reference (all of it evaluated there through nested
// ForwardDiff duals and heap-allocated 3x3 matrices):
//   phase_state  : cons2prim HyperelasticityMPh.jl:106-133 (rho from det, :113-114),
//                  finger Strains.jl:26-32, invariants Strains.jl:46-52,
//                  entropy EquationsOfState.jl:139-156 (kept as S' = exp(S/cv), clamp included),
//                  stress  EquationsOfState.jl:179-190 (row 1 only; closed form of the gradient),
//                  temperature = derivative(energy, S) HyperelasticityMPh.jl:212
//   phase_flux   : flux HyperelasticityMPh.jl:146-175
//   phase_cmax2  : acoustic EquationsOfState.jl:223-246 for n = (1,0,0) (closed form of the
//                  nested jacobian) + eigvals HyperelasticityMPh.jl:263 (largest |eigenvalue|)
// Exact identities used (valid for every state, not only consistent ones):
//   rho^2 = det(A)/(alpha^3 rho0)  =>  det F = rho0/rho,  I3 = det G = (rho/rho0)^2,
//   so I3^(x/2) = (rho/rho0)^x: no pow/exp/log on the default exponents (alpha,beta,gamma)=(1,3,2),
//   one log + three exp otherwise (GEN = true).
// The header also compiles as plain C++ (tests/hostmath) so the closed forms can be checked
// against the dual-number oracle on a machine without a GPU.  It is never a product CPU path.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define HS_HD __host__ __device__ __forceinline__
#define HS_HD_COLD __host__ __device__ __noinline__
#else
#define HS_HD inline
#define HS_HD_COLD inline
#endif

namespace hs {

// FP64 literals whose low word is not zero cannot be immediates of a DFMA / DMUL: the compiler builds each of them in
// a register pair with two moves at every use (it will not keep them live at 128 registers).  As constant-bank
// operands they cost nothing.
#ifdef __CUDACC__
struct HotLits { double third, sixth, th_clamp, sp_clamp, tiny, r_hi, r_lo, mu0, mu1, mu2, mu3; };
__constant__ HotLits c_lit = {1.0 / 3.0, 1.0 / 6.0, 1e-6 - 1.0, 1e-6, 1e-280, 1.0000001, -0.875,
                              1.7464452327513027, 0.32800957660022223, -0.1425897443947186, 0.08116787571408166};
#endif
#ifdef __CUDA_ARCH__
#define HS_LIT(field, value) (c_lit.field)
#else
#define HS_LIT(field, value) (value)
#endif

// Barton2009 block exactly as the C ABI passes it (EquationsOfState.jl:71-85 field order) ...
struct EosAbi {
  double rho0, c0, cv, t0, b0, alpha, beta, gamma, b0sq, k0;
};
// ... and the derived constants the kernels use (computed once on the host).
struct EosDev {
  double rho0, inv_rho0, cvt0, inv_cvt0, t0, cv;
  double t0c, hb;         // t0 * 1e-6 (temperature factor of a clamped state), b0^2/2
  double kA, kA1;         // k0/(2 alpha^2), k0/(2 alpha)
  double ea, eb, eg;      // alpha, beta, gamma
  double hbeta, hg;       // beta/2, gamma/2
  // products the closed forms use as they stand (one FP64 instruction each time they would be formed in a kernel)
  double kA1ha;           // (k0/2alpha)(alpha/2)
  double c_a;             // -b0^2/6:      a = e1 + e2 I1 = c_a rB I1   (e1 = b0^2 rB I1/3, e2 = -b0^2 rB/2, so e1 = -2a)
  double c_kg;            // 2 (1 + beta): kg = c_kg a,  kh = -c_kg e2  (phase_acoustic_sym)
  double hg2, hb2;        // (gamma/2)^2, (beta/2)^2
};

inline EosDev make_eos_dev(const EosAbi& e) {
  EosDev d;
  d.rho0 = e.rho0; d.inv_rho0 = 1.0 / e.rho0;
  d.cvt0 = e.cv * e.t0; d.inv_cvt0 = 1.0 / (e.cv * e.t0);
  d.t0 = e.t0; d.cv = e.cv;
  d.t0c = e.t0 * 1e-6; d.hb = 0.5 * e.b0sq;
  d.kA = 0.5 * e.k0 / (e.alpha * e.alpha); d.kA1 = 0.5 * e.k0 / e.alpha;
  d.ea = e.alpha; d.eb = e.beta; d.eg = e.gamma;
  d.hbeta = 0.5 * e.beta; d.hg = 0.5 * e.gamma;
  d.kA1ha = d.kA1 * (0.5 * e.alpha);
  d.c_a = -e.b0sq / 6.0;
  d.c_kg = 2.0 * (1.0 + e.beta);
  d.hg2 = d.hg * d.hg; d.hb2 = d.hbeta * d.hbeta;
  return d;
}
inline bool eos_is_default_exponents(const EosAbi& e) { return e.alpha == 1.0 && e.beta == 3.0 && e.gamma == 2.0; }

// Branch-free reciprocal / reciprocal square root / square root for the hot path: hardware seed
// (MUFU.RCP64H / MUFU.RSQ64H read the high word only: ~20 good bits) + ONE third-order correction
// (residual e ~ 1e-6 -> e^3 ~ 1e-18, i.e. <= 1-2 ulp after the final rounding; one FP64 instruction fewer
// per reciprocal and three fewer per reciprocal square root than two Newton steps).  The CUDA library
// versions are correctly rounded but carry a slow-path call with register shuffling around every use;
// every argument here is an O(1) positive quantity of an admissible state (inadmissible states are
// flagged separately through PhaseState::bad).
HS_HD double hs_rcp(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);          // 1 - x y
  return fma(y, fma(e, e, e), y);            // y (1 + e + e^2)
#else
  return 1.0 / x;
#endif
}
HS_HD double hs_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);    // 1 - x y^2
  return fma(y * e, fma(0.375, e, 0.5), y);  // y (1 + e/2 + 3 e^2/8)
#else
  return 1.0 / sqrt(x);
#endif
}
HS_HD double hs_sqrt(double x) {
#ifdef __CUDA_ARCH__
  const double y = hs_rsqrt(x);
  const double sq = x * y;
  const double r = fma(0.5 * y, fma(-sq, sq, x), sq);   // one more correction on the product
  return x > 0.0 ? r : (x == 0.0 ? 0.0 : r);            // sqrt(0) = 0 (the seed is inf there); negative -> NaN
#else
  return sqrt(x);
#endif
}

// Everything downstream needs from one phase record.  Symmetric tensors are stored as
// [11,12,13,22,23,33].
struct PhaseState {
  double alpha, inv_alpha, rho, den, inv_den;
  double u[3], Etot;
  double G[6], G2r1[3];   // Finger tensor, row 1 of G^2
  double h22, h33;        // (G^2)_22, (G^2)_33 (only when phase_state is asked for them: WITH_H)
  double I1, J;           // tr G ;  I1^2/3 - I2
  double rB;              // (rho/rho0)^beta = I3^(beta/2)
  double W;               // shear energy (b0^2/2) I3^(beta/2) J
  double th;              // cv t0 I3^(gamma/2) (S' - 1)
  double uc1, uc2;        // (rA-1) rA ;  (2 rA - 1) rA     (cold-compression pieces)
  double Sp;              // S' = exp(S/cv), clamped at 1e-6 (EquationsOfState.jl:152-154)
  double T;               // temperature de/dS = t0 I3^(gamma/2) S'
  double a, e2, E3;       // M = a G - e2 G^2 + E3 I ,  sigma = -2 rho M
  double sig1[3];         // row 1 of sigma (true stress of the phase, not alpha-weighted)
  int bad;                // 1 where Julia would throw DomainError (sqrt of a negative) or NaN
};

// ONE: alpha is the literal 1 (single-phase model) -- its reciprocal and powers fold away at compile time
// (the reciprocal below is inline PTX, which the compiler cannot fold by itself; rcp(1) = 1 exactly, so
// the result is bit-identical).
// WITH_H: the caller goes on to the acoustic tensor, which needs the diagonal of G^2 anyway: tr(G^2) is then their
// sum instead of a separate contraction.
// TWICE: m, E, A hold TWICE the record (the sum of two records: the caller wants the state at their mean, NumFluxes.jl:86).
// Scaling by a power of two is exact, so folding the 1/2 into three scalar factors (1/8 on det, 1/4 on kappa, 1/2 on
// 1/den) gives bit-identical results to halving the 13 inputs first.
template <bool GEN, bool ONE = false, bool WITH_H = false, bool TWICE = false>
HS_HD void phase_state(const EosDev& eos, double alpha, const double* m, double E, const double* A, PhaseState& s) {
  // cofactors of A:  inv(A) = C^T / det A
  const double C11 = A[4] * A[8] - A[7] * A[5], C12 = A[7] * A[2] - A[1] * A[8], C13 = A[1] * A[5] - A[4] * A[2];
  const double C21 = A[6] * A[5] - A[3] * A[8], C22 = A[0] * A[8] - A[6] * A[2], C23 = A[3] * A[2] - A[0] * A[5];
  const double C31 = A[3] * A[7] - A[6] * A[4], C32 = A[6] * A[1] - A[0] * A[7], C33 = A[0] * A[4] - A[3] * A[1];
  // (A is column-major: A[i+3j]; C_ij = cofactor of A_ij.)  det by first row: A11 C11 + A12 C12 + A13 C13
  const double detA = A[0] * C11 + A[3] * C12 + A[6] * C13;
  const double ia = ONE ? 1.0 : hs_rcp(alpha);
  if (ONE) alpha = 1.0;
  const double ia2r = ia * ia * eos.inv_rho0;                                          // 1/(alpha^2 rho0)
  const double x = (TWICE ? 0.125 * detA : detA) * (ia * ia2r);                        // det(A/alpha)/rho0 = rho^2
  s.bad = !(x > 0.0);
  const double rs = hs_rsqrt(x);                              // 1/rho
  const double rho = x * rs;
  s.alpha = alpha; s.inv_alpha = ia; s.rho = rho; s.den = alpha * rho; s.inv_den = ia * rs;
  const double idm = TWICE ? 0.5 * s.inv_den : s.inv_den;
  s.u[0] = m[0] * idm; s.u[1] = m[1] * idm; s.u[2] = m[2] * idm;
  s.Etot = E * idm;
  const double e_int = s.Etot - 0.5 * (s.u[0] * s.u[0] + s.u[1] * s.u[1] + s.u[2] * s.u[2]);
  // G = (F F^T)^-1 = kappa^2 C C^T with kappa = den/det A = 1/(alpha^2 rho0 rho)
  const double kap = (TWICE ? 0.25 * rs : rs) * ia2r, k2 = kap * kap;
  s.G[0] = k2 * (C11 * C11 + C12 * C12 + C13 * C13);
  s.G[1] = k2 * (C11 * C21 + C12 * C22 + C13 * C23);
  s.G[2] = k2 * (C11 * C31 + C12 * C32 + C13 * C33);
  s.G[3] = k2 * (C21 * C21 + C22 * C22 + C23 * C23);
  s.G[4] = k2 * (C21 * C31 + C22 * C32 + C23 * C33);
  s.G[5] = k2 * (C31 * C31 + C32 * C32 + C33 * C33);
  const double* G = s.G;
  s.I1 = G[0] + G[3] + G[5];
  // J = I1^2/3 - I2 with I2 = (I1^2 - tr G^2)/2 (Strains.jl:49-50), i.e. J = tr(G^2)/2 - I1^2/6
  s.G2r1[0] = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
  s.G2r1[1] = G[0] * G[1] + G[1] * G[3] + G[2] * G[4];
  s.G2r1[2] = G[0] * G[2] + G[1] * G[4] + G[2] * G[5];
  double trG2;
  if (WITH_H) {
    s.h22 = G[1] * G[1] + G[3] * G[3] + G[4] * G[4];
    s.h33 = G[2] * G[2] + G[4] * G[4] + G[5] * G[5];
    trG2 = s.G2r1[0] + s.h22 + s.h33;
  } else {
    s.h22 = s.h33 = 0.0;
    trG2 = G[0] * G[0] + G[3] * G[3] + G[5] * G[5] + 2.0 * (G[1] * G[1] + G[2] * G[2] + G[4] * G[4]);
  }
  s.J = fma(0.5, trG2, -(s.I1 * s.I1) * HS_LIT(sixth, 1.0 / 6.0));
  // powers of I3 = r^2
  const double r = rho * eos.inv_rho0;
  double rA, rB, rC, irC;
  if (GEN) {
    // generic exponents: one log, then one exp per exponent that is not a small integer (the shipped
    // alternative copper set, main.jl:133, has alpha = 1: r itself)
    const double L = log(r);
    rA = (eos.ea == 1.0) ? r : ((eos.ea == 2.0) ? r * r : exp(eos.ea * L));
    rB = (eos.eb == 3.0) ? r * r * r : ((eos.eb == 2.0) ? r * r : exp(eos.eb * L));
    rC = (eos.eg == 2.0) ? r * r : ((eos.eg == 1.0) ? r : exp(eos.eg * L));
    irC = hs_rcp(rC);
  } else {
    const double ir = eos.rho0 * rs;
    rA = r; rB = r * r * r; rC = r * r; irC = ir * ir;
  }
  s.rB = rB;
  const double am1 = rA - 1.0;
  s.uc1 = am1 * rA; s.uc2 = (2.0 * rA - 1.0) * rA;
  const double W = eos.hb * rB * s.J;
  s.W = W;
  // Thermal energy th = cv t0 I3^(gamma/2) (S' - 1).  entropy (EquationsOfState.jl:139-156) forms
  // S' = th_raw / (cv t0 I3^(gamma/2)) + 1 from th_raw = e - W - U_cold and clamps it at 1e-6; multiplying the clamp
  // through by the positive cv t0 I3^(gamma/2) gives th = max(th_raw, (1e-6 - 1) cv t0 I3^(gamma/2)) without the
  // round trip through S' (and without the cancellation in S' - 1).  S' itself is only formed where it is asked for
  // (cons2prim's entropy output); the temperature de/dS = t0 I3^(gamma/2) S' = t0 (I3^(gamma/2) + th / (cv t0)).
  const double th_raw = e_int - W - eos.kA * am1 * am1;
  if (th_raw != th_raw) s.bad = 1;
  const double cr = eos.cvt0 * rC;
  const double th_min = cr * HS_LIT(th_clamp, 1e-6 - 1.0);
  // Clamped branch (EquationsOfState.jl:152-154, S' = 1e-6): S' and T = t0 I3^(gamma/2) S' are SELECTED, never formed
  // through the sums below -- with th = th_min those cancel six digits (T off by ~3e-10 against the reference's
  // exp(log(1e-6))).
  const bool clamped = th_raw < th_min;
  s.th = clamped ? th_min : th_raw;
  s.T = clamped ? eos.t0c * rC : eos.t0 * fma(th_raw, eos.inv_cvt0, rC);
  s.Sp = clamped ? HS_LIT(sp_clamp, 1e-6) : fma(th_raw * eos.inv_cvt0, irC, 1.0);
  // first derivatives of e(I1,I2,I3;S):  e1 = b0^2 rB I1/3, e2 = -b0^2 rB/2, E3 = e3*I3 (its shear part is (beta/2) W);
  // a = e1 + e2 I1 = -(b0^2/6) rB I1
  s.e2 = -eos.hb * rB;
  s.E3 = eos.kA1 * s.uc1 + eos.hg * s.th + eos.hbeta * W;
  s.a = eos.c_a * (rB * s.I1);
  const double m2r = -2.0 * rho, am = m2r * s.a, em = m2r * s.e2;   // sigma = -2 rho (a G - e2 G^2 + E3 I)
  s.sig1[0] = fma(am, G[0], fma(-em, s.G2r1[0], m2r * s.E3));
  s.sig1[1] = fma(am, G[1], -em * s.G2r1[1]);
  s.sig1[2] = fma(am, G[2], -em * s.G2r1[2]);
}

// Row-1 flavour of phase_state for the quadrature states of the path integrals (HS_PHASE_CH, the default since round 2).  The non-conservative column only needs u, T and row 1 of sigma, i.e. row 1 of G and
// of G^2, tr G and J.  They come from B = A A^T without forming G:  G = kappa^2 cof(B) (row 1 and the diagonal of the cofactor
// matrix only), G^-1 = B / den^2, I2 = I3 tr(G^-1) with I3 / den^2 = 1/(alpha rho0)^2, and by Cayley-Hamilton
// G^2 = I1 G - I2 1 + I3 G^-1: about 10 FP64 instructions fewer per state for a few ulp of cancellation in G^2
// (I1 G - I2 + I3 G^-1 ~ 3 - 3 + 1).  Sets everything noncons_accumulate reads; G[3..5], h22, h33, uc2, Sp are NOT set.
template <bool GEN>
HS_HD void phase_state_row1(const EosDev& eos, double alpha, const double* m, double E, const double* A, PhaseState& s) {
  const double C11 = A[4] * A[8] - A[7] * A[5], C12 = A[7] * A[2] - A[1] * A[8], C13 = A[1] * A[5] - A[4] * A[2];
  const double detA = A[0] * C11 + A[3] * C12 + A[6] * C13;
  const double ia = hs_rcp(alpha);
  const double ia2r = ia * ia * eos.inv_rho0;                 // 1/(alpha^2 rho0)
  const double x = detA * (ia * ia2r);                        // det(A/alpha)/rho0 = rho^2
  s.bad = !(x > 0.0);
  const double rs = hs_rsqrt(x);                              // 1/rho
  const double rho = x * rs;
  s.alpha = alpha; s.inv_alpha = ia; s.rho = rho; s.den = alpha * rho; s.inv_den = ia * rs;
  s.u[0] = m[0] * s.inv_den; s.u[1] = m[1] * s.inv_den; s.u[2] = m[2] * s.inv_den;
  s.Etot = E * s.inv_den;
  const double e_int = s.Etot - 0.5 * (s.u[0] * s.u[0] + s.u[1] * s.u[1] + s.u[2] * s.u[2]);
  const double kap = rs * ia2r, k2 = kap * kap;
  // B = A A^T (B_ij = sum_k A_ik A_jk), symmetric; K = the cofactors of B that are needed
  const double B11 = A[0] * A[0] + A[3] * A[3] + A[6] * A[6], B12 = A[0] * A[1] + A[3] * A[4] + A[6] * A[7];
  const double B13 = A[0] * A[2] + A[3] * A[5] + A[6] * A[8], B22 = A[1] * A[1] + A[4] * A[4] + A[7] * A[7];
  const double B23 = A[1] * A[2] + A[4] * A[5] + A[7] * A[8], B33 = A[2] * A[2] + A[5] * A[5] + A[8] * A[8];
  const double K11 = B22 * B33 - B23 * B23, K22 = B11 * B33 - B13 * B13, K33 = B11 * B22 - B12 * B12;
  const double K12 = B13 * B23 - B12 * B33, K13 = B12 * B23 - B13 * B22;
  s.G[0] = k2 * K11; s.G[1] = k2 * K12; s.G[2] = k2 * K13;
  s.G[3] = s.G[4] = s.G[5] = 0.0;
  s.h22 = s.h33 = 0.0;
  s.I1 = k2 * (K11 + K22 + K33);
  const double gI = ia2r * eos.inv_rho0;                      // I3 / den^2:  I3 G^-1 = gI B
  const double I2 = gI * (B11 + B22 + B33);
  s.G2r1[0] = fma(s.I1, s.G[0], fma(gI, B11, -I2));
  s.G2r1[1] = fma(s.I1, s.G[1], gI * B12);
  s.G2r1[2] = fma(s.I1, s.G[2], gI * B13);
  s.J = fma(s.I1 * s.I1, HS_LIT(third, 1.0 / 3.0), -I2);      // I1^2/3 - I2, EquationsOfState.jl:134 as written
  const double r = rho * eos.inv_rho0;
  double rA, rB, rC;
  if (GEN) {
    const double L = log(r);
    rA = (eos.ea == 1.0) ? r : ((eos.ea == 2.0) ? r * r : exp(eos.ea * L));
    rB = (eos.eb == 3.0) ? r * r * r : ((eos.eb == 2.0) ? r * r : exp(eos.eb * L));
    rC = (eos.eg == 2.0) ? r * r : ((eos.eg == 1.0) ? r : exp(eos.eg * L));
  } else {
    rA = r; rB = r * r * r; rC = r * r;
  }
  s.rB = rB;
  const double am1 = rA - 1.0;
  s.uc1 = am1 * rA; s.uc2 = 0.0; s.Sp = 0.0;
  const double W = eos.hb * rB * s.J;
  s.W = W;
  const double th_raw = e_int - W - eos.kA * am1 * am1;
  if (th_raw != th_raw) s.bad = 1;
  const double th_min = (eos.cvt0 * rC) * HS_LIT(th_clamp, 1e-6 - 1.0);
  const bool clamped = th_raw < th_min;        // (same selection as phase_state)
  s.th = clamped ? th_min : th_raw;
  s.T = clamped ? eos.t0c * rC : eos.t0 * fma(th_raw, eos.inv_cvt0, rC);
  s.e2 = -eos.hb * rB;
  s.E3 = eos.kA1 * s.uc1 + eos.hg * s.th + eos.hbeta * W;
  s.a = eos.c_a * (rB * s.I1);
  const double m2r = -2.0 * rho, am = m2r * s.a, em = m2r * s.e2;   // sigma = -2 rho (a G - e2 G^2 + E3 I)
  s.sig1[0] = fma(am, s.G[0], fma(-em, s.G2r1[0], m2r * s.E3));
  s.sig1[1] = fma(am, s.G[1], -em * s.G2r1[1]);
  s.sig1[2] = fma(am, s.G[2], -em * s.G2r1[2]);
}

// Physical x-flux of one phase in the 15-slot MPh order [0, den u1, mom(3), energy, A-block(9)].
// HyperelasticityMPh.jl:168-172;  den*(u1 F_ij - u_i F_1j) == u1 A_ij - u_i A_1j.
HS_HD void phase_flux(const PhaseState& s, const double* A, double* f) {
  const double du1 = s.den * s.u[0];
  const double as0 = s.alpha * s.sig1[0], as1 = s.alpha * s.sig1[1], as2 = s.alpha * s.sig1[2];
  f[0] = 0.0;
  f[1] = du1;
  f[2] = du1 * s.u[0] - as0;
  f[3] = du1 * s.u[1] - as1;
  f[4] = du1 * s.u[2] - as2;
  f[5] = du1 * s.Etot - (s.u[0] * as0 + s.u[1] * as1 + s.u[2] * as2);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double A1j = A[3 * j];
    f[6 + 3 * j] = 0.0;  // u1 A_1j - u1 A_1j
    f[7 + 3 * j] = s.u[0] * A[1 + 3 * j] - s.u[1] * A1j;
    f[8 + 3 * j] = s.u[0] * A[2 + 3 * j] - s.u[2] * A1j;
  }
}

// Eigenvalues of a symmetric 3x3 [11,12,13,22,23,33] by the trigonometric closed form,
// ascending.  (The reference calls LAPACK on a symmetric-to-roundoff matrix.)
HS_HD void sym3_eigs(const double* a, double* ev) {
  const double q = (a[0] + a[3] + a[5]) * (1.0 / 3.0);
  const double p1 = a[1] * a[1] + a[2] * a[2] + a[4] * a[4];
  const double b0 = a[0] - q, b3 = a[3] - q, b5 = a[5] - q;
  const double p2 = b0 * b0 + b3 * b3 + b5 * b5 + 2.0 * p1;
  if (!(p2 > 0.0)) { ev[0] = ev[1] = ev[2] = q; return; }
  const double p = sqrt(p2 * (1.0 / 6.0));
  const double ip = 1.0 / p;
  const double c0 = b0 * ip, c1 = a[1] * ip, c2 = a[2] * ip, c3 = b3 * ip, c4 = a[4] * ip, c5 = b5 * ip;
  double r = 0.5 * (c0 * (c3 * c5 - c4 * c4) - c1 * (c1 * c5 - c4 * c2) + c2 * (c1 * c4 - c3 * c2));
  r = fmin(1.0, fmax(-1.0, r));
  const double phi = acos(r) * (1.0 / 3.0);
  ev[2] = q + 2.0 * p * cos(phi);
  ev[0] = q + 2.0 * p * cos(phi + 2.0943951023931954923);
  ev[1] = 3.0 * q - ev[0] - ev[2];
}
// All three eigenvalues, robust for (near-)degenerate pairs: cyclic Jacobi, fixed 6 sweeps.
// Used by the full get_eigvals API (the trigonometric form loses ~sqrt(eps) on a degenerate pair,
// e.g. the two shear speeds at F = I); the hot path only needs the largest one, which the
// trigonometric form delivers to full precision when it is simple.
HS_HD void sym3_eigs_jacobi(const double* s6, double* ev) {
  double a00 = s6[0], a01 = s6[1], a02 = s6[2], a11 = s6[3], a12 = s6[4], a22 = s6[5];
#define HS_JROT(app, aqq, apq, apr, aqr)                                                  \
  if (apq != 0.0) {                                                                      \
    const double th = (aqq - app) / (2.0 * apq);                                         \
    const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));        \
    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;                                 \
    app -= t * apq; aqq += t * apq; apq = 0.0;                                            \
    const double r1 = c * apr - sn * aqr, r2 = sn * apr + c * aqr;                        \
    apr = r1; aqr = r2;                                                                   \
  }
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    HS_JROT(a00, a11, a01, a02, a12)
    HS_JROT(a00, a22, a02, a01, a12)
    HS_JROT(a11, a22, a12, a01, a02)
  }
#undef HS_JROT
  ev[0] = fmin(a00, fmin(a11, a22));
  ev[1] = fmax(fmin(a00, a11), fmin(fmax(a00, a11), a22));   // median
  ev[2] = fmax(a00, fmax(a11, a22));
}
// rare path of sym3_max_abs_eig: full Jacobi solve (accurate for any degeneracy), kept out of line
HS_HD_COLD double sym3_max_abs_eig_cold(double a0, double a1, double a2, double a3, double a4, double a5) {
  const double a[6] = {a0, a1, a2, a3, a4, a5};   // by value: the hot path must not spill the tensor to local memory
  double ev[3];
  sym3_eigs_jacobi(a, ev);
  return fmax(fabs(ev[0]), fabs(ev[2]));
}

// approximate reciprocal (about 20 good bits); only used inside self-correcting Newton iterations
HS_HD double hs_rcp_approx(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return (double)(1.0f / (float)x);
#endif
}

// Largest |eigenvalue| of a symmetric 3x3 [11,12,13,22,23,33] -- the hot-path eigen-solve.
// With q = tr/3, p^2 = |A - qI|_F^2 / 6, r = det(A - qI) / (2 p^3) the eigenvalues are q + p mu_k,
// mu^3 - 3 mu - 2 r = 0, mu_k = 2 cos((acos r + 2 pi k)/3).  The largest root mu in [1, 2] is found
// by Newton from above (monotone: f is convex and increasing right of it), started from a cubic
// over-estimate of 2cos(acos(r)/3); the division uses an approximate reciprocal because the
// iteration is self-correcting.  lambda_max >= |lambda_min| is guaranteed when p <= 2q (lambda_min
// >= q - 2p, lambda_max >= q + p), which always holds for a positive-definite acoustic tensor; any
// other case, and the neighbourhood of a degenerate largest pair (r -> -1, where every
// characteristic-polynomial method loses sqrt(eps)), takes the out-of-line Jacobi solve.
HS_HD double sym3_max_abs_eig(const double* a) {
  const double q = (a[0] + a[3] + a[5]) * HS_LIT(third, 1.0 / 3.0);
  const double p1 = a[1] * a[1] + a[2] * a[2] + a[4] * a[4];
  const double b0 = a[0] - q, b3 = a[3] - q, b5 = a[5] - q;
  const double p2 = (b0 * b0 + b3 * b3 + b5 * b5 + 2.0 * p1) * HS_LIT(sixth, 1.0 / 6.0);
  if (!(p2 > HS_LIT(tiny, 1e-280))) return fabs(q);   // isotropic tensor (also keeps the flush-to-zero reciprocal-sqrt seed away from denormals)
  const double ip = hs_rsqrt(p2);
  const double p = p2 * ip;
  const double detb = b0 * (b3 * b5 - a[4] * a[4]) - a[1] * (a[1] * b5 - a[4] * a[2]) + a[2] * (a[1] * a[4] - b3 * a[2]);
  // |r| <= 1 up to roundoff; a value a few ulp above 1 only moves the root a few ulp above 2, and anything outside
  // [-0.875, 1 + 1e-7] (NaN, or the garbage r of a numerically isotropic tensor) takes the cold path: no clamp needed
  const double r = 0.5 * detb * (ip * ip * ip);
  if (r >= HS_LIT(r_lo, -0.875) && r <= HS_LIT(r_hi, 1.0000001) && p <= 2.0 * q) {
    double mu = fma(r, fma(r, fma(r, HS_LIT(mu3, 0.08116787571408166), HS_LIT(mu2, -0.1425897443947186)), HS_LIT(mu1, 0.32800957660022223)),
                    HS_LIT(mu0, 1.7464452327513027));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const double m2 = mu * mu;
      const double f = mu * (m2 - 3.0) - 2.0 * r;
      mu = mu - f * hs_rcp_approx(3.0 * m2 - 3.0);
    }
    return q + p * mu;
  }
  return sym3_max_abs_eig_cold(a[0], a[1], a[2], a[3], a[4], a[5]);
}

// Symmetrised acoustic tensor for n = (1,0,0):
//   Omega_ij = (1/rho) sum_l d sigma_1i / dF_jl F_1l, entropy held fixed (EquationsOfState.jl:223-246).
// With dF = e_j (x) row1(F):  d rho = -rho d_1j,  dG = -(g_j e_1^T + e_1 g_j^T),
// dI1 = -2 G_1j, dI2 = I1 dI1 + 2 (G^2)_1j, dI3/I3 = -2 d_1j  (SURVEY.md A.5); only G, G^2 and
// the scalar energy derivatives enter -- F itself drops out.
template <bool WITH_H = false>
HS_HD void phase_acoustic_sym(const EosDev& eos, const PhaseState& s, double* S6) {
  // Collecting the terms of -2 (dM^(j)_1i - d_1j M_1i) by tensor structure and symmetrising
  // (g1, h1 = first rows of G, G^2; b0^2 rB = -2 e2, e1 = -(2/3) e2 I1):
  //   Omega = -2 { (e2 G11 - a) G + e2 G^2 + (e2/3) g1 g1^T + kg (g1 e_1^T + e_1 g1^T)
  //                + kh (h1 e_1^T + e_1 h1^T) - (2 dE3c + E3) e_1 e_1^T }
  //   kg = -a (1 + beta/2) + (beta/4) e1,   kh = e2 (1 + beta),
  //   dE3c = (k0/2alpha)(alpha/2)(2 rA - 1) rA + (gamma/2)^2 th + (b0^2/2)(beta/2)^2 rB J.
  const double* G = s.G;
  const double h11 = s.G2r1[0], h12 = s.G2r1[1], h13 = s.G2r1[2];
  const double h22 = WITH_H ? s.h22 : G[1] * G[1] + G[3] * G[3] + G[4] * G[4];
  const double h23 = G[1] * G[2] + G[3] * G[4] + G[4] * G[5];
  const double h33 = WITH_H ? s.h33 : G[2] * G[2] + G[4] * G[4] + G[5] * G[5];
  // with e1 = -2a:  kg = 2 (1 + beta) a,  kh = -2 (1 + beta) e2;  the shear part of dE3c is (beta/2)^2 W
  const double cG = -2.0 * (s.e2 * G[0] - s.a);
  const double cH = -2.0 * s.e2;
  const double cg = cH * HS_LIT(third, 1.0 / 3.0);
  const double kg = eos.c_kg * s.a;
  const double kh = -eos.c_kg * s.e2;
  const double dE3c = eos.kA1ha * s.uc2 + eos.hg2 * s.th + eos.hb2 * s.W;
  const double k11 = 2.0 * (2.0 * dE3c + s.E3);
  // entries collected by G / G^2 component (g1 = row 1 of G):
  //   S11 = G11 (cG + 2 kg + cg G11) + h11 (cH + 2 kh) + k11,   S1j = G1j (cG + kg + cg G11) + h1j (cH + kh)
  const double t1 = fma(cg, G[0], cG + kg);
  const double c3 = cH + kh;
  S6[0] = fma(G[0], t1 + kg, fma(h11, fma(2.0, kh, cH), k11));
  S6[1] = fma(G[1], t1, h12 * c3);
  S6[2] = fma(G[2], t1, h13 * c3);
  const double cg2 = cg * G[1], cg3 = cg * G[2];
  S6[3] = fma(cG, G[3], fma(cH, h22, cg2 * G[1]));
  S6[4] = fma(cG, G[4], fma(cH, h23, cg2 * G[2]));
  S6[5] = fma(cG, G[5], fma(cH, h33, cg3 * G[2]));
}

// The same tensor for an arbitrary unit normal n (EquationsOfState.jl:223-246 takes n; the 1-D driver
// only ever passes (1,0,0), main.jl:208).  With dF = e_j (x) (F^T n):  d rho = -rho n_j,
// dG = -(g_j n^T + n g_j^T), so the closed form above holds with e_1 -> n, g1 -> G n, h1 -> G^2 n,
// G11 -> n.G n.  Needs the full state (not on the hot path: used by hs_get_eigvals only).
HS_HD void phase_acoustic_sym_n(const EosDev& eos, const PhaseState& s, const double* n, double* S6) {
  const double* G = s.G;
  const double Gm[3][3] = {{G[0], G[1], G[2]}, {G[1], G[3], G[4]}, {G[2], G[4], G[5]}};
  double H[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = Gm[i][0] * Gm[0][j] + Gm[i][1] * Gm[1][j] + Gm[i][2] * Gm[2][j];
  double gn[3], hn[3];
  for (int i = 0; i < 3; ++i) {
    gn[i] = Gm[i][0] * n[0] + Gm[i][1] * n[1] + Gm[i][2] * n[2];
    hn[i] = H[i][0] * n[0] + H[i][1] * n[1] + H[i][2] * n[2];
  }
  const double nGn = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
  const double cG = -2.0 * (s.e2 * nGn - s.a), cH = -2.0 * s.e2, cg = cH * (1.0 / 3.0);
  const double kg = eos.c_kg * s.a;
  const double kh = -eos.c_kg * s.e2;
  const double dE3c = eos.kA1ha * s.uc2 + eos.hg2 * s.th + eos.hb2 * s.W;
  const double k11 = 2.0 * (2.0 * dE3c + s.E3);
  const int ix[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
  for (int k = 0; k < 6; ++k) {
    const int i = ix[k][0], j = ix[k][1];
    S6[k] = cG * Gm[i][j] + cH * H[i][j] + cg * gn[i] * gn[j] + kg * (gn[i] * n[j] + n[i] * gn[j]) +
            kh * (hn[i] * n[j] + n[i] * hn[j]) + k11 * n[i] * n[j];
  }
}

// c_max = sqrt(max_k |eig_k(Omega)|): the only thing any consumer of get_eigvals keeps
// (main.jl:210, NumFluxes.jl:90-91 take min / max / max|.| of u1 +- c_k).
template <bool WITH_H = false>
HS_HD double phase_cmax(const EosDev& eos, const PhaseState& s) {
  double S6[6];
  phase_acoustic_sym<WITH_H>(eos, s, S6);
  return hs_sqrt(sym3_max_abs_eig(S6));
}

}  // namespace hs
