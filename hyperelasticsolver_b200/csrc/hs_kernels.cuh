// Device kernels of the finite-volume hot path (sm_100a, FP64, structure-of-arrays).
//
// Work decomposition: one thread per (cell, phase).  The two phases of a cell sit in adjacent
// lanes, so the only cross-phase coupling of the model -- the interface velocity / interface
// stress of the non-conservative matrix (HyperelasticityMPh.jl:212-217) and the min/max of the
// wave bounds over phases -- is a lane-xor-1 shuffle.  The single-phase model is the same code
// with one thread per cell.
//
// k_step is ONE kernel per time step (main.jl:204-227 fused):
//   load tile (+1 halo cell each side) -> per-cell state + physical flux (shared memory)
//   -> per-face HLL / LxF path-conservative fluctuations (3 x 6 quadrature states per face for
//      HLL, each evaluated once instead of the reference's twice) -> conservative update
//   -> wave bounds of the NEW state + block max + one atomicMax per block, which is the
//      lambda_max that the NEXT step's dt = cfl dx / lambda_max needs.
// Intermediate fields (primitives, stress, fluxes, Q_hll, fluctuations) never touch HBM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

#include "hs_phase.cuh"
#include "hs_hank.cuh"

namespace hs {

// quadrature nodes evaluated per loop trip of a path integral (tuning knob; 1 = rolled)
#ifndef HS_NODE_UNROLL
#define HS_NODE_UNROLL 1
#endif

// single-phase: 1 = the step reads / writes the cached 1/rho + stress rows (the aux rows after the wave-speed row(s)), 0 = it recovers the
// state at its head instead (less traffic, more arithmetic) -- tuning knob
#ifndef HS_SP_USE_CACHE
#define HS_SP_USE_CACHE 1
#endif

constexpr int MODEL_SP13 = 0, MODEL_MPH30 = 1;
constexpr int FLUX_LXF = 0, FLUX_HLL = 1;
constexpr unsigned FULL = 0xffffffffu;

struct EosPair { EosDev e[2]; };

// max over the warp of a NON-NEGATIVE double: such doubles order like their bit patterns, so two
// 32-bit redux.sync (high word, then low word among the lanes that hold the winning high word)
// replace five shuffle + compare rounds.
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(FULL, hi);
  const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// gausslegendre(6) / gausslobatto(6) mapped to [0,1] as NumFluxes.jl:95 / :37 do.
__constant__ double c_gleg_x[6] = {(-0.9324695142031520278 + 1.0) / 2.0, (-0.6612093864662645137 + 1.0) / 2.0,
                                   (-0.2386191860831969086 + 1.0) / 2.0, (0.2386191860831969086 + 1.0) / 2.0,
                                   (0.6612093864662645137 + 1.0) / 2.0,  (0.9324695142031520278 + 1.0) / 2.0};
__constant__ double c_gleg_w[6] = {0.1713244923791703450 / 2.0, 0.3607615730481386076 / 2.0, 0.4679139345726910474 / 2.0,
                                   0.4679139345726910474 / 2.0, 0.3607615730481386076 / 2.0, 0.1713244923791703450 / 2.0};
__constant__ double c_glob_x[6] = {(-1.0 + 1.0) / 2.0, (-0.7650553239294646929 + 1.0) / 2.0, (-0.2852315164806450963 + 1.0) / 2.0,
                                   (0.2852315164806450963 + 1.0) / 2.0, (0.7650553239294646929 + 1.0) / 2.0, (1.0 + 1.0) / 2.0};
__constant__ double c_glob_w[6] = {0.06666666666666666667 / 2.0, 0.3784749562978469803 / 2.0, 0.5548583770354863530 / 2.0,
                                   0.5548583770354863530 / 2.0,  0.3784749562978469803 / 2.0, 0.06666666666666666667 / 2.0};

// Canonical per-phase record (15 slots): 0 alpha, 1 alpha*rho, 2-4 momentum, 5 energy,
// 6-14 A = alpha*rho*F column-major.  SP13 variable v lives in slot sp_slot(v); slots 0,1 unused.
__host__ __device__ constexpr int sp_slot(int v) { return v < 3 ? 2 + v : (v == 12 ? 5 : 6 + ((v - 3) / 3) + 3 * ((v - 3) % 3)); }

// Physical-flux slots that are identically zero (alpha, and the A_1j row: u1 A_1j - u1 A_1j) are not
// stored: the shared flux tile has 11 rows (slots 1..5, 7, 8, 10, 11, 13, 14).
__host__ __device__ constexpr bool flux_is_zero(int j) { return j == 0 || j == 6 || j == 9 || j == 12; }
__host__ __device__ constexpr int flux_row(int j) { return j - 1 - (j > 6) - (j > 9) - (j > 12); }
#define HS_FLUX(F, j) (flux_is_zero(j) ? 0.0 : (F)[flux_row(j) * T])

template <int MODEL> struct ModelTraits;
// NAUX: cached per-cell rows next to the state.  Rows 0,1 = wave bounds lo / hi (all any consumer of
// get_eigvals keeps).  The single-phase model also caches 1/rho and row 1 of the stress (the rows after them):
// the CFL sweep at the end of a step computes them anyway, and with them the next step's physical
// flux needs no state recovery (saves ~20 % of the FP64 work for 32 B/cell more traffic each way).
// HS_SP_CROW = 1: the single-phase model caches ONE wave-speed row, c_max, instead of the two bounds lo / hi = u1 -+ c_max
// (rows: 0 c_max, 1 1/rho, 2..4 stress row 1).  u1 = m1 * (1/rho) is the very product the flux forms anyway and 1/rho is cached, so
// the bounds come back bit-identical (the product is rounded on its own, __dmul_rn, exactly as phase_state rounds u1 before it
// subtracts) for 16 B less DRAM traffic per cell-update and no extra instruction.
#ifndef HS_SP_CROW
#define HS_SP_CROW 1   // (include/hyperelastic_b200.h sets the same default; hs_api.cu asserts that the two agree)
#endif
constexpr bool SP_CROW = HS_SP_CROW != 0;
constexpr int SP_R_ID = SP_CROW ? 1 : 2;   // cache row of 1/rho
constexpr int SP_R_SG = SP_CROW ? 2 : 3;   // first cache row of the stress row
template <> struct ModelTraits<MODEL_SP13> { static constexpr int NPH = 1, NVAR = 13, J0 = 2, NAUX = SP_CROW ? 5 : 6; };
template <> struct ModelTraits<MODEL_MPH30> { static constexpr int NPH = 2, NVAR = 30, J0 = 0, NAUX = 2; };

// physical flux of the single-phase record from the cached 1/rho and stress row (same expressions as
// phase_flux with alpha = 1; den*u1 is the momentum itself)
__device__ __forceinline__ void sp_flux_cached(const double* rec, double inv_den, const double* sig1, double* f) {
  const double* m = rec + 2;
  const double* A = rec + 6;
  const double u0 = m[0] * inv_den, u1 = m[1] * inv_den, u2 = m[2] * inv_den;
  f[2] = m[0] * u0 - sig1[0];
  f[3] = m[0] * u1 - sig1[1];
  f[4] = m[0] * u2 - sig1[2];
  f[5] = m[0] * (rec[5] * inv_den) - (u0 * sig1[0] + u1 * sig1[1] + u2 * sig1[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double A1j = A[3 * j];
    f[6 + 3 * j] = 0.0;
    f[7 + 3 * j] = u0 * A[1 + 3 * j] - u1 * A1j;
    f[8 + 3 * j] = u0 * A[2 + 3 * j] - u2 * A1j;
  }
}

// state of the record stored in a shared-memory column (row stride T)
template <int MODEL, bool GEN, int T>
__device__ __forceinline__ void column_state(const EosDev& eos, const double* col, PhaseState& st) {
  double m[3] = {col[2 * T], col[3 * T], col[4 * T]};
  double A[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = col[(6 + k) * T];
  const double alpha = (MODEL == MODEL_MPH30) ? col[0] : 1.0;
  phase_state<GEN, MODEL == MODEL_SP13>(eos, alpha, m, col[5 * T], A, st);
}

// Column 1 of the non-conservative block of this thread's phase at one state (the other phase's
// temperature, velocity and stress come from the adjacent lane).  HyperelasticityMPh.jl:212-230
// with omega = 0, k = (1/2, 1/2), beta = 0.  c[1] (the alpha*rho row) is identically zero.
__device__ __forceinline__ void noncons_column(const PhaseState& st, const double* A, double* c) {
  const double To = __shfl_xor_sync(FULL, st.T, 1);
  double uo[3], so[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { uo[k] = __shfl_xor_sync(FULL, st.u[k], 1); so[k] = __shfl_xor_sync(FULL, st.sig1[k], 1); }
  const double uI[3] = {0.5 * st.u[0] + 0.5 * uo[0], 0.5 * st.u[1] + 0.5 * uo[1], 0.5 * st.u[2] + 0.5 * uo[2]};  // :213
  const double inv = hs_rcp(st.T + To);
  const double sI[3] = {(To * st.sig1[0] + st.T * so[0]) * inv, (To * st.sig1[1] + st.T * so[1]) * inv,
                        (To * st.sig1[2] + st.T * so[2]) * inv};                                            // :217
  c[0] = uI[0];                                                                                              // :223
  c[1] = 0.0;
  c[2] = sI[0]; c[3] = sI[1]; c[4] = sI[2];                                                                  // :224
  c[5] = sI[0] * uI[0] + sI[1] * uI[1] + sI[2] * uI[2];                                                      // :225
  const double dv[3] = {uI[0] - st.u[0], uI[1] - st.u[1], uI[2] - st.u[2]};
  const double ia = st.inv_alpha;  // rho F = A / alpha
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double rF1j = A[3 * j] * ia;
    c[6 + 3 * j] = rF1j * st.u[0] + ia * (A[3 * j] * dv[0] + A[3 * j + 1] * dv[1] + A[3 * j + 2] * dv[2]);   // :228,:230
    c[7 + 3 * j] = rF1j * st.u[1];
    c[8 + 3 * j] = rF1j * st.u[2];
  }
}

// HS_PHASE_CH = 1 (default since round 2: 1.642 vs 1.591 G cell-updates/s on the same box, GPU parity suite green): quadrature
// states through phase_state_row1 (B = A A^T + Cayley-Hamilton, hs_phase.cuh) -- ~10 FP64 instructions fewer per state, results
// differ by a few ulp.  0 = the full phase_state.
#ifndef HS_PHASE_CH
#define HS_PHASE_CH 1
#endif
// w * (column 1 of the non-conservative block) as 14 products factor x value (slot 1, the alpha*rho row, is identically zero):
// same arithmetic as noncons_column with the quadrature weight folded into the common factors.  (Multiply by the reciprocal: a
// true FP64 division here costs ~10 % of the whole two-phase step.)  The products are NOT formed here: the caller accumulates
// acc = fma(factor, value, acc), node after node, so that the fused kernel (one thread walks over the nodes) and the
// quadrature-parallel kernel for small grids (one warp per node, k_step_qp) round identically.
//   factor index: 0 = w, 1 = w/(T1+T2), 2 = w/alpha, 3..5 = (w/alpha) A_1j;   value index k <-> slot (k == 0 ? 0 : k + 1)
struct NcTerms { double f[6]; double X[14]; };
__host__ __device__ constexpr int nc_factor(int k) { return k == 0 ? 0 : (k <= 4 ? 1 : ((k - 5) % 3 == 0 ? 2 : 3 + (k - 5) / 3)); }
__host__ __device__ constexpr int nc_slot(int k) { return k == 0 ? 0 : k + 1; }
__device__ __forceinline__ void noncons_terms(const PhaseState& st, const double* A, double w, NcTerms& t) {
  const double To = __shfl_xor_sync(FULL, st.T, 1);
  double uo[3], so[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { uo[k] = __shfl_xor_sync(FULL, st.u[k], 1); so[k] = __shfl_xor_sync(FULL, st.sig1[k], 1); }
  const double uI[3] = {0.5 * st.u[0] + 0.5 * uo[0], 0.5 * st.u[1] + 0.5 * uo[1], 0.5 * st.u[2] + 0.5 * uo[2]};
  const double wi = w * hs_rcp(st.T + To);
  const double n0 = To * st.sig1[0] + st.T * so[0], n1 = To * st.sig1[1] + st.T * so[1], n2 = To * st.sig1[2] + st.T * so[2];
  t.f[0] = w; t.f[1] = wi;
  t.X[0] = uI[0];
  t.X[1] = n0; t.X[2] = n1; t.X[3] = n2;
  t.X[4] = n0 * uI[0] + n1 * uI[1] + n2 * uI[2];
  // rho F_1j u_1 + rho (F^T (u_I - u))_j: the two A_1j terms combine, A_1j (u_1 + (u_I - u)_1) = A_1j u_I1
  const double dv1 = uI[1] - st.u[1], dv2 = uI[2] - st.u[2];
  const double wia = w * st.inv_alpha;
  t.f[2] = wia;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    t.f[3 + j] = wia * A[3 * j];
    t.X[5 + 3 * j] = fma(A[3 * j], uI[0], fma(A[3 * j + 1], dv1, A[3 * j + 2] * dv2));
    t.X[6 + 3 * j] = st.u[1];
    t.X[7 + 3 * j] = st.u[2];
  }
}
__device__ __forceinline__ void noncons_accumulate(const PhaseState& st, const double* A, double w, double* acc) {
  NcTerms t;
  noncons_terms(st, A, w, t);
#pragma unroll
  for (int k = 0; k < 14; ++k) acc[nc_slot(k)] = fma(t.f[nc_factor(k)], t.X[k], acc[nc_slot(k)]);
}

// The state of this thread's phase at the point base + s * delta of a straight path between two records (14 components: alpha,
// momentum(3), energy, A(9) = slots 0, 2..14; slot 1 is never read).  One fused multiply-add per component.
template <bool GEN>
__device__ __forceinline__ void path_point(const EosDev& eos, const double* base14, const double* d14, double s, PhaseState& st, double* A) {
  const double alpha = fma(s, d14[0], base14[0]);
  double m[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = fma(s, d14[1 + k], base14[1 + k]);
  const double E = fma(s, d14[4], base14[4]);
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = fma(s, d14[5 + k], base14[5 + k]);
#if HS_PHASE_CH
  phase_state_row1<GEN>(eos, alpha, m, E, A, st);
#else
  phase_state<GEN>(eos, alpha, m, E, A, st);
#endif
}

// acc[j] += dalpha * sum_q w_q c_j(psi(s_q)) along the straight path between two records
// (NumFluxes.jl:97-107).  The path is given as a base column and a DELTA column d = end - start
// (shared memory, row stride T; d[0] = dalpha):
//   FROM_END = false: psi(s) = base + s d          (base = start of the segment)
//   FROM_END = true : psi(s) = base - (1 - s) d    (base = end of the segment)
// i.e. one fused multiply-add per component instead of the two of Q_l (1-s) + Q_r s.  MPh only.
template <bool GEN, int T, bool FROM_END>
__device__ __forceinline__ void path_integral(const EosDev& eos, const double* base, const double* d, const double* xs,
                                              const double* ws, double* acc, int& bad) {
  const double dalpha = d[0];
  constexpr int NODE_UNROLL = HS_NODE_UNROLL;
#pragma unroll NODE_UNROLL
  for (int q = 0; q < 6; ++q) {
    const double s = FROM_END ? xs[q] - 1.0 : xs[q], w = ws[q] * dalpha;
    double b14[14], d14[14], A[9];
    b14[0] = base[0]; d14[0] = dalpha;
#pragma unroll
    for (int k = 1; k < 14; ++k) { b14[k] = base[(1 + k) * T]; d14[k] = d[(1 + k) * T]; }
    PhaseState st;
    path_point<GEN>(eos, b14, d14, s, st, A);
    bad |= st.bad;
    noncons_accumulate(st, A, w, acc);
  }
}

// Component-wise pieces of hll_pathcons / lxf / update_cell with the rounding spelled out (explicit fused multiply-adds and
// individually rounded products), shared by the fused kernel and the quadrature-parallel kernel for small grids so that both
// produce the same bits.
//   Q_hll = (Q_r s_r - Q_l s_l - (B_int + F_r - F_l)) / (s_r - s_l)                                      NumFluxes.jl:109-111
__device__ __forceinline__ double hll_qhll(double a, double b, double Fa, double Fb, double acc, double s_l, double s_r, double inv_ds) {
  const double path = (acc + Fb) - Fa;
  return __dmul_rn(fma(b, s_r, -__dmul_rn(a, s_l)) - path, inv_ds);
}
//   D- = -s_l/(s_r-s_l) [F_r - F_l + B_int(Q_l,Q_hll) + B_int(Q_hll,Q_r)] + s_l s_r/(s_r-s_l) (Q_r - Q_l),  D+ likewise   :128-129
__device__ __forceinline__ void hll_fluct(double a, double b, double Fa, double Fb, double acc, double k_q, double k_m, double k_p,
                                          double& dm, double& dp) {
  const double br = (Fb - Fa) + acc;
  const double dq = __dmul_rn(k_q, b - a);
  dm = fma(k_m, br, dq);
  dp = fma(k_p, br, -dq);
}
//   (F(Q_l) + F(Q_r))/2 - lambda (Q_r - Q_l)/2                                                             NumFluxes.jl:30
__device__ __forceinline__ double lxf_cons(double a, double b, double Fa, double Fb, double lambda) {
  return fma(-(0.5 * lambda), b - a, 0.5 * (Fa + Fb));
}
//   Q - dt/dx ((F_r - F_l) + (NF_r + NF_l)) with the face terms already combined per side                  main.jl:59 / :40
__device__ __forceinline__ double cell_update(double q, double upd, double from_right_face, double from_left_face) {
  return fma(-upd, from_right_face + from_left_face, q);
}

// HLL wave-speed bounds of a face, NumFluxes.jl:86-91: min / max over 0, the speeds at Q_m = (Q_l + Q_r)/2 and the cached
// bounds of the left / right cell.  a, b: record columns (row stride T).
template <int MODEL, bool GEN, int T>
__device__ __forceinline__ void face_speeds(const EosDev& eos, const double* a, const double* b, double lo_l, double hi_r,
                                            double& s_l, double& s_r, int& bad) {
  constexpr bool MPH = MODEL == MODEL_MPH30;
  double m[3], A[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = a[(2 + k) * T] + b[(2 + k) * T];
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = a[(6 + k) * T] + b[(6 + k) * T];
  const double alpha = MPH ? 0.5 * (a[0] + b[0]) : 1.0;
  PhaseState sm;   // (m, E, A are the sums: phase_state folds the halving in, exactly)
  phase_state<GEN, !MPH, true, true>(eos, alpha, m, a[5 * T] + b[5 * T], A, sm);
  bad |= sm.bad;
  const double cm = phase_cmax<true>(eos, sm);
  double lo_m = sm.u[0] - cm, hi_m = sm.u[0] + cm;
  if (MPH) {
    lo_m = fmin(lo_m, __shfl_xor_sync(FULL, lo_m, 1));
    hi_m = fmax(hi_m, __shfl_xor_sync(FULL, hi_m, 1));
  }
  s_l = fmin(0.0, fmin(lo_m, lo_l));
  s_r = fmax(0.0, fmax(hi_m, hi_r));
}

// One face between the records in columns a (left) and b (right) with their physical fluxes Fa, Fb.
// lo_l / hi_r: cached wave bounds of the left / right cell (the `eigvals` argument of hll,
// NumFluxes.jl:90-91).  H: this thread's private scratch column (Q_hll).  For every slot j the
// functor receives (j, cons, dm, dp) = the three return values of hll / lxf for that component.
template <int MODEL, int FLUX, bool GEN, int T, class Emit>
__device__ __forceinline__ void face_eval(const EosDev& eos, const double* a, const double* b, const double* Fa,
                                          const double* Fb, double lo_l, double hi_r, double lambda, double* H,
                                          int& bad, double* s_out, Emit emit) {
  constexpr int J0 = ModelTraits<MODEL>::J0;
  constexpr bool MPH = MODEL == MODEL_MPH30;
  if (FLUX == FLUX_HLL) {
    double s_l, s_r;
    face_speeds<MODEL, GEN, T>(eos, a, b, lo_l, hi_r, s_l, s_r, bad);
    if (s_out) { s_out[0] = s_l; s_out[1] = s_r; }
    const double inv_ds = hs_rcp(s_r - s_l);
    const double k_q = s_l * s_r * inv_ds;
    if (MPH) {
      double acc[15];
#pragma unroll
      for (int j = 0; j < 15; ++j) acc[j] = 0.0;
      // H (this thread's private column) holds the delta of the segment being integrated
#pragma unroll
      for (int j = 0; j < 15; ++j)
        if (j != 1) H[j * T] = b[j * T] - a[j * T];      // alpha*rho (slot 1) is never read by the physics
      path_integral<GEN, T, false>(eos, a, H, c_gleg_x, c_gleg_w, acc, bad);           // B_int(Q_l, Q_r)
#pragma unroll
      for (int j = 0; j < 15; ++j) {
        if (j == 1) continue;
        const double qh = hll_qhll(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), acc[j], s_l, s_r, inv_ds);   // :109-111  Q_hll
        H[j * T] = qh - a[j * T];
        acc[j] = 0.0;
      }
      path_integral<GEN, T, false>(eos, a, H, c_gleg_x, c_gleg_w, acc, bad);           // B_int(Q_l, Q_hll)
#pragma unroll
      for (int j = 0; j < 15; ++j)
        if (j != 1) H[j * T] = (b[j * T] - a[j * T]) - H[j * T];                      // Q_r - Q_hll
      path_integral<GEN, T, true>(eos, b, H, c_gleg_x, c_gleg_w, acc, bad);            // B_int(Q_hll, Q_r)
      const double k_m = -s_l * inv_ds, k_p = s_r * inv_ds;
#pragma unroll
      for (int j = 0; j < 15; ++j) {                                                   // :128-129
        double dm, dp;
        hll_fluct(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), j == 1 ? 0.0 : acc[j], k_q, k_m, k_p, dm, dp);
        emit(j, 0.0, dm, dp);
      }
    } else {
      const double w_l = s_r * inv_ds, w_r = s_l * inv_ds;                             // weights of F_l and F_r
#pragma unroll
      for (int j = J0; j < 15; ++j)                                                    // NumFluxes.jl:78 (one-phase form)
        emit(j, fma(w_l, HS_FLUX(Fa, j), fma(-w_r, HS_FLUX(Fb, j), k_q * (b[j * T] - a[j * T]))), 0.0, 0.0);
    }
  } else {
    double acc[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) acc[j] = 0.0;
    if (MPH) {                                                                         // NumFluxes.jl:35-47
#pragma unroll
      for (int j = 0; j < 15; ++j)
        if (j != 1) H[j * T] = b[j * T] - a[j * T];
      path_integral<GEN, T, false>(eos, a, H, c_glob_x, c_glob_w, acc, bad);
    }
#pragma unroll
    for (int j = J0; j < 15; ++j) {
      const double cons = lxf_cons(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), lambda);          // :30
      const double d = (MPH && j != 1) ? 0.5 * acc[j] : 0.0;                                       // :50-51
      emit(j, cons, d, d);
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct StepArgs {
  const double* Qin; double* Qout;
  const double* aux_in; double* aux_out;   // [NAUX][stride]
  unsigned long long* lam;   // [3][nprob] bit patterns of lambda_max (non-negative doubles order like u64)
  double* tt;                // [3][nprob] time
  long long* steps;          // [nprob]
  int* status;
  double* dt_hist; long long hist_k, hist_cap;   // dt_hist[prob*hist_cap + hist_k];  hist_k < 0: index from the device counter hist_n[prob]
  long long* hist_n;         // [nprob] per-problem count of recorded steps (lets a captured CUDA graph of steps be replayed)
  long long stride;
  int ncells, nprob, tiles_per_prob;
  int cur, nxt, clr;
  int ghost;                 // bit 0 / 1: the first / last cell is a halo copy owned by a neighbouring slab
  double cfl, dx, t_end;
  unsigned long long spin_ns;   // time limit of the tile-copy wait of k_step_sp (never hang the GPU on a lost copy)
  const double* dt_shared;      // not null: every problem steps with this dt (device scalar) instead of cfl dx / its own max(lambda)
  EosPair eos;
};

// dt of the step just taken by problem `prob` into the history (called by one thread per problem and step)
__device__ __forceinline__ void hist_record(const StepArgs& g, int prob, double dt) {
  if (!g.dt_hist) return;
  long long k = g.hist_k;
  if (k < 0) { k = g.hist_n[prob]; g.hist_n[prob] = k + 1; }
  if (k < g.hist_cap) g.dt_hist[(size_t)prob * g.hist_cap + k] = dt;
}

// rows J0..14 of the record and scratch tiles, the non-zero flux rows, cached bounds, reduction scratch
template <int MODEL, int T> constexpr size_t step_smem_bytes() {
  return sizeof(double) * ((2 * (15 - ModelTraits<MODEL>::J0) + 11 - flux_row(ModelTraits<MODEL>::J0 < 1 ? 1 : ModelTraits<MODEL>::J0)) * T + 2 * T + 32);
}

// resident blocks per SM the register allocation is capped for (tuned on B200, profiles/)
#ifndef HS_MINB_MPH
#define HS_MINB_MPH 4
#endif
#ifndef HS_MINB_SP
#define HS_MINB_SP 4
#endif
// SAME: both phases share one equation of state (the shipped configuration, main.jl:134): the EoS constants
// are then uniform kernel parameters (constant-bank operands) instead of per-lane constant loads.
template <int MODEL, int FLUX, bool GEN, int T, bool SAME>
__global__ void __launch_bounds__(T, (MODEL == MODEL_MPH30 ? HS_MINB_MPH : HS_MINB_SP)) k_step(const StepArgs g) {
  using MT = ModelTraits<MODEL>;
  constexpr int NPH = MT::NPH, CPB = T / NPH, J0 = MT::J0;
  extern __shared__ double smem[];
  // Three tile arrays of NROW = 15 - J0 rows (the single-phase model has no slots 0,1); the
  // pointers are biased by -J0 rows so that slot j is always row j.
  constexpr int NROW = 15 - J0;
  constexpr int F0 = flux_row(J0 < 1 ? 1 : J0), NFROW = 11 - F0;   // first stored flux row, number of flux rows
  double* Rs = smem - J0 * T;                  // records        [J0..14][T]
  double* Hs = smem + NROW * T - J0 * T;       // Q_hll / path deltas, then the fluctuation handed to the left cell
  double* Fs = smem + 2 * NROW * T - F0 * T;   // physical flux, non-zero slots only (flux_row)
  double* lo_s = smem + (2 * NROW + NFROW) * T;  // [CPB]
  double* hi_s = lo_s + T;      // [CPB]
  double* red = hi_s + T;       // [T/32]

  const int tid = threadIdx.x, l = tid / NPH, ph = tid % NPH;
  const int prob = blockIdx.x / g.tiles_per_prob, tile = blockIdx.x % g.tiles_per_prob;
  const int c = tile * (CPB - 2) + l;
  const bool valid = c < g.ncells;
  const long long gi = (long long)prob * g.ncells + (valid ? c : g.ncells - 1);
  const EosDev& eos = g.eos.e[SAME ? 0 : ph];

  // Issue every global load of the block back to back (two per-problem scalars, the tile, the
  // cached bounds) before anything consumes them, so their latencies overlap.
  const unsigned long long lam_bits = __ldg(g.lam + (size_t)g.cur * g.nprob + prob);
  const double t_cur = __ldg(g.tt + (size_t)g.cur * g.nprob + prob);
  double rin[15], lo_in_r = 0.0, hi_in_r = 0.0, ax[(MODEL == MODEL_SP13) ? 4 : 1];
  constexpr bool SPC = (MODEL == MODEL_SP13) && (HS_SP_USE_CACHE != 0);   // single-phase: use the cached 1/rho + stress row
  if (MODEL == MODEL_MPH30) {
#pragma unroll
    for (int j = 0; j < 15; ++j) rin[j] = __ldg(g.Qin + (size_t)(15 * ph + j) * g.stride + gi);
  } else {
#pragma unroll
    for (int v = 0; v < 13; ++v) rin[sp_slot(v)] = __ldg(g.Qin + (size_t)v * g.stride + gi);
  }
  constexpr bool CROW = SP_CROW && MODEL == MODEL_SP13;
  static_assert(!CROW || SPC, "the c_max row needs the cached 1/rho");
  double c_in_r = 0.0;
  if (CROW) c_in_r = __ldg(g.aux_in + gi);
  else if (ph == 0) { lo_in_r = __ldg(g.aux_in + gi); hi_in_r = __ldg(g.aux_in + g.stride + gi); }
  if (SPC) {
#pragma unroll
    for (int r = 0; r < 4; ++r) ax[r] = __ldg(g.aux_in + (size_t)(SP_R_ID + r) * g.stride + gi);
  }
  if (CROW) {
    const double u1 = __dmul_rn(rin[2], ax[0]);   // u1 as phase_state rounds it (no contraction into the add)
    lo_in_r = u1 - c_in_r; hi_in_r = u1 + c_in_r;
  }

  const bool own_interior = valid && l >= 1 && l <= CPB - 2 && c <= g.ncells - 2;
  // frozen physical boundary cells, main.jl:219-220 (halo cells of a slab are neither written nor
  // counted in lambda_max: their owner does both)
  const bool own_frozen = valid && ((c == 0 && !(g.ghost & 1)) || (c == g.ncells - 1 && !(g.ghost & 2)));

  // ---- stage the tile in shared memory -------------------------------------------------------
#pragma unroll
  for (int j = J0; j < 15; ++j) Rs[j * T + tid] = rin[j];
  if (ph == 0) { lo_s[l] = lo_in_r; hi_s[l] = hi_in_r; }

  const double lam_cur = __longlong_as_double((long long)lam_bits);
  const bool active = t_cur < g.t_end;                   // while t < T, main.jl:202
  // dt and the update factor are per-problem scalars: one thread does the (IEEE, correctly rounded)
  // divisions, the block reads them after the next barrier
  double* sc = red + 8;                                  // [dt, update factor, dx/dt]
  if (tid == 0) {
    const double dt0 = g.dt_shared ? *g.dt_shared : g.cfl * g.dx / lam_cur;   // main.jl:212 (dt_shared: a dimension-split sweep takes the grid's dt)
    const double lambda0 = g.dx / dt0;                   // main.jl:223
    sc[0] = dt0;
    sc[1] = (FLUX == FLUX_HLL) ? dt0 / g.dx : 1.0 / lambda0;   // main.jl:225,59 / :40
    sc[2] = lambda0;
  }

  if (!active) {  // this problem already reached t_end: carry the state through unchanged
    if (own_interior || own_frozen) {
      if (MODEL == MODEL_MPH30) {
#pragma unroll
        for (int j = 0; j < 15; ++j) g.Qout[(size_t)(15 * ph + j) * g.stride + gi] = Rs[j * T + tid];
      } else {
#pragma unroll
        for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = Rs[sp_slot(v) * T + tid];
      }
      if (CROW) g.aux_out[gi] = c_in_r;
      else if (ph == 0) { g.aux_out[gi] = lo_s[l]; g.aux_out[g.stride + gi] = hi_s[l]; }
      if (SPC) {
#pragma unroll
        for (int r = 0; r < 4; ++r) g.aux_out[(size_t)(SP_R_ID + r) * g.stride + gi] = ax[r];
      }
    }
    if (tile == 0 && tid == 0) {
      g.tt[(size_t)g.nxt * g.nprob + prob] = t_cur;
      g.lam[(size_t)g.nxt * g.nprob + prob] = (unsigned long long)__double_as_longlong(lam_cur);
      g.lam[(size_t)g.clr * g.nprob + prob] = 0ull;
    }
    return;
  }

  int bad = 0;
  // ---- per-cell physical flux (flux_mph, HyperelasticityMPh.jl:140-175) -----------------------
  if (SPC) {
    double f[15];
    sp_flux_cached(rin, ax[0], ax + 1, f);   // state recovery was done by the previous step's CFL sweep
#pragma unroll
    for (int j = J0; j < 15; ++j)
      if (!flux_is_zero(j)) Fs[flux_row(j) * T + tid] = f[j];
  } else {
    PhaseState st;
    column_state<MODEL, GEN, T>(eos, Rs + tid, st);
    if (valid) bad |= st.bad;
    double A[9], f[15];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = Rs[(6 + k) * T + tid];
    phase_flux(st, A, f);
#pragma unroll
    for (int j = J0; j < 15; ++j)
      if (!flux_is_zero(j)) Fs[flux_row(j) * T + tid] = f[j];
  }
  __syncthreads();

  const double dt = sc[0], upd = sc[1], lambda = sc[2];

  // ---- face between cell l-1 and cell l (hll / lxf, NumFluxes.jl) ---------------------------
  const int tl = (l >= 1) ? tid - NPH : tid;   // halo threads evaluate a dummy face against themselves
  // Two-phase: CR = fluctuation kept by this (right) cell.  Single-phase: the face only has a
  // conservative flux, so the right cell's share is minus the left cell's (read back from Hs).
  double CR[MODEL == MODEL_MPH30 ? 15 : 1];
  {
    int fbad = 0;
    double* H = Hs + tid;
    auto emit = [&](int j, double cons, double dm, double dp) {
      H[j * T] = cons + dm;                              // F_r + NF_r of the left cell  (update_cell, main.jl:57-59)
      if (MODEL == MODEL_MPH30) CR[j] = dp - cons;       // - F_l + NF_l of this cell
    };
    face_eval<MODEL, FLUX, GEN, T>(eos, Rs + tl, Rs + tid, Fs + tl, Fs + tid, lo_s[(l >= 1) ? l - 1 : l], hi_s[l],
                                   lambda, H, fbad, nullptr, emit);
    if (valid && l >= 1) bad |= fbad;
  }
  __syncthreads();

  // ---- conservative update (update_cell, main.jl:59 / :40) + wave bounds of the new state ----
  double lamv = 0.0;
  {
    double qn[15];
    const int tr = own_interior ? tid + NPH : tid;
#pragma unroll
    for (int j = J0; j < 15; ++j) {
      const double q = Rs[j * T + tid];
      const double mine = (MODEL == MODEL_MPH30) ? CR[j] : -Hs[j * T + tid];
      qn[j] = own_interior ? cell_update(q, upd, Hs[j * T + tr], mine) : q;
    }
    if (own_interior) {
      if (MODEL == MODEL_MPH30) {
#pragma unroll
        for (int j = 0; j < 15; ++j) g.Qout[(size_t)(15 * ph + j) * g.stride + gi] = qn[j];
      } else {
#pragma unroll
        for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = qn[sp_slot(v)];
      }
    } else if (own_frozen) {
      if (MODEL == MODEL_MPH30) {
#pragma unroll
        for (int j = 0; j < 15; ++j) g.Qout[(size_t)(15 * ph + j) * g.stride + gi] = qn[j];
      } else {
#pragma unroll
        for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = qn[sp_slot(v)];
      }
    }
    // CFL sweep of the next step (get_eigvals, main.jl:204-211) on the state just produced
    PhaseState sn;
    phase_state<GEN, MODEL == MODEL_SP13, true>(eos, (MODEL == MODEL_MPH30) ? qn[0] : 1.0, qn + 2, qn[5], qn + 6, sn);
    const double cn = phase_cmax<true>(eos, sn);
    double lo_n = sn.u[0] - cn, hi_n = sn.u[0] + cn;
    if (NPH == 2) {
      lo_n = fmin(lo_n, __shfl_xor_sync(FULL, lo_n, 1));
      hi_n = fmax(hi_n, __shfl_xor_sync(FULL, hi_n, 1));
    }
    if (own_interior) {
      bad |= sn.bad;
      if (CROW) g.aux_out[gi] = cn;
      else if (ph == 0) { g.aux_out[gi] = lo_n; g.aux_out[g.stride + gi] = hi_n; }
      if (SPC) {
        g.aux_out[(size_t)SP_R_ID * g.stride + gi] = sn.inv_den;
#pragma unroll
        for (int r = 0; r < 3; ++r) g.aux_out[(size_t)(SP_R_SG + r) * g.stride + gi] = sn.sig1[r];
      }
      lamv = fmax(fabs(lo_n), fabs(hi_n));
    } else if (own_frozen) {
      if (CROW) g.aux_out[gi] = c_in_r;
      else if (ph == 0) { g.aux_out[gi] = lo_s[l]; g.aux_out[g.stride + gi] = hi_s[l]; }
      if (SPC) {
#pragma unroll
        for (int r = 0; r < 4; ++r) g.aux_out[(size_t)(SP_R_ID + r) * g.stride + gi] = ax[r];
      }
      lamv = fmax(fabs(lo_s[l]), fabs(hi_s[l]));
    }
  }
  // ---- block max of lambda, one atomic per block ---------------------------------------------
  lamv = warp_max_nonneg(lamv);
  if ((tid & 31) == 0) red[tid >> 5] = lamv;
  bad = __any_sync(FULL, bad);
  if (bad && (tid & 31) == 0) atomicOr(g.status, 1);
  __syncthreads();
  if (tid == 0) {
    double mx = red[0];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) mx = fmax(mx, red[w]);
    atomicMax(g.lam + (size_t)g.nxt * g.nprob + prob, (unsigned long long)__double_as_longlong(mx));
    if (tile == 0) {
      g.tt[(size_t)g.nxt * g.nprob + prob] = t_cur + dt;   // main.jl:214
      g.steps[prob] += 1;                                   // main.jl:215
      hist_record(g, prob, dt);
      g.lam[(size_t)g.clr * g.nprob + prob] = 0ull;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_step_qp: the two-phase time step for SMALL grids, quadrature-parallel.
//
// k_step gives every (cell, phase) one thread that walks through the 18 quadrature states of its face one after the other
// (~6 300 dependent-ish instructions): on a grid that cannot fill the GPU -- the run main.jl ships is nx = 1000 -- the step time
// is that thread's latency (~10 us), not throughput.  Here a block of 12 warps owns a tile of 16 cells (32 (cell, phase) pairs =
// the 32 lanes of a warp, same lane layout as k_step: the other phase sits in lane ^ 1) and the independent pieces of a face run
// in DIFFERENT WARPS at the same time:
//   A   warps 0-5: quadrature node q = warp of B_int(Q_l, Q_r);  warp 6: physical flux of the cells;  warp 7: wave speeds at Q_m
//   B   all warps: Q_hll from the six node terms (components shared out over the warps)
//   C   warps 0-5: node q of B_int(Q_l, Q_hll);  warps 6-11: node q of B_int(Q_hll, Q_r)
//   D   all warps: fluctuations D-, D+        E   all warps: conservative update        F   warp 0: CFL sweep of the new state
// so the critical path is ~3 state evaluations + 2 eigen-solves instead of 20.  The node terms are handed over as (factor, value)
// pairs and accumulated with the same fused multiply-adds in the same order as k_step's loop, and every other formula is the
// shared inline function k_step uses: the result is BIT-IDENTICAL to k_step (tests/test_gpu_small_grid.py), so which kernel runs
// is purely a matter of grid size (hsd_step: two-phase grids of <= HS_QP_MAX_CELLS cells, default 2048).
// ------------------------------------------------------------------------------------------------
constexpr int QP_NP = 32, QP_CPB = 16, QP_WARPS = 12, QP_T = QP_WARPS * 32;
constexpr size_t qp_smem_doubles() {
  return (size_t)QP_NP * (15 /*Rs*/ + 11 /*Fs*/ + 15 + 15 /*H1 H2*/ + 12 * 6 /*PF*/ + 12 * 14 /*PX*/ + 15 + 15 /*Hm CR*/ + 15 /*Rn*/ + 4 /*lo hi sl sr*/) + 32;
}

// one quadrature node: state of this lane's phase at base + s delta, the 6 factors and 14 values of w * (non-conservative column)
template <bool GEN>
__device__ __forceinline__ int qp_node(const EosDev& eos, const double* b14, const double* d14, double s, double w_q, double* PFq,
                                       double* PXq, int lane) {
  PhaseState st;
  double A[9];
  path_point<GEN>(eos, b14, d14, s, st, A);
  NcTerms t;
  noncons_terms(st, A, w_q * d14[0], t);
#pragma unroll
  for (int k = 0; k < 6; ++k) PFq[k * QP_NP + lane] = t.f[k];
#pragma unroll
  for (int k = 0; k < 14; ++k) PXq[k * QP_NP + lane] = t.X[k];
  return st.bad;
}

// One step of k_step_qp on the buffers / scalar slots given (the persistent variant k_step_qp_loop rotates them itself).
// COHERENT: the inputs were written by other blocks of the SAME launch (previous step of the persistent loop): loads must
// come from L2 (ld.global.cg), not through the non-coherent path.
template <int FLUX, bool GEN, bool SAME, bool COHERENT>
__device__ __forceinline__ void qp_step_body(const StepArgs& g, const double* Qin, double* Qout, const double* aux_in, double* aux_out,
                                             int cur, int nxt, int clr, double* smem) {
  constexpr int T = QP_NP, CPB = QP_CPB;   // (T: row stride of the shared tiles, as the HS_FLUX macro expects)
  auto ld = [](const double* ptr) { return COHERENT ? __ldcg(ptr) : __ldg(ptr); };
  double* Rs = smem;                 // records              [15][32]
  double* Fs = Rs + 15 * T;          // physical flux        [11][32]  (flux_row)
  double* H1 = Fs + 11 * T;          // Q_hll - Q_l          [15][32]
  double* H2 = H1 + 15 * T;          // Q_r - Q_hll          [15][32]
  double* PF = H2 + 15 * T;          // node factors         [12][6][32]
  double* PX = PF + 12 * 6 * T;      // node values          [12][14][32]
  double* Hm = PX + 12 * 14 * T;     // face term handed to the left cell  [15][32]
  double* CR = Hm + 15 * T;          // face term kept by the right cell   [15][32]
  double* Rn = CR + 15 * T;          // new records          [15][32]
  double* lo_s = Rn + 15 * T;        // [16] cached bounds
  double* hi_s = lo_s + T;
  double* sl_s = hi_s + T;           // [32] s_l, s_r of this pair's face
  double* sr_s = sl_s + T;
  double* sc = sr_s + T;             // scalars

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = lane >> 1, ph = lane & 1;
  const int prob = blockIdx.x / g.tiles_per_prob, tile = blockIdx.x % g.tiles_per_prob;
  const int c = tile * (CPB - 2) + l;
  const bool valid = c < g.ncells;
  const long long gi = (long long)prob * g.ncells + (valid ? c : g.ncells - 1);
  const EosDev& eos = g.eos.e[SAME ? 0 : ph];

  for (int j = warp; j < 15; j += QP_WARPS) Rs[j * T + lane] = ld(Qin + (size_t)(15 * ph + j) * g.stride + gi);
  if (warp == QP_WARPS - 1 && ph == 0) { lo_s[l] = ld(aux_in + gi); hi_s[l] = ld(aux_in + g.stride + gi); }
  if (tid == 0) {
    const double lam_cur = ld(reinterpret_cast<const double*>(g.lam) + (size_t)cur * g.nprob + prob);
    const double dt0 = g.dt_shared ? *g.dt_shared : g.cfl * g.dx / lam_cur;   // main.jl:212 (dt_shared: a dimension-split sweep takes the grid's dt)
    const double lambda0 = g.dx / dt0;                   // main.jl:223
    sc[0] = dt0;
    sc[1] = (FLUX == FLUX_HLL) ? dt0 / g.dx : 1.0 / lambda0;   // main.jl:225,59 / :40
    sc[2] = lambda0;
    sc[3] = ld(g.tt + (size_t)cur * g.nprob + prob);
    sc[4] = lam_cur;
  }
  __syncthreads();
  const double dt = sc[0], upd = sc[1], lambda = sc[2], t_cur = sc[3];
  const bool active = t_cur < g.t_end;                   // while t < T, main.jl:202
  const bool own_interior = valid && l >= 1 && l <= CPB - 2 && c <= g.ncells - 2;
  const bool own_frozen = valid && ((c == 0 && !(g.ghost & 1)) || (c == g.ncells - 1 && !(g.ghost & 2)));
  if (!active) {  // this problem already reached t_end: carry the state through unchanged
    if (own_interior || own_frozen) {
      for (int j = warp; j < 15; j += QP_WARPS) Qout[(size_t)(15 * ph + j) * g.stride + gi] = Rs[j * T + lane];
      if (warp == QP_WARPS - 1 && ph == 0) { aux_out[gi] = lo_s[l]; aux_out[g.stride + gi] = hi_s[l]; }
    }
    if (tile == 0 && tid == 0) {
      g.tt[(size_t)nxt * g.nprob + prob] = t_cur;
      g.lam[(size_t)nxt * g.nprob + prob] = (unsigned long long)__double_as_longlong(sc[4]);
      g.lam[(size_t)clr * g.nprob + prob] = 0ull;
    }
    return;
  }

  const int tl = (l >= 1) ? lane - 2 : lane;   // the face of pair `lane` lies between cell l-1 and cell l (l = 0: dummy face)
  const double* a = Rs + tl;
  const double* b = Rs + lane;
  const double* Fa = Fs + tl;
  const double* Fb = Fs + lane;
  const double* xs = (FLUX == FLUX_HLL) ? c_gleg_x : c_glob_x;
  const double* ws = (FLUX == FLUX_HLL) ? c_gleg_w : c_glob_w;
  int bad = 0;

  // ---- A: first path integral (one node per warp) | physical flux | wave speeds at Q_m ---------------------------------
  if (warp < 6) {
    double b14[14], d14[14];
    b14[0] = a[0]; d14[0] = b[0] - a[0];
#pragma unroll
    for (int k = 1; k < 14; ++k) { b14[k] = a[(1 + k) * T]; d14[k] = b[(1 + k) * T] - a[(1 + k) * T]; }
    const int nb = qp_node<GEN>(eos, b14, d14, xs[warp], ws[warp], PF + warp * 6 * T, PX + warp * 14 * T, lane);
    if (valid && l >= 1) bad |= nb;
  } else if (warp == 6) {
    PhaseState st;
    column_state<MODEL_MPH30, GEN, T>(eos, Rs + lane, st);
    if (valid) bad |= st.bad;
    double A[9], f[15];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = Rs[(6 + k) * T + lane];
    phase_flux(st, A, f);
#pragma unroll
    for (int j = 1; j < 15; ++j)
      if (!flux_is_zero(j)) Fs[flux_row(j) * T + lane] = f[j];
  } else if (warp == 7 && FLUX == FLUX_HLL) {
    double s_l, s_r;
    int fbad = 0;
    face_speeds<MODEL_MPH30, GEN, T>(eos, a, b, lo_s[(l >= 1) ? l - 1 : l], hi_s[l], s_l, s_r, fbad);
    sl_s[lane] = s_l; sr_s[lane] = s_r;
    if (valid && l >= 1) bad |= fbad;
  }
  __syncthreads();

  if (FLUX == FLUX_HLL) {
    const double s_l = sl_s[lane], s_r = sr_s[lane];
    const double inv_ds = hs_rcp(s_r - s_l);
    // ---- B: Q_hll (NumFluxes.jl:109-111) and the deltas of the two remaining segments ------------------------------------
    for (int k = warp; k < 14; k += QP_WARPS) {
      const int j = nc_slot(k), fk = nc_factor(k);
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) acc = fma(PF[(q * 6 + fk) * T + lane], PX[(q * 14 + k) * T + lane], acc);
      const double qh = hll_qhll(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), acc, s_l, s_r, inv_ds);
      const double h1 = qh - a[j * T];
      H1[j * T + lane] = h1;
      H2[j * T + lane] = (b[j * T] - a[j * T]) - h1;
    }
    __syncthreads();
    // ---- C: B_int(Q_l, Q_hll) in warps 0-5, B_int(Q_hll, Q_r) in warps 6-11 -----------------------------------------------
    {
      const bool second = warp >= 6;
      const int q = second ? warp - 6 : warp;
      const double* base = second ? b : a;
      const double* dl = (second ? H2 : H1) + lane;
      double b14[14], d14[14];
      b14[0] = base[0]; d14[0] = dl[0];
#pragma unroll
      for (int k = 1; k < 14; ++k) { b14[k] = base[(1 + k) * T]; d14[k] = dl[(1 + k) * T]; }
      const int nb = qp_node<GEN>(eos, b14, d14, second ? xs[q] - 1.0 : xs[q], ws[q], PF + warp * 6 * T, PX + warp * 14 * T, lane);
      if (valid && l >= 1) bad |= nb;
    }
    __syncthreads();
    // ---- D: fluctuations (NumFluxes.jl:128-129) -----------------------------------------------------------------------------
    const double k_q = s_l * s_r * inv_ds, k_m = -s_l * inv_ds, k_p = s_r * inv_ds;
    for (int j = warp; j < 15; j += QP_WARPS) {
      double acc = 0.0;
      if (j != 1) {
        const int k = j == 0 ? 0 : j - 1, fk = nc_factor(k);
#pragma unroll
        for (int q = 0; q < 12; ++q) acc = fma(PF[(q * 6 + fk) * T + lane], PX[(q * 14 + k) * T + lane], acc);
      }
      double dm, dp;
      hll_fluct(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), acc, k_q, k_m, k_p, dm, dp);
      Hm[j * T + lane] = 0.0 + dm;      // F_r + NF_r of the left cell  (update_cell, main.jl:57-59; hll's conservative part is zero)
      CR[j * T + lane] = dp - 0.0;      // - F_l + NF_l of this cell
    }
  } else {
    // ---- D (LxF): NumFluxes.jl:30, :50-51 ---------------------------------------------------------------------------------
    for (int j = warp; j < 15; j += QP_WARPS) {
      double acc = 0.0;
      if (j != 1) {
        const int k = j == 0 ? 0 : j - 1, fk = nc_factor(k);
#pragma unroll
        for (int q = 0; q < 6; ++q) acc = fma(PF[(q * 6 + fk) * T + lane], PX[(q * 14 + k) * T + lane], acc);
      }
      const double cons = lxf_cons(a[j * T], b[j * T], HS_FLUX(Fa, j), HS_FLUX(Fb, j), lambda);
      const double d = (j != 1) ? 0.5 * acc : 0.0;
      Hm[j * T + lane] = cons + d;
      CR[j * T + lane] = d - cons;
    }
  }
  __syncthreads();

  // ---- E: conservative update (update_cell, main.jl:59 / :40) -------------------------------------------------------------
  {
    const int tr = own_interior ? lane + 2 : lane;
    for (int j = warp; j < 15; j += QP_WARPS) {
      const double q = Rs[j * T + lane];
      const double qn = own_interior ? cell_update(q, upd, Hm[j * T + tr], CR[j * T + lane]) : q;
      Rn[j * T + lane] = qn;
      if (own_interior || own_frozen) Qout[(size_t)(15 * ph + j) * g.stride + gi] = qn;
    }
  }
  __syncthreads();

  // ---- F: CFL sweep of the next step on the state just produced (get_eigvals, main.jl:204-211) ----------------------------
  if (warp == 0) {
    double qn[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) qn[j] = Rn[j * T + lane];
    PhaseState sn;
    phase_state<GEN, false, true>(eos, qn[0], qn + 2, qn[5], qn + 6, sn);
    const double cn = phase_cmax<true>(eos, sn);
    double lo_n = sn.u[0] - cn, hi_n = sn.u[0] + cn;
    lo_n = fmin(lo_n, __shfl_xor_sync(FULL, lo_n, 1));
    hi_n = fmax(hi_n, __shfl_xor_sync(FULL, hi_n, 1));
    double lamv = 0.0;
    if (own_interior) {
      bad |= sn.bad;
      if (ph == 0) { aux_out[gi] = lo_n; aux_out[g.stride + gi] = hi_n; }
      lamv = fmax(fabs(lo_n), fabs(hi_n));
    } else if (own_frozen) {
      if (ph == 0) { aux_out[gi] = lo_s[l]; aux_out[g.stride + gi] = hi_s[l]; }
      lamv = fmax(fabs(lo_s[l]), fabs(hi_s[l]));
    }
    lamv = warp_max_nonneg(lamv);
    if (lane == 0) {
      atomicMax(g.lam + (size_t)nxt * g.nprob + prob, (unsigned long long)__double_as_longlong(lamv));
      if (tile == 0) {
        g.tt[(size_t)nxt * g.nprob + prob] = t_cur + dt;   // main.jl:214
        g.steps[prob] += 1;                                   // main.jl:215
        hist_record(g, prob, dt);
        g.lam[(size_t)clr * g.nprob + prob] = 0ull;
      }
    }
  }
  bad = __any_sync(FULL, bad);
  if (bad && lane == 0) atomicOr(g.status, 1);
}

template <int FLUX, bool GEN, bool SAME>
__global__ void __launch_bounds__(QP_T, 1) k_step_qp(const StepArgs g) {
  extern __shared__ double smem[];
  qp_step_body<FLUX, GEN, SAME, false>(g, g.Qin, g.Qout, g.aux_in, g.aux_out, g.cur, g.nxt, g.clr, smem);
}

// The device-resident loop of a small grid in ONE launch: nsteps steps of k_step_qp separated by grid-wide barriers
// (cooperative launch: every block is resident), buffers and scalar slots rotated inside the kernel.  A step costs its own
// latency (~3 us for nx = 1000) plus a barrier instead of a kernel launch (~5 us between dependent launches, graph or not).
// Steps past t_end are no-ops, exactly as the host-launched sequence (the host sizes nsteps from the clock).
template <int FLUX, bool GEN, bool SAME>
__global__ void __launch_bounds__(QP_T, 1) k_step_qp_loop(const StepArgs g, const int nsteps) {
  extern __shared__ double smem[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  double* const Qa = const_cast<double*>(g.Qin);
  double* const Aa = const_cast<double*>(g.aux_in);
  int cur = g.cur, nxt = g.nxt, clr = g.clr;
  for (int s = 0; s < nsteps; ++s) {
    const bool odd = s & 1;
    qp_step_body<FLUX, GEN, SAME, true>(g, odd ? g.Qout : Qa, odd ? Qa : g.Qout, odd ? g.aux_out : Aa, odd ? Aa : g.aux_out, cur, nxt, clr, smem);
    const int c0 = cur; cur = nxt; nxt = clr; clr = c0;
    grid.sync();      // all writes of step s (state, cache rows, max(lambda), clock) visible to every block; shared tiles free again
  }
}

// ------------------------------------------------------------------------------------------------
// k_step_sp: the single-phase time step as a TMA-fed tile pipeline.
//
// Same arithmetic as k_step<MODEL_SP13> (same inline functions, same expressions), different data
// movement.  In k_step a block spends a quarter of its life waiting for its own tile (4 blocks/SM at
// 128 registers: nothing else to run), so here a block walks over `kper` tiles (grid-stride, so the
// hardware block scheduler still balances the SMs -- a static persistent partition loses ~10 %,
// profiles/r01_experiments.md) and, while it computes tile k, the TMA engine (cp.async.bulk, one
// 1040-byte row copy per state / cache row, completion on an mbarrier) fills the other shared-memory
// stage with tile k+1.  No register or instruction cost for the loads, no load latency on the
// critical path after the first tile.
//   * rows are copied from the 16-byte-aligned address at or below the tile start (130 doubles per
//     row), so any ncells / stride parity works: slot j of cell `col` sits at column col + parity(row);
//   * the physical flux of BOTH cells of a face comes straight from the cached 1/rho and stress rows
//     (23 FP64 instructions) instead of a shared flux tile + barrier;
//   * the face flux is handed to the left cell by a warp shuffle (shared memory only across the
//     three warp boundaries), so nothing reads a stage after the one barrier of a tile and the
//     stage can be refilled while the block is still in its update phase;
//   * max(lambda) is kept per thread across the tiles of a problem: one atomicMax per block
//     instead of one per tile.
// Tiles whose 130-element window would run past the end of the arrays (the last one or two of the
// launch) are loaded by the threads themselves.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// one 2-D tile of a row-major (rows x stride) array through a tensor map: box = 128 columns x all rows, any start column
__device__ __forceinline__ void tensor2d_g2s(void* dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory access through a 32-bit shared-window address (the compiler otherwise re-derives the window base for
// every predicated access of the warp-boundary hand-off: 4 extra instructions per load, issued by every lane)
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
}  // namespace tma

constexpr int SP_TS = 130;                       // doubles per stage row: 128 cells + alignment slack
constexpr int SP_NAX = ModelTraits<MODEL_SP13>::NAUX;   // cache rows per cell (6, or 5 with HS_SP_CROW)
constexpr int SP_NST = 13 + SP_NAX;              // rows per stage: 13 state rows (slots 2..14) + the cache rows
constexpr unsigned SP_ROW_BYTES = SP_TS * 8, SP_STAGE_BYTES = SP_NST * SP_ROW_BYTES;
// SP13 variable stored in record slot j (inverse of sp_slot)
__host__ __device__ constexpr int sp_var(int j) { return j < 5 ? j - 2 : (j == 5 ? 12 : 3 + 3 * ((j - 6) % 3) + (j - 6) / 3); }
// shared-memory stages of the tensor-map flavour (tile k + HS_SP_STAGES - 1 is in flight while tile k is computed); the
// row-copy flavour always has two
#ifndef HS_SP_STAGES
#define HS_SP_STAGES 2
#endif
template <bool TM2D> __host__ __device__ constexpr int sp_stages() { return TM2D ? HS_SP_STAGES : 2; }
template <int T, bool TM2D> __host__ __device__ constexpr size_t sp_stage_doubles() {
  return (size_t)sp_stages<TM2D>() * SP_NST * (TM2D ? T : SP_TS) > (size_t)2 * SP_NST * SP_TS ? (size_t)sp_stages<TM2D>() * SP_NST * (TM2D ? T : SP_TS)
                                                                                             : (size_t)2 * SP_NST * SP_TS;
}
template <int T, bool TM2D = false> constexpr size_t step_sp_smem_bytes() {
  return sizeof(double) * (sp_stage_doubles<T, TM2D>() + 2 * (T / 32) * 13 + T / 32 + 24 + 4);
}

// SINGLE: one problem (nprob == 1, every grid config): the problem index, the per-problem scalars, the column parities
// and the 64-bit tile offsets are then loop invariants or 32-bit, and max(lambda) is flushed once per block.
// TM2D: the state and the cached rows of a tile arrive as TWO tensor-map copies (13 x 128 and 6 x 128 doubles, any start
// column, out-of-range columns zero-filled) issued by one thread, instead of one row copy per state / cache row spread over 20 lanes: ~40 fewer
// issue slots per warp and tile, no column parities, no thread-loaded last tile.  Needs an even stride (row pitch multiple of
// 16 bytes) and a stride below 2^31; the host falls back to the row copies otherwise.
template <int FLUX, bool GEN, int T, bool SINGLE, bool TM2D>
__global__ void __launch_bounds__(T, HS_MINB_SP) k_step_sp(const StepArgs g, const int kper, const int bpp, const __grid_constant__ CUtensorMap tmQ,
                                                           const __grid_constant__ CUtensorMap tmA) {
  static_assert(T == 128, "stage rows hold 128 cells");
  constexpr int TS = TM2D ? T : SP_TS;          // doubles per stage row
  constexpr int STAGE = SP_NST * TS;            // doubles per stage
  constexpr int NSTG = sp_stages<TM2D>(), PD = NSTG - 1;         // stages, prefetch distance in tiles
  extern __shared__ __align__(128) double smem[];
  double* const stage0 = smem;                                   // [NSTG][SP_NST][TS]
  double* const Hb = smem + sp_stage_doubles<T, TM2D>();         // [2][T/32][13] face flux of lane 0 of every warp
  double* const red = Hb + 2 * (T / 32) * 13;                    // [T/32]
  double* const sc = red + T / 32;                               // [3][8]: dt, update factor, dx/dt, t, lambda_max
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(sc + 24);   // [NSTG]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned ntiles = (unsigned)g.tiles_per_prob * (unsigned)g.nprob;
  const EosDev& eos = g.eos.e[0];

  auto write_scalars = [&](double* dst, unsigned long long lam_bits, double t) {
    const double lam_cur = __longlong_as_double((long long)lam_bits);
    const double dt0 = g.dt_shared ? *g.dt_shared : g.cfl * g.dx / lam_cur;   // main.jl:212 (dt_shared: a dimension-split sweep takes the grid's dt)
    const double lambda0 = g.dx / dt0;                   // main.jl:223
    dst[0] = dt0;
    dst[1] = (FLUX == FLUX_HLL) ? dt0 / g.dx : 1.0 / lambda0;   // main.jl:225,59 / :40
    dst[2] = lambda0;
    dst[3] = t;
    dst[4] = lam_cur;
  };
  // SP_NST row copies per tile, spread over the four warps (a bulk copy is a uniform-datapath instruction: the lanes of a
  // warp issue theirs one after the other, so one warp doing all 19 would reach the tile's barrier ~15 % late); the
  // window starts at the even element at or below the tile start.  Copies of the other warps may complete before thread
  // 0 has posted the expected byte count: the phase cannot complete before that arrival, and tx-count is signed.
  auto issue = [&](long long off, double* st, uint64_t* bar) {
    if (TM2D) {
      if (tid == 0) {
        tma::mbar_arrive_expect_tx(bar, (unsigned)(STAGE * sizeof(double)));
        tma::tensor2d_g2s(st, &tmQ, (int)off, 0, bar);
        tma::tensor2d_g2s(st + 13 * TS, &tmA, (int)off, 0, bar);
      }
      return;
    }
    if (tid == 0) tma::mbar_arrive_expect_tx(bar, SP_STAGE_BYTES);
    const int rr = lane * (T / 32) + warp;
    if (lane < (SP_NST + T / 32 - 1) / (T / 32) && rr < SP_NST) {
      const bool isq = rr < 13;
      const int r = isq ? rr : rr - 13;
      const int row = isq ? sp_slot(rr) - 2 : rr;
      const long long e = (long long)r * g.stride + off;
      const double* src = (isq ? g.Qin : g.aux_in) + (e - (e & 1));
      tma::bulk_g2s(st + row * SP_TS, src, SP_ROW_BYTES, bar);
    }
  };

  // SINGLE (the problem is a block invariant): bpp blocks share one problem, block b of them takes its tiles b, b + bpp, ...
  // (blocks resident at the same time work on neighbouring tiles); one grid is the case bpp = gridDim.  Since round 2 ensembles of
  // problems with at least a few tiles run this way too (bpp = ceil(tiles_per_prob / 8) blocks per problem): the per-problem
  // indexing of the other flavour cost ~110 instructions per cell-update (ncu: 1150 vs 1039).
  // !SINGLE (ensembles of tiny problems): block b takes the kper consecutive tiles from b * kper across problem boundaries (new
  // scalars and one more barrier for the max(lambda) flush whenever the problem changes).
  int prob = SINGLE ? (int)(blockIdx.x / (unsigned)bpp) : 0;
  unsigned id = SINGLE ? blockIdx.x - (unsigned)prob * (unsigned)bpp : blockIdx.x * (unsigned)kper;
  const unsigned id_step = SINGLE ? (unsigned)bpp : 1u;
  if (!SINGLE) prob = (int)(id / (unsigned)g.tiles_per_prob);
  int tile = SINGLE ? (int)id : (int)(id % (unsigned)g.tiles_per_prob);
  const int pbase = SINGLE ? prob * g.ncells : 0;          // first cell of the block's problem (< 2^31: hsd_problem_init)
  const unsigned nt_lim = SINGLE ? (unsigned)g.tiles_per_prob : ntiles;
  long long off = (long long)prob * g.ncells + (long long)tile * (T - 2);
  bool cur_tma = TM2D || off + SP_TS <= g.stride;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSTG; ++i) tma::mbar_init(&mbar[i], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  // start column of the tile d places after the current one in this block's sequence (false: there is none)
  auto tile_after = [&](int kk, int d, unsigned id_, int prob_, int tile_, long long& offd) -> bool {
    const unsigned idd = id_ + (unsigned)d * id_step;
    if (kk + d >= kper || idd >= nt_lim) return false;
    if (SINGLE) { offd = (long long)(pbase + (int)idd * (T - 2)); return true; }
    int td = tile_ + d, pd = prob_;
    while (td >= g.tiles_per_prob) { td -= g.tiles_per_prob; ++pd; }
    offd = (long long)pd * g.ncells + (long long)td * (T - 2);
    return true;
  };
  if (cur_tma) issue(off, stage0, &mbar[0]);
  if (NSTG > 2) {   // deeper pipeline: tiles 1 .. PD-1 of this block are requested up front as well
#pragma unroll
    for (int d = 1; d < PD; ++d) {
      long long offd = 0;
      if (tile_after(0, d, id, prob, tile, offd)) issue(offd, stage0 + d * STAGE, &mbar[d]);
    }
  }
  // the per-problem scalars (three IEEE divisions) are the job of one thread of warp 1, and only when the problem changes
  if (tid == 32) write_scalars(sc, __ldg(g.lam + (size_t)g.cur * g.nprob + prob), __ldg(g.tt + (size_t)g.cur * g.nprob + prob));
  __syncthreads();

  const unsigned hb0 = tma::smem_u32(Hb) + (unsigned)warp * (13u * 8u);   // this warp's slot of the boundary-flux buffer
  unsigned phase_bits = 0;   // bit s: parity of the next completion of stage s
  int sci = 0;               // scalar slot of the current problem
  double lam_run = 0.0;
  int bad = 0;
  for (int k = 0; k < kper; ++k) {
    const int s = NSTG == 2 ? (k & 1) : k % NSTG;
    double* const st = stage0 + s * STAGE;
    const double* const scv = sc + (SINGLE ? 0 : sci * 8);
    // ---- start fetching the next tile of this block ---------------------------------------------
    const unsigned idn = id + id_step;
    const bool has_next = (k + 1 < kper) && idn < nt_lim;
    int probn = prob, tilen = 0;
    long long offn = 0;
    bool next_tma = false;
    if (has_next) {
      if (SINGLE) {
        tilen = (int)idn;
        offn = (long long)(pbase + tilen * (T - 2));  // < ncells * nprob < 2^31
      } else {
        tilen = tile + 1;
        if (tilen >= g.tiles_per_prob) { tilen = 0; ++probn; }
        offn = (long long)probn * g.ncells + (long long)tilen * (T - 2);
      }
      next_tma = TM2D || offn + SP_TS <= g.stride;
    }
    // (every thread passed the barrier of tile k-1, after which nobody reads the stage of tile k-1 any more: it takes tile k+PD)
    if (NSTG == 2) {
      if (next_tma) issue(offn, stage0 + (s ^ 1) * STAGE, &mbar[s ^ 1]);
    } else {
      long long offp = 0;
      const int sp = (k + PD) % NSTG;
      if (tile_after(k, PD, id, prob, tile, offp)) issue(offp, stage0 + sp * STAGE, &mbar[sp]);
    }
    const bool new_prob = !SINGLE && has_next && probn != prob;
    unsigned long long lam_n = 0ull;
    double t_n = 0.0;
    if (tid == 32 && new_prob) {
      lam_n = __ldg(g.lam + (size_t)g.cur * g.nprob + probn);
      t_n = __ldg(g.tt + (size_t)g.cur * g.nprob + probn);
    }

    const int c = tile * (T - 2) + tid;
    const bool valid = c < g.ncells;
    // (32-bit cell index: ncells * nprob < 2^31 is guaranteed by hsd_problem_init, and a 32-bit index lets every global access be one
    // wide multiply-add on a uniform row base -- the ensemble flavour spent ~100 instructions per cell-update on 64-bit indexing)
    const int gi = prob * g.ncells + (valid ? c : g.ncells - 1);
    const int pe = SINGLE ? (pbase & 1) : (int)(off & 1), po = SINGLE ? (int)((pbase + g.stride) & 1) : (int)((off + g.stride) & 1);   // (tiles start on even cells of their problem)
    // slot j / cache row r of the cell in stage column `col`
    // (row copies: stage row = slot - 2, column shifted by the row's parity; tensor-map copies: stage row = variable, no shift)
#define SQ(j, col) st[(TM2D ? sp_var(j) : (j) - 2) * TS + (col) + (TM2D ? 0 : ((sp_var(j) & 1) ? po : pe))]
#define SA(r, col) st[(13 + (r)) * TS + (col) + (TM2D ? 0 : (((r) & 1) ? po : pe))]
    bool lost = false;   // the tile copy did not arrive in time: flag it and write NOTHING of this tile (the input buffer stays intact)
    if (cur_tma) {
      const unsigned par = (phase_bits >> s) & 1u;
      if (!tma::mbar_try_wait(&mbar[s], par)) {
        // wall-clock limit (%globaltimer, several seconds: time-slicing, MPS, a debugger or the sanitizer cannot trip it)
        const unsigned long long t0 = global_timer_ns();
        while (!tma::mbar_try_wait(&mbar[s], par)) {
          if (global_timer_ns() - t0 > g.spin_ns) { atomicOr(g.status, 4); lost = true; break; }
        }
      }
      phase_bits ^= 1u << s;
    } else {
#pragma unroll
      for (int v = 0; v < 13; ++v) SQ(sp_slot(v), tid) = __ldg(g.Qin + (size_t)v * g.stride + gi);
#pragma unroll
      for (int r = 0; r < SP_NAX; ++r) SA(r, tid) = __ldg(g.aux_in + (size_t)r * g.stride + gi);
      __syncthreads();
    }

    const bool own_interior = valid && !lost && tid >= 1 && tid <= T - 2 && c <= g.ncells - 2;
    // frozen physical boundary cells, main.jl:219-220 (halo cells of a slab belong to the neighbour)
    const bool own_frozen = valid && !lost && ((c == 0 && !(g.ghost & 1)) || (c == g.ncells - 1 && !(g.ghost & 2)));
    const double dt = scv[0], upd = scv[1], lambda = scv[2], t_cur = scv[3];
    const bool active = t_cur < g.t_end;                  // while t < T, main.jl:202
    (void)lambda;

    double q[15], F[15];   // own record and the numerical flux through the face left of it (slots 2..14)
    double lamv = 0.0;
    if (active) {
      const int L = (tid >= 1) ? tid - 1 : tid;           // thread 0 evaluates a dummy face against itself
      int fbad = 0;
      double s_l = 0.0, s_r = 0.0, inv_ds = 0.0, k_q = 0.0;
      if (FLUX == FLUX_HLL) {
        // wave-speed bounds at Q_m = (Q_l + Q_r)/2, NumFluxes.jl:86-91
        double m[3], A[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) m[i] = SQ(2 + i, L) + SQ(2 + i, tid);
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = SQ(6 + i, L) + SQ(6 + i, tid);
        PhaseState sm;   // state at the mean of the two records (the halving is folded into phase_state, exactly)
        phase_state<GEN, true, true, true>(eos, 1.0, m, SQ(5, L) + SQ(5, tid), A, sm);
        fbad = sm.bad;
        const double cm = phase_cmax<true>(eos, sm);
        const double lo_m = sm.u[0] - cm, hi_m = sm.u[0] + cm;
        // cached bounds of the two cells (the `eigvals` argument, NumFluxes.jl:90-91); with the c_max row: u1 -+ c_max
        const double lo_l = SP_CROW ? __dmul_rn(SQ(2, L), SA(SP_R_ID, L)) - SA(0, L) : SA(0, L);
        const double hi_r = SP_CROW ? __dmul_rn(SQ(2, tid), SA(SP_R_ID, tid)) + SA(0, tid) : SA(1, tid);
        s_l = fmin(0.0, fmin(lo_m, lo_l));
        s_r = fmax(0.0, fmax(hi_m, hi_r));
        inv_ds = hs_rcp(s_r - s_l);
        k_q = s_l * s_r * inv_ds;
        s_l *= inv_ds; s_r *= inv_ds;                     // weights of F_l and F_r in the HLL flux
      }
      // physical flux of both cells from their cached 1/rho and stress row (flux, Hyperelasticity.jl:99-114)
      double ra[15], fa[15], fb[15];
#pragma unroll
      for (int j = 2; j < 15; ++j) { ra[j] = SQ(j, L); q[j] = SQ(j, tid); }
      {
        const double sl[3] = {SA(SP_R_SG, L), SA(SP_R_SG + 1, L), SA(SP_R_SG + 2, L)};
        const double sr[3] = {SA(SP_R_SG, tid), SA(SP_R_SG + 1, tid), SA(SP_R_SG + 2, tid)};
        sp_flux_cached(ra, SA(SP_R_ID, L), sl, fa);
        sp_flux_cached(q, SA(SP_R_ID, tid), sr, fb);
      }
#pragma unroll
      for (int j = 2; j < 15; ++j) {
        if (flux_is_zero(j)) {   // rho F_1j: the physical flux is identically zero (u1 A_1j - u1 A_1j); only the jump term is left
          F[j] = (FLUX == FLUX_HLL) ? k_q * (q[j] - ra[j]) : -(0.5 * lambda) * (q[j] - ra[j]);
          continue;
        }
        const double Fa = fa[j], Fb = fb[j];
        if (FLUX == FLUX_HLL) F[j] = fma(s_r, Fa, fma(-s_l, Fb, k_q * (q[j] - ra[j])));      // NumFluxes.jl:78
        else F[j] = 0.5 * (Fa + Fb) - 0.5 * lambda * (q[j] - ra[j]);                          // NumFluxes.jl:30
      }
      if (valid && tid >= 1) bad |= fbad;
      if (lane == 0 && warp > 0) {
        const unsigned hw = hb0 + (unsigned)(k & 1) * ((T / 32) * 13u * 8u);
#pragma unroll
        for (int j = 2; j < 15; ++j) tma::sts_f64(hw + 8u * (j - 2), F[j]);
      }
      if (own_frozen) {
#pragma unroll
        for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = q[sp_slot(v)];
#pragma unroll
        for (int r = 0; r < SP_NAX; ++r) g.aux_out[(size_t)r * g.stride + gi] = SA(r, tid);
        if (SP_CROW) {
          const double u1 = __dmul_rn(q[2], SA(SP_R_ID, tid));
          lamv = fmax(fabs(u1 - SA(0, tid)), fabs(u1 + SA(0, tid)));
        } else {
          lamv = fmax(fabs(SA(0, tid)), fabs(SA(1, tid)));
        }
      }
    } else if (own_interior || own_frozen) {   // this problem already reached t_end: carry it through unchanged
#pragma unroll
      for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = SQ(sp_slot(v), tid);
#pragma unroll
      for (int r = 0; r < SP_NAX; ++r) g.aux_out[(size_t)r * g.stride + gi] = SA(r, tid);
    }
    if (tid == 32 && new_prob) write_scalars(sc + ((sci + 1) % 3) * 8, lam_n, t_n);
    if (tid == 0) {
      if (!active && tile == 0) {
        g.tt[(size_t)g.nxt * g.nprob + prob] = t_cur;
        g.lam[(size_t)g.nxt * g.nprob + prob] = (unsigned long long)__double_as_longlong(scv[4]);
        g.lam[(size_t)g.clr * g.nprob + prob] = 0ull;
      }
    }
#undef SQ
#undef SA
    __syncthreads();   // the only barrier of a tile: boundary fluxes visible, stage s free for the copy after next

    if (active) {
      // ---- conservative update (update_cell, main.jl:59 / :40) + wave bounds of the new state ----
      double qn[15];
      const double upd_own = own_interior ? upd : 0.0;   // halo / boundary cells keep their state: q - 0 * (finite) = q
      const bool edge = lane == 31 && warp < T / 32 - 1;  // the right neighbour sits in the next warp
      const unsigned hr = hb0 + (unsigned)(k & 1) * ((T / 32) * 13u * 8u) + 13u * 8u;
#pragma unroll
      for (int j = 2; j < 15; ++j) {
        double Fr = __shfl_down_sync(FULL, F[j], 1);
        if (edge) Fr = tma::lds_f64(hr + 8u * (j - 2));
        qn[j] = q[j] - upd_own * (Fr + (-F[j]));
      }
      if (own_interior) {
#pragma unroll
        for (int v = 0; v < 13; ++v) g.Qout[(size_t)v * g.stride + gi] = qn[sp_slot(v)];
      }
      // CFL sweep of the next step (get_eigvals, main.jl:204-211) on the state just produced
      PhaseState sn;
      phase_state<GEN, true, true>(eos, 1.0, qn + 2, qn[5], qn + 6, sn);
      const double cn = phase_cmax<true>(eos, sn);
      const double lo_n = sn.u[0] - cn, hi_n = sn.u[0] + cn;
      if (own_interior) {
        bad |= sn.bad;
        if (SP_CROW) g.aux_out[gi] = cn;
        else { g.aux_out[gi] = lo_n; g.aux_out[g.stride + gi] = hi_n; }
        g.aux_out[(size_t)SP_R_ID * g.stride + gi] = sn.inv_den;
#pragma unroll
        for (int r = 0; r < 3; ++r) g.aux_out[(size_t)(SP_R_SG + r) * g.stride + gi] = sn.sig1[r];
        lamv = fabs(sn.u[0]) + cn;                        // = max(|lo_n|, |hi_n|), bit for bit (rounding is monotone and odd)
      }
      lam_run = fmax(lam_run, lamv);
      if (tile == 0 && tid == 0) {
        g.tt[(size_t)g.nxt * g.nprob + prob] = t_cur + dt;   // main.jl:214
        g.steps[prob] += 1;                                   // main.jl:215
        hist_record(g, prob, dt);
        g.lam[(size_t)g.clr * g.nprob + prob] = 0ull;
      }
    }
    // ---- max(lambda) of this block's share of the problem: one atomic when the problem changes ----
    if (!has_next || (!SINGLE && probn != prob)) {
      lam_run = warp_max_nonneg(lam_run);
      if (lane == 0) red[warp] = lam_run;
      bad = __any_sync(FULL, bad);
      if (bad && lane == 0) atomicOr(g.status, 1);
      __syncthreads();
      if (tid == 0 && active) {
        double mx = red[0];
#pragma unroll
        for (int w = 1; w < T / 32; ++w) mx = fmax(mx, red[w]);
        atomicMax(g.lam + (size_t)g.nxt * g.nprob + prob, (unsigned long long)__double_as_longlong(mx));
      }
      lam_run = 0.0;
      bad = 0;
    }
    if (!has_next) break;
    if (new_prob) sci = (sci + 1) % 3;
    id = idn; tile = tilen; cur_tma = next_tma;
    if (SINGLE) off = (long long)(pbase + tile * (T - 2));
    else { prob = probn; off = offn; }
  }
}

// ------------------------------------------------------------------------------------------------
// CFL sweep on a resident state (main.jl:204-212): lo/hi per cell, lambda_max per problem, and
// optionally the full get_eigvals output (6 per phase) in Julia layout (6*NPH, ncells*nprob).
// ------------------------------------------------------------------------------------------------
template <int MODEL, bool GEN, int T>
__global__ void __launch_bounds__(T) k_bounds(const double* __restrict__ Q, double* __restrict__ aux,
                                              unsigned long long* lam_slot, double* eig_full, int* status, long long stride,
                                              int ncells, int nprob, int tiles_per_prob, const EosPair eosp) {
  using MT = ModelTraits<MODEL>;
  constexpr int NPH = MT::NPH, CPB = T / NPH;
  __shared__ double red[T / 32];
  const int tid = threadIdx.x, l = tid / NPH, ph = tid % NPH;
  const int prob = blockIdx.x / tiles_per_prob, tile = blockIdx.x % tiles_per_prob;
  const int c = tile * CPB + l;
  const bool valid = c < ncells;
  const long long gi = (long long)prob * ncells + (valid ? c : ncells - 1);
  const EosDev& eos = eosp.e[ph];
  double rec[15];
  if (MODEL == MODEL_MPH30) {
#pragma unroll
    for (int j = 0; j < 15; ++j) rec[j] = __ldg(Q + (size_t)(15 * ph + j) * stride + gi);
  } else {
    rec[0] = 1.0; rec[1] = 0.0;
#pragma unroll
    for (int v = 0; v < 13; ++v) rec[sp_slot(v)] = __ldg(Q + (size_t)v * stride + gi);
  }
  // The same instruction sequence as the CFL sweep at the tail of the fused step kernels (phase_state<.., WITH_H> + the Newton
  // largest-eigenvalue solve), so that the cache rows of an uploaded state are BIT-IDENTICAL to the rows the step that produced
  // that state left behind: a host-resident loop (upload + step + download per step) then reproduces the device-resident loop bit
  // for bit.  Only the full get_eigvals output (all three speeds) takes the Jacobi solve.
  PhaseState st;
  phase_state<GEN, MODEL == MODEL_SP13, true>(eos, (MODEL == MODEL_MPH30) ? rec[0] : 1.0, rec + 2, rec[5], rec + 6, st);
  double ev[3] = {0.0, 0.0, 0.0};
  double cm;
  if (eig_full) {
    double S6[6];
    phase_acoustic_sym<true>(eos, st, S6);
    sym3_eigs_jacobi(S6, ev);
    cm = sqrt(fmax(fabs(ev[2]), fabs(ev[0])));
  } else {
    cm = phase_cmax<true>(eos, st);
  }
  double lo_c = st.u[0] - cm, hi_c = st.u[0] + cm;
  if (NPH == 2) {
    lo_c = fmin(lo_c, __shfl_xor_sync(FULL, lo_c, 1));
    hi_c = fmax(hi_c, __shfl_xor_sync(FULL, hi_c, 1));
  }
  double lamv = 0.0;
  int bad = 0;
  if (valid) {
    bad = st.bad;
    if (SP_CROW && MODEL == MODEL_SP13) aux[gi] = cm;
    else if (ph == 0) { aux[gi] = lo_c; aux[stride + gi] = hi_c; }
    if (MODEL == MODEL_SP13) {
      aux[(size_t)SP_R_ID * stride + gi] = st.inv_den;
#pragma unroll
      for (int r = 0; r < 3; ++r) aux[(size_t)(SP_R_SG + r) * stride + gi] = st.sig1[r];
    }
    lamv = fmax(fabs(lo_c), fabs(hi_c));   // (single-phase: == |u1| + c_max bit for bit, as the step kernel forms it)
    if (eig_full) {  // [u1 + c_k (ascending), u1 - c_k], HyperelasticityMPh.jl:263-265
      double* e = eig_full + (size_t)gi * (6 * NPH) + 6 * ph;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double ck = sqrt(fabs(ev[k]));
        e[k] = st.u[0] + ck;
        e[3 + k] = st.u[0] - ck;
      }
    }
  }
  lamv = warp_max_nonneg(lamv);
  if ((tid & 31) == 0) red[tid >> 5] = lamv;
  bad = __any_sync(FULL, bad);
  if (bad && (tid & 31) == 0) atomicOr(status, 1);
  __syncthreads();
  if (tid == 0) {
    double mx = red[0];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) mx = fmax(mx, red[w]);
    atomicMax(lam_slot + prob, (unsigned long long)__double_as_longlong(mx));
  }
}

// ------------------------------------------------------------------------------------------------
// Julia layout (nvar, n) <-> structure of arrays [nvar][stride]; both sides coalesced through a
// padded shared-memory tile of TC cells.
// ------------------------------------------------------------------------------------------------
template <int NVAR, int TC>
__global__ void __launch_bounds__(256) k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, long long n, long long stride) {
  __shared__ double tile[TC * NVAR + TC];  // [cell][var] with pad
  const long long c0 = (long long)blockIdx.x * TC;
  const int ncell = (int)((n - c0 < TC) ? (n - c0) : TC);
  const int tot = ncell * NVAR;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int cell = i / NVAR, v = i % NVAR;
    tile[cell * (NVAR + 1) + v] = __ldg(aos + c0 * NVAR + i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NVAR * TC; i += blockDim.x) {
    const int v = i / TC, cell = i % TC;
    if (cell < ncell) soa[(size_t)v * stride + c0 + cell] = tile[cell * (NVAR + 1) + v];
  }
}

template <int NVAR, int TC>
__global__ void __launch_bounds__(256) k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, long long n, long long stride) {
  __shared__ double tile[TC * NVAR + TC];
  const long long c0 = (long long)blockIdx.x * TC;
  const int ncell = (int)((n - c0 < TC) ? (n - c0) : TC);
  for (int i = threadIdx.x; i < NVAR * TC; i += blockDim.x) {
    const int v = i / TC, cell = i % TC;
    if (cell < ncell) tile[cell * (NVAR + 1) + v] = __ldg(soa + (size_t)v * stride + c0 + cell);
  }
  __syncthreads();
  const int tot = ncell * NVAR;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int cell = i / NVAR, v = i % NVAR;
    aos[c0 * NVAR + i] = tile[cell * (NVAR + 1) + v];
  }
}

// Halo pack / unpack for the slab decomposition: buf = [Q(nvar), aux(naux)] of one cell.
// pack: first owned cell (index 1) -> to_left, last owned (ncells-2) -> to_right.
// unpack: from_left -> cell 0, from_right -> cell ncells-1.   mask bit 0 / 1 = left / right neighbour exists.
__global__ void k_halo(double* Q, double* aux, double* left, double* right, long long stride, int ncells, int nvar, int naux,
                       int mask, int unpack) {
  const int v = threadIdx.x;
  if (v >= nvar + naux) return;
  const int side = blockIdx.x;   // 0 left, 1 right
  if (!(mask & (1 << side))) return;
  double* buf = side ? right : left;
  const long long cell = unpack ? (side ? ncells - 1 : 0) : (side ? ncells - 2 : 1);
  double* p = v < nvar ? Q + (size_t)v * stride + cell : aux + (size_t)(v - nvar) * stride + cell;
  if (unpack) *p = buf[v]; else buf[v] = *p;
}

// ------------------------------------------------------------------------------------------------
// k_exchange_p2p: the per-step exchange of a slab-decomposed grid in ONE small kernel over NVLink
// peer memory (no NCCL call on the step path).  Every rank owns a mailbox in symmetric memory that
// all peers can address:
//   mbox[parity] = { from_left[W], from_right[W], lam[8], flag[8] }     W = nvar + naux, parity = seq & 1
// post:   my first / last owned cell -> the neighbours' from_right / from_left, my local lambda_max
//         -> lam[rank] of EVERY peer, __threadfence_system(), then flag[rank] = seq on every peer;
// wait:   spin until flag[p] == seq for every peer p of my own mailbox;
// unpack: halo cells <- from_left / from_right, lambda_max slot <- max_p lam[p].
// max is exact and nothing else crosses ranks, so the result is bit-identical to the NCCL path and to
// the single-GPU run.  Two parities suffice: a rank can only post step n+1 after every peer has
// posted step n (it waited for their flags), and no peer can post step n+2 before it has my n+1 flag.
// ------------------------------------------------------------------------------------------------
constexpr int MBOX_MAXW = 40, MBOX_MAXR = 8;
constexpr int MBOX_STRIDE = 2 * MBOX_MAXW + 2 * MBOX_MAXR;    // doubles per parity
struct PeerPtrs { double* p[MBOX_MAXR]; };

__global__ void __launch_bounds__(64) k_exchange_p2p(double* Q, double* aux, unsigned long long* lam_slot, const PeerPtrs peers,
                                                     long long stride, int ncells, int nvar, int naux, int rank, int world,
                                                     unsigned long long seq, int* status, unsigned long long timeout_ns) {
  const int v = threadIdx.x, W = nvar + naux;
  const int par = (int)(seq & 1ull);
  const size_t base = (size_t)par * MBOX_STRIDE;
  // ---- post ---------------------------------------------------------------------------------
  if (v < W) {
    const double* src = v < nvar ? Q + (size_t)v * stride : aux + (size_t)(v - nvar) * stride;
    if (rank > 0) peers.p[rank - 1][base + MBOX_MAXW + v] = src[1];                  // my first owned cell -> left peer's from_right
    if (rank < world - 1) peers.p[rank + 1][base + v] = src[ncells - 2];             // my last owned cell -> right peer's from_left
  }
  if (v < world) reinterpret_cast<unsigned long long*>(peers.p[v] + base + 2 * MBOX_MAXW)[rank] = *lam_slot;
  __threadfence_system();
  __syncthreads();
  if (v < world) {
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(peers.p[v] + base + 2 * MBOX_MAXW + MBOX_MAXR);
    f[rank] = seq;
  }
  // ---- wait ---------------------------------------------------------------------------------
  double* mine = peers.p[rank] + base;
  if (v < world) {
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(mine + 2 * MBOX_MAXW + MBOX_MAXR);
    const unsigned long long t0 = global_timer_ns();
    while (f[v] != seq) {
      // a peer that died never posts: give up instead of hanging the GPU, and say so in the status word
      if (global_timer_ns() - t0 > timeout_ns) { atomicOr(status, 2); break; }
    }
  }
  __threadfence_system();
  __syncthreads();
  // ---- unpack -------------------------------------------------------------------------------
  if (v < W) {
    double* dst = v < nvar ? Q + (size_t)v * stride : aux + (size_t)(v - nvar) * stride;
    if (rank > 0) dst[0] = __ldcv(mine + v);
    if (rank < world - 1) dst[ncells - 1] = __ldcv(mine + MBOX_MAXW + v);
  }
  if (v == 0) {
    const unsigned long long* l = reinterpret_cast<const unsigned long long*>(mine + 2 * MBOX_MAXW);
    unsigned long long m = 0ull;
    for (int p = 0; p < world; ++p) { const unsigned long long x = __ldcv(l + p); m = x > m ? x : m; }
    *lam_slot = m;
  }
}

// ------------------------------------------------------------------------------------------------
// Stateless batches over Julia-layout arrays: literal drop-ins for the per-cell / per-face
// reference functions.  One thread per (item, phase).
// ------------------------------------------------------------------------------------------------
enum { OP_CONS2PRIM = 0, OP_PRIM2CONS = 1, OP_FLUX = 2, OP_NONCONS = 3 };

// energy(eos, S, finger(F)) for an arbitrary (rho-independent) F: prim2cons, HyperelasticityMPh.jl:75-76
template <bool GEN>
__device__ __forceinline__ double energy_of_F(const EosDev& eos, double S, const double* F) {
  const double C11 = F[4] * F[8] - F[7] * F[5], C12 = F[7] * F[2] - F[1] * F[8], C13 = F[1] * F[5] - F[4] * F[2];
  const double C21 = F[6] * F[5] - F[3] * F[8], C22 = F[0] * F[8] - F[6] * F[2], C23 = F[3] * F[2] - F[0] * F[5];
  const double C31 = F[3] * F[7] - F[6] * F[4], C32 = F[6] * F[1] - F[0] * F[7], C33 = F[0] * F[4] - F[3] * F[1];
  const double det = F[0] * C11 + F[3] * C12 + F[6] * C13;
  const double id = 1.0 / det, k2 = id * id;   // G = C C^T / det^2,  I3 = 1/det^2
  const double G0 = k2 * (C11 * C11 + C12 * C12 + C13 * C13), G1 = k2 * (C11 * C21 + C12 * C22 + C13 * C23);
  const double G2 = k2 * (C11 * C31 + C12 * C32 + C13 * C33), G3 = k2 * (C21 * C21 + C22 * C22 + C23 * C23);
  const double G4 = k2 * (C21 * C31 + C22 * C32 + C23 * C33), G5 = k2 * (C31 * C31 + C32 * C32 + C33 * C33);
  const double I1 = G0 + G3 + G5;
  const double I2 = 0.5 * (I1 * I1 - (G0 * G0 + G3 * G3 + G5 * G5 + 2.0 * (G1 * G1 + G2 * G2 + G4 * G4)));
  const double r = fabs(id);  // I3^(1/2)
  double rA, rB, rC;
  if (GEN) { const double L = log(r); rA = exp(eos.ea * L); rB = exp(eos.eb * L); rC = exp(eos.eg * L); }
  else { rA = r; rB = r * r * r; rC = r * r; }
  const double U = eos.kA * (rA - 1.0) * (rA - 1.0) + eos.cvt0 * rC * (exp(S / eos.cv) - 1.0);   // EquationsOfState.jl:129-132
  const double W = eos.hb * rB * (I1 * I1 * (1.0 / 3.0) - I2);                                     // :134
  return U + W;
}

// Hank2016 (EquationsOfState.jl:301-356) as stateless batches: one thread per item; the (NT, n) Julia-layout tensor
// argument (NT = 9 entries of G / of the distortion, or 3 invariants) is staged through shared memory so that the
// global loads and the stress stores are contiguous.
enum { HANK_ENERGY = 0, HANK_PRESSURE = 1, HANK_STRESS = 2 };
template <int OP>
__global__ void __launch_bounds__(128) k_hank(const HankAbi eos, const double* __restrict__ s0, const double* __restrict__ s1,
                                              const double* __restrict__ ten, double* __restrict__ out, long long n, int* status) {
  constexpr int NT = (OP == HANK_PRESSURE) ? 3 : 9, NO = (OP == HANK_STRESS) ? 9 : 1;
  __shared__ double sh[128 * 9];
  const long long base = (long long)blockIdx.x * 128;
  const int cnt = (int)((n - base) < 128 ? (n - base) : 128);
  for (int k = threadIdx.x; k < cnt * NT; k += 128) sh[k] = ten[base * NT + k];
  __syncthreads();
  const int l = threadIdx.x;
  const bool valid = l < cnt;
  double x[9];
#pragma unroll
  for (int k = 0; k < NT; ++k) x[k] = sh[(valid ? l : 0) * NT + k];
  const double a0 = valid ? s0[base + l] : 1.0;
  const double a1 = (valid && OP != HANK_STRESS) ? s1[base + l] : 0.0;   // the pressure argument of stress() does not enter
  int bad = 0;
  double y[9];
  if (OP == HANK_ENERGY) y[0] = hank_energy(eos, a0, a1, x, &bad);
  else if (OP == HANK_PRESSURE) y[0] = hank_pressure(eos, a0, a1, x, &bad);
  else hank_stress(eos, a0, x, y, &bad);
  __syncthreads();
  if (valid) {
#pragma unroll
    for (int k = 0; k < NO; ++k) sh[l * NO + k] = y[k];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < cnt * NO; k += 128) out[base * NO + k] = sh[k];
  if (valid && bad) atomicOr(status, 1);
}

template <int MODEL, bool GEN, int OP>
__global__ void __launch_bounds__(128) k_cellop(const double* __restrict__ in, double* __restrict__ out, long long n,
                                                const EosPair eosp, int* status) {
  using MT = ModelTraits<MODEL>;
  constexpr int NPH = MT::NPH, NVAR = MT::NVAR;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / NPH;
  const int ph = (int)(t % NPH);
  const bool valid = i < n;
  const long long ii = valid ? i : n - 1;
  const EosDev& eos = eosp.e[ph];
  const double* x = in + ii * NVAR + 15 * ph;   // SP: ph == 0
  double* y = out + ii * NVAR + 15 * ph;
  int bad = 0;
  if (OP == OP_PRIM2CONS) {
    if (MODEL == MODEL_MPH30) {   // HyperelasticityMPh.jl:66-87
      const double frac = x[0], den = frac * x[1];
      const double e_tot = energy_of_F<GEN>(eos, x[5], x + 6) + 0.5 * (x[2] * x[2] + x[3] * x[3] + x[4] * x[4]);
      if (valid) {
        y[0] = frac; y[1] = den;
        y[2] = den * x[2]; y[3] = den * x[3]; y[4] = den * x[4];
        y[5] = den * e_tot;
        for (int k = 0; k < 9; ++k) y[6 + k] = den * x[6 + k];
      }
    } else {                      // Hyperelasticity.jl:70-93: P = [u(3), F(9 row-major), S]
      double Fc[9];
      for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) Fc[r + 3 * cc] = x[3 + 3 * r + cc];
      const double det = Fc[0] * (Fc[4] * Fc[8] - Fc[7] * Fc[5]) + Fc[3] * (Fc[7] * Fc[2] - Fc[1] * Fc[8]) + Fc[6] * (Fc[1] * Fc[5] - Fc[4] * Fc[2]);
      const double den = eos.rho0 / det;
      const double e_tot = energy_of_F<GEN>(eos, x[12], Fc) + 0.5 * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      if (valid) {
        y[0] = den * x[0]; y[1] = den * x[1]; y[2] = den * x[2];
        for (int k = 0; k < 9; ++k) y[3 + k] = den * x[3 + k];
        y[12] = den * e_tot;
      }
    }
    return;
  }
  double rec[15];
  if (MODEL == MODEL_MPH30) { for (int j = 0; j < 15; ++j) rec[j] = x[j]; }
  else { rec[0] = 1.0; rec[1] = 0.0; for (int v = 0; v < 13; ++v) rec[sp_slot(v)] = x[v]; }
  PhaseState st;
  phase_state<GEN>(eos, (MODEL == MODEL_MPH30) ? rec[0] : 1.0, rec + 2, rec[5], rec + 6, st);
  bad = valid ? st.bad : 0;
  if (OP == OP_CONS2PRIM) {       // HyperelasticityMPh.jl:106-133
    const double S = eos.cv * log(st.Sp);   // EquationsOfState.jl:155
    if (valid) {
      if (MODEL == MODEL_MPH30) {
        y[0] = rec[0]; y[1] = st.rho; y[2] = st.u[0]; y[3] = st.u[1]; y[4] = st.u[2]; y[5] = S;
        for (int k = 0; k < 9; ++k) y[6 + k] = rec[6 + k] * st.inv_den;
      } else {
        y[0] = st.u[0]; y[1] = st.u[1]; y[2] = st.u[2];
        for (int v = 3; v < 12; ++v) y[v] = rec[sp_slot(v)] * st.inv_den;
        y[12] = S;
      }
    }
  } else if (OP == OP_FLUX) {     // HyperelasticityMPh.jl:146-175 / Hyperelasticity.jl:99-114
    double f[15];
    phase_flux(st, rec + 6, f);
    if (valid) {
      if (MODEL == MODEL_MPH30) { for (int j = 0; j < 15; ++j) y[j] = f[j]; }
      else { for (int v = 0; v < 13; ++v) y[v] = f[sp_slot(v)]; }
    }
  } else if (OP == OP_NONCONS) {  // HyperelasticityMPh.jl:178-250 (column 1 of each block)
    double cc[15];
    noncons_column(st, rec + 6, cc);
    if (valid) for (int j = 0; j < 15; ++j) y[j] = cc[j];
  }
  bad = __any_sync(FULL, bad);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

// hll / lxf over a batch of faces (NumFluxes.jl:25-132).  eig_l / eig_r (6*NPH, n): the cached
// get_eigvals of the adjacent cells; only their min / max are read (NumFluxes.jl:90-91).
template <int MODEL, int FLUX, bool GEN, int T>
__global__ void __launch_bounds__(T) k_faceop(const double* __restrict__ Ql, const double* __restrict__ Qr,
                                              const double* __restrict__ eig_l, const double* __restrict__ eig_r,
                                              double lambda, double* cons, double* dm, double* dp, double* s_out,
                                              long long n, const EosPair eosp, int* status) {
  using MT = ModelTraits<MODEL>;
  constexpr int NPH = MT::NPH, NVAR = MT::NVAR, J0 = MT::J0;
  extern __shared__ double smem[];
  double* Ra = smem; double* Rb = Ra + 15 * T; double* Fa = Rb + 15 * T; double* Fb = Fa + 15 * T; double* Hs = Fb + 15 * T;
  const int tid = threadIdx.x, ph = tid % NPH;
  const long long i = ((long long)blockIdx.x * T + tid) / NPH;
  const bool valid = i < n;
  const long long ii = valid ? i : n - 1;
  const EosDev& eos = eosp.e[ph];
  int bad = 0;
  for (int side = 0; side < 2; ++side) {
    const double* x = (side ? Qr : Ql) + ii * NVAR + 15 * ph;
    double* R = (side ? Rb : Ra) + tid;
    double* F = (side ? Fb : Fa) + tid;
    if (MODEL == MODEL_MPH30) { for (int j = 0; j < 15; ++j) R[j * T] = x[j]; }
    else { R[0] = 1.0; R[T] = 0.0; for (int v = 0; v < 13; ++v) R[sp_slot(v) * T] = x[v]; }
    PhaseState st;
    column_state<MODEL, GEN, T>(eos, R, st);
    bad |= st.bad;
    double A[9], f[15];
    for (int k = 0; k < 9; ++k) A[k] = R[(6 + k) * T];
    phase_flux(st, A, f);
    for (int j = 1; j < 15; ++j)
      if (!flux_is_zero(j)) F[flux_row(j) * T] = f[j];
  }
  double lo_l = 0.0, hi_r = 0.0;
  if (FLUX == FLUX_HLL) {
    const double* el = eig_l + ii * (6 * NPH);
    const double* er = eig_r + ii * (6 * NPH);
    lo_l = el[0]; hi_r = er[0];
    for (int k = 1; k < 6 * NPH; ++k) { lo_l = fmin(lo_l, el[k]); hi_r = fmax(hi_r, er[k]); }
  }
  double sl_sr[2] = {0.0, 0.0};
  const long long base = ii * NVAR + 15 * ph;
  auto emit = [&](int j, double c_, double dm_, double dp_) {
    if (!valid) return;
    const int v = (MODEL == MODEL_MPH30) ? j : (j < 5 ? j - 2 : (j == 5 ? 12 : 3 + 3 * ((j - 6) % 3) + (j - 6) / 3));
    if (cons) cons[base + v] = c_;
    if (dm) dm[base + v] = dm_;
    if (dp) dp[base + v] = dp_;
  };
  face_eval<MODEL, FLUX, GEN, T>(eos, Ra + tid, Rb + tid, Fa + tid, Fb + tid, lo_l, hi_r, lambda, Hs + tid, bad, sl_sr, emit);
  if (valid && s_out && ph == 0) { s_out[2 * ii] = sl_sr[0]; s_out[2 * ii + 1] = sl_sr[1]; }
  bad = __any_sync(FULL, valid ? bad : 0);
  if (bad && (tid & 31) == 0) atomicOr(status, 1);
  (void)J0;
}

// ------------------------------------------------------------------------------------------------
// Dimension-split 2-D stepping.  The physics is frame-indifferent (u -> R u, F -> R F; oracle anchor
// "wave speeds rotate with the frame"), so a sweep along y IS the x-sweep applied to the state seen from a frame rotated by
// R e_2 = e_1: R = [[0,1,0],[-1,0,0],[0,0,1]], i.e. a signed permutation of the components -- exact in floating point.
// k_transpose_rot: per variable a tiled 2-D transpose ([rows][cols] -> [cols][rows], both coalesced through a padded
// shared-memory tile) with that signed permutation folded in, so that the rows of the result are the grid's columns in the
// rotated frame and the 1-D kernels run on them unchanged (as an ensemble of independent rows with one shared dt).
// ------------------------------------------------------------------------------------------------
struct RotMap { int src[30]; double sign[30]; int nvar; };
__global__ void __launch_bounds__(256) k_transpose_rot(const double* __restrict__ in, double* __restrict__ out, int rows, int cols,
                                                       const RotMap map) {
  __shared__ double tile[32][33];
  const int v = blockIdx.z;
  const double* src = in + (size_t)map.src[v] * rows * cols;
  double* dst = out + (size_t)v * rows * cols;
  const double sg = map.sign[v];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[k][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + threadIdx.x;      // out is [cols][rows]
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = sg * tile[threadIdx.x][k];
  }
}
// dt = min(cfl dx / max_rows lambda_x, cfl dy / max_cols lambda_y), clock and step count of the 2-D grid; one block.
// clock: [t, dt, steps (as double), lambda_x, lambda_y]
__global__ void __launch_bounds__(256) k_dt2d(const unsigned long long* __restrict__ lam_x, int nlx, const unsigned long long* __restrict__ lam_y,
                                              int nly, double cfl, double dx, double dy, double t_end, double* clock) {
  __shared__ unsigned long long red[2][256];
  unsigned long long mx = 0ull, my = 0ull;
  for (int i = threadIdx.x; i < nlx; i += 256) mx = max(mx, lam_x[i]);     // non-negative doubles order like their bit patterns
  for (int i = threadIdx.x; i < nly; i += 256) my = max(my, lam_y[i]);
  red[0][threadIdx.x] = mx; red[1][threadIdx.x] = my;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] = max(red[0][threadIdx.x], red[0][threadIdx.x + s]);
      red[1][threadIdx.x] = max(red[1][threadIdx.x], red[1][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double lx = __longlong_as_double((long long)red[0][0]), ly = __longlong_as_double((long long)red[1][0]);
    const double dtx = cfl * dx / lx, dty = cfl * dy / ly;
    const double t = clock[0];
    const bool active = t < t_end;                 // while t < T, main.jl:202 (a finished grid takes dt = 0 sweeps: no-ops)
    const double dt = active ? fmin(dtx, dty) : 0.0;
    clock[1] = dt;
    clock[0] = t + dt;
    clock[2] += active ? 1.0 : 0.0;
    clock[3] = lx; clock[4] = ly;
  }
}

// self-test of the branch-free device math (hs_rcp / hs_rsqrt / hs_sqrt / largest eigenvalue) on caller data
__global__ void k_selftest_math(const double* __restrict__ x, double* rcp, double* rsq, double* sq, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rcp[i] = hs_rcp(x[i]); rsq[i] = hs_rsqrt(x[i]); sq[i] = hs_sqrt(x[i]);
}
__global__ void k_selftest_eig(const double* __restrict__ s6, double* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = sym3_max_abs_eig(s6 + 6 * i);
}

// full get_eigvals over a Julia-layout batch (HyperelasticityMPh.jl:252-266)
struct Normal3 { double n[3]; };
template <int MODEL, bool GEN>
__global__ void __launch_bounds__(128) k_eigvals(const double* __restrict__ in, double* __restrict__ eig, long long n,
                                                 const EosPair eosp, const Normal3 nrm, int* status) {
  using MT = ModelTraits<MODEL>;
  constexpr int NPH = MT::NPH, NVAR = MT::NVAR;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / NPH;
  const int ph = (int)(t % NPH);
  const bool valid = i < n;
  const long long ii = valid ? i : n - 1;
  const EosDev& eos = eosp.e[ph];
  const double* x = in + ii * NVAR + 15 * ph;
  double rec[15];
  if (MODEL == MODEL_MPH30) { for (int j = 0; j < 15; ++j) rec[j] = x[j]; }
  else { rec[0] = 1.0; rec[1] = 0.0; for (int v = 0; v < 13; ++v) rec[sp_slot(v)] = x[v]; }
  PhaseState st;
  phase_state<GEN>(eos, (MODEL == MODEL_MPH30) ? rec[0] : 1.0, rec + 2, rec[5], rec + 6, st);
  double S6[6], ev[3];
  phase_acoustic_sym_n(eos, st, nrm.n, S6);
  sym3_eigs_jacobi(S6, ev);
  const double spd = st.u[0] * nrm.n[0] + st.u[1] * nrm.n[1] + st.u[2] * nrm.n[2];   // dot(P[3:5], n), HyperelasticityMPh.jl:264
  if (valid) {
    double* e = eig + ii * (6 * NPH) + 6 * ph;
    for (int k = 0; k < 3; ++k) { const double ck = sqrt(fabs(ev[k])); e[k] = spd + ck; e[3 + k] = spd - ck; }
  }
  int bad = __any_sync(FULL, valid ? st.bad : 0);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

}  // namespace hs
