// Test harness only: compiles hs_phase.cuh as plain C++ so the closed forms the CUDA kernels
// use can be compared with the dual-number oracle on a machine without a GPU.
// Not part of the product; nothing in hyperelasticsolver_b200/ loads it.
#include "../../hyperelasticsolver_b200/csrc/hs_phase.cuh"
#include "../../hyperelasticsolver_b200/csrc/hs_hank.cuh"
using namespace hs;
extern "C" {
// out: rho, u(3), Etot, Sp, T, sig1(3), G(6), cmax, flux(15), S6(6), bad
void hm_phase(const double* eos_abi, int gen, double alpha, const double* m, double E, const double* A, double* out) {
  EosDev e = make_eos_dev(*reinterpret_cast<const EosAbi*>(eos_abi));
  PhaseState s;
  if (gen) phase_state<true>(e, alpha, m, E, A, s); else phase_state<false>(e, alpha, m, E, A, s);
  int k = 0;
  out[k++] = s.rho; for (int i = 0; i < 3; ++i) out[k++] = s.u[i];
  out[k++] = s.Etot; out[k++] = s.Sp; out[k++] = s.T;
  for (int i = 0; i < 3; ++i) out[k++] = s.sig1[i];
  for (int i = 0; i < 6; ++i) out[k++] = s.G[i];
  out[k++] = phase_cmax(e, s);
  phase_flux(s, A, out + k); k += 15;
  phase_acoustic_sym(e, s, out + k); k += 6;
  out[k++] = s.bad;
}
// row-1 flavour of phase_state (Cayley-Hamilton): out = rho, u(3), Etot, T, sig1(3), I1, J, G2r1(3), bad
void hm_phase_row1(const double* eos_abi, int gen, double alpha, const double* m, double E, const double* A, double* out) {
  EosDev e = make_eos_dev(*reinterpret_cast<const EosAbi*>(eos_abi));
  PhaseState s;
  if (gen) phase_state_row1<true>(e, alpha, m, E, A, s); else phase_state_row1<false>(e, alpha, m, E, A, s);
  int k = 0;
  out[k++] = s.rho; for (int i = 0; i < 3; ++i) out[k++] = s.u[i];
  out[k++] = s.Etot; out[k++] = s.T;
  for (int i = 0; i < 3; ++i) out[k++] = s.sig1[i];
  out[k++] = s.I1; out[k++] = s.J;
  for (int i = 0; i < 3; ++i) out[k++] = s.G2r1[i];
  out[k++] = s.bad;
}
// symmetrised acoustic tensor for a general unit normal
void hm_acoustic_n(const double* eos_abi, int gen, double alpha, const double* m, double E, const double* A, const double* n, double* S6) {
  EosDev e = make_eos_dev(*reinterpret_cast<const EosAbi*>(eos_abi));
  PhaseState s;
  if (gen) phase_state<true>(e, alpha, m, E, A, s); else phase_state<false>(e, alpha, m, E, A, s);
  phase_acoustic_sym_n(e, s, n, S6);
}
// Hank2016 closed forms (hs_hank.cuh); return value = domain flag
int hm_hank_energy(const double* eos, double den, double pres, const double* G, double* e) {
  int bad = 0; *e = hank_energy(*reinterpret_cast<const HankAbi*>(eos), den, pres, G, &bad); return bad; }
int hm_hank_pressure(const double* eos, double den, double e_int, const double* inv3, double* p) {
  int bad = 0; *p = hank_pressure(*reinterpret_cast<const HankAbi*>(eos), den, e_int, inv3, &bad); return bad; }
int hm_hank_stress(const double* eos, double den, const double* A, double* sig) {
  int bad = 0; hank_stress(*reinterpret_cast<const HankAbi*>(eos), den, A, sig, &bad); return bad; }
void hm_sym3_eigs(const double* a, double* ev) { sym3_eigs(a, ev); }
void hm_sym3_eigs_jacobi(const double* a, double* ev) { sym3_eigs_jacobi(a, ev); }
double hm_sym3_max_abs(const double* a) { return sym3_max_abs_eig(a); }
}
