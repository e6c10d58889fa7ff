/* examples/abi_smoke.c -- the C ABI driven from plain C, the way a Julia `ccall` (or any FFI) drives it.
 * Reproduces SURVEY.md Appendix B.5: two-phase test case 6, nx = 16, cfl 0.6, five HLL steps; checks the dt
 * history and two cells against the golden values.  No Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/abi_smoke.c -o /tmp/abi_smoke \
 *       -Lhyperelasticsolver_b200 -lhyperelastic_b200 -Wl,-rpath,$PWD/hyperelasticsolver_b200 -lm && /tmp/abi_smoke
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hyperelastic_b200.h"

static hs_barton2009_t barton2009(void) { /* EquationsOfState.jl:90-115 defaults */
  hs_barton2009_t e = {8.93, 4.6, 3.9e-4, 300.0, 2.1, 1.0, 3.0, 2.0, 0.0, 0.0};
  e.b0sq = e.b0 * e.b0;
  e.k0 = e.c0 * e.c0 - (4.0 / 3.0) * e.b0 * e.b0;
  return e;
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != HS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, hs_last_error()); return 1; } \
  } while (0)

int main(void) {
  enum { NX = 16, NVAR = 30, STEPS = 5 };
  hs_barton2009_t eos[2] = {barton2009(), barton2009()};
  /* primitive Riemann states of test case 6 (HyperelasticityMPh.jl:348-367): [alpha, rho, u(3), S, F(9 column-major)] x 2 */
  const double detl = 0.98, detr = 1.0;
  double P[2][NVAR] = {
      {0.1, 8.9 / detl, 0.0, 0.5, 1.0, 1.0e-3, 0.98, 0.02, 0.0, 0.0, 1.0, 0.0, 0.0, 0.1, 1.0,
       0.9, 8.9 / detl, 0.0, 0.5, 1.0, 1.0e-3, 0.98, 0.02, 0.0, 0.0, 1.0, 0.0, 0.0, 0.1, 1.0},
      {0.9, 8.9 / detr, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.1, 1.0,
       0.1, 8.9 / detr, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.1, 1.0}};
  double Qlr[2][NVAR];
  CHECK(hs_prim2cons(HS_MODEL_MPH30, eos, 2, &P[0][0], &Qlr[0][0], 2, 0));          /* prim2cons_mph */
  double Q[NX][NVAR];                                                                 /* Q0 :: Array{Float64}(30, nx) */
  for (int i = 0; i < NX; ++i)
    for (int v = 0; v < NVAR; ++v) Q[i][v] = (i < NX / 2.0) ? Qlr[0][v] : Qlr[1][v]; /* initial_condition, main.jl:99-106 */

  hs_ctx_t* ctx = NULL;
  CHECK(hs_create(&ctx, HS_MODEL_MPH30, eos, 2, NX, 1, 0));
  CHECK(hs_upload(ctx, &Q[0][0]));
  const double golden_dt[STEPS] = {0.006525871150502139, 0.006525871150502139, 0.006446700759844828, 0.005958085061464478,
                                   0.005706267106528782};
  double t = 0.0, worst = 0.0;
  for (int n = 0; n < STEPS; ++n) {                                                   /* while t < T, main.jl:202-227 */
    double dt = 0.0;
    CHECK(hs_step(ctx, HS_FLUX_HLL, 0.6, 1.0 / NX, &dt));
    t += dt;
    worst = fmax(worst, fabs(dt - golden_dt[n]) / golden_dt[n]);
  }
  CHECK(hs_download(ctx, &Q[0][0]));
  CHECK(hs_destroy(ctx));
  const double golden_c8[6] = {0.38297853292399647, 3.5431983793351409, 1.1321498895965438, 0.32995939011061942, 0.78173554155610159,
                               1.7603676372817543};
  for (int v = 0; v < 6; ++v) worst = fmax(worst, fabs(Q[7][v] - golden_c8[v]) / fabs(golden_c8[v]));
  printf("%s: t = %.17g after %d steps, max relative deviation from the golden values %.2e (%s)\n", hs_version(), t, STEPS, worst,
         worst < 1e-11 ? "ABI-SMOKE-OK" : "ABI-SMOKE-FAIL");
  return worst < 1e-11 ? 0 : 2;
}
