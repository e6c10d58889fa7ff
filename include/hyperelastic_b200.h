/*
 * hyperelastic_b200.h -- C ABI of the B200-native finite-volume hot path of
 * BlackSiberian/HyperelasticSolver (libhyperelastic_b200.so).
 *
 * The reference has no FFI: its "operator API" is a set of pure Julia functions on heap
 * Vector{Float64}/Matrix{Float64}.  Every entry point below names the
 * reference interface it replaces (file:line in the reference tree).  A Julia driver reaches
 * them with `ccall` (julia/HyperelasticB200.jl); the Python package
 * hyperelasticsolver_b200 binds the same symbols with ctypes.
 *
 * Conventions
 *  - All host arrays are Float64, "Julia layout": an (nvar, n) column-major matrix, i.e. one
 *    cell / face per contiguous record of nvar doubles -- exactly Q0::Array{Float64}(30, nx)
 *    of main.jl:101.  The library transposes to structure-of-arrays on the device.
 *  - model HS_MODEL_MPH30: 2 phases x [alpha, alpha*rho, alpha*rho*u(3), alpha*rho*E,
 *    alpha*rho*F(9, column-major)] (HyperelasticityMPh.jl:80-84).  HS_MODEL_SP13:
 *    [rho*u(3), rho*F(9, row-major), rho*E] (Hyperelasticity.jl:81-91).
 *  - Every function returns an int status (HS_OK == 0).  HS_ERR_DOMAIN is returned where the
 *    Julia code would throw DomainError (sqrt/log of a negative: HyperelasticityMPh.jl:114,
 *    EquationsOfState.jl:129-134) or produce NaN; results are still written.
 *  - There is no CPU fallback: without a CUDA device every call returns HS_ERR_CUDA.
 *  - A context is used from one host thread at a time; calls return when results are valid.
 */
#ifndef HYPERELASTIC_B200_H
#define HYPERELASTIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* HS_ERR_EXCHANGE: the multi-GPU halo / max(lambda) exchange of a slab-decomposed grid did not complete (a peer stopped stepping:
 * the exchange kernel gave up after HS_EXCHANGE_TIMEOUT_S seconds instead of hanging the GPU).  The library itself never calls
 * NCCL: the one-process-per-GPU driver (hyperelasticsolver_b200/slab.py) may, as its fallback exchange, through torch.distributed. */
enum { HS_OK = 0, HS_ERR_ARG = 1, HS_ERR_CUDA = 2, HS_ERR_DOMAIN = 3, HS_ERR_EXCHANGE = 4 };
enum { HS_MODEL_SP13 = 0, HS_MODEL_MPH30 = 1 };
enum { HS_FLUX_LXF = 0, HS_FLUX_HLL = 1 };

/* Barton2009 -- EquationsOfState.jl:71-85, same field order (b0sq = b0^2, k0 = c0^2 - 4/3 b0^2,
 * EquationsOfState.jl:111-112).  One block per phase (phases may differ, main.jl:133). */
typedef struct hs_barton2009 {
  double rho0, c0, cv, t0, b0, alpha, beta, gamma, b0sq, k0;
} hs_barton2009_t;

typedef struct hs_ctx hs_ctx_t;

const char* hs_version(void);
/* message of the last failure on this thread ("" if none) */
const char* hs_last_error(void);
/* number of CUDA devices visible (0 => every compute call fails with HS_ERR_CUDA) */
int hs_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Stateful solver: replaces the body of the `while t < T` loop, main.jl:202-227.
 * ------------------------------------------------------------------------------------------ */

/* Allocate device state for `nprob` independent problems of `ncells` cells each on CUDA device
 * `device`.  eos: nphase blocks (2 for MPH30, 1 for SP13). */
int hs_create(hs_ctx_t** ctx, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells,
              int64_t nprob, int device);
/* The same over `ndev` (<= 8) devices of this process -- what a single-process driver such as
 * `julia main.jl` uses to reach an 8-GPU box.  All other calls (upload / step / advance / download ...)
 * are unchanged and the results are bit-identical to the single-device context.
 *   nprob == 1: ONE grid slab-decomposed over the devices (contiguous slabs, one halo cell per side); the
 *     devices must be able to access each other's memory (NVLink / PCIe peer access).  Every step is one fused
 *     kernel per device plus one small peer-memory exchange kernel per device (halo cells and max(lambda)),
 *     no NCCL, no host synchronisation.
 *   nprob  > 1: the independent problems are shared out over the devices; no exchange at all. */
int hs_create_multi(hs_ctx_t** ctx, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob,
                    const int* devices, int ndev);
int hs_destroy(hs_ctx_t* ctx);

/* Q0 of main.jl:101,179: (nvar, ncells, nprob) column-major.  Upload also evaluates the
 * per-cell wave speeds of the new state (the CFL sweep main.jl:204-211) and resets t, steps. */
int hs_upload(hs_ctx_t* ctx, const double* Q);
int hs_download(hs_ctx_t* ctx, double* Q);
/* set / get the clock of every problem (restart: main.jl:185-186) */
int hs_set_time(hs_ctx_t* ctx, double t, int64_t step);

/* CFL sweep main.jl:204-212: lambda_max[nprob] = max_i max|eigvals_i|; eig (6*nphase, ncells,
 * nprob) receives get_eigvals of every cell if not NULL (HyperelasticityMPh.jl:252-266,
 * per phase [u1+c_k, u1-c_k], c_k ascending). */
int hs_wave_speeds(hs_ctx_t* ctx, double* eig_or_null, double* lambda_max);

/* One time step, main.jl:204-227: dt = cfl*dx/lambda_max, boundary cells frozen
 * (main.jl:219-220), interior cells updated with update_cell (main.jl:30-60).
 * dt_out[nprob] (may be NULL). */
int hs_step(hs_ctx_t* ctx, int flux, double cfl, double dx, double* dt_out);

/* Device-resident loop: steps until t >= t_end (no clipping of the last step, main.jl:202,214)
 * or max_steps more steps were taken.  Per problem: t_io[nprob], step_io[nprob] in/out (NULL =
 * keep the context's own clock), dt_hist (max_steps, nprob) column-major or NULL. */
int hs_advance(hs_ctx_t* ctx, int flux, double cfl, double dx, double t_end, int64_t max_steps,
               double* t_io, int64_t* step_io, double* dt_hist);

/* One step on HOST arrays: upload Qin, step, download into Qout (may alias Qin).  This is the
 * literal drop-in for one iteration of main.jl:202-227 with Q0 living in Julia memory; the whole state crosses
 * the host link once in each direction per call.  One grid on one device is processed chunk by chunk with the
 * H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlapping (full-duplex link):
 * dt = cfl dx / max(lambda) needs the CFL sweep of the whole uploaded state, so the step runs speculatively with the
 * max(lambda) the previous call's fused step produced for the state it returned, the sweep of the uploaded data
 * confirms it bit for bit at the end, and on a mismatch (first call, state edited in between) the step is redone on
 * the device with the right dt -- the result is always bit-identical to hs_upload + hs_step + hs_download.
 * An ensemble on one device is pipelined the same way by groups of whole problems with per-problem hints (any refuted hint redoes
 * the step); contexts spanning several devices and odd cell counts take upload + step + download.
 * The copies only overlap from page-locked memory: register the arrays once with hs_host_register (a Julia Array,
 * a numpy array ... are pageable).  HS_HOST_PIPELINE=0 forces upload + step + download; HS_HOST_CHUNK = cells per chunk. */
int hs_step_host(hs_ctx_t* ctx, int flux, double cfl, double dx, const double* Qin, double* Qout,
                 double* dt_out);
/* page-lock / release a host array (cudaHostRegister, portable) so that hs_upload / hs_download / hs_step_host copy
 * at the link rate and asynchronously; registering twice / releasing an unregistered array is not an error */
int hs_host_register(void* ptr, size_t bytes);
int hs_host_unregister(void* ptr);
/* bookkeeping of hs_step_host on this context: calls that took the pipelined form, and how many of them found the
 * hinted max(lambda) confirmed (no redo) */
int hs_step_host_stats(hs_ctx_t* ctx, int64_t* pipelined_calls, int64_t* speculation_hits);

/* ------------------------------------------------------------------------------------------
 * Dimension-split 2-D solver on an nx x ny grid (the reference driver is 1-D, its physics takes a normal:
 * EquationsOfState.jl:223, HyperelasticityMPh.jl:264).  Q^{n+1} = Y(dt) X(dt) Q^n with X / Y the 1-D step of main.jl:204-227
 * along the grid rows / columns -- the y-sweep runs the same kernels on the state seen from the frame rotated by R e_2 = e_1
 * (u -> R u, F -> R F: a signed permutation, exact) --, dt = cfl min(dx / max lambda_x, dy / max lambda_y).  Each sweep is the 1-D
 * step on every grid line including its boundary rule (first / last cell of the line frozen, main.jl:219-220).
 * Q: (nvar, nx, ny) column-major (cell (i, j) = record i + nx j).  A grid that is uniform in y reproduces the 1-D solver in every
 * row (bit for bit, except that the two-phase HLL flux leaves ~1e-16 behind between equal states: its Q_hll is Q only to roundoff).
 * ------------------------------------------------------------------------------------------ */
typedef struct hs2d_ctx hs2d_ctx_t;
int hs2d_create(hs2d_ctx_t** ctx, int model, const hs_barton2009_t* eos, int nphase, int64_t nx, int64_t ny, int device);
int hs2d_destroy(hs2d_ctx_t* ctx);
int hs2d_upload(hs2d_ctx_t* ctx, const double* Q);       /* also resets the clock */
int hs2d_download(hs2d_ctx_t* ctx, double* Q);
int hs2d_step(hs2d_ctx_t* ctx, int flux, double cfl, double dx, double dy, double* dt_out);
/* steps while t < t_end (no clipping of the last step), at most max_steps; t_out / steps_out: clock of the grid since upload */
int hs2d_advance(hs2d_ctx_t* ctx, int flux, double cfl, double dx, double dy, double t_end, int64_t max_steps, double* t_out,
                 int64_t* steps_out);

/* ------------------------------------------------------------------------------------------
 * Stateless batches: literal drop-ins for the per-cell / per-face Julia functions.
 * n = number of cells (faces).  device = CUDA device ordinal.
 * ------------------------------------------------------------------------------------------ */
/* cons2prim_mph HyperelasticityMPh.jl:99-133 / prim2cons_mph :63-90.  SP13 primitives are
 * [u(3), F(9 row-major), S] = the arguments of prim2cons, Hyperelasticity.jl:70. */
int hs_cons2prim(int model, const hs_barton2009_t* eos, int nphase, const double* Q, double* P, int64_t n, int device);
int hs_prim2cons(int model, const hs_barton2009_t* eos, int nphase, const double* P, double* Q, int64_t n, int device);
/* flux_mph HyperelasticityMPh.jl:140-175 / flux Hyperelasticity.jl:99-114 */
int hs_flux(int model, const hs_barton2009_t* eos, int nphase, const double* Q, double* F, int64_t n, int device);
/* noncons_flux HyperelasticityMPh.jl:178-250.  col (30, n): column 1 of each 15x15 diagonal
 * block (the only non-zero entries, :223-230).  Bdense (30, 30, n) gets the full matrix if not NULL. */
int hs_noncons_flux(const hs_barton2009_t* eos, const double* Q, double* col, double* Bdense, int64_t n, int device);
/* get_eigvals HyperelasticityMPh.jl:252-266: eig (6*nphase, n), per phase [u.n + c_k, u.n - c_k].
 * normal: unit vector (3 doubles) or NULL for (1,0,0), the only normal the 1-D driver uses (main.jl:208). */
int hs_get_eigvals(int model, const hs_barton2009_t* eos, int nphase, const double* Q, const double* normal, double* eig,
                   int64_t n, int device);
/* hll NumFluxes.jl:70-132: Ql, Qr (30, n); eig_l, eig_r (12, n) = get_eigvals of the two cells
 * (the `eigvals` argument, main.jl:56-57).  cons = zeros (NumFluxes.jl:82), dm = D^-, dp = D^+,
 * s (2, n) = [s_l, s_r] (may be NULL).  For SP13 (nvar 13, eig 6 x n) cons is the conservative
 * HLL flux of NumFluxes.jl:75-78 and dm = dp = 0. */
int hs_hll(int model, const hs_barton2009_t* eos, int nphase, const double* Ql, const double* Qr,
           const double* eig_l, const double* eig_r, double* cons, double* dm, double* dp, double* s,
           int64_t n, int device);
/* lxf NumFluxes.jl:25-60 (lambda = dx/dt) */
int hs_lxf(int model, const hs_barton2009_t* eos, int nphase, const double* Ql, const double* Qr, double lambda,
           double* cons, double* dm, double* dp, int64_t n, int device);

/* ------------------------------------------------------------------------------------------
 * Hank2016 equation of state, EquationsOfState.jl:301-364.  Dead code in the
 * reference (nothing calls it; `energy` and `stress` pass a 3x3 Matrix to the Vector-only `invariants` /
 * `finger`, Strains.jl:26,46, and cannot run as written): built as the stateless batches the three
 * functions spell out, a 3x3 tensor being its 9 column-major entries.  The law is parametrised by
 * (density, pressure), not by entropy, so it does not plug into the step kernels (whose state carries S).
 * ------------------------------------------------------------------------------------------ */
/* struct Hank2016, EquationsOfState.jl:305-319 (same field order; defaults 2.7, 26e9, 3.4, 21.5e9, 0.5) */
typedef struct hs_hank2016 {
  double rho0, mu, gamma, pres_inf, a;
} hs_hank2016_t;
/* energy(eos::Hank2016, den, pres, G) :317-331: den[n], pres[n], G (9, n) -> e_int[n] */
int hs_hank2016_energy(const hs_hank2016_t* eos, const double* den, const double* pres, const double* G, double* e_int,
                       int64_t n, int device);
/* pressure(eos::Hank2016, den, e_int, i) :333-346: inv3 (3, n) = invariants(G), Strains.jl:46-52 -> pres[n] */
int hs_hank2016_pressure(const hs_hank2016_t* eos, const double* den, const double* e_int, const double* inv3, double* pres,
                         int64_t n, int device);
/* stress(eos::Hank2016, den, pressure, distortion) :348-356: distortion (9, n) -> sigma (9, n) = -2 den G de/dG with
 * G = finger(inv(distortion)).  pres may be NULL: the hydrodynamic energy does not depend on G. */
int hs_hank2016_stress(const hs_hank2016_t* eos, const double* den, const double* pres, const double* distortion, double* sigma,
                       int64_t n, int device);

/* Device self-test hooks used by tests/: the hot path's branch-free reciprocal / reciprocal square root /
 * square root (x > 0) and its largest-|eigenvalue| solve of symmetric 3x3 tensors s6 = [11,12,13,22,23,33] (6, n). */
int hs_selftest_math(const double* x, double* rcp, double* rsq, double* sq, int64_t n, int device);
int hs_selftest_eig(const double* s6, double* lam_max_abs, int64_t n, int device);

/* ------------------------------------------------------------------------------------------
 * Device-pointer layer (structure-of-arrays, caller-owned device memory and stream): what the
 * one-process-per-GPU driver uses so that halo exchange / allreduce (NCCL through
 * torch.distributed) can be enqueued on the same stream between steps.
 *   Q      : [nvar][stride] doubles, stride = ncells*nprob (cell index = prob*ncells + i)
 *   aux    : [HS_NAUX(model)][stride] cached per-cell rows.  Two-phase: 0/1 = wave bounds min/max over phases of
 *            u1 -+ c_max.  Single-phase: row 0 = c_max (the bounds are u1 -+ c_max with u1 = m1 * (1/rho), bit-identical
 *            to caching them), rows 1..4 = 1/rho and row 1 of the stress (what the next step's physical flux
 *            needs, so it can skip the state recovery)
 *   scal   : HS_SCAL_DOUBLES(nprob) doubles of per-problem scalars (lambda_max x3 slots, t x3
 *            slots, step count, status); opaque, zero-initialise, see hsd_scal_* helpers
 * ------------------------------------------------------------------------------------------ */
/* HS_SP_CROW (build option, default 1): the single-phase model caches ONE wave-speed row, c_max, instead of the two bounds
 * lo / hi = u1 -+ c_max -- rows [c_max, 1/rho, stress row 1(3)]; 0 restores [lo, hi, 1/rho, stress row 1(3)]. */
#ifndef HS_SP_CROW
#define HS_SP_CROW 1
#endif
#define HS_NAUX_SP (HS_SP_CROW ? 5 : 6)
#define HS_NAUX(model) ((model) == HS_MODEL_SP13 ? HS_NAUX_SP : 2)
#define HS_SCAL_SLOTS 8
#define HS_SCAL_DOUBLES(nprob) (HS_SCAL_SLOTS * (nprob) + 8)

typedef struct hsd_problem {
  int model, nphase, gen;   /* gen: 1 = generic exponents (exp/log path), 0 = (alpha,beta,gamma)=(1,3,2) */
  int reserved;
  int64_t ncells, nprob, stride;
  double eos_dev[2][20];    /* derived EoS constants, passed to the kernels by value */
} hsd_problem_t;

/* fill a problem descriptor (host only, no device work) */
int hsd_problem_init(hsd_problem_t* prob_out, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob);
/* AoS (nvar, n) <-> SoA [nvar][stride] on the device */
int hsd_aos_to_soa(const hsd_problem_t* p, const double* aos_dev, double* soa_dev, void* stream);
int hsd_soa_to_aos(const hsd_problem_t* p, const double* soa_dev, double* aos_dev, void* stream);
/* CFL sweep: fills aux and lambda_max slot `slot` of scal (after zeroing it) */
int hsd_wave_bounds(const hsd_problem_t* p, const double* Q, double* aux, double* scal, int slot, void* stream);
/* the same without zeroing the slot first: max(lambda) accumulates over several calls (sweeps of the windows of one grid)
 *
 * WINDOWS.  Every hsd_* call works on a sub-range of the cells of a grid when it is given a window descriptor: a copy of the
 * grid's hsd_problem_t with ncells = cells of the window (nprob = 1, stride = row pitch of the full arrays unchanged) and array
 * pointers advanced to the window's first cell.  With ghost_mask bits set for the ends that are not physical boundaries, hsd_step
 * on the windows [b_i - 2, b_i+1) updates exactly the cells [b_i - 1, b_i+1 - 1): this is how hs_step_host and
 * SlabSolver.step_host process a grid chunk by chunk while it is still arriving over the host link.  (Even window starts keep the
 * single-phase step on its tensor-map tile copies.) */
int hsd_wave_bounds_acc(const hsd_problem_t* p, const double* Q, double* aux, double* scal, int slot, void* stream);
/* one fused step n: reads Qin/aux_in and slot n%3, writes Qout/aux_out and slot (n+1)%3;
 * if dt_hist != NULL the step's dt of problem p is stored at dt_hist[p*hist_cap + hist_k] (hist_k < 0: at the index the
 * per-problem counter inside `scal` holds, which the kernel then increments -- what lets a captured CUDA graph of steps be replayed).
 * Two-phase grids of at most HS_QP_MAX_CELLS cells (environment, default 2048; 0 = never) are stepped by the quadrature-parallel
 * kernel k_step_qp (one warp per quadrature node: low latency on grids too small to fill the GPU); its results are bit-identical
 * to the fused kernel's, so the switch is invisible.
 * ghost_mask bit 0 / bit 1: the first / last cell of the array is a halo copy of a neighbouring
 * slab's cell (not written, not counted in lambda_max) instead of a frozen physical boundary cell
 * (main.jl:219-220). */
int hsd_step(const hsd_problem_t* p, int flux, double cfl, double dx, double t_end, int64_t n,
             const double* Qin, const double* aux_in, double* Qout, double* aux_out,
             double* scal, double* dt_hist, int64_t hist_k, int64_t hist_cap, int ghost_mask, void* stream);
/* halo of a slab (nprob == 1): unpack == 0 packs [Q(nvar), aux(HS_NAUX)] of the first / last OWNED
 * cell (index 1 / ncells-2) into left / right (nvar+HS_NAUX doubles each); unpack == 1 stores left /
 * right into the halo cells (index 0 / ncells-1).  mask bit 0 / 1: a left / right neighbour exists. */
int hsd_halo(const hsd_problem_t* p, double* Q, double* aux, double* left, double* right, int mask, int unpack, void* stream);
/* The same exchange plus the max-reduction of lambda_max in ONE kernel over NVLink peer memory (no NCCL
 * on the step path).  mailboxes[world]: device pointers, valid in THIS process, of every rank's mailbox
 * (hsd_mailbox_doubles() doubles each, zero-initialised, allocated as peer-accessible / symmetric
 * memory, e.g. torch.distributed._symmetric_memory); lam_slot = hsd_scal_lambda_next(scal, 1, n) of the
 * step just enqueued; seq = 1, 2, 3, ... identical on all ranks and never reused.  <= 8 ranks, one node.
 * Results are bit-identical to hsd_halo + all-reduce(max).  A peer that never posts (crashed rank) does not
 * hang the GPU: after HS_EXCHANGE_TIMEOUT_S seconds (default 20) the kernel gives up and sets bit 1 of the
 * status word of `scal`. */
int hsd_mailbox_doubles(void);
/* HS_NAUX(model) of the library as built (a host binding sizes its aux arrays with this) */
int hsd_naux(int model);
int hsd_exchange_p2p(const hsd_problem_t* p, double* Q, double* aux, double* lam_slot, void* const* mailboxes, int rank, int world,
                     uint64_t seq, double* scal, void* stream);
/* address (device pointer) of the lambda_max slot that step n WRITES, as doubles [nprob]:
 * the buffer to all-reduce(max) across ranks between step n and n+1 */
double* hsd_scal_lambda_next(double* scal, int64_t nprob, int64_t n);
double* hsd_scal_lambda_cur(double* scal, int64_t nprob, int64_t n);
double* hsd_scal_time(double* scal, int64_t nprob, int64_t n);     /* t slot READ by step n */
double* hsd_scal_steps(double* scal, int64_t nprob);               /* int64 step counts */
double* hsd_scal_status(double* scal, int64_t nprob);              /* int32 status word */
/* number of kernels launched by this library in this process (for bench accounting) */
int64_t hs_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HYPERELASTIC_B200_H */
