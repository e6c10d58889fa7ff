// C ABI of libhyperelastic_b200.so (see include/hyperelastic_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host side only: argument checks, device
// memory, launches.  There is no CPU compute path in this file by design.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hyperelastic_b200.h"
#include "hs_kernels.cuh"

using namespace hs;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof b_, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      return fail(HS_ERR_CUDA, b_);                                                                \
    }                                                                                              \
  } while (0)

// Switches the calling thread to `dev` and back on scope exit, so the host-buffer entry points never
// disturb the caller's current device (torch, a Julia CUDA allocator, ...).
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevBufRaw {   // scoped device allocation on the current device
  double* p = nullptr;
  ~DevBufRaw() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(double)); }
};

static_assert(sizeof(EosDev) <= sizeof(double) * 20, "EosDev must fit hsd_problem_t::eos_dev");
static_assert(sizeof(hs_barton2009_t) == sizeof(EosAbi), "ABI struct mismatch");

EosPair eos_pair(const hsd_problem_t* p) {
  EosPair e;
  std::memcpy(&e.e[0], p->eos_dev[0], sizeof(EosDev));
  std::memcpy(&e.e[1], p->eos_dev[1], sizeof(EosDev));
  return e;
}

// per-problem scalar block layout inside `scal` (doubles): lam[3][nprob] | t[3][nprob] | steps[nprob] | hist_n[nprob] | status
unsigned long long* scal_lam(double* s) { return reinterpret_cast<unsigned long long*>(s); }
double* scal_t(double* s, int64_t nprob) { return s + 3 * nprob; }
long long* scal_steps(double* s, int64_t nprob) { return reinterpret_cast<long long*>(s + 6 * nprob); }
long long* scal_hist_n(double* s, int64_t nprob) { return reinterpret_cast<long long*>(s + 7 * nprob); }   // recorded-dt counters
int* scal_status(double* s, int64_t nprob) { return reinterpret_cast<int*>(s + HS_SCAL_SLOTS * nprob); }

// threads per block of the fused step / sweep kernels
#ifndef HS_T_STEP_MPH
#define HS_T_STEP_MPH 128
#endif
#ifndef HS_T_STEP_SP
#define HS_T_STEP_SP 128
#endif
constexpr int T_STEP_MPH = HS_T_STEP_MPH, T_STEP_SP = HS_T_STEP_SP, T_FACE = 64;

template <int MODEL, int FLUX, bool GEN, int T, bool SAME>
int launch_step_s(const StepArgs& a, int64_t nblocks, cudaStream_t st) {
  constexpr size_t smem = step_smem_bytes<MODEL, T>();
  static std::atomic<unsigned> attr_dev_mask{0};   // (the attribute is per device and per kernel instantiation)
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load() & (1u << (dev & 31)))) {
    CU(cudaFuncSetAttribute(k_step<MODEL, FLUX, GEN, T, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev_mask.fetch_or(1u << (dev & 31));
  }
  k_step<MODEL, FLUX, GEN, T, SAME><<<(unsigned)nblocks, T, smem, st>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

template <int MODEL, int FLUX, bool GEN, int T>
int launch_step_t(const StepArgs& a, int64_t nblocks, cudaStream_t st) {
  // one thread per cell (single phase) always reads phase 0; two phases: specialise when the EoS blocks are identical
  const bool same = MODEL == MODEL_SP13 || std::memcmp(&a.eos.e[0], &a.eos.e[1], sizeof(EosDev)) == 0;
  if (MODEL == MODEL_SP13 || same) return launch_step_s<MODEL, FLUX, GEN, T, true>(a, nblocks, st);
  return launch_step_s<MODEL, FLUX, GEN, T, (MODEL == MODEL_SP13)>(a, nblocks, st);
}

// Tensor maps for the 2-D tile copies of k_step_sp<TM2D>: a (rows x stride) row-major FP64 array, box = 128 columns x all rows.
// cuTensorMapEncodeTiled is a pure host function of the driver; it is reached through the runtime's entry-point query so that
// the library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// width: columns that exist from `base` on (the whole array, or a window of it when the step runs on a sub-range of the cells:
// then base points into the array and the rows are still `stride` doubles apart); columns >= width are zero-filled.
bool make_tile_map(CUtensorMap* m, const double* base, long long width, long long stride, int rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)stride * sizeof(double)};
  const cuuint32_t box[2] = {128u, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1u, 1u};
  // L2 promotion of the tile reads: 128 B by default; HS_SP_L2PROMO=0 / 64 / 256 for tuning runs
  static const CUtensorMapL2promotion promo = [] {
    const char* e = std::getenv("HS_SP_L2PROMO");
    const int v = e ? std::atoi(e) : 128;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  }();
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Single-phase TMA tile pipeline (k_step_sp): every block walks over `kper` tiles.
template <int FLUX, bool GEN, int T, bool SINGLE, bool TM2D>
int launch_step_sp_s(const StepArgs& a, int64_t ntiles, int kper_in, const CUtensorMap& mq, const CUtensorMap& ma, cudaStream_t st) {
  // SINGLE: bpp blocks per problem, each walking over its tiles b, b + bpp, ...; otherwise kper consecutive tiles per block
  int kper = kper_in, bpp = 1;
  int64_t nblocks;
  if (SINGLE) {
    // ensembles with enough problems to fill the GPU twice over: one block per problem (the prologue of the tile pipeline is paid once
    // per problem: 17.99 vs 17.73 G cell-updates/s with blocks of 7 tiles on the 65 536 x 4 096 ensemble)
    if (a.nprob >= 2 * 4 * 148 && !std::getenv("HS_SP_TILES")) kper = a.tiles_per_prob;
    bpp = (int)((a.tiles_per_prob + kper - 1) / kper);
    kper = (a.tiles_per_prob + bpp - 1) / bpp;
    nblocks = (int64_t)bpp * a.nprob;
  } else {
    nblocks = (ntiles + kper - 1) / kper;
  }
  constexpr size_t smem = step_sp_smem_bytes<T, TM2D>();
  static std::atomic<unsigned> attr_dev_mask{0};   // (the attribute is per device and per kernel instantiation)
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load() & (1u << (dev & 31)))) {
    CU(cudaFuncSetAttribute(k_step_sp<FLUX, GEN, T, SINGLE, TM2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev_mask.fetch_or(1u << (dev & 31));
  }
  k_step_sp<FLUX, GEN, T, SINGLE, TM2D><<<(unsigned)nblocks, T, smem, st>>>(a, kper, bpp, mq, ma);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

template <int FLUX, bool GEN, int T>
int launch_step_sp(const StepArgs& a, int64_t ntiles, int kper, cudaStream_t st) {
  // one problem (every grid configuration): problem index, scalars and column parities are loop invariants
  // (ensembles: one problem per block as soon as a problem has a few tiles; HS_SP_SINGLE=0 forces the cross-problem flavour)
  const char* e = std::getenv("HS_SP_SINGLE");
  const bool single = (a.nprob == 1 || a.tiles_per_prob >= 4) && !(e && e[0] == '0');
  // tensor-map tile copies need a row pitch that is a multiple of 16 bytes and 32-bit column coordinates
  const char* e2 = std::getenv("HS_SP_TMA2D");
  CUtensorMap mq, ma;
  std::memset(&mq, 0, sizeof mq); std::memset(&ma, 0, sizeof ma);
  // (every tile must start on a 16-byte boundary: even row pitch, and even problem length when there are several problems)
  const bool tm2d = !(e2 && e2[0] == '0') && (a.stride % 2 == 0) && (a.nprob == 1 || a.ncells % 2 == 0) && a.stride < 0x7fffffffLL &&
                    make_tile_map(&mq, a.Qin, (long long)a.ncells * a.nprob, a.stride, 13) &&
                    make_tile_map(&ma, a.aux_in, (long long)a.ncells * a.nprob, a.stride, SP_NAX);
  if (tm2d) {
    if (single) return launch_step_sp_s<FLUX, GEN, T, true, true>(a, ntiles, kper, mq, ma, st);
    return launch_step_sp_s<FLUX, GEN, T, false, true>(a, ntiles, kper, mq, ma, st);
  }
  if (single) return launch_step_sp_s<FLUX, GEN, T, true, false>(a, ntiles, kper, mq, ma, st);
  return launch_step_sp_s<FLUX, GEN, T, false, false>(a, ntiles, kper, mq, ma, st);
}

// tiles per block of the pipeline: 8, fewer on grids too small to fill the GPU twice over; HS_SP_TILES=k forces k
int sp_tiles_per_block(int64_t ntiles) {
  int64_t k;
  if (const char* e = std::getenv("HS_SP_TILES")) {
    k = std::atoi(e);
    if (k > 4096) k = 4096;
  } else {
    // at least two blocks per resident slot, at most ~sqrt(ntiles)/30 tiles per block (12 at 2^24 cells, 24 from 2^26 on): longer
    // blocks amortise the unprefetched first tile, shorter ones keep the tail of the launch short (measured: profiles/r02_experiments.md)
    k = ntiles / (2 * 4 * 148);
    int64_t cap = (int64_t)(std::sqrt((double)ntiles) / 30.0);
    if (cap < 8) cap = 8;
    if (cap > 24) cap = 24;
    if (k > cap) k = cap;
  }
  if (k > ntiles) k = ntiles;
  if (k < 1) k = 1;
  return (int)k;
}

// quadrature-parallel two-phase step for small grids (k_step_qp): tiles of 16 cells, 12 warps per tile
template <int FLUX, bool GEN, bool SAME>
int launch_step_qp_s(StepArgs a, cudaStream_t st) {
  constexpr size_t smem = qp_smem_doubles() * sizeof(double);
  static std::atomic<unsigned> attr_dev_mask{0};
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (!(attr_dev_mask.load() & (1u << (dev & 31)))) {
    CU(cudaFuncSetAttribute(k_step_qp<FLUX, GEN, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev_mask.fetch_or(1u << (dev & 31));
  }
  a.tiles_per_prob = (a.ncells - 2 + (QP_CPB - 2) - 1) / (QP_CPB - 2);
  const unsigned nb = (unsigned)a.tiles_per_prob * (unsigned)a.nprob;
  k_step_qp<FLUX, GEN, SAME><<<nb, QP_T, smem, st>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}
int launch_step_qp(int flux, int gen, const StepArgs& a, cudaStream_t st) {
  const bool same = std::memcmp(&a.eos.e[0], &a.eos.e[1], sizeof(EosDev)) == 0;
#define HS_QP(F, G)  (same ? launch_step_qp_s<F, G, true>(a, st) : launch_step_qp_s<F, G, false>(a, st))
  if (flux == HS_FLUX_HLL) return gen ? HS_QP(FLUX_HLL, true) : HS_QP(FLUX_HLL, false);
  return gen ? HS_QP(FLUX_LXF, true) : HS_QP(FLUX_LXF, false);
#undef HS_QP
}
// nsteps steps in one cooperative launch (k_step_qp_loop); HS_ERR_ARG-free "not possible" answer: returns 1 when the grid cannot be
// co-resident or cooperative launches are unsupported, so that the caller falls back to one launch per step
template <int FLUX, bool GEN, bool SAME>
int launch_step_qp_loop_s(StepArgs a, int nsteps, cudaStream_t st, bool* done) {
  constexpr size_t smem = qp_smem_doubles() * sizeof(double);
  *done = false;
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  CU(cudaGetDevice(&dev));
  static std::atomic<unsigned> attr_dev_mask{0};
  if (!(attr_dev_mask.load() & (1u << (dev & 31)))) {
    CU(cudaFuncSetAttribute(k_step_qp_loop<FLUX, GEN, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev_mask.fetch_or(1u << (dev & 31));
  }
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_step_qp_loop<FLUX, GEN, SAME>, QP_T, smem));
  a.tiles_per_prob = (a.ncells - 2 + (QP_CPB - 2) - 1) / (QP_CPB - 2);
  const long long nb = (long long)a.tiles_per_prob * a.nprob;
  if (!coop || nb > (long long)per_sm * sms) return HS_OK;
  void* args[] = {&a, &nsteps};
  CU(cudaLaunchCooperativeKernel((void*)k_step_qp_loop<FLUX, GEN, SAME>, dim3((unsigned)nb), dim3(QP_T), args, smem, st));
  g_launches++;
  *done = true;
  return HS_OK;
}
int launch_step_qp_loop(int flux, int gen, const StepArgs& a, int nsteps, cudaStream_t st, bool* done) {
  const bool same = std::memcmp(&a.eos.e[0], &a.eos.e[1], sizeof(EosDev)) == 0;
#define HS_QPL(F, G)  (same ? launch_step_qp_loop_s<F, G, true>(a, nsteps, st, done) : launch_step_qp_loop_s<F, G, false>(a, nsteps, st, done))
  if (flux == HS_FLUX_HLL) return gen ? HS_QPL(FLUX_HLL, true) : HS_QPL(FLUX_HLL, false);
  return gen ? HS_QPL(FLUX_LXF, true) : HS_QPL(FLUX_LXF, false);
#undef HS_QPL
}
// two-phase grids up to this many cells (all problems together) take k_step_qp; HS_QP_MAX_CELLS=0 disables it
int64_t qp_max_cells() {
  const char* e = std::getenv("HS_QP_MAX_CELLS");
  return e ? std::atoll(e) : 2048;
}

template <int MODEL, int T>
int launch_step_m(int flux, int gen, const StepArgs& a, int64_t nb, cudaStream_t st) {
  if constexpr (MODEL == MODEL_SP13) {
    // the pipeline needs 16-byte aligned array bases (any cudaMalloc / torch allocation); HS_SP_TMA=0 forces the plain kernel
    const char* e = std::getenv("HS_SP_TMA");
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.Qin) | reinterpret_cast<uintptr_t>(a.aux_in)) & 15u) == 0;
    if (aligned && !(e && e[0] == '0')) {
      const int kper = sp_tiles_per_block(nb);
      if (flux == HS_FLUX_HLL) return gen ? launch_step_sp<FLUX_HLL, true, T>(a, nb, kper, st) : launch_step_sp<FLUX_HLL, false, T>(a, nb, kper, st);
      return gen ? launch_step_sp<FLUX_LXF, true, T>(a, nb, kper, st) : launch_step_sp<FLUX_LXF, false, T>(a, nb, kper, st);
    }
  }
  if (flux == HS_FLUX_HLL) return gen ? launch_step_t<MODEL, FLUX_HLL, true, T>(a, nb, st) : launch_step_t<MODEL, FLUX_HLL, false, T>(a, nb, st);
  return gen ? launch_step_t<MODEL, FLUX_LXF, true, T>(a, nb, st) : launch_step_t<MODEL, FLUX_LXF, false, T>(a, nb, st);
}

}  // namespace

extern "C" {

const char* hs_version(void) { return "hyperelastic_b200 0.2 (sm_100a)"; }
const char* hs_last_error(void) { return g_err.c_str(); }
int64_t hs_kernel_launch_count(void) { return g_launches.load(); }

int hs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---------------------------------------------------------------------------------------------
// device-pointer layer
// ---------------------------------------------------------------------------------------------
int hsd_problem_init(hsd_problem_t* p, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob) {
  if (!p || !eos) return fail(HS_ERR_ARG, "null argument");
  if (model != HS_MODEL_SP13 && model != HS_MODEL_MPH30) return fail(HS_ERR_ARG, "unknown model");
  const int want = model == HS_MODEL_MPH30 ? 2 : 1;
  if (nphase != want) return fail(HS_ERR_ARG, "nphase must be 2 for MPH30 and 1 for SP13");
  if (ncells < 3 || nprob < 1) return fail(HS_ERR_ARG, "need ncells >= 3 and nprob >= 1");
  if (ncells > 0x7fffffff || nprob > 0x7fffffff || ncells * nprob > 0x7fffffffLL) return fail(HS_ERR_ARG, "ncells * nprob exceeds 2^31-1 (more cells than one device can hold)");
  std::memset(p, 0, sizeof *p);
  p->model = model; p->nphase = nphase; p->ncells = ncells; p->nprob = nprob; p->stride = ncells * nprob;
  p->gen = 0;
  for (int k = 0; k < 2; ++k) {
    EosAbi a;
    std::memcpy(&a, &eos[k < nphase ? k : 0], sizeof a);
    if (!eos_is_default_exponents(a)) p->gen = 1;
    EosDev d = make_eos_dev(a);
    std::memcpy(p->eos_dev[k], &d, sizeof d);
  }
  return HS_OK;
}

// n cells starting at aos / soa (which may point into larger arrays), rows of the SoA array `stride` doubles apart
static int transpose_range(int model, bool to_soa, const double* src, double* dst, long long n, long long stride, cudaStream_t st) {
  constexpr int TC = 64;
  const unsigned nb = (unsigned)((n + TC - 1) / TC);
  if (to_soa) {
    if (model == HS_MODEL_MPH30) k_aos_to_soa<30, TC><<<nb, 256, 0, st>>>(src, dst, n, stride);
    else k_aos_to_soa<13, TC><<<nb, 256, 0, st>>>(src, dst, n, stride);
  } else {
    if (model == HS_MODEL_MPH30) k_soa_to_aos<30, TC><<<nb, 256, 0, st>>>(src, dst, n, stride);
    else k_soa_to_aos<13, TC><<<nb, 256, 0, st>>>(src, dst, n, stride);
  }
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

// (ncells * nprob cells; for a WINDOW descriptor -- fewer cells than the row pitch `stride`, pointers into larger arrays -- only the window)
int hsd_aos_to_soa(const hsd_problem_t* p, const double* aos, double* soa, void* stream) {
  return transpose_range(p->model, true, aos, soa, p->ncells * p->nprob, p->stride, (cudaStream_t)stream);
}

int hsd_soa_to_aos(const hsd_problem_t* p, const double* soa, double* aos, void* stream) {
  return transpose_range(p->model, false, soa, aos, p->ncells * p->nprob, p->stride, (cudaStream_t)stream);
}

// accumulate: keep the slot's current value and max into it (the sweep of one window of a grid that arrives chunk by chunk)
static int wave_bounds_impl(const hsd_problem_t* p, const double* Q, double* aux, double* scal, int slot,
                            double* eig_full, cudaStream_t st, bool accumulate = false) {
  if (slot < 0 || slot > 2) return fail(HS_ERR_ARG, "slot must be 0..2");
  unsigned long long* lam = scal_lam(scal) + (size_t)slot * p->nprob;
  if (!accumulate) CU(cudaMemsetAsync(lam, 0, sizeof(double) * p->nprob, st));
  int* status = scal_status(scal, p->nprob);
  const EosPair e = eos_pair(p);
  constexpr int T = 128;
  const int cpb = T / p->nphase;
  const int tiles = (int)((p->ncells + cpb - 1) / cpb);
  const unsigned nb = (unsigned)(tiles * p->nprob);
  if (p->model == HS_MODEL_MPH30) {
    if (p->gen) k_bounds<MODEL_MPH30, true, T><<<nb, T, 0, st>>>(Q, aux, lam, eig_full, status, p->stride, (int)p->ncells, (int)p->nprob, tiles, e);
    else k_bounds<MODEL_MPH30, false, T><<<nb, T, 0, st>>>(Q, aux, lam, eig_full, status, p->stride, (int)p->ncells, (int)p->nprob, tiles, e);
  } else {
    if (p->gen) k_bounds<MODEL_SP13, true, T><<<nb, T, 0, st>>>(Q, aux, lam, eig_full, status, p->stride, (int)p->ncells, (int)p->nprob, tiles, e);
    else k_bounds<MODEL_SP13, false, T><<<nb, T, 0, st>>>(Q, aux, lam, eig_full, status, p->stride, (int)p->ncells, (int)p->nprob, tiles, e);
  }
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

int hsd_wave_bounds(const hsd_problem_t* p, const double* Q, double* aux, double* scal, int slot, void* stream) {
  return wave_bounds_impl(p, Q, aux, scal, slot, nullptr, (cudaStream_t)stream);
}
int hsd_wave_bounds_acc(const hsd_problem_t* p, const double* Q, double* aux, double* scal, int slot, void* stream) {
  return wave_bounds_impl(p, Q, aux, scal, slot, nullptr, (cudaStream_t)stream, true);
}

static int make_step_args(StepArgs& a, const hsd_problem_t* p, int flux, double cfl, double dx, double t_end, int64_t n, const double* Qin,
                          const double* aux_in, double* Qout, double* aux_out, double* scal,
                          double* dt_hist, int64_t hist_k, int64_t hist_cap, int ghost_mask) {
  if (flux != HS_FLUX_HLL && flux != HS_FLUX_LXF) return fail(HS_ERR_ARG, "unknown flux");
  if (ghost_mask < 0 || ghost_mask > 3) return fail(HS_ERR_ARG, "ghost_mask must be 0..3");
  a.Qin = Qin; a.Qout = Qout; a.aux_in = aux_in; a.aux_out = aux_out;
  a.lam = scal_lam(scal); a.tt = scal_t(scal, p->nprob); a.steps = scal_steps(scal, p->nprob);
  a.status = scal_status(scal, p->nprob);
  a.dt_hist = dt_hist; a.hist_k = hist_k; a.hist_cap = dt_hist ? hist_cap : 0;
  a.hist_n = scal_hist_n(scal, p->nprob);
  a.stride = p->stride; a.ncells = (int)p->ncells; a.nprob = (int)p->nprob;
  a.cur = (int)(n % 3); a.nxt = (int)((n + 1) % 3); a.clr = (int)((n + 2) % 3);
  a.ghost = ghost_mask;
  a.cfl = cfl; a.dx = dx; a.t_end = t_end;
  static const unsigned long long spin_ns = 1000000000ull * (unsigned long long)(std::getenv("HS_TMA_TIMEOUT_S") ? std::atoi(std::getenv("HS_TMA_TIMEOUT_S")) : 10);
  a.spin_ns = spin_ns;
  a.eos = eos_pair(p);
  a.tiles_per_prob = 0;
  a.dt_shared = nullptr;
  return HS_OK;
}

static int step_impl(const hsd_problem_t* p, int flux, double cfl, double dx, double t_end, int64_t n, const double* Qin,
                     const double* aux_in, double* Qout, double* aux_out, double* scal,
                     double* dt_hist, int64_t hist_k, int64_t hist_cap, int ghost_mask, const double* dt_shared, void* stream);

int hsd_step(const hsd_problem_t* p, int flux, double cfl, double dx, double t_end, int64_t n, const double* Qin,
             const double* aux_in, double* Qout, double* aux_out, double* scal,
             double* dt_hist, int64_t hist_k, int64_t hist_cap, int ghost_mask, void* stream) {
  return step_impl(p, flux, cfl, dx, t_end, n, Qin, aux_in, Qout, aux_out, scal, dt_hist, hist_k, hist_cap, ghost_mask, nullptr, stream);
}

// dt_shared (device scalar, may be null): every problem of the launch steps with this dt -- the sweeps of the dimension-split 2-D solver
static int step_impl(const hsd_problem_t* p, int flux, double cfl, double dx, double t_end, int64_t n, const double* Qin,
                     const double* aux_in, double* Qout, double* aux_out, double* scal,
                     double* dt_hist, int64_t hist_k, int64_t hist_cap, int ghost_mask, const double* dt_shared, void* stream) {
  StepArgs a;
  const int rca = make_step_args(a, p, flux, cfl, dx, t_end, n, Qin, aux_in, Qout, aux_out, scal, dt_hist, hist_k, hist_cap, ghost_mask);
  if (rca) return rca;
  a.dt_shared = dt_shared;
  if (p->model == HS_MODEL_MPH30) {
    if (p->ncells * p->nprob <= qp_max_cells()) return launch_step_qp(flux, p->gen, a, (cudaStream_t)stream);
    constexpr int T = T_STEP_MPH, CPB = T / 2;
    a.tiles_per_prob = (int)((p->ncells - 2 + (CPB - 2) - 1) / (CPB - 2));
    return launch_step_m<MODEL_MPH30, T>(flux, p->gen, a, (int64_t)a.tiles_per_prob * p->nprob, (cudaStream_t)stream);
  } else {
    constexpr int T = T_STEP_SP, CPB = T;
    a.tiles_per_prob = (int)((p->ncells - 2 + (CPB - 2) - 1) / (CPB - 2));
    return launch_step_m<MODEL_SP13, T>(flux, p->gen, a, (int64_t)a.tiles_per_prob * p->nprob, (cudaStream_t)stream);
  }
}

int hsd_halo(const hsd_problem_t* p, double* Q, double* aux, double* left, double* right, int mask, int unpack, void* stream) {
  if (p->nprob != 1) return fail(HS_ERR_ARG, "halo exchange applies to a single slab-decomposed grid");
  if (!mask) return HS_OK;
  const int nvar = p->model == HS_MODEL_MPH30 ? 30 : 13, naux = HS_NAUX(p->model);
  k_halo<<<2, 64, 0, (cudaStream_t)stream>>>(Q, aux, left, right, p->stride, (int)p->ncells, nvar, naux, mask, unpack);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

static_assert(HS_NAUX(HS_MODEL_SP13) == SP_NAX && HS_NAUX(HS_MODEL_MPH30) == ModelTraits<MODEL_MPH30>::NAUX, "header and kernels disagree on the cache rows");
int hsd_naux(int model) { return HS_NAUX(model); }
int hsd_mailbox_doubles(void) { return 2 * MBOX_STRIDE; }

int hsd_exchange_p2p(const hsd_problem_t* p, double* Q, double* aux, double* lam_slot, void* const* mailboxes, int rank, int world,
                     uint64_t seq, double* scal, void* stream) {
  if (p->nprob != 1) return fail(HS_ERR_ARG, "halo exchange applies to a single slab-decomposed grid");
  if (world < 1 || world > MBOX_MAXR || rank < 0 || rank >= world) return fail(HS_ERR_ARG, "p2p exchange supports 1..8 ranks on one node");
  if (seq == 0) return fail(HS_ERR_ARG, "seq must start at 1 (mailboxes are zero-initialised)");
  const int nvar = p->model == HS_MODEL_MPH30 ? 30 : 13, naux = HS_NAUX(p->model);
  if (nvar + naux > MBOX_MAXW) return fail(HS_ERR_ARG, "mailbox too small");
  PeerPtrs pp;
  for (int i = 0; i < MBOX_MAXR; ++i) pp.p[i] = i < world ? static_cast<double*>(mailboxes[i]) : nullptr;
  static const unsigned long long timeout_ns = 1000000000ull * (unsigned long long)(std::getenv("HS_EXCHANGE_TIMEOUT_S") ? std::atoi(std::getenv("HS_EXCHANGE_TIMEOUT_S")) : 20);
  k_exchange_p2p<<<1, 64, 0, (cudaStream_t)stream>>>(Q, aux, reinterpret_cast<unsigned long long*>(lam_slot), pp, p->stride,
                                                    (int)p->ncells, nvar, naux, rank, world, (unsigned long long)seq,
                                                    scal_status(scal, p->nprob), timeout_ns);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

double* hsd_scal_lambda_next(double* scal, int64_t nprob, int64_t n) { return scal + ((n + 1) % 3) * nprob; }
double* hsd_scal_lambda_cur(double* scal, int64_t nprob, int64_t n) { return scal + (n % 3) * nprob; }
double* hsd_scal_time(double* scal, int64_t nprob, int64_t n) { return scal + 3 * nprob + (n % 3) * nprob; }
double* hsd_scal_steps(double* scal, int64_t nprob) { return scal + 6 * nprob; }
double* hsd_scal_status(double* scal, int64_t nprob) { return scal + HS_SCAL_SLOTS * nprob; }

// ---------------------------------------------------------------------------------------------
// stateful context (host-buffer ABI).  One context = one grid (or ensemble) on one device, or one
// grid slab-decomposed over several devices of this process (hs_create_multi): the drop-in for a
// single-process driver such as `julia main.jl`.  With several devices every step is one fused
// kernel per device followed by one k_exchange_p2p per device (peer access, no NCCL).
// ---------------------------------------------------------------------------------------------
namespace {
struct Part {                 // one slab (or one share of an ensemble) on one device
  int device = 0;
  cudaStream_t stream = nullptr;
  hsd_problem_t prob;         // slab: ncells = local cells (halo included), nprob = 1; ensemble share: nprob = local problems
  int64_t a = 0, b = 0;       // slab: owned global cells [a, b);  ensemble share: owned problems [a, b)
  int64_t lo_g = 0;           // slab: global index of local cell 0
  int ghost = 0;
  double* Q[2] = {nullptr, nullptr};
  double* aux[2] = {nullptr, nullptr};
  double* scal = nullptr;
  double* stage = nullptr;    // AoS staging, nvar * local cells
  double* mbox = nullptr;     // exchange mailbox (slab mode only)
  double* dt_hist = nullptr;  // grown on demand
  int64_t hist_cap = 0;
  // chunk-pipelined host step (hs_step_host): copy streams, output staging, the sweep's own scalar block, per-chunk events
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  double* stage_out = nullptr;
  double* scal_sweep = nullptr;
  double* scal_win = nullptr;     // ensembles: per-chunk scalar blocks of the speculative steps and of the sweeps
  size_t scal_win_doubles = 0;
  std::vector<cudaEvent_t> ev_in, ev_out;
};
}  // namespace

struct hs_ctx {
  int model = 0, nvar = 0, nphase = 0;
  int64_t ncells = 0, nprob = 0;
  bool slabs = false;         // several devices share ONE grid (halo exchange); otherwise they share the problems
  std::vector<Part> parts;
  std::vector<void*> mailboxes;
  int64_t n = 0;              // launch counter since the last upload (selects buffers and scalar slots)
  uint64_t xseq = 0;          // exchange sequence number (never reused)
  // hs_advance: six consecutive steps (the period of the buffer / scalar-slot rotation) captured as a CUDA graph and replayed;
  // valid for one (flux, cfl, dx, t_end, history buffer) and for launches that start at n % 6 == n0
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int flux = -1, n0 = -1;
    double cfl = 0, dx = 0, t_end = 0;
    const double* hist = nullptr;
    int64_t hist_cap = 0;
  } graph;
  bool has_state = false;     // the device holds a state with a valid max(lambda) slot (after upload / step / step_host)
  int64_t pipelined_calls = 0, speculation_hits = 0;   // hs_step_host bookkeeping (hs_step_host_stats)
  // host range (in cells / in problems) a part reads and writes
  int64_t first_cell(const Part& p) const { return slabs ? p.lo_g : p.a * ncells; }
  int64_t owned_first(const Part& p) const { return slabs ? p.a : p.a * ncells; }
  int64_t owned_cells(const Part& p) const { return slabs ? p.b - p.a : (p.b - p.a) * ncells; }
  int64_t first_prob(const Part& p) const { return slabs ? 0 : p.a; }
  int64_t nprob_local(const Part& p) const { return p.prob.nprob; }
};

#define PART_ENTER(p)                                              \
  DeviceGuard guard_((p).device);                                  \
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed")

static int create_impl(hs_ctx_t** out, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob,
                       const int* devices, int ndev) {
  if (!out) return fail(HS_ERR_ARG, "null ctx pointer");
  *out = nullptr;
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  if (ndev < 1 || ndev > MBOX_MAXR || !devices) return fail(HS_ERR_ARG, "1..8 devices");
  hsd_problem_t whole;
  int rc = hsd_problem_init(&whole, model, eos, nphase, ncells, nprob);
  if (rc) return rc;
  const bool slabs = ndev > 1 && nprob == 1;
  if (slabs && ncells / ndev < 3) return fail(HS_ERR_ARG, "slabs too small: need >= 3 cells per device");
  if (ndev > 1 && !slabs && nprob < ndev) return fail(HS_ERR_ARG, "fewer problems than devices");
  hs_ctx* c = new hs_ctx();
  c->model = model; c->nvar = model == HS_MODEL_MPH30 ? 30 : 13; c->nphase = nphase; c->ncells = ncells; c->nprob = nprob;
  c->slabs = slabs;
  c->parts.resize(ndev);
  auto bail = [&](int code) { hs_destroy(c); return code; };
  for (int r = 0; r < ndev; ++r) {
    Part& p = c->parts[r];
    p.device = devices[r];
    if (slabs) {
      // same partition as slab.py::slab_bounds: interior cuts are odd, so that every slab plus its halo cells is an even-sized
      // array when ncells is even (an even row pitch is what the tensor-map tile copies of the single-phase step need)
      auto cut = [&](int64_t k) -> int64_t {
        if (k <= 0) return 0;
        if (k >= ndev) return ncells;
        int64_t x = ncells * k / ndev;
        if (x % 2 == 0 && ncells / ndev >= 64) --x;
        return x;
      };
      p.a = cut(r); p.b = cut(r + 1);
      p.lo_g = p.a - (r > 0 ? 1 : 0);
      const int64_t hi_g = p.b + (r < ndev - 1 ? 1 : 0);
      p.ghost = (r > 0 ? 1 : 0) | (r < ndev - 1 ? 2 : 0);
      rc = hsd_problem_init(&p.prob, model, eos, nphase, hi_g - p.lo_g, 1);
    } else {
      p.a = nprob * r / ndev; p.b = nprob * (r + 1) / ndev;                    // problems [a, b)
      rc = hsd_problem_init(&p.prob, model, eos, nphase, ncells, p.b - p.a);
    }
    if (rc) return bail(rc);
    DeviceGuard g(p.device);
    if (!g.ok) return bail(fail(HS_ERR_CUDA, "cudaSetDevice failed"));
    cudaError_t e = cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking);
    const size_t nq = (size_t)c->nvar * p.prob.stride * sizeof(double), nb = (size_t)HS_NAUX(model) * p.prob.stride * sizeof(double);
    const size_t ns = sizeof(double) * HS_SCAL_DOUBLES(p.prob.nprob);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
      e = cudaMalloc(&p.Q[k], nq);
      if (e == cudaSuccess) e = cudaMalloc(&p.aux[k], nb);
    }
    if (e == cudaSuccess) e = cudaMalloc(&p.scal, ns);
    if (e == cudaSuccess) e = cudaMemset(p.scal, 0, ns);
    if (e == cudaSuccess && slabs) {
      e = cudaMalloc(&p.mbox, sizeof(double) * hsd_mailbox_doubles());
      if (e == cudaSuccess) e = cudaMemset(p.mbox, 0, sizeof(double) * hsd_mailbox_doubles());
      for (int q = 0; q < ndev && e == cudaSuccess; ++q) {
        if (devices[q] == p.device) continue;
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, p.device, devices[q]);
        if (e == cudaSuccess && !can) { g_err = "devices cannot access each other's memory (no peer access)"; return bail(HS_ERR_CUDA); }
        if (e == cudaSuccess) {
          e = cudaDeviceEnablePeerAccess(devices[q], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) { e = cudaSuccess; cudaGetLastError(); }
        }
      }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { g_err = std::string("device allocation failed: ") + cudaGetErrorString(e); return bail(HS_ERR_CUDA); }
  }
  for (auto& p : c->parts) c->mailboxes.push_back(p.mbox);
  *out = c;
  return HS_OK;
}

int hs_create(hs_ctx_t** out, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob, int device) {
  return create_impl(out, model, eos, nphase, ncells, nprob, &device, 1);
}

int hs_create_multi(hs_ctx_t** out, int model, const hs_barton2009_t* eos, int nphase, int64_t ncells, int64_t nprob,
                    const int* devices, int ndev) {
  return create_impl(out, model, eos, nphase, ncells, nprob, devices, ndev);
}

int hs_destroy(hs_ctx_t* c) {
  if (!c) return HS_OK;
  if (c->graph.exec) { cudaGraphExecDestroy(c->graph.exec); c->graph.exec = nullptr; }
  for (auto& p : c->parts) {
    DeviceGuard g(p.device);
    for (int k = 0; k < 2; ++k) { cudaFree(p.Q[k]); cudaFree(p.aux[k]); }
    cudaFree(p.scal); cudaFree(p.stage); cudaFree(p.mbox); cudaFree(p.dt_hist); cudaFree(p.stage_out); cudaFree(p.scal_sweep); cudaFree(p.scal_win);
    for (auto e : p.ev_in) cudaEventDestroy(e);
    for (auto e : p.ev_out) cudaEventDestroy(e);
    if (p.s_h2d) cudaStreamDestroy(p.s_h2d);
    if (p.s_d2h) cudaStreamDestroy(p.s_d2h);
    if (p.stream) cudaStreamDestroy(p.stream);
  }
  delete c;
  return HS_OK;
}

static int ensure_stage(hs_ctx* c, Part& p) {
  if (!p.stage) CU(cudaMalloc(&p.stage, (size_t)c->nvar * p.prob.stride * sizeof(double)));
  return HS_OK;
}

static int sync_all(hs_ctx* c) {
  for (auto& p : c->parts) { PART_ENTER(p); CU(cudaStreamSynchronize(p.stream)); }
  return HS_OK;
}

static int read_status(hs_ctx* c) {
  int bad = 0;
  for (auto& p : c->parts) {
    PART_ENTER(p);
    int st = 0;
    CU(cudaMemcpyAsync(&st, scal_status(p.scal, p.prob.nprob), sizeof(int), cudaMemcpyDeviceToHost, p.stream));
    CU(cudaStreamSynchronize(p.stream));
    bad |= st;
  }
  if (bad & 2) return fail(HS_ERR_EXCHANGE, "peer-memory exchange timed out: a device of the context stopped stepping");
  if (bad & 4) return fail(HS_ERR_CUDA, "tile copy (TMA) did not complete: internal error of the single-phase step kernel");
  if (bad) return fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError");
  return HS_OK;
}

int hs_upload(hs_ctx_t* c, const double* Q) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (!Q) return fail(HS_ERR_ARG, "null Q");
  c->n = 0;
  for (auto& p : c->parts) {
    PART_ENTER(p);
    int rc = ensure_stage(c, p); if (rc) return rc;
    const size_t nq = (size_t)c->nvar * p.prob.stride * sizeof(double);
    CU(cudaMemcpyAsync(p.stage, Q + (size_t)c->first_cell(p) * c->nvar, nq, cudaMemcpyHostToDevice, p.stream));
    CU(cudaMemsetAsync(p.scal, 0, sizeof(double) * HS_SCAL_DOUBLES(p.prob.nprob), p.stream));
    rc = hsd_aos_to_soa(&p.prob, p.stage, p.Q[0], p.stream); if (rc) return rc;
    rc = hsd_wave_bounds(&p.prob, p.Q[0], p.aux[0], p.scal, 0, p.stream); if (rc) return rc;
  }
  if (c->slabs) {   // lambda_max over the slabs (setup path: through the host)
    double lmax = 0.0;
    for (auto& p : c->parts) {
      PART_ENTER(p);
      double l = 0.0;
      CU(cudaMemcpyAsync(&l, p.scal, sizeof(double), cudaMemcpyDeviceToHost, p.stream));
      CU(cudaStreamSynchronize(p.stream));
      lmax = l > lmax ? l : lmax;
    }
    for (auto& p : c->parts) {
      PART_ENTER(p);
      CU(cudaMemcpyAsync(p.scal, &lmax, sizeof(double), cudaMemcpyHostToDevice, p.stream));
      CU(cudaStreamSynchronize(p.stream));
    }
  }
  c->has_state = true;
  return read_status(c);
}

int hs_download(hs_ctx_t* c, double* Q) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (!Q) return fail(HS_ERR_ARG, "null Q");
  for (auto& p : c->parts) {
    PART_ENTER(p);
    int rc = ensure_stage(c, p); if (rc) return rc;
    rc = hsd_soa_to_aos(&p.prob, p.Q[c->n & 1], p.stage, p.stream); if (rc) return rc;
    const int64_t first = c->owned_first(p);
    CU(cudaMemcpyAsync(Q + (size_t)first * c->nvar, p.stage + (size_t)(first - c->first_cell(p)) * c->nvar,
                       (size_t)c->owned_cells(p) * c->nvar * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
  }
  return sync_all(c);
}

int hs_set_time(hs_ctx_t* c, double t, int64_t step) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  for (auto& p : c->parts) {
    PART_ENTER(p);
    const int64_t np = p.prob.nprob;
    std::vector<double> tv(np, t);
    std::vector<long long> sv(np, step);
    CU(cudaMemcpyAsync(hsd_scal_time(p.scal, np, c->n), tv.data(), sizeof(double) * np, cudaMemcpyHostToDevice, p.stream));
    CU(cudaMemcpyAsync(scal_steps(p.scal, np), sv.data(), sizeof(long long) * np, cudaMemcpyHostToDevice, p.stream));
    CU(cudaStreamSynchronize(p.stream));
  }
  return HS_OK;
}

int hs_wave_speeds(hs_ctx_t* c, double* eig, double* lambda_max) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  const int cur = (int)(c->n & 1);
  const int neig = 6 * c->nphase;
  for (auto& p : c->parts) {
    PART_ENTER(p);
    const int64_t np = p.prob.nprob;
    if (eig) {
      DevBufRaw d_eig, d_scal;   // the sweep runs against a scratch scalar block: the context's lambda slots stay untouched
      CU(d_eig.alloc((size_t)neig * p.prob.stride));
      CU(d_scal.alloc((size_t)HS_SCAL_DOUBLES(np)));
      CU(cudaMemsetAsync(d_scal.p, 0, sizeof(double) * HS_SCAL_DOUBLES(np), p.stream));
      int rc = wave_bounds_impl(&p.prob, p.Q[cur], p.aux[cur], d_scal.p, 0, d_eig.p, p.stream); if (rc) return rc;
      const int64_t first = c->owned_first(p);
      CU(cudaMemcpyAsync(eig + (size_t)first * neig, d_eig.p + (size_t)(first - c->first_cell(p)) * neig,
                         (size_t)c->owned_cells(p) * neig * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
      CU(cudaStreamSynchronize(p.stream));
    }
    if (lambda_max && (!c->slabs || &p == &c->parts[0])) {
      CU(cudaMemcpyAsync(lambda_max + c->first_prob(p), hsd_scal_lambda_cur(p.scal, np, c->n), sizeof(double) * np, cudaMemcpyDeviceToHost, p.stream));
      CU(cudaStreamSynchronize(p.stream));
    }
  }
  return read_status(c);
}

// hist_cap > 0: every part records the step's dt of its problems at dt_hist[prob * hist_cap + hist_k]
static int enqueue_step(hs_ctx* c, int flux, double cfl, double dx, double t_end, int64_t hist_k, int64_t hist_cap) {
  const int a = (int)(c->n & 1), b = a ^ 1;
  const int ndev = (int)c->parts.size();
  for (int r = 0; r < ndev; ++r) {
    Part& p = c->parts[r];
    PART_ENTER(p);
    double* hist = (hist_cap > 0 && (!c->slabs || r == 0)) ? p.dt_hist : nullptr;
    int rc = hsd_step(&p.prob, flux, cfl, dx, t_end, c->n, p.Q[a], p.aux[a], p.Q[b], p.aux[b], p.scal, hist, hist_k, hist_cap, p.ghost,
                      p.stream);
    if (rc) return rc;
  }
  if (c->slabs) {   // halo cells + max(lambda) over the slabs: one peer-memory kernel per device
    c->xseq += 1;
    for (int r = 0; r < ndev; ++r) {
      Part& p = c->parts[r];
      PART_ENTER(p);
      int rc = hsd_exchange_p2p(&p.prob, p.Q[b], p.aux[b], hsd_scal_lambda_next(p.scal, 1, c->n), c->mailboxes.data(), r, ndev, c->xseq,
                                p.scal, p.stream);
      if (rc) return rc;
    }
  }
  c->n += 1;
  return HS_OK;
}

static int ensure_hist(hs_ctx* c, int64_t cap) {
  for (auto& p : c->parts) {
    PART_ENTER(p);
    if (p.hist_cap < cap || !p.dt_hist) {
      if (p.dt_hist) { cudaFree(p.dt_hist); p.dt_hist = nullptr; }
      CU(cudaMalloc(&p.dt_hist, sizeof(double) * cap * p.prob.nprob));
      p.hist_cap = cap;
    }
    CU(cudaMemsetAsync(p.dt_hist, 0, sizeof(double) * cap * p.prob.nprob, p.stream));
  }
  return HS_OK;
}

// gather per-problem doubles (dt history columns, ...) from the parts into a host array
static int gather_hist(hs_ctx* c, double* dst, int64_t cap) {
  for (auto& p : c->parts) {
    if (c->slabs && &p != &c->parts[0]) continue;
    PART_ENTER(p);
    CU(cudaMemcpyAsync(dst + (size_t)c->first_prob(p) * cap, p.dt_hist, sizeof(double) * cap * p.prob.nprob, cudaMemcpyDeviceToHost, p.stream));
  }
  return HS_OK;
}

int hs_step(hs_ctx_t* c, int flux, double cfl, double dx, double* dt_out) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  int rc = ensure_hist(c, 1); if (rc) return rc;
  rc = enqueue_step(c, flux, cfl, dx, 1.0e300, 0, 1); if (rc) return rc;
  if (dt_out) { rc = gather_hist(c, dt_out, 1); if (rc) return rc; }
  return read_status(c);
}

// Six steps from the current n as one graph launch (single-device contexts).  The graph is (re)captured when its key changes.
static int launch_six_steps(hs_ctx* c, int flux, double cfl, double dx, double t_end, int64_t hist_cap) {
  Part& p = c->parts[0];
  PART_ENTER(p);
  hs_ctx::StepGraph& G = c->graph;
  const double* hist = hist_cap > 0 ? p.dt_hist : nullptr;
  const int n0 = (int)(c->n % 6);
  if (!G.exec || G.flux != flux || G.n0 != n0 || G.cfl != cfl || G.dx != dx || G.t_end != t_end || G.hist != hist || G.hist_cap != hist_cap) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    const int64_t n_save = c->n;
    CU(cudaStreamBeginCapture(p.stream, cudaStreamCaptureModeThreadLocal));
    int rc = HS_OK;
    for (int k = 0; k < 6 && rc == HS_OK; ++k) rc = enqueue_step(c, flux, cfl, dx, t_end, -1, hist_cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(p.stream, &graph);
    c->n = n_save;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    CU(e);
    const cudaError_t e2 = cudaGraphInstantiate(&G.exec, graph, 0);
    cudaGraphDestroy(graph);
    CU(e2);
    G.flux = flux; G.n0 = n0; G.cfl = cfl; G.dx = dx; G.t_end = t_end; G.hist = hist; G.hist_cap = hist_cap;
  }
  CU(cudaGraphLaunch(G.exec, p.stream));
  g_launches += 6;
  c->n += 6;
  return HS_OK;
}

int hs_advance(hs_ctx_t* c, int flux, double cfl, double dx, double t_end, int64_t max_steps, double* t_io,
               int64_t* step_io, double* dt_hist) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (max_steps < 0) return fail(HS_ERR_ARG, "max_steps < 0");
  if (flux != HS_FLUX_HLL && flux != HS_FLUX_LXF) return fail(HS_ERR_ARG, "unknown flux");
  for (auto& p : c->parts) {
    PART_ENTER(p);
    const int64_t np = p.prob.nprob, off = c->first_prob(p);
    if (t_io) CU(cudaMemcpyAsync(hsd_scal_time(p.scal, np, c->n), t_io + off, sizeof(double) * np, cudaMemcpyHostToDevice, p.stream));
    if (step_io) CU(cudaMemcpyAsync(scal_steps(p.scal, np), step_io + off, sizeof(long long) * np, cudaMemcpyHostToDevice, p.stream));
    CU(cudaMemsetAsync(scal_hist_n(p.scal, np), 0, sizeof(long long) * np, p.stream));   // the kernels index the dt history with these counters
  }
  const bool record = dt_hist && max_steps > 0;
  if (record) { int rc = ensure_hist(c, max_steps); if (rc) return rc; }
  const int64_t hist_cap = record ? max_steps : 0;
  // The loop stays on the device: steps are enqueued without waiting for them (on one device as replays of a six-step CUDA
  // graph, so a grid too small to fill the GPU does not pay a launch per step), and the host only looks at the clock to decide
  // how many more to enqueue.  One grid: the number is estimated from the current dt (it changes slowly), so a run of N steps
  // needs a handful of synchronisations; kernels launched past t_end are no-ops (same overshoot semantics as main.jl:202,214).
  const char* eg = std::getenv("HS_GRAPH");
  const bool use_graph = c->parts.size() == 1 && !(eg && eg[0] == '0');
  const char* el = std::getenv("HS_QP_LOOP");
  const bool use_loop = c->parts.size() == 1 && c->model == HS_MODEL_MPH30 && c->ncells * c->nprob <= qp_max_cells() && !(el && el[0] == '0');
  std::vector<double> tv;
  int64_t done = 0;
  while (done < max_steps) {
    bool any = false;
    double est = -1.0;
    for (auto& p : c->parts) {
      if (c->slabs && &p != &c->parts[0]) continue;
      PART_ENTER(p);
      const int64_t np = p.prob.nprob;
      tv.resize(np + 1);
      CU(cudaMemcpyAsync(tv.data(), hsd_scal_time(p.scal, np, c->n), sizeof(double) * np, cudaMemcpyDeviceToHost, p.stream));
      if (c->nprob == 1) CU(cudaMemcpyAsync(tv.data() + 1, hsd_scal_lambda_cur(p.scal, 1, c->n), sizeof(double), cudaMemcpyDeviceToHost, p.stream));
      CU(cudaStreamSynchronize(p.stream));
      for (int64_t i = 0; i < np && !any; ++i) any = tv[i] < t_end;
      if (c->nprob == 1 && any && tv[1] > 0.0) est = (t_end - tv[0]) / (cfl * dx / tv[1]);   // steps left at the current dt
    }
    if (!any) break;
    int64_t m = 32;
    if (est >= 0.0 && est < 1.0e15) {
      m = (int64_t)est - 2;            // stop short of the estimate and look again (dt drifts by a few per cent over such a stretch)
      if (m > 16384) m = 16384;
      if (m < 1) m = 1;
    }
    if (m > max_steps - done) m = max_steps - done;
    int64_t k = 0;
    if (use_loop && m >= 2) {   // small two-phase grid: the whole stretch in one cooperative launch (grid barrier between the steps)
      Part& p = c->parts[0];
      PART_ENTER(p);
      const int a = (int)(c->n & 1), b = a ^ 1;
      StepArgs sa;
      int rc = make_step_args(sa, &p.prob, flux, cfl, dx, t_end, c->n, p.Q[a], p.aux[a], p.Q[b], p.aux[b], p.scal, hist_cap > 0 ? p.dt_hist : nullptr,
                              -1, hist_cap, p.ghost);
      if (rc) return rc;
      bool launched = false;
      rc = launch_step_qp_loop(flux, p.prob.gen, sa, (int)m, p.stream, &launched);
      if (rc) return rc;
      if (launched) { c->n += m; k = m; }
    }
    if (use_graph && m - k >= 12) {
      for (; k + 6 <= m; k += 6) { int rc = launch_six_steps(c, flux, cfl, dx, t_end, hist_cap); if (rc) return rc; }
    }
    for (; k < m; ++k) { int rc = enqueue_step(c, flux, cfl, dx, t_end, -1, hist_cap); if (rc) return rc; }
    done += m;
  }
  for (auto& p : c->parts) {
    if (c->slabs && &p != &c->parts[0]) continue;
    PART_ENTER(p);
    const int64_t np = p.prob.nprob, off = c->first_prob(p);
    if (t_io) CU(cudaMemcpyAsync(t_io + off, hsd_scal_time(p.scal, np, c->n), sizeof(double) * np, cudaMemcpyDeviceToHost, p.stream));
    if (step_io) CU(cudaMemcpyAsync(step_io + off, scal_steps(p.scal, np), sizeof(long long) * np, cudaMemcpyDeviceToHost, p.stream));
  }
  if (record) { int rc = gather_hist(c, dt_hist, max_steps); if (rc) return rc; }
  int rc = sync_all(c); if (rc) return rc;
  return read_status(c);
}

static int step_host_plain(hs_ctx_t* c, int flux, double cfl, double dx, const double* Qin, double* Qout, double* dt_out) {
  int rc = hs_upload(c, Qin); if (rc && rc != HS_ERR_DOMAIN) return rc;
  int rc2 = hs_step(c, flux, cfl, dx, dt_out); if (rc2 && rc2 != HS_ERR_DOMAIN) return rc2;
  int rc3 = hs_download(c, Qout); if (rc3) return rc3;
  return rc ? rc : rc2;
}

// ---------------------------------------------------------------------------------------------
// Chunk-pipelined host step (one grid on one device).
//
// A host-resident step is H2D(Q0) -> CFL sweep -> step -> D2H(Q1), and dt = cfl dx / max(lambda) needs the sweep of the WHOLE
// uploaded state before any cell can be updated, so the two copies cannot overlap: 2 x state / link bandwidth per step.  In a
// solver loop, however, the state passed in is the state the previous call returned, whose max(lambda) the previous fused step
// already produced (it sits in the context's scalar slot).  That value is used as a HINT:
//   for chunk i of the grid (even boundaries b_i, so that every window below is an even-sized array starting on an even cell --
//   what the tensor-map tile copies need):
//     copy stream (H2D):  Qin[b_i, b_i+1)                         -> AoS staging
//     compute stream:     transpose chunk i -> Q[0]; CFL sweep of chunk i -> cache rows, TRUE max(lambda) accumulates in its own slot;
//                         fused step of the window [b_i - 2, b_i+1) with ghost cells at both ends, dt from the HINT
//                         -> cells [b_i - 1, b_i+1 - 1) of Q[1] are final; transpose them -> AoS output staging
//     copy stream (D2H):  -> Qout
//   so that H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1 run at the same time (full-duplex link).
// At the end the true max(lambda) of the uploaded data is compared with the hint, bit for bit.  Equal (every call of a solver loop
// but the first): done -- dt depends on nothing else, so the result is bit-identical to upload + step + download.  Different
// (first call, state edited by the caller): the step is redone from the intact structure-of-arrays input on the device with the
// right dt, and downloaded chunk by chunk (transpose of chunk i+1 overlapping D2H of chunk i).
// ---------------------------------------------------------------------------------------------
static int64_t host_chunk_cells() {   // (read at every call: tests shrink it)
  const char* e = std::getenv("HS_HOST_CHUNK");
  long long x = e ? std::atoll(e) : (1ll << 19);
  if (x < 1024) x = 1024;
  return (int64_t)(x & ~1ll);
}

static int ensure_pipeline(hs_ctx* c, Part& p, int nchunks) {
  if (!p.s_h2d) CU(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
  if (!p.s_d2h) CU(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
  if (!p.stage_out) CU(cudaMalloc(&p.stage_out, (size_t)c->nvar * p.prob.stride * sizeof(double)));
  if (!p.scal_sweep) CU(cudaMalloc(&p.scal_sweep, sizeof(double) * HS_SCAL_DOUBLES(1)));
  while ((int)p.ev_in.size() < nchunks) {
    cudaEvent_t a, b;
    CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    p.ev_in.push_back(a); p.ev_out.push_back(b);
  }
  return ensure_stage(c, p);
}

static hsd_problem_t window_of(const hsd_problem_t& whole, int64_t ncells) {
  hsd_problem_t w = whole;   // same model / EoS / row pitch (stride), fewer cells
  w.ncells = ncells; w.nprob = 1;
  return w;
}

static int step_host_pipelined(hs_ctx_t* c, int flux, double cfl, double dx, const double* Qin, double* Qout, double* dt_out) {
  Part& p = c->parts[0];
  PART_ENTER(p);
  const int64_t N = c->ncells, nvar = c->nvar, CH = host_chunk_cells();
  // Chunk boundaries (all even).  The first chunks are short (CH/32, CH/16, ... then CH), so the kernels and the D2H stream start
  // working early; from then on equal chunks: D2H of chunk i runs beside H2D of chunk i+1, so the drain after the last H2D is one
  // chunk's D2H (shrinking chunks at the end would only let the D2H stream fall behind).  Every copy costs ~25 us of link idle
  // time, a chunk c/N of the ~36 ms the link needs: ~16-32 chunks is the optimum for a 2^24-cell single-phase grid (measured on
  // B200 / PCIe Gen5: H2D alone 55.5 GB/s; both directions at once 48 GB/s each, i.e. 36.4 ms for 2 x 1.745 GB; this call 38.5 ms).
  std::vector<int64_t> bnd;
  {
    int64_t lo = 0;
    for (int64_t sz = CH / 32; sz < CH && N - lo > 2 * CH + 4 * sz; sz *= 2) { bnd.push_back(lo); lo += sz & ~1ll; }
    for (int64_t x = lo; x < N; x += CH) bnd.push_back(x);
    if (bnd.size() > 1 && N - bnd.back() < CH / 4) bnd.pop_back();   // a short last chunk joins its neighbour
    bnd.push_back(N);
  }
  const int K = (int)bnd.size() - 1;
  int rc = ensure_pipeline(c, p, K); if (rc) return rc;
  rc = ensure_hist(c, 1); if (rc) return rc;
  auto b_of = [&](int i) -> int64_t { return bnd[i > K ? K : i]; };

  // the hint: max(lambda) of the state the context holds now (what the previous call returned)
  unsigned long long hint_bits = 0;
  if (c->has_state) {
    CU(cudaMemcpyAsync(&hint_bits, hsd_scal_lambda_cur(p.scal, 1, c->n), sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    CU(cudaStreamSynchronize(p.stream));
  }
  double hint;
  std::memcpy(&hint, &hint_bits, sizeof hint);
  const bool spec = c->has_state && hint > 0.0 && hint < 1.0e300;
  c->has_state = false;   // (set again on success)

  // scalar blocks: the context's own block drives the step (slot 0 = hint), the sweep accumulates the true max(lambda) in its own
  CU(cudaMemsetAsync(p.scal, 0, sizeof(double) * HS_SCAL_DOUBLES(1), p.stream));
  CU(cudaMemsetAsync(p.scal_sweep, 0, sizeof(double) * HS_SCAL_DOUBLES(1), p.stream));
  if (spec) CU(cudaMemcpyAsync(p.scal, &hint_bits, sizeof(double), cudaMemcpyHostToDevice, p.stream));
  cudaEvent_t ev0 = p.ev_out[0];   // (reused below only after the copy streams have waited on it)
  CU(cudaEventRecord(ev0, p.stream));
  CU(cudaStreamWaitEvent(p.s_h2d, ev0, 0));   // staging buffers of the previous call are free (its D2H was synchronised before it returned)

  // HS_HOST_TRACE=1: per-call timeline of the three streams on stderr (development aid)
  const bool trace = std::getenv("HS_HOST_TRACE") != nullptr;
  cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};
  const auto host_t0 = std::chrono::steady_clock::now();
  if (trace) {
    for (auto& e : tr) cudaEventCreate(&e);
    cudaEventRecord(tr[0], p.stream);
  }
  const size_t cellb = (size_t)nvar * sizeof(double);
  for (int i = 0; i < K; ++i) {
    const int64_t b0 = b_of(i), b1 = b_of(i + 1);
    CU(cudaMemcpyAsync(p.stage + b0 * nvar, Qin + b0 * nvar, (size_t)(b1 - b0) * cellb, cudaMemcpyHostToDevice, p.s_h2d));
    CU(cudaEventRecord(p.ev_in[i], p.s_h2d));
    CU(cudaStreamWaitEvent(p.stream, p.ev_in[i], 0));
    rc = transpose_range(c->model, true, p.stage + b0 * nvar, p.Q[0] + b0, b1 - b0, p.prob.stride, p.stream); if (rc) return rc;
    const hsd_problem_t wsw = window_of(p.prob, b1 - b0);
    rc = wave_bounds_impl(&wsw, p.Q[0] + b0, p.aux[0] + b0, p.scal_sweep, 0, nullptr, p.stream, true); if (rc) return rc;
    if (spec) {
      const int64_t w0 = i ? b0 - 2 : 0;                       // window [w0, b1): ghost cell at each end that is not a physical boundary
      const int ghost = (i ? 1 : 0) | (i < K - 1 ? 2 : 0);
      const hsd_problem_t wst = window_of(p.prob, b1 - w0);
      rc = hsd_step(&wst, flux, cfl, dx, 1.0e300, 0, p.Q[0] + w0, p.aux[0] + w0, p.Q[1] + w0, p.aux[1] + w0, p.scal,
                    i == 0 ? p.dt_hist : nullptr, 0, 1, ghost, p.stream);
      if (rc) return rc;
      const int64_t u0 = i ? b0 - 1 : 0, u1 = (i < K - 1) ? b1 - 1 : N;   // cells this window made final
      rc = transpose_range(c->model, false, p.Q[1] + u0, p.stage_out + u0 * nvar, u1 - u0, p.prob.stride, p.stream); if (rc) return rc;
      CU(cudaEventRecord(p.ev_out[i], p.stream));
      CU(cudaStreamWaitEvent(p.s_d2h, p.ev_out[i], 0));
      CU(cudaMemcpyAsync(Qout + u0 * nvar, p.stage_out + u0 * nvar, (size_t)(u1 - u0) * cellb, cudaMemcpyDeviceToHost, p.s_d2h));
    }
  }
  if (trace) {
    cudaEventRecord(tr[1], p.s_h2d); cudaEventRecord(tr[2], p.stream); cudaEventRecord(tr[3], p.s_d2h);
    const double enq = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    cudaEventSynchronize(tr[1]); cudaEventSynchronize(tr[2]); cudaEventSynchronize(tr[3]);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, tr[0], tr[1]); cudaEventElapsedTime(&b, tr[0], tr[2]); cudaEventElapsedTime(&d, tr[0], tr[3]);
    fprintf(stderr, "[hs_step_host] chunks %d spec %d: host enqueue %.2f ms; since start: H2D done %.2f, kernels done %.2f, D2H done %.2f ms\n",
            K, (int)spec, enq, a, b, d);
    for (auto& e : tr) cudaEventDestroy(e);
  }
  unsigned long long true_bits = 0;
  CU(cudaMemcpyAsync(&true_bits, p.scal_sweep, sizeof(double), cudaMemcpyDeviceToHost, p.stream));
  int st_sweep = 0;
  CU(cudaMemcpyAsync(&st_sweep, scal_status(p.scal_sweep, 1), sizeof(int), cudaMemcpyDeviceToHost, p.stream));
  CU(cudaStreamSynchronize(p.stream));
  c->pipelined_calls += 1;
  if (spec && true_bits == hint_bits) {
    // every window counted one step: the grid took one
    const long long one = 1;
    CU(cudaMemcpyAsync(scal_steps(p.scal, 1), &one, sizeof one, cudaMemcpyHostToDevice, p.stream));
    if (dt_out) CU(cudaMemcpyAsync(dt_out, p.dt_hist, sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    CU(cudaStreamSynchronize(p.stream));
    CU(cudaStreamSynchronize(p.s_d2h));
    c->n = 1;
    c->has_state = true;
    c->speculation_hits += 1;
    const int rs = read_status(c);
    if (rs) return rs;
    return st_sweep ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
  }
  // no usable hint: the step is (re)done from the intact input with dt from the true max(lambda)
  if (spec) CU(cudaStreamSynchronize(p.s_d2h));
  CU(cudaMemsetAsync(p.scal, 0, sizeof(double) * HS_SCAL_DOUBLES(1), p.stream));
  CU(cudaMemcpyAsync(p.scal, p.scal_sweep, sizeof(double), cudaMemcpyDeviceToDevice, p.stream));
  c->n = 0;
  rc = enqueue_step(c, flux, cfl, dx, 1.0e300, 0, 1); if (rc) return rc;
  if (dt_out) CU(cudaMemcpyAsync(dt_out, p.dt_hist, sizeof(double), cudaMemcpyDeviceToHost, p.stream));
  for (int i = 0; i < K; ++i) {
    const int64_t b0 = b_of(i), b1 = b_of(i + 1);
    rc = transpose_range(c->model, false, p.Q[1] + b0, p.stage_out + b0 * nvar, b1 - b0, p.prob.stride, p.stream); if (rc) return rc;
    CU(cudaEventRecord(p.ev_out[i], p.stream));
    CU(cudaStreamWaitEvent(p.s_d2h, p.ev_out[i], 0));
    CU(cudaMemcpyAsync(Qout + b0 * nvar, p.stage_out + b0 * nvar, (size_t)(b1 - b0) * cellb, cudaMemcpyDeviceToHost, p.s_d2h));
  }
  CU(cudaStreamSynchronize(p.stream));
  CU(cudaStreamSynchronize(p.s_d2h));
  c->has_state = true;
  const int rs = read_status(c);
  if (rs) return rs;
  return st_sweep ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
}

// ---------------------------------------------------------------------------------------------
// The same pipeline for an ENSEMBLE on one device: a chunk is a group of whole problems (no ghost cells: problems are independent),
// the hints are the per-problem max(lambda) of the state the previous call returned, and every group steps on its own compact
// scalar block (the kernels index the per-problem scalars by the problem's position inside the launch).  All nprob true values are
// compared with the hints at the end; any mismatch redoes the whole step on the device.
// ---------------------------------------------------------------------------------------------
static int step_host_pipelined_ensemble(hs_ctx_t* c, int flux, double cfl, double dx, const double* Qin, double* Qout, double* dt_out) {
  Part& p = c->parts[0];
  PART_ENTER(p);
  const int64_t NC = c->ncells, NP = c->nprob, nvar = c->nvar;
  int64_t G = host_chunk_cells() / NC;          // problems per chunk
  if (G < 1) G = 1;
  const int K = (int)((NP + G - 1) / G);
  int rc = ensure_pipeline(c, p, K); if (rc) return rc;
  rc = ensure_hist(c, 1); if (rc) return rc;
  // two compact scalar blocks per chunk (step, sweep)
  const size_t blk = (size_t)HS_SCAL_DOUBLES(G);
  if (p.scal_win_doubles < 2 * blk * K) {
    if (p.scal_win) { cudaFree(p.scal_win); p.scal_win = nullptr; }
    CU(cudaMalloc(&p.scal_win, sizeof(double) * 2 * blk * K));
    p.scal_win_doubles = 2 * blk * K;
  }
  const bool spec = c->has_state;
  c->has_state = false;
  const int cur_main = (int)(c->n % 3);
  CU(cudaMemsetAsync(p.scal_win, 0, sizeof(double) * 2 * blk * K, p.stream));
  // hints: the context's current max(lambda) slot, problem by problem, into slot 0 of the groups' step blocks
  if (spec) {
    for (int i = 0; i < K; ++i) {
      const int64_t p0 = (int64_t)i * G, g = std::min<int64_t>(G, NP - p0);
      CU(cudaMemcpyAsync(p.scal_win + (size_t)(2 * i) * blk, hsd_scal_lambda_cur(p.scal, NP, c->n) + p0, sizeof(double) * g, cudaMemcpyDeviceToDevice, p.stream));
    }
  }
  std::vector<double> hint(NP, 0.0), truth(NP, 0.0);
  if (spec) CU(cudaMemcpyAsync(hint.data(), p.scal + (size_t)cur_main * NP, sizeof(double) * NP, cudaMemcpyDeviceToHost, p.stream));
  cudaEvent_t ev0 = p.ev_out[0];
  CU(cudaEventRecord(ev0, p.stream));
  CU(cudaStreamWaitEvent(p.s_h2d, ev0, 0));
  const size_t cellb = (size_t)nvar * sizeof(double);
  for (int i = 0; i < K; ++i) {
    const int64_t p0 = (int64_t)i * G, g = std::min<int64_t>(G, NP - p0), c0 = p0 * NC, nc = g * NC;
    double* sstep = p.scal_win + (size_t)(2 * i) * blk;
    double* ssweep = sstep + blk;
    CU(cudaMemcpyAsync(p.stage + c0 * nvar, Qin + c0 * nvar, (size_t)nc * cellb, cudaMemcpyHostToDevice, p.s_h2d));
    CU(cudaEventRecord(p.ev_in[i], p.s_h2d));
    CU(cudaStreamWaitEvent(p.stream, p.ev_in[i], 0));
    rc = transpose_range(c->model, true, p.stage + c0 * nvar, p.Q[0] + c0, nc, p.prob.stride, p.stream); if (rc) return rc;
    hsd_problem_t w = p.prob;   // g problems of NC cells, rows still `stride` apart
    w.nprob = g;
    rc = wave_bounds_impl(&w, p.Q[0] + c0, p.aux[0] + c0, ssweep, 0, nullptr, p.stream); if (rc) return rc;
    if (spec) {
      rc = hsd_step(&w, flux, cfl, dx, 1.0e300, 0, p.Q[0] + c0, p.aux[0] + c0, p.Q[1] + c0, p.aux[1] + c0, sstep, p.dt_hist + p0, 0, 1, 0, p.stream);
      if (rc) return rc;
      rc = transpose_range(c->model, false, p.Q[1] + c0, p.stage_out + c0 * nvar, nc, p.prob.stride, p.stream); if (rc) return rc;
      CU(cudaEventRecord(p.ev_out[i], p.stream));
      CU(cudaStreamWaitEvent(p.s_d2h, p.ev_out[i], 0));
      CU(cudaMemcpyAsync(Qout + c0 * nvar, p.stage_out + c0 * nvar, (size_t)nc * cellb, cudaMemcpyDeviceToHost, p.s_d2h));
    }
  }
  for (int i = 0; i < K; ++i) {   // (after the loop: a copy into pageable memory would stall the enqueueing thread at every chunk)
    const int64_t p0 = (int64_t)i * G, g = std::min<int64_t>(G, NP - p0);
    CU(cudaMemcpyAsync(truth.data() + p0, p.scal_win + (size_t)(2 * i + 1) * blk, sizeof(double) * g, cudaMemcpyDeviceToHost, p.stream));
  }
  CU(cudaStreamSynchronize(p.stream));
  c->pipelined_calls += 1;
  int st_all = 0;
  bool same = spec && std::memcmp(hint.data(), truth.data(), sizeof(double) * NP) == 0;
  // status words of the groups' blocks
  {
    std::vector<int> stw(2 * K, 0);
    for (int i = 0; i < 2 * K; ++i) {
      const int64_t g = std::min<int64_t>(G, NP - (int64_t)(i / 2) * G);
      CU(cudaMemcpyAsync(&stw[i], scal_status(p.scal_win + (size_t)i * blk, g), sizeof(int), cudaMemcpyDeviceToHost, p.stream));
    }
    CU(cudaStreamSynchronize(p.stream));
    for (int i = 0; i < 2 * K; ++i) if (same || (i & 1)) st_all |= stw[i];
  }
  CU(cudaMemsetAsync(p.scal, 0, sizeof(double) * HS_SCAL_DOUBLES(NP), p.stream));
  if (same) {
    // assemble the context's scalar block as upload + one step leave it: slot 1 = max(lambda) of the new states, t = dt, one step
    for (int i = 0; i < K; ++i) {
      const int64_t p0 = (int64_t)i * G, g = std::min<int64_t>(G, NP - p0);
      double* sstep = p.scal_win + (size_t)(2 * i) * blk;
      CU(cudaMemcpyAsync(p.scal + (size_t)1 * NP + p0, sstep + (size_t)1 * g, sizeof(double) * g, cudaMemcpyDeviceToDevice, p.stream));                 // lambda slot 1
      CU(cudaMemcpyAsync(scal_t(p.scal, NP) + (size_t)1 * NP + p0, scal_t(sstep, g) + (size_t)1 * g, sizeof(double) * g, cudaMemcpyDeviceToDevice, p.stream));   // t slot 1
      CU(cudaMemcpyAsync(scal_steps(p.scal, NP) + p0, scal_steps(sstep, g), sizeof(long long) * g, cudaMemcpyDeviceToDevice, p.stream));
    }
    if (dt_out) CU(cudaMemcpyAsync(dt_out, p.dt_hist, sizeof(double) * NP, cudaMemcpyDeviceToHost, p.stream));
    CU(cudaStreamSynchronize(p.stream));
    CU(cudaStreamSynchronize(p.s_d2h));
    c->n = 1;
    c->has_state = true;
    c->speculation_hits += 1;
  } else {
    if (spec) CU(cudaStreamSynchronize(p.s_d2h));
    CU(cudaMemcpyAsync(p.scal, truth.data(), sizeof(double) * NP, cudaMemcpyHostToDevice, p.stream));   // slot 0 = true max(lambda) per problem
    c->n = 0;
    rc = enqueue_step(c, flux, cfl, dx, 1.0e300, 0, 1); if (rc) return rc;
    if (dt_out) CU(cudaMemcpyAsync(dt_out, p.dt_hist, sizeof(double) * NP, cudaMemcpyDeviceToHost, p.stream));
    for (int i = 0; i < K; ++i) {
      const int64_t p0 = (int64_t)i * G, g = std::min<int64_t>(G, NP - p0), c0 = p0 * NC, nc = g * NC;
      rc = transpose_range(c->model, false, p.Q[1] + c0, p.stage_out + c0 * nvar, nc, p.prob.stride, p.stream); if (rc) return rc;
      CU(cudaEventRecord(p.ev_out[i], p.stream));
      CU(cudaStreamWaitEvent(p.s_d2h, p.ev_out[i], 0));
      CU(cudaMemcpyAsync(Qout + c0 * nvar, p.stage_out + c0 * nvar, (size_t)nc * cellb, cudaMemcpyDeviceToHost, p.s_d2h));
    }
    CU(cudaStreamSynchronize(p.stream));
    CU(cudaStreamSynchronize(p.s_d2h));
    c->has_state = true;
    const int rs = read_status(c);
    if (rs) return rs;
  }
  if (st_all & 4) return fail(HS_ERR_CUDA, "tile copy (TMA) did not complete: internal error of the single-phase step kernel");
  return st_all ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
}

int hs_step_host(hs_ctx_t* c, int flux, double cfl, double dx, const double* Qin, double* Qout, double* dt_out) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (!Qin || !Qout) return fail(HS_ERR_ARG, "null Q");
  if (flux != HS_FLUX_HLL && flux != HS_FLUX_LXF) return fail(HS_ERR_ARG, "unknown flux");
  const char* eoff = std::getenv("HS_HOST_PIPELINE");
  const bool off = eoff && eoff[0] == '0';
  // one grid on one device (at least half a chunk of cells): the pipelined form; ensembles on one device: the same by groups of whole
  // problems; everything else (several devices, small grids, odd cell counts) takes upload + step + download
  const bool pipe = !off && c->parts.size() == 1 && c->nprob == 1 && c->ncells % 2 == 0 && c->ncells >= host_chunk_cells() / 2 && c->ncells >= 4096;
  // an ensemble on one device: groups of whole problems, at least two groups (even problem length keeps the groups on the tensor-map copies)
  const bool pipe_ens = !off && c->parts.size() == 1 && c->nprob > 1 && c->ncells % 2 == 0 && c->ncells * c->nprob >= 4096 &&
                        c->ncells * c->nprob >= 2 * std::max<int64_t>(host_chunk_cells(), c->ncells);
  if (pipe) return step_host_pipelined(c, flux, cfl, dx, Qin, Qout, dt_out);
  if (pipe_ens) return step_host_pipelined_ensemble(c, flux, cfl, dx, Qin, Qout, dt_out);
  return step_host_plain(c, flux, cfl, dx, Qin, Qout, dt_out);
}

int hs_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(HS_ERR_ARG, "null buffer");
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return HS_OK; }
  CU(e);
  return HS_OK;
}

int hs_host_unregister(void* ptr) {
  if (!ptr) return fail(HS_ERR_ARG, "null buffer");
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return HS_OK; }
  CU(e);
  return HS_OK;
}

int hs_step_host_stats(hs_ctx_t* c, int64_t* pipelined_calls, int64_t* speculation_hits) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (pipelined_calls) *pipelined_calls = c->pipelined_calls;
  if (speculation_hits) *speculation_hits = c->speculation_hits;
  return HS_OK;
}

// ---------------------------------------------------------------------------------------------
// Dimension-split 2-D solver: Godunov splitting Q^{n+1} = Y(dt) X(dt) Q^n on an nx x ny grid.  X = the 1-D
// step on every grid row (an ensemble of ny rows of nx cells with ONE dt), Y = the same kernels on the columns of the grid seen
// from the frame rotated by R e_2 = e_1 (k_transpose_rot).  dt = cfl min(dx / max lambda_x, dy / max lambda_y); every sweep is the
// reference's 1-D step on each grid line INCLUDING its boundary rule (first and last cell of the line frozen, main.jl:219-220): the
// x-sweep freezes the first / last column, the y-sweep the first / last row, so only the four corner cells never change.  The reference driver is 1-D; this is the "natural growth of
// the same kernels" the physics' normal argument (EquationsOfState.jl:223, HyperelasticityMPh.jl:264) points to, validated by
// what can be validated exactly: a grid uniform in y reproduces the 1-D solver bit for bit in every row, a grid uniform in x
// with states rotated by R^T reproduces it in every column, and the first dt equals the one from get_eigvals with normals e_1, e_2.
// ---------------------------------------------------------------------------------------------
struct hs2d_ctx {
  int model = 0, nvar = 0, nphase = 0, device = 0;
  int64_t nx = 0, ny = 0;
  hsd_problem_t px, py;          // ensembles: ny rows of nx cells / nx columns of ny cells
  cudaStream_t stream = nullptr;
  double* Qx[2] = {nullptr, nullptr};  double* Qy[2] = {nullptr, nullptr};
  double* ax[2] = {nullptr, nullptr};  double* ay[2] = {nullptr, nullptr};
  double* sx = nullptr; double* sy = nullptr;   // scalar blocks of the two ensembles
  double* clock = nullptr;       // [t, dt, steps, lambda_x, lambda_y]
  double* stage = nullptr;
  int64_t n = 0;                 // steps taken since upload (selects the scalar slots of both ensembles)
  RotMap fwd, back;
};

static RotMap make_rotmap(int model, bool forward) {
  RotMap m;
  std::memset(&m, 0, sizeof m);
  m.nvar = model == HS_MODEL_MPH30 ? 30 : 13;
  for (int v = 0; v < 30; ++v) { m.src[v] = v; m.sign[v] = 1.0; }
  // forward (state in the frame rotated by R = [[0,1,0],[-1,0,0],[0,0,1]]): component 1' = component 2, component 2' = -component 1
  // back (R^T): component 1 = -component 2', component 2 = component 1'
  auto pair = [&](int i1, int i2) {   // variables holding the first / second spatial component of a vector (u, or a column of F)
    if (forward) { m.src[i1] = i2; m.sign[i1] = 1.0; m.src[i2] = i1; m.sign[i2] = -1.0; }
    else { m.src[i1] = i2; m.sign[i1] = -1.0; m.src[i2] = i1; m.sign[i2] = 1.0; }
  };
  if (model == HS_MODEL_MPH30) {
    for (int p = 0; p < 2; ++p) {
      pair(15 * p + 2, 15 * p + 3);                                         // momentum
      for (int j = 0; j < 3; ++j) pair(15 * p + 6 + 3 * j, 15 * p + 7 + 3 * j);   // A column-major: rows 1, 2 of column j
    }
  } else {
    pair(0, 1);                                                             // momentum
    for (int j = 0; j < 3; ++j) pair(3 + j, 6 + j);                         // rho F row-major: rows 1, 2, column j
  }
  return m;
}

static int transpose_rot(hs2d_ctx* c, const double* in, double* out, int64_t rows, int64_t cols, const RotMap& map) {
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)c->nvar), block(32, 8);
  k_transpose_rot<<<grid, block, 0, c->stream>>>(in, out, (int)rows, (int)cols, map);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

int hs2d_destroy(hs2d_ctx_t* c) {
  if (!c) return HS_OK;
  DeviceGuard g(c->device);
  for (int k = 0; k < 2; ++k) { cudaFree(c->Qx[k]); cudaFree(c->Qy[k]); cudaFree(c->ax[k]); cudaFree(c->ay[k]); }
  cudaFree(c->sx); cudaFree(c->sy); cudaFree(c->clock); cudaFree(c->stage);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return HS_OK;
}

int hs2d_create(hs2d_ctx_t** out, int model, const hs_barton2009_t* eos, int nphase, int64_t nx, int64_t ny, int device) {
  if (!out) return fail(HS_ERR_ARG, "null ctx pointer");
  *out = nullptr;
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  if (nx < 3 || ny < 3) return fail(HS_ERR_ARG, "need nx >= 3 and ny >= 3");
  if (nx * ny > 0x7fffffffLL) return fail(HS_ERR_ARG, "nx * ny exceeds 2^31-1");
  hs2d_ctx* c = new hs2d_ctx();
  int rc = hsd_problem_init(&c->px, model, eos, nphase, nx, ny);
  if (!rc) rc = hsd_problem_init(&c->py, model, eos, nphase, ny, nx);
  if (rc) { delete c; return rc; }
  c->model = model; c->nvar = model == HS_MODEL_MPH30 ? 30 : 13; c->nphase = nphase; c->nx = nx; c->ny = ny; c->device = device;
  c->fwd = make_rotmap(model, true); c->back = make_rotmap(model, false);
  DeviceGuard g(device);
  if (!g.ok) { delete c; return fail(HS_ERR_CUDA, "cudaSetDevice failed"); }
  const size_t nq = (size_t)c->nvar * nx * ny * sizeof(double), na = (size_t)HS_NAUX(model) * nx * ny * sizeof(double);
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = cudaMalloc(&c->Qx[k], nq);
    if (e == cudaSuccess) e = cudaMalloc(&c->Qy[k], nq);
    if (e == cudaSuccess) e = cudaMalloc(&c->ax[k], na);
    if (e == cudaSuccess) e = cudaMalloc(&c->ay[k], na);
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->sx, sizeof(double) * HS_SCAL_DOUBLES(ny));
  if (e == cudaSuccess) e = cudaMalloc(&c->sy, sizeof(double) * HS_SCAL_DOUBLES(nx));
  if (e == cudaSuccess) e = cudaMalloc(&c->clock, sizeof(double) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&c->stage, nq);
  if (e != cudaSuccess) { g_err = std::string("device allocation failed: ") + cudaGetErrorString(e); hs2d_destroy(c); return HS_ERR_CUDA; }
  *out = c;
  return HS_OK;
}

static int hs2d_status(hs2d_ctx* c) {
  int stx = 0, sty = 0;
  CU(cudaMemcpyAsync(&stx, scal_status(c->sx, c->ny), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&sty, scal_status(c->sy, c->nx), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if ((stx | sty) & 4) return fail(HS_ERR_CUDA, "tile copy (TMA) did not complete: internal error of the single-phase step kernel");
  if (stx | sty) return fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError");
  return HS_OK;
}

/* Q: (nvar, nx, ny) column-major = Julia Array{Float64,3}(nvar, nx, ny): cell (i, j) is record i + nx j */
int hs2d_upload(hs2d_ctx_t* c, const double* Q) {
  if (!c || !Q) return fail(HS_ERR_ARG, "null argument");
  DeviceGuard g(c->device);
  const size_t nq = (size_t)c->nvar * c->nx * c->ny * sizeof(double);
  CU(cudaMemcpyAsync(c->stage, Q, nq, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->sx, 0, sizeof(double) * HS_SCAL_DOUBLES(c->ny), c->stream));
  CU(cudaMemsetAsync(c->sy, 0, sizeof(double) * HS_SCAL_DOUBLES(c->nx), c->stream));
  CU(cudaMemsetAsync(c->clock, 0, sizeof(double) * 8, c->stream));
  c->n = 0;
  int rc = hsd_aos_to_soa(&c->px, c->stage, c->Qx[0], c->stream); if (rc) return rc;
  rc = hsd_wave_bounds(&c->px, c->Qx[0], c->ax[0], c->sx, 0, c->stream); if (rc) return rc;               // lambda_x per row, cache rows
  rc = transpose_rot(c, c->Qx[0], c->Qy[0], c->ny, c->nx, c->fwd); if (rc) return rc;
  rc = hsd_wave_bounds(&c->py, c->Qy[0], c->ay[0], c->sy, 0, c->stream); if (rc) return rc;               // lambda_y per column
  return hs2d_status(c);
}

int hs2d_download(hs2d_ctx_t* c, double* Q) {
  if (!c || !Q) return fail(HS_ERR_ARG, "null argument");
  DeviceGuard g(c->device);
  int rc = hsd_soa_to_aos(&c->px, c->Qx[0], c->stage, c->stream); if (rc) return rc;
  CU(cudaMemcpyAsync(Q, c->stage, (size_t)c->nvar * c->nx * c->ny * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return HS_OK;
}

static int hs2d_enqueue_step(hs2d_ctx* c, int flux, double cfl, double dx, double dy, double t_end) {
  const int64_t n = c->n;
  const int cur = (int)(n % 3), nxt = (int)((n + 1) % 3);
  k_dt2d<<<1, 256, 0, c->stream>>>(scal_lam(c->sx) + (size_t)cur * c->ny, (int)c->ny, scal_lam(c->sy) + (size_t)cur * c->nx, (int)c->nx,
                                  cfl, dx, dy, t_end, c->clock);
  g_launches++;
  CU(cudaGetLastError());
  const double* dt = c->clock + 1;
  // X(dt): rows
  int rc = step_impl(&c->px, flux, cfl, dx, t_end, n, c->Qx[0], c->ax[0], c->Qx[1], c->ax[1], c->sx, nullptr, 0, 0, 0, dt, c->stream); if (rc) return rc;
  // columns of the intermediate state in the rotated frame, their cache rows
  rc = transpose_rot(c, c->Qx[1], c->Qy[0], c->ny, c->nx, c->fwd); if (rc) return rc;
  rc = hsd_wave_bounds(&c->py, c->Qy[0], c->ay[0], c->sy, cur, c->stream); if (rc) return rc;
  // Y(dt): columns; the tail of the step leaves lambda_y of the NEW state in slot nxt
  rc = step_impl(&c->py, flux, cfl, dy, t_end, n, c->Qy[0], c->ay[0], c->Qy[1], c->ay[1], c->sy, nullptr, 0, 0, 0, dt, c->stream); if (rc) return rc;
  // back to rows; cache rows and lambda_x of the new state (slot nxt of the row ensemble)
  rc = transpose_rot(c, c->Qy[1], c->Qx[0], c->nx, c->ny, c->back); if (rc) return rc;
  rc = hsd_wave_bounds(&c->px, c->Qx[0], c->ax[0], c->sx, nxt, c->stream); if (rc) return rc;
  c->n += 1;
  return HS_OK;
}

int hs2d_step(hs2d_ctx_t* c, int flux, double cfl, double dx, double dy, double* dt_out) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (flux != HS_FLUX_HLL && flux != HS_FLUX_LXF) return fail(HS_ERR_ARG, "unknown flux");
  DeviceGuard g(c->device);
  int rc = hs2d_enqueue_step(c, flux, cfl, dx, dy, 1.0e300); if (rc) return rc;
  if (dt_out) CU(cudaMemcpyAsync(dt_out, c->clock + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return hs2d_status(c);
}

int hs2d_advance(hs2d_ctx_t* c, int flux, double cfl, double dx, double dy, double t_end, int64_t max_steps, double* t_out, int64_t* steps_out) {
  if (!c) return fail(HS_ERR_ARG, "null context");
  if (flux != HS_FLUX_HLL && flux != HS_FLUX_LXF) return fail(HS_ERR_ARG, "unknown flux");
  if (max_steps < 0) return fail(HS_ERR_ARG, "max_steps < 0");
  DeviceGuard g(c->device);
  double clk[3] = {0, 0, 0};
  int64_t done = 0;
  while (done < max_steps) {
    CU(cudaMemcpyAsync(clk, c->clock, sizeof clk, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!(clk[0] < t_end)) break;
    int64_t m = 16;
    if (clk[1] > 0.0) { m = (int64_t)((t_end - clk[0]) / clk[1]) - 2; if (m > 4096) m = 4096; if (m < 1) m = 1; }
    if (m > max_steps - done) m = max_steps - done;
    for (int64_t k = 0; k < m; ++k) { int rc = hs2d_enqueue_step(c, flux, cfl, dx, dy, t_end); if (rc) return rc; }
    done += m;
  }
  CU(cudaMemcpyAsync(clk, c->clock, sizeof clk, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (t_out) *t_out = clk[0];
  if (steps_out) *steps_out = (int64_t)clk[2];
  return hs2d_status(c);
}

// ---------------------------------------------------------------------------------------------
// stateless batches
// ---------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(double)); }
};

int stateless_prolog(int model, const hs_barton2009_t* eos, int nphase, int64_t n, int device, hsd_problem_t* p) {
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  if (n < 1) return fail(HS_ERR_ARG, "n < 1");
  return hsd_problem_init(p, model, eos, nphase, 3, 1);
}

template <int OP>
int cellop(int model, const hs_barton2009_t* eos, int nphase, const double* in, double* out, int64_t n, int device) {
  hsd_problem_t p;
  int rc = stateless_prolog(model, eos, nphase, n, device, &p); if (rc) return rc;
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  if (!in || !out) return fail(HS_ERR_ARG, "null array");
  const int nvar = model == HS_MODEL_MPH30 ? 30 : 13;
  DevBuf din, dout, dst;
  CU(din.alloc((size_t)nvar * n)); CU(dout.alloc((size_t)nvar * n)); CU(dst.alloc(1));
  CU(cudaMemcpy(din.p, in, sizeof(double) * nvar * n, cudaMemcpyHostToDevice));
  CU(cudaMemset(dst.p, 0, sizeof(double)));
  const EosPair e = eos_pair(&p);
  const unsigned nb = (unsigned)((n * nphase + 127) / 128);
  int* st = reinterpret_cast<int*>(dst.p);
  if (model == HS_MODEL_MPH30) {
    if (p.gen) k_cellop<MODEL_MPH30, true, OP><<<nb, 128>>>(din.p, dout.p, n, e, st);
    else k_cellop<MODEL_MPH30, false, OP><<<nb, 128>>>(din.p, dout.p, n, e, st);
  } else {
    if (p.gen) k_cellop<MODEL_SP13, true, OP><<<nb, 128>>>(din.p, dout.p, n, e, st);
    else k_cellop<MODEL_SP13, false, OP><<<nb, 128>>>(din.p, dout.p, n, e, st);
  }
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, dout.p, sizeof(double) * nvar * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  CU(cudaMemcpy(&bad, st, sizeof(int), cudaMemcpyDeviceToHost));
  return bad ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
}

template <int MODEL, int FLUX>
int faceop_launch(const hsd_problem_t& p, const double* ql, const double* qr, const double* el, const double* er, double lambda,
                  double* cons, double* dm, double* dp, double* s, int64_t n, int* st) {
  constexpr int T = T_FACE;
  const size_t smem = sizeof(double) * 75 * T;
  const EosPair e = eos_pair(&p);
  const unsigned nb = (unsigned)((n * ModelTraits<MODEL>::NPH + T - 1) / T);
  if (p.gen) k_faceop<MODEL, FLUX, true, T><<<nb, T, smem>>>(ql, qr, el, er, lambda, cons, dm, dp, s, n, e, st);
  else k_faceop<MODEL, FLUX, false, T><<<nb, T, smem>>>(ql, qr, el, er, lambda, cons, dm, dp, s, n, e, st);
  g_launches++;
  CU(cudaGetLastError());
  return HS_OK;
}

int faceop(int model, int flux, const hs_barton2009_t* eos, int nphase, const double* Ql, const double* Qr, const double* eig_l,
           const double* eig_r, double lambda, double* cons, double* dm, double* dp, double* s, int64_t n, int device) {
  hsd_problem_t p;
  int rc = stateless_prolog(model, eos, nphase, n, device, &p); if (rc) return rc;
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  if (!Ql || !Qr) return fail(HS_ERR_ARG, "null array");
  if (flux == HS_FLUX_HLL && (!eig_l || !eig_r)) return fail(HS_ERR_ARG, "hll needs the cached eigvals of both cells");
  const int nvar = model == HS_MODEL_MPH30 ? 30 : 13, neig = 6 * nphase;
  DevBuf dl, dr, del, der, dc, dmm, dpp, ds, dst;
  CU(dl.alloc((size_t)nvar * n)); CU(dr.alloc((size_t)nvar * n));
  CU(dc.alloc((size_t)nvar * n)); CU(dmm.alloc((size_t)nvar * n)); CU(dpp.alloc((size_t)nvar * n));
  CU(ds.alloc((size_t)2 * n)); CU(dst.alloc(1));
  CU(cudaMemcpy(dl.p, Ql, sizeof(double) * nvar * n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dr.p, Qr, sizeof(double) * nvar * n, cudaMemcpyHostToDevice));
  if (flux == HS_FLUX_HLL) {
    CU(del.alloc((size_t)neig * n)); CU(der.alloc((size_t)neig * n));
    CU(cudaMemcpy(del.p, eig_l, sizeof(double) * neig * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(der.p, eig_r, sizeof(double) * neig * n, cudaMemcpyHostToDevice));
  }
  CU(cudaMemset(dst.p, 0, sizeof(double)));
  CU(cudaMemset(ds.p, 0, sizeof(double) * 2 * n));
  int* st = reinterpret_cast<int*>(dst.p);
  if (model == HS_MODEL_MPH30) {
    rc = flux == HS_FLUX_HLL ? faceop_launch<MODEL_MPH30, FLUX_HLL>(p, dl.p, dr.p, del.p, der.p, lambda, dc.p, dmm.p, dpp.p, ds.p, n, st)
                             : faceop_launch<MODEL_MPH30, FLUX_LXF>(p, dl.p, dr.p, del.p, der.p, lambda, dc.p, dmm.p, dpp.p, ds.p, n, st);
  } else {
    rc = flux == HS_FLUX_HLL ? faceop_launch<MODEL_SP13, FLUX_HLL>(p, dl.p, dr.p, del.p, der.p, lambda, dc.p, dmm.p, dpp.p, ds.p, n, st)
                             : faceop_launch<MODEL_SP13, FLUX_LXF>(p, dl.p, dr.p, del.p, der.p, lambda, dc.p, dmm.p, dpp.p, ds.p, n, st);
  }
  if (rc) return rc;
  if (cons) CU(cudaMemcpy(cons, dc.p, sizeof(double) * nvar * n, cudaMemcpyDeviceToHost));
  if (dm) CU(cudaMemcpy(dm, dmm.p, sizeof(double) * nvar * n, cudaMemcpyDeviceToHost));
  if (dp) CU(cudaMemcpy(dp, dpp.p, sizeof(double) * nvar * n, cudaMemcpyDeviceToHost));
  if (s) CU(cudaMemcpy(s, ds.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  CU(cudaMemcpy(&bad, st, sizeof(int), cudaMemcpyDeviceToHost));
  return bad ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
}
}  // namespace

extern "C" {
// device self-test hooks (tests/ only): the hot path's branch-free reciprocal / rsqrt / sqrt and its
// largest-eigenvalue solve evaluated on caller data
int hs_selftest_math(const double* x, double* rcp, double* rsq, double* sq, int64_t n, int device) {
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  DevBuf dx, d1, d2, d3;
  CU(dx.alloc(n)); CU(d1.alloc(n)); CU(d2.alloc(n)); CU(d3.alloc(n));
  CU(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_selftest_math<<<(unsigned)((n + 127) / 128), 128>>>(dx.p, d1.p, d2.p, d3.p, n);
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy(rcp, d1.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(rsq, d2.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(sq, d3.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return HS_OK;
}

int hs_selftest_eig(const double* s6, double* lam_max_abs, int64_t n, int device) {
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  DevBuf di, dout;
  CU(di.alloc(6 * n)); CU(dout.alloc(n));
  CU(cudaMemcpy(di.p, s6, sizeof(double) * 6 * n, cudaMemcpyHostToDevice));
  k_selftest_eig<<<(unsigned)((n + 127) / 128), 128>>>(di.p, dout.p, n);
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy(lam_max_abs, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return HS_OK;
}

}  // extern "C"
namespace {
// Hank2016 batches: scalars s0[n], s1[n] (s1 may be NULL for the stress: its pressure argument does not enter),
// tensor (NT, n), result (NO, n)
template <int OP>
int hank_batch(const hs_hank2016_t* eos, const double* s0, const double* s1, const double* ten, double* out, int64_t n, int device) {
  constexpr int NT = (OP == HANK_PRESSURE) ? 3 : 9, NO = (OP == HANK_STRESS) ? 9 : 1;
  if (hs_device_count() <= 0) return fail(HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
  if (n < 1) return fail(HS_ERR_ARG, "n < 1");
  if (!eos || !s0 || !ten || !out || (OP != HANK_STRESS && !s1)) return fail(HS_ERR_ARG, "null array");
  if (!(eos->rho0 > 0.0) || eos->gamma == 1.0) return fail(HS_ERR_ARG, "Hank2016 needs rho0 > 0 and gamma != 1");
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  DevBuf d0, d1, dt, dout, dst;
  CU(d0.alloc(n)); CU(dt.alloc((size_t)NT * n)); CU(dout.alloc((size_t)NO * n)); CU(dst.alloc(1));
  CU(cudaMemcpy(d0.p, s0, sizeof(double) * n, cudaMemcpyHostToDevice));
  if (OP != HANK_STRESS) { CU(d1.alloc(n)); CU(cudaMemcpy(d1.p, s1, sizeof(double) * n, cudaMemcpyHostToDevice)); }
  CU(cudaMemcpy(dt.p, ten, sizeof(double) * NT * n, cudaMemcpyHostToDevice));
  CU(cudaMemset(dst.p, 0, sizeof(double)));
  HankAbi e = {eos->rho0, eos->mu, eos->gamma, eos->pres_inf, eos->a};
  int* st = reinterpret_cast<int*>(dst.p);
  k_hank<OP><<<(unsigned)((n + 127) / 128), 128>>>(e, d0.p, d1.p, dt.p, dout.p, n, st);
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, dout.p, sizeof(double) * NO * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  CU(cudaMemcpy(&bad, st, sizeof(int), cudaMemcpyDeviceToHost));
  return bad ? fail(HS_ERR_DOMAIN, "det G <= 0: Julia's fractional power would throw DomainError") : HS_OK;
}
}  // namespace
extern "C" {
int hs_hank2016_energy(const hs_hank2016_t* eos, const double* den, const double* pres, const double* G, double* e_int, int64_t n, int device) {
  return hank_batch<HANK_ENERGY>(eos, den, pres, G, e_int, n, device);
}
int hs_hank2016_pressure(const hs_hank2016_t* eos, const double* den, const double* e_int, const double* inv3, double* pres, int64_t n, int device) {
  return hank_batch<HANK_PRESSURE>(eos, den, e_int, inv3, pres, n, device);
}
int hs_hank2016_stress(const hs_hank2016_t* eos, const double* den, const double* pres, const double* distortion, double* sigma, int64_t n, int device) {
  return hank_batch<HANK_STRESS>(eos, den, pres, distortion, sigma, n, device);
}

int hs_cons2prim(int model, const hs_barton2009_t* eos, int nphase, const double* Q, double* P, int64_t n, int device) {
  return cellop<OP_CONS2PRIM>(model, eos, nphase, Q, P, n, device);
}
int hs_prim2cons(int model, const hs_barton2009_t* eos, int nphase, const double* P, double* Q, int64_t n, int device) {
  return cellop<OP_PRIM2CONS>(model, eos, nphase, P, Q, n, device);
}
int hs_flux(int model, const hs_barton2009_t* eos, int nphase, const double* Q, double* F, int64_t n, int device) {
  return cellop<OP_FLUX>(model, eos, nphase, Q, F, n, device);
}

int hs_noncons_flux(const hs_barton2009_t* eos, const double* Q, double* col, double* Bdense, int64_t n, int device) {
  std::vector<double> tmp;
  double* c = col;
  if (!c) { tmp.resize((size_t)30 * n); c = tmp.data(); }
  int rc = cellop<OP_NONCONS>(HS_MODEL_MPH30, eos, 2, Q, c, n, device);
  if (rc && rc != HS_ERR_DOMAIN) return rc;
  if (Bdense) {  // block-diagonal, only column 1 of each block is non-zero (HyperelasticityMPh.jl:221-248)
    std::memset(Bdense, 0, sizeof(double) * 900 * n);
    for (int64_t i = 0; i < n; ++i)
      for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 15; ++r) Bdense[900 * i + (15 * p + r) + 30 * (15 * p)] = c[30 * i + 15 * p + r];
  }
  return rc;
}

int hs_get_eigvals(int model, const hs_barton2009_t* eos, int nphase, const double* Q, const double* normal, double* eig,
                   int64_t n, int device) {
  Normal3 nrm = {{1.0, 0.0, 0.0}};
  if (normal) {
    const double n2 = normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2];
    if (!(fabs(n2 - 1.0) < 1e-12)) return fail(HS_ERR_ARG, "the normal must be a unit vector");
    nrm.n[0] = normal[0]; nrm.n[1] = normal[1]; nrm.n[2] = normal[2];
  }
  hsd_problem_t p;
  int rc = stateless_prolog(model, eos, nphase, n, device, &p); if (rc) return rc;
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail(HS_ERR_CUDA, "cudaSetDevice failed");
  if (!Q || !eig) return fail(HS_ERR_ARG, "null array");
  const int nvar = model == HS_MODEL_MPH30 ? 30 : 13, neig = 6 * nphase;
  DevBuf din, dout, dst;
  CU(din.alloc((size_t)nvar * n)); CU(dout.alloc((size_t)neig * n)); CU(dst.alloc(1));
  CU(cudaMemcpy(din.p, Q, sizeof(double) * nvar * n, cudaMemcpyHostToDevice));
  CU(cudaMemset(dst.p, 0, sizeof(double)));
  const EosPair e = eos_pair(&p);
  const unsigned nb = (unsigned)((n * nphase + 127) / 128);
  int* st = reinterpret_cast<int*>(dst.p);
  if (model == HS_MODEL_MPH30) {
    if (p.gen) k_eigvals<MODEL_MPH30, true><<<nb, 128>>>(din.p, dout.p, n, e, nrm, st); else k_eigvals<MODEL_MPH30, false><<<nb, 128>>>(din.p, dout.p, n, e, nrm, st);
  } else {
    if (p.gen) k_eigvals<MODEL_SP13, true><<<nb, 128>>>(din.p, dout.p, n, e, nrm, st); else k_eigvals<MODEL_SP13, false><<<nb, 128>>>(din.p, dout.p, n, e, nrm, st);
  }
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy(eig, dout.p, sizeof(double) * neig * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  CU(cudaMemcpy(&bad, st, sizeof(int), cudaMemcpyDeviceToHost));
  return bad ? fail(HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError") : HS_OK;
}

int hs_hll(int model, const hs_barton2009_t* eos, int nphase, const double* Ql, const double* Qr, const double* eig_l,
           const double* eig_r, double* cons, double* dm, double* dp, double* s, int64_t n, int device) {
  return faceop(model, HS_FLUX_HLL, eos, nphase, Ql, Qr, eig_l, eig_r, 0.0, cons, dm, dp, s, n, device);
}
int hs_lxf(int model, const hs_barton2009_t* eos, int nphase, const double* Ql, const double* Qr, double lambda, double* cons,
           double* dm, double* dp, int64_t n, int device) {
  return faceop(model, HS_FLUX_LXF, eos, nphase, Ql, Qr, nullptr, nullptr, lambda, cons, dm, dp, nullptr, n, device);
}

}  // extern "C"
