// lwfb90.cu — kernels and C ABI of liblwfb90 (see include/lwfb90.h).
//
// Kernels:
//   integrate_lanes<NMAX,EQ> (lwfb90_lanes.cuh)  production integrator: one lane per soil layer, 32/N members per warp,
//                          Tsit5 stages + RHS + error norm + PI controller + per-step isotope transport + daily callbacks;
//                          launched in work quanta by lwfb_run (checkpoint anywhere, regrouping by progress).
//   integrate_kernel<NT> (lwfb90_member.cu)  thread-per-member fallback for N < 3 or N > 32, and rhs_kernel (one RHS
//                          evaluation per member, unit parity tests).
//   placement_keys / spread_placement  sort keys and thin slot assignment of the regrouping (with cub's radix sort).
//   score_sums_kernel / score_finalize_kernel  per-member RMSE / KGE against observations from the daily record or from
//                          the sums the integrator accumulated itself (lwfb_set_observations).
// All device blocks are member-major (the ABI's own layouts): no transposes at the boundary.
//   water_partitioning_kernel, soil_state_kernel (lwfb90_post.cuh)  post-processing of the resident daily record.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "lwfb90_device.cuh"
#include "lwfb90_lanes.cuh"
#include "lwfb90_launch.h"
#include "lwfb90_post.cuh"
#include "lwfb90_setup.cuh"

using namespace lwfb;

// does any member use one of the optional flow paths of the right-hand side (bypass flow, source-area flow, downslope
// flow, groundwater storage feedback)?  If none does, the lean instantiation of the lane kernel applies.
__global__ void optional_paths_kernel(const double* __restrict__ scal, int64_t M, int* __restrict__ used) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool u = false;
  if (m < M) {
    const double* s = scal + m * LWFB_PS;
    const bool slope = !(s[LWFB_S_LENGTH_SLOPE] == 0 || s[LWFB_S_DSLOPE] == 0);
    u = !(s[LWFB_S_GSC] < 1.0e-8) || s[LWFB_S_BYPAR] == 1.0 || (int)s[LWFB_S_QLAYER] > 0 || slope;
  }
  if (__any_sync(0xffffffffu, u) && (threadIdx.x & 31) == 0) atomicOr(used, 1);
}

// sort keys of the placement (ascending radix sort, stable: ties keep the member order)
//   by_cost: RHS evaluations of the previous run, most first
//   else   : progress of the unfinished members (simulation time reached, least first); finished members get the
//            largest key and are counted out of *n_left
__global__ void placement_keys(const double* __restrict__ ck, int ck_rows, const int32_t* __restrict__ status,
                               const long long* __restrict__ counters, int64_t M, double t0, double tend, int by_cost,
                               uint32_t* __restrict__ key, int32_t* __restrict__ idx, int* __restrict__ n_left) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool unfinished = false;
  if (i < M) {
    uint32_t k;
    if (by_cost) {
      const long long c = counters[i * LWFB_NCOUNTERS + LWFB_C_NF];
      k = 0xffffffffu - (uint32_t)(c < 0 ? 0 : (c > 0xfffffffeLL ? 0xfffffffeLL : c));
    } else {
      const double ti = ck[i * ck_rows + LWFB_CK_T];
      unfinished = status[i] == LWFB_ST_OK && ti < tend;
      k = unfinished ? __float_as_uint((float)(ti - t0)) : 0xffffffffu;   // >= 0: the bit pattern orders like the value
    }
    key[i] = k;
    idx[i] = (int32_t)i;
  }
  const unsigned b = __ballot_sync(0xffffffffu, unfinished);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_left, __popc(b));
}

// thin placement for a device that is not full: sorted member j goes to slot (j / per_block) * slots_per_block +
// (j % per_block), i.e. every block gets members for its first per_block slots only and the rest stays empty (-1)
__global__ void spread_placement(const int32_t* __restrict__ sorted, int64_t n, int per_block, int slots_per_block,
                                 int32_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[(j / per_block) * slots_per_block + (j % per_block)] = sorted[j];
}

// thin FRONT blocks of a full device (experimental, LWFB_THIN_FRONT): the first n_front blocks get per_front members
// each (the rest of their slots stays empty), the other members are packed behind them
__global__ void spread_front(const int32_t* __restrict__ sorted, int64_t n, int64_t n_front, int per_front,
                             int slots_per_block, int32_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t thin = n_front * per_front;
  out[j < thin ? (j / per_front) * slots_per_block + (j % per_front) : n_front * slots_per_block + (j - thin)] = sorted[j];
}

// per-member goodness-of-fit sums against observed theta / delta18O from the resident daily record: the same
// expressions in the same order (days outer, observed layers inner) as the integrator's own accumulation (score_rec in
// lwfb90_lanes.cuh), so that both ways give bit-identical scores.  One thread per member (a few hundred scattered reads).
__global__ void score_sums_kernel(DevData d, int n_obs_days, const int32_t* __restrict__ obs_day, int n_obs_layers,
                                  const int32_t* __restrict__ obs_layer, const double* __restrict__ obs_theta,
                                  const double* __restrict__ obs_d18o, double* __restrict__ sacc) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.M) return;
  const int N = d.N, L = d.L;
  double a[2 * LWFB_SA];
  for (int k = 0; k < 2 * LWFB_SA; k++) a[k] = 0.0;
  for (int k = 0; k < n_obs_days; k++) {
    const double* r = d.rec + (m * (int64_t)(d.n_days + 1) + obs_day[k]) * L;
    for (int l = 0; l < n_obs_layers; l++) {
      const int i = obs_layer[l];
      if (obs_theta) {
        const double ths = __ldg(d.lay + (m * LWFB_PL + LWFB_L_THSAT) * N + i);
        const double thr = __ldg(d.lay + (m * LWFB_PL + LWFB_L_THR) * N + i);
        const double smax = __ldg(d.lay + (m * LWFB_PL + LWFB_L_SWATMAX) * N + i);
        score_add(a, score_theta(ths, thr, smax, r[i]), obs_theta[(int64_t)k * n_obs_layers + l]);
      }
      if (obs_d18o && d.iso) score_add(a + LWFB_SA, r[N + i], obs_d18o[(int64_t)k * n_obs_layers + l]);
    }
  }
  for (int k = 0; k < 2 * LWFB_SA; k++) sacc[m * (2 * LWFB_SA) + k] = a[k];
}

// sums -> RMSE / KGE of theta and delta18O
__global__ void score_finalize_kernel(const double* __restrict__ sacc, int64_t M, double* __restrict__ scores) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double res[LWFB_NSCORES];
  score_finish(sacc + m * (2 * LWFB_SA), &res[LWFB_SC_RMSE_THETA], &res[LWFB_SC_KGE_THETA]);
  score_finish(sacc + m * (2 * LWFB_SA) + LWFB_SA, &res[LWFB_SC_RMSE_D18O], &res[LWFB_SC_KGE_D18O]);
  for (int s = 0; s < LWFB_NSCORES; s++) scores[m * LWFB_NSCORES + s] = res[s];
}

// elementary functions of include/lwfb90_elem.h evaluated on the device (host/device bit-equality tests)
__global__ void elem_kernel(int fn, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                            double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (fn) {
    case 0: r = lwfb_exp(x[i]); break;
    case 1: r = lwfb_log(x[i]); break;
    case 2: r = lwfb_pow(x[i], y[i]); break;
    case 3: r = lwfb_sin(x[i]); break;
    case 4: r = lwfb_cos(x[i]); break;
    case 5: r = lwfb_tan(x[i]); break;
    case 6: r = lwfb_asin(x[i]); break;
    case 7: r = lwfb_acos(x[i]); break;
    case 8: r = fastpow(x[i], y[i]); break;
    default: r = NAN;
  }
  out[i] = r;
}

// FP64 pipe peak: 16 independent DFMA chains per thread, no memory traffic
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =============================================================================================
// host side
// =============================================================================================
struct lwfb_handle {
  lwfb_opts o;
  int L = 0;
  int n_sites = 0, n_rows = 0;
  int64_t M = 0;
  double *d_forc = nullptr, *d_scal = nullptr, *d_lay = nullptr, *d_u0 = nullptr, *d_rec = nullptr, *d_fin = nullptr;
  int32_t *d_site = nullptr, *d_status = nullptr;
  long long* d_counters = nullptr;
  double* d_stage = nullptr;  // staging for layout transposes
  size_t stage_bytes = 0;
  double* d_scores = nullptr;
  // observations registered for scoring during the run (lwfb_set_observations) and the running sums [M][2][LWFB_SA]
  int n_obs_days = 0, n_obs_layers = 0;
  int32_t *d_obs_of_rec = nullptr, *d_obs_layer = nullptr;
  double *d_obs_theta = nullptr, *d_obs_d18o = nullptr;
  double *d_sacc = nullptr, *d_sacc_rec = nullptr;   // sums of the run / scratch of the record-based scores
  bool sums_valid = false;     // d_sacc holds the sums of a completed run
  double *d_ntfli0 = nullptr, *d_ntfli_out = nullptr;   // du_NTFLI cache in (lwfb_set_ntfli0) / out (lwfb_get_ntfli), [M][N]
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int n_launches = 0;
  float last_ms = 0.f;
  bool ran = false;
  bool pending = false;        // lwfb_run_async launched, not yet synchronised
  void* pin[2] = {nullptr, nullptr};   // pinned bounce buffers of the downloads
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  bool force_legacy = false;   // testing: run the thread-per-member kernel also for 3 <= N <= 32
  bool lean = false;           // no member uses an optional flow path of the RHS (optional_paths_kernel): lean lane kernel
  bool force_general = false;  // testing: never use the lean / fixed-N instantiations
  // placement of members in the lane kernel's slots (lwfb_set_placement)
  int placement = 1;           // 0 identity, 1 adaptive (cost-sorted from the previous run), 2 the same for any launch size
  int32_t* d_perm = nullptr;
  double* d_ck = nullptr;      // checkpoint block of the work-quantum execution [LWFB_CK + LWFB_CKL * N][M]
  int* d_stop = nullptr;       // {blocks exited, stop flag} of the current launch
  // device-side sorting of the placement: keys in/out, member indices in (out = d_perm), cub scratch, unfinished count
  uint32_t* d_key = nullptr;
  uint32_t* d_key2 = nullptr;
  int32_t* d_idx = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int* d_nleft = nullptr;
  int32_t* d_perm_thin = nullptr;   // placement with empty slots (spread_placement / spread_front)
  size_t perm_thin_len = 0;
  bool perm_valid = false;     // d_perm holds a permutation for the current member count
  bool cost_valid = false;     // d_counters holds the RHS-evaluation counts of a completed lane-kernel run
  std::string err;
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return LWFB_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

static int ensure_done(lwfb_handle* h);
static int alloc_members(lwfb_handle* h, int64_t n_members);
static int classify_members(lwfb_handle* h);
static int download(lwfb_handle* h, void* dst, const void* src, size_t bytes);

static int ensure_stage(lwfb_handle* h, size_t bytes) {
  if (h->stage_bytes >= bytes) return LWFB_OK;
  if (h->d_stage) cudaFree(h->d_stage);
  h->d_stage = nullptr; h->stage_bytes = 0;
  CK(cudaMalloc(&h->d_stage, bytes));
  h->stage_bytes = bytes;
  return LWFB_OK;
}

extern "C" void lwfb_default_opts(lwfb_opts* o) {
  memset(o, 0, sizeof(*o));
  o->n_layers = 0; o->simulate_isotopes = 1; o->t0_day = 0; o->n_days = 0;
  o->reltol = 1e-5; o->abstol = 1e-6; o->dt_init = 1e-3; o->dtmax = 60.0 / 60.0 / 24.0;
  o->maxiters = 1000000; o->device = 0; o->record_daily = 1;
}

extern "C" int lwfb_create(lwfb_handle** out, const lwfb_opts* opts) {
  if (!out || !opts) return LWFB_ERR_ARG;
  *out = nullptr;
  if (opts->n_layers < 1 || opts->n_layers > LWFB_MAX_LAYERS) return LWFB_ERR_UNSUPPORTED;
  if (opts->n_days < 1 || opts->t0_day < 0) return LWFB_ERR_ARG;
  if (opts->n_layers > 32 && opts->n_layers > wide_max_layers(opts->device)) return LWFB_ERR_UNSUPPORTED;   // member block > shared memory of an SM
  lwfb_handle* h = new lwfb_handle();
  h->o = *opts;
  h->L = LWFB_STATE_LEN(opts->n_layers, opts->simulate_isotopes);
  cudaError_t e = cudaSetDevice(opts->device);
  if (e != cudaSuccess) { delete h; return LWFB_ERR_CUDA; }
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  *out = h;
  return LWFB_OK;
}

static void free_members(lwfb_handle* h) {
  cudaFree(h->d_scal); cudaFree(h->d_lay); cudaFree(h->d_u0); cudaFree(h->d_rec); cudaFree(h->d_fin);
  cudaFree(h->d_site); cudaFree(h->d_status); cudaFree(h->d_counters); cudaFree(h->d_scores); cudaFree(h->d_perm); cudaFree(h->d_ck); cudaFree(h->d_stop);
  cudaFree(h->d_key); cudaFree(h->d_key2); cudaFree(h->d_idx); cudaFree(h->d_sort_tmp); cudaFree(h->d_nleft);
  cudaFree(h->d_perm_thin); h->d_perm_thin = nullptr; h->perm_thin_len = 0;
  cudaFree(h->d_sacc); cudaFree(h->d_sacc_rec); h->d_sacc = h->d_sacc_rec = nullptr; h->sums_valid = false;
  cudaFree(h->d_ntfli0); cudaFree(h->d_ntfli_out); h->d_ntfli0 = h->d_ntfli_out = nullptr;
  h->d_key = h->d_key2 = nullptr; h->d_idx = nullptr; h->d_sort_tmp = nullptr; h->d_nleft = nullptr; h->sort_tmp_bytes = 0;
  h->d_perm = nullptr; h->d_ck = nullptr; h->d_stop = nullptr; h->perm_valid = false; h->cost_valid = false;
  h->d_scal = h->d_lay = h->d_u0 = h->d_rec = h->d_fin = h->d_scores = nullptr;
  h->d_site = h->d_status = nullptr; h->d_counters = nullptr;
  h->M = 0; h->ran = false;
}

extern "C" void lwfb_destroy(lwfb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->o.device);
  free_members(h);
  cudaFree(h->d_forc); cudaFree(h->d_stage);
  cudaFree(h->d_obs_of_rec); cudaFree(h->d_obs_layer); cudaFree(h->d_obs_theta); cudaFree(h->d_obs_d18o);
  for (int k = 0; k < 2; k++) { if (h->pin[k]) cudaFreeHost(h->pin[k]); if (h->pin_ev[k]) cudaEventDestroy(h->pin_ev[k]); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
}

extern "C" const char* lwfb_last_error(const lwfb_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int lwfb_set_sites(lwfb_handle* h, int32_t n_sites, int32_t n_rows, const double* forcing) {
  if (!h || !forcing || n_sites < 1) return LWFB_ERR_ARG;
  if (n_rows < h->o.t0_day + h->o.n_days + 1) { h->err = "forcing table shorter than t0_day+n_days+1 rows"; return LWFB_ERR_ARG; }
  CK(cudaSetDevice(h->o.device));
  cudaFree(h->d_forc); h->d_forc = nullptr;
  size_t bytes = (size_t)n_sites * n_rows * LWFB_F * sizeof(double);
  CK(cudaMalloc(&h->d_forc, bytes));
  CK(cudaMemcpy(h->d_forc, forcing, bytes, cudaMemcpyHostToDevice));
  h->n_sites = n_sites; h->n_rows = n_rows;
  return LWFB_OK;
}

extern "C" int lwfb_set_members(lwfb_handle* h, int64_t n_members, const int32_t* site_of_member,
                                const double* scal, const double* lay, const double* u0) {
  if (!h || n_members < 1 || !site_of_member || !scal || !lay || !u0) return LWFB_ERR_ARG;
  if (!h->d_forc) { h->err = "lwfb_set_sites must be called first"; return LWFB_ERR_ARG; }
  for (int64_t i = 0; i < n_members; i++)
    if (site_of_member[i] < 0 || site_of_member[i] >= h->n_sites) { h->err = "site index out of range"; return LWFB_ERR_ARG; }
  CK(cudaSetDevice(h->o.device));
  const int N = h->o.n_layers, L = h->L;
  { int rc = alloc_members(h, n_members); if (rc) return rc; }
  CK(cudaMemcpy(h->d_site, site_of_member, (size_t)h->M * sizeof(int32_t), cudaMemcpyHostToDevice));
  // the device blocks are member-major like the ABI's: plain copies, no transposes
  CK(cudaMemcpy(h->d_scal, scal, (size_t)h->M * LWFB_PS * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_lay, lay, (size_t)h->M * LWFB_PL * N * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_u0, u0, (size_t)h->M * L * sizeof(double), cudaMemcpyHostToDevice));
  { int rc = classify_members(h); if (rc) return rc; }
  cudaFree(h->d_ntfli0); h->d_ntfli0 = nullptr;      // a warm du_NTFLI cache belongs to the members it was set for
  h->ran = false; h->sums_valid = false;
  h->cost_valid = false; h->perm_valid = false;   // counters / placement of the previous members say nothing about these
  return LWFB_OK;
}

// device blocks of an ensemble of n_members members (re)allocated; h->M set
static int alloc_members(lwfb_handle* h, int64_t n_members) {
  const int N = h->o.n_layers, L = h->L;
  if (n_members == h->M) return LWFB_OK;
  free_members(h);
  int64_t M = n_members;
  CK(cudaMalloc(&h->d_scal, (size_t)M * LWFB_PS * sizeof(double)));
  CK(cudaMalloc(&h->d_lay, (size_t)M * LWFB_PL * N * sizeof(double)));
  CK(cudaMalloc(&h->d_u0, (size_t)M * L * sizeof(double)));
  CK(cudaMalloc(&h->d_fin, (size_t)M * L * sizeof(double)));
  CK(cudaMalloc(&h->d_site, (size_t)M * sizeof(int32_t)));
  CK(cudaMalloc(&h->d_status, (size_t)M * sizeof(int32_t)));
  CK(cudaMalloc(&h->d_counters, (size_t)M * LWFB_NCOUNTERS * sizeof(long long)));
  CK(cudaMalloc(&h->d_scores, (size_t)M * LWFB_NSCORES * sizeof(double)));
  CK(cudaMalloc(&h->d_sacc, (size_t)M * 2 * LWFB_SA * sizeof(double)));
  CK(cudaMalloc(&h->d_ntfli_out, (size_t)M * N * sizeof(double)));
  CK(cudaMalloc(&h->d_perm, (size_t)M * sizeof(int32_t)));
  CK(cudaMalloc(&h->d_ck, (size_t)M * (LWFB_CK + LWFB_CKL * N) * sizeof(double)));
  CK(cudaMalloc(&h->d_stop, 2 * sizeof(int)));
  CK(cudaMalloc(&h->d_key, (size_t)M * sizeof(uint32_t)));
  CK(cudaMalloc(&h->d_key2, (size_t)M * sizeof(uint32_t)));
  CK(cudaMalloc(&h->d_idx, (size_t)M * sizeof(int32_t)));
  CK(cudaMalloc(&h->d_nleft, sizeof(int)));
  h->sort_tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, h->sort_tmp_bytes, h->d_key, h->d_key2, h->d_idx, h->d_perm, (int)M));
  CK(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes ? h->sort_tmp_bytes : 16));
  if (h->o.record_daily) {
    size_t bytes = (size_t)M * (h->o.n_days + 1) * L * sizeof(double);
    cudaError_t e = cudaMalloc(&h->d_rec, bytes);
    if (e != cudaSuccess) { h->err = "cannot allocate the daily record (" + std::to_string(bytes >> 20) + " MiB)"; free_members(h); return LWFB_ERR_NOMEM; }
  }
  h->M = M;
  return LWFB_OK;
}

// replaces remakeSPAC -> setup for every member (src/LWFBrook90.jl:214-345, :365-587): builds the member blocks ON THE
// DEVICE from one base member and a knob row per member (include/lwfb90_setup.h); only the knobs cross the bus
extern "C" int lwfb_setup_members(lwfb_handle* h, int64_t n_members, const int32_t* site_of_member, const double* base_scal,
                                  const double* base_lay, const double* base_u0, const double* knobs, int32_t n_over,
                                  const int32_t* over_idx, int32_t n_profiles, const double* profiles,
                                  const double* psim_base, int32_t ilayer, int32_t* n_invalid) {
  if (!h || n_members < 1 || !base_scal || !base_lay || !base_u0 || !knobs || n_over < 0 || (n_over > 0 && !over_idx) ||
      n_profiles < 0 || (n_profiles > 0 && !profiles) || !psim_base)
    return LWFB_ERR_ARG;
  if (!h->d_forc) { h->err = "lwfb_set_sites must be called first"; return LWFB_ERR_ARG; }
  if (site_of_member)
    for (int64_t i = 0; i < n_members; i++)
      if (site_of_member[i] < 0 || site_of_member[i] >= h->n_sites) { h->err = "site index out of range"; return LWFB_ERR_ARG; }
  for (int j = 0; j < n_over; j++)
    if (over_idx[j] < 0 || over_idx[j] >= LWFB_PS) { h->err = "scalar index out of range"; return LWFB_ERR_ARG; }
  CK(cudaSetDevice(h->o.device));
  const int N = h->o.n_layers, L = h->L;
  int rc = alloc_members(h, n_members);
  if (rc) return rc;
  const int64_t M = h->M;
  if (site_of_member) CK(cudaMemcpy(h->d_site, site_of_member, (size_t)M * sizeof(int32_t), cudaMemcpyHostToDevice));
  else CK(cudaMemset(h->d_site, 0, (size_t)M * sizeof(int32_t)));
  // small inputs in one staging block: base blocks, index list, root profiles, initial potentials; then the knob rows
  const size_t n_small = (size_t)LWFB_PS + (size_t)LWFB_PL * N + L + (size_t)n_profiles * (2 * N + 1) + N;
  const size_t knob_doubles = (size_t)M * (LWFB_K + n_over);
  rc = ensure_stage(h, (n_small + knob_doubles) * sizeof(double) + (size_t)(n_over + 2) * sizeof(int32_t));
  if (rc) return rc;
  double* p = h->d_stage;
  double* d_bs = p; p += LWFB_PS;
  double* d_bl = p; p += (size_t)LWFB_PL * N;
  double* d_bu = p; p += L;
  double* d_pr = p; p += (size_t)n_profiles * (2 * N + 1);
  double* d_ps = p; p += N;
  double* d_kn = p; p += knob_doubles;
  int32_t* d_oi = (int32_t*)p;
  CK(cudaMemcpy(d_bs, base_scal, LWFB_PS * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bl, base_lay, (size_t)LWFB_PL * N * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bu, base_u0, (size_t)L * sizeof(double), cudaMemcpyHostToDevice));
  if (n_profiles > 0) CK(cudaMemcpy(d_pr, profiles, (size_t)n_profiles * (2 * N + 1) * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ps, psim_base, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_kn, knobs, knob_doubles * sizeof(double), cudaMemcpyHostToDevice));
  if (n_over > 0) CK(cudaMemcpy(d_oi, over_idx, (size_t)n_over * sizeof(int32_t), cudaMemcpyHostToDevice));
  setup_members_kernel<<<(unsigned)((M + 127) / 128), 128>>>(M, N, h->o.simulate_isotopes, d_bs, d_bl, d_bu, d_kn, n_over, d_oi,
                                                              n_profiles, d_pr, d_ps, ilayer, h->d_scal, h->d_lay, h->d_u0,
                                                              h->d_status);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (n_invalid) {
    // members without a field-capacity wetness (the reference throws "Computed invalid p_WETF" for those)
    std::vector<int32_t> st((size_t)M);
    CK(cudaMemcpy(st.data(), h->d_status, (size_t)M * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int32_t n = 0;
    for (int64_t i = 0; i < M; i++) n += st[(size_t)i] != LWFB_ST_OK;
    *n_invalid = n;
  }
  { int rc2 = classify_members(h); if (rc2) return rc2; }
  h->ran = false; h->sums_valid = false;
  h->cost_valid = false; h->perm_valid = false;
  return LWFB_OK;
}

// after the member blocks have changed: may the lean lane kernel be used?
static int classify_members(lwfb_handle* h) {
  int* d_used = h->d_nleft;                            // (scratch word)
  CK(cudaMemset(d_used, 0, sizeof(int)));
  optional_paths_kernel<<<(unsigned)((h->M + 255) / 256), 256>>>(h->d_scal, h->M, d_used);
  CK(cudaGetLastError());
  int used = 1;
  CK(cudaMemcpy(&used, d_used, sizeof(int), cudaMemcpyDeviceToHost));
  h->lean = used == 0;
  return LWFB_OK;
}

static DevData make_dev(lwfb_handle* h) {
  DevData d;
  d.M = h->M; d.N = h->o.n_layers; d.iso = h->o.simulate_isotopes; d.t0_day = h->o.t0_day; d.n_days = h->o.n_days;
  d.n_rows = h->n_rows; d.L = h->L;
  d.scal = h->d_scal; d.lay = h->d_lay; d.forc = h->d_forc; d.site = h->d_site; d.u0 = h->d_u0;
  d.perm = (h->placement >= 1 && h->perm_valid) ? h->d_perm : nullptr;
  d.rec = h->d_rec; d.fin = h->d_fin; d.status = h->d_status; d.counters = h->d_counters;
  d.reltol = h->o.reltol; d.abstol = h->o.abstol; d.dt_init = h->o.dt_init; d.dtmax = h->o.dtmax;
  d.maxiters = h->o.maxiters;
  d.iter_budget = 0; d.prio_blocks = 0; d.prio_budget = 0; d.n_slots = 0; d.resume = 0; d.ck = nullptr;
  d.stop = nullptr; d.resident = 0; d.cb_wait = 0;
  d.n_obs_layers = h->n_obs_days > 0 ? h->n_obs_layers : 0;
  d.obs_of_rec = h->d_obs_of_rec; d.obs_layer = h->d_obs_layer; d.obs_theta = h->d_obs_theta; d.obs_d18o = h->d_obs_d18o;
  d.sacc = h->d_sacc;
  d.ntfli0 = h->d_ntfli0; d.ntfli_out = h->d_ntfli_out;
  d.probe_frozen = nullptr; d.probe_ntfli = nullptr; d.probe_t = 0; d.probe_out = nullptr;
  if (const char* e = getenv("LWFB_CB_WAIT")) d.cb_wait = atoi(e);
  return d;
}

static bool use_lanes(const lwfb_handle* h) { return h->o.n_layers >= 3 && h->o.n_layers <= 32 && !h->force_legacy; }
// deep profiles: one warp per member, per-layer arrays in shared memory (lwfb90_wide.cuh)
static bool use_wide(const lwfb_handle* h) { return h->o.n_layers > 32 && !h->force_legacy; }

// Sorting concentrates the expensive members in few blocks: good when the blocks outnumber the resident slots several
// times (the block scheduler balances), bad for a launch of one or two waves (it only makes the slowest block slower).
static bool many_waves(const lwfb_handle* h) {
  if (h->placement == 2) return true;
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->o.device);
  const int64_t members_per_block = (int64_t)(LWFB_LANES_BLOCK / 32) * (32 / h->o.n_layers);
  const int64_t n_blocks = (h->M + members_per_block - 1) / members_per_block;
  return n_blocks >= 3 * (int64_t)n_sm * LWFB_LANES_MINBLOCKS;
}

// placement (d_perm) := the members sorted on the device, see placement_keys; *n_left (optional): unfinished members
static int sort_placement(lwfb_handle* h, bool by_cost, cudaStream_t st, int64_t* n_left) {
  const int64_t M = h->M;
  if (M > 0x7fffffffLL) { h->err = "more than 2^31 members per handle"; return LWFB_ERR_ARG; }
  CK(cudaMemsetAsync(h->d_nleft, 0, sizeof(int), st));
  placement_keys<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(
      h->d_ck, LWFB_CK + LWFB_CKL * h->o.n_layers, h->d_status, h->d_counters, M, (double)h->o.t0_day,
      (double)(h->o.t0_day + h->o.n_days), by_cost ? 1 : 0, h->d_key, h->d_idx, h->d_nleft);
  CK(cudaGetLastError());
  CK(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp, h->sort_tmp_bytes, h->d_key, h->d_key2, h->d_idx, h->d_perm, (int)M, 0, 32, st));
  if (n_left) {
    int n = 0;
    CK(cudaMemcpyAsync(&n, h->d_nleft, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_left = n;
  }
  h->perm_valid = true;
  return LWFB_OK;
}

// placement := members sorted by the number of RHS evaluations counted so far (most expensive first), so that the
// members that share a block (and wait for each other at every step) have similar step counts
static int sort_members_by_cost(lwfb_handle* h, cudaStream_t st) { return sort_placement(h, true, st, nullptr); }

// one launch of the integrator over the slots [0, d.n_slots); latency: the 1-block-per-SM (255 register) build of the lane
// kernel, for launches that do not fill the device
static int launch_integrator(lwfb_handle* h, const DevData& d, cudaStream_t st, bool latency = false) {
  if (use_lanes(h)) {
    // production path: one lane per soil layer, 32/N members per warp; per-thread constants in shared memory
    const int mpw = 32 / d.N;
    const int64_t warps = ((d.n_slots > 0 ? d.n_slots : d.M) + mpw - 1) / mpw;
    const int wpb = LWFB_LANES_BLOCK / 32;
    const unsigned lgrid = (unsigned)((warps + wpb - 1) / wpb);
    const bool lean = h->lean && !h->force_general;
    const bool lean_kernel = lean && (d.N == 5 || d.N == 7);
    size_t smem = LWFB_SMEM_DOUBLES((size_t)wpb * mpw, lean_kernel) * sizeof(double);   // columns + log table + member slots + forcing rows
    if (const char* pad = getenv("LWFB_LANES_SMEM_PAD_KB")) smem += (size_t)atoi(pad) * 1024;   // profiling knob: fewer resident blocks
    // instantiations: general (run-time N, every flow path) and, for ensembles without optional flow paths, lean ones
    // with the layer count fixed at compile time for the layer counts of BASELINE.json's configurations
    // shared-memory carve-out: just enough for the resident blocks the kernel is built for (the rest of the 256 KB is
    // L1, which the register spills live in; left to its own heuristics the driver has been seen to pick a carve-out
    // that holds ONE block, halving the occupancy)
    const size_t per_block = smem + 1024, sm_total = 228 * 1024;
    const size_t want_res = latency ? 1 : LWFB_LANES_MINBLOCKS;
    const size_t n_res = per_block * want_res <= sm_total ? want_res : 1;
    // the carve-out comes in steps (KB); ask for the smallest step that holds the resident blocks, as a percentage that
    // rounds UP to exactly that step (a percentage computed from the bytes can land just above a step: 86 % of 228 KB is
    // 196.08 KB, which selects 228 KB and leaves 28 KB of L1 instead of 60)
    static const int carve_steps_kb[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    int carve_kb = 228;
    for (int k = 0; k < 10; k++) if ((size_t)carve_steps_kb[k] * 1024 >= per_block * n_res) { carve_kb = carve_steps_kb[k]; break; }
    const int carve_pct = carve_kb * 100 / 228;
#define LWFB_LAUNCH1(...)                                                                                       \
    do {                                                                                                       \
      if (smem > 48 * 1024) cudaFuncSetAttribute(integrate_lanes<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      cudaFuncSetAttribute(integrate_lanes<__VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_pct > 100 ? 100 : carve_pct); \
      integrate_lanes<__VA_ARGS__><<<lgrid, LWFB_LANES_BLOCK, smem, st>>>(d);                                    \
    } while (0)
#define LWFB_LAUNCH(NMAX_, EQ_, NX_, LEAN_)                                                                     \
    do {                                                                                                       \
      if (latency) LWFB_LAUNCH1(NMAX_, EQ_, NX_, LEAN_, 1);                                                     \
      else LWFB_LAUNCH1(NMAX_, EQ_, NX_, LEAN_, LWFB_LANES_MINBLOCKS);                                          \
    } while (0)
#ifndef LWFB_FIXED_N
#define LWFB_FIXED_N 1           // experiment knob: 0 = lean instantiations with a run-time layer count
#endif
    if (lean && d.N == 5) LWFB_LAUNCH(8, 2, LWFB_FIXED_N ? 5 : 0, true);
    else if (lean && d.N == 7) LWFB_LAUNCH(8, 2, LWFB_FIXED_N ? 7 : 0, true);
    else if (d.N == 3) LWFB_LAUNCH(8, 3, 0, false);
    else if (d.N < 8) LWFB_LAUNCH(8, 2, 0, false);
    else if (d.N == 8) LWFB_LAUNCH(8, 1, 0, false);
    else LWFB_LAUNCH(32, 1, 0, false);
#undef LWFB_LAUNCH
#undef LWFB_LAUNCH1
  } else if (use_wide(h)) {
    cudaError_t e = launch_wide_integrate(d, h->o.device, st);   // warp-per-member, 32 < N <= LWFB_MAX_LAYERS
    if (e != cudaSuccess) { h->err = std::string("wide kernel: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  } else {
    launch_member_integrate(d, st);   // thread-per-member, runtime layer count (N < 3; testing: N <= LWFB_MEMBER_MAX_LAYERS)
  }
  CK(cudaGetLastError());
  h->n_launches++;
  return LWFB_OK;
}

// single launch over the whole record; with adaptive placement the members are sorted by the cost of the previous run
extern "C" int lwfb_run_async(lwfb_handle* h, void* cuda_stream) {
  if (!h || h->M < 1) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool lanes = use_lanes(h);
  const bool sortable = lanes && h->placement >= 1 && many_waves(h);
  if (lanes && !sortable) h->perm_valid = false;
  if (sortable && h->cost_valid) {
    int rc = sort_members_by_cost(h, st);
    if (rc) return rc;
    h->cost_valid = false;
  }
  DevData d = make_dev(h);
  h->n_launches = 0;
  CK(cudaEventRecord(h->ev0, st));
  if (d.n_obs_layers > 0) CK(cudaMemsetAsync(h->d_sacc, 0, (size_t)h->M * 2 * LWFB_SA * sizeof(double), st));
  h->sums_valid = d.n_obs_layers > 0;
  // an ensemble that does not fill one block per SM is a latency problem: the 255-register build (see integrate_lanes)
  bool latency = false;
  if (lanes && LWFB_LANES_MINBLOCKS > 1) {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->o.device);
    const int64_t members_per_block = (int64_t)(LWFB_LANES_BLOCK / 32) * (32 / h->o.n_layers);
    latency = (h->M + members_per_block - 1) / members_per_block <= (int64_t)n_sm;
    if (const char* e = getenv("LWFB_LATENCY_KERNEL")) latency = latency && atoi(e) != 0;
  }
  int rc = launch_integrator(h, d, st, latency);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, st));
  h->ran = true;
  h->pending = true;
  h->cost_valid = lanes;
  return LWFB_OK;
}

extern "C" int lwfb_sync(lwfb_handle* h) {
  if (!h) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->pending = false;
  return LWFB_OK;
}

// placement := the unfinished members sorted by progress (simulation time reached, least first); returns their number.
// Members of a block then are at the same day (they share the daily callback phases and step at a similar pace), the
// finished ones drop out of the launch, and the front blocks hold the members that decide when the run ends.  (Sorting
// by an estimate of the remaining steps instead was measured slower: it mixes days within a block.)
static int regroup_by_progress(lwfb_handle* h, cudaStream_t st, int64_t* n_left) { return sort_placement(h, false, st, n_left); }

// Synchronous run.  With adaptive placement and enough members for regrouping to pay, the record is integrated in WORK
// QUANTA: a block ends after LWFB_QUANTUM passes of its loop (default 128; 0 = one launch to the end), every member
// checkpoints where it is (fin + ck), the host regroups the unfinished members by progress and launches
// again.  The front blocks of a launch (least progress) keep their slots for the whole launch while the other slots
// turn over once per quantum, so the members that decide the end of the run never wait for a slot.  Nothing has to be
// known about the members in advance; results are identical to a single launch.
extern "C" int lwfb_run(lwfb_handle* h, void* cuda_stream) {
  if (!h || h->M < 1) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  // passes of a block per launch while the device is full.  Round 1 settled on 128; since a callback phase costs a third
  // of what it did (block-level callback), regrouping twice as often pays: 1,022,976 DAV2020 members 20.94 -> 19.90 s,
  // 127872 members 2593 -> 2547 ms, isoBEA2016 unchanged; 48 is slower again (profiles/r02_exp_quantum.log).  A device that
  // is not full keeps the longer quanta (BEA2016 10000: 4.37 s with 128, 4.44 s with 64 there).
  long long quantum = 64;
  if (const char* e = getenv("LWFB_QUANTUM")) quantum = atoll(e);
  int prio_pct = 15;                                  // share of the resident slots given to the front blocks
  if (const char* e = getenv("LWFB_PRIO_PCT")) prio_pct = atoi(e);
  bool stop_flag = true;                              // end a launch when its first slot runs out of work
  if (const char* e = getenv("LWFB_STOP_FLAG")) stop_flag = atoi(e) != 0;
  int min_waves = 1;                                  // launches that do not fill the device once are not cut into quanta
  if (const char* e = getenv("LWFB_MIN_WAVES")) min_waves = atoi(e);
  bool spread = true;                                 // a device that is not full gets its members spread thinly over all slots
  if (const char* e = getenv("LWFB_SPREAD")) spread = atoi(e) != 0;
  int thin_front = 0;                                 // experimental: warps with members per FRONT block of a full device (0: all)
  if (const char* e = getenv("LWFB_THIN_FRONT")) thin_front = atoi(e);
  bool latency_kernel = LWFB_LANES_MINBLOCKS > 1;     // what no longer fills ONE block per SM runs in the 255-register build
  if (const char* e = getenv("LWFB_LATENCY_KERNEL")) latency_kernel = latency_kernel && atoi(e) != 0;
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->o.device);
  const int64_t R_full = (int64_t)n_sm * LWFB_LANES_MINBLOCKS;                  // resident blocks of the throughput build
  const int64_t mpw = 32 / (h->o.n_layers > 0 && h->o.n_layers <= 32 ? h->o.n_layers : 32);   // members per warp
  const int64_t wpb = LWFB_LANES_BLOCK / 32;
  const int64_t mpb = wpb * mpw;                                                 // members per block
  const bool enough = h->placement == 2 || (h->M + mpb - 1) / mpb >= (int64_t)min_waves * R_full;
  const bool quanta = use_lanes(h) && h->placement >= 1 && quantum > 0 && enough;
  if (!quanta) {
    int rc = lwfb_run_async(h, cuda_stream);
    if (rc) return rc;
    return lwfb_sync(h);
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // first launch: sorted by the cost of the previous run of the handle if there was one, else as given
  if (h->cost_valid) {
    int rc = sort_members_by_cost(h, st);
    if (rc) return rc;
  } else h->perm_valid = false;
  h->n_launches = 0;
  CK(cudaEventRecord(h->ev0, st));
  if (h->n_obs_days > 0) CK(cudaMemsetAsync(h->d_sacc, 0, (size_t)h->M * 2 * LWFB_SA * sizeof(double), st));
  h->sums_valid = h->n_obs_days > 0;
  int64_t n_left = h->M;
  const bool cold = !h->cost_valid;
  // diagnostics: LWFB_TRACE=<file> writes one line per launch (members left, blocks, budget, front blocks, slots, ms)
  const char* trace_path = getenv("LWFB_TRACE");
  struct TraceRow { int64_t n_left, nb, budget, prio, slots; cudaEvent_t e0, e1; };
  std::vector<TraceRow> trace;
  for (int launch = 0; n_left > 0; launch++) {
    // a cold handle knows nothing about its members: the first quanta are short (128, 256, ... passes, no front blocks)
    // so that the first regroupings come early
    long long q = quantum;
    if (cold && launch < 16 && (128LL << launch) < quantum) q = 128LL << launch;
    const bool ramp = q < quantum;
    DevData d = make_dev(h);
    d.ck = h->d_ck;
    d.resume = launch > 0;
    d.n_slots = n_left;
    const int64_t nb = (n_left + mpb - 1) / mpb;
    const bool latency = latency_kernel && nb <= (int64_t)n_sm;
    const int64_t R = latency ? (int64_t)n_sm : R_full;                             // slots of this launch
    if (nb <= R) {
      // every block has a slot: longer quanta the emptier the device (regrouping only thins the blocks out)
      int64_t f = R / nb;
      d.iter_budget = (ramp ? q : 2 * q) * (f > 16 ? 16 : f);
      // ... and thinner blocks: the members are dealt to all resident slots, w warps per block instead of all eight,
      // so that a pass runs at the latency of few warps per SM (what is left at the end of a run is its critical path)
      const int64_t w = (n_left + R * mpw - 1) / (R * mpw);
      if (spread && h->perm_valid && w < wpb) {
        const int64_t per_block = w * mpw, nbt = (n_left + per_block - 1) / per_block;
        if (h->perm_thin_len < (size_t)(R * mpb)) {
          cudaFree(h->d_perm_thin); h->d_perm_thin = nullptr; h->perm_thin_len = 0;
          CK(cudaMalloc(&h->d_perm_thin, (size_t)(R * mpb) * sizeof(int32_t)));
          h->perm_thin_len = (size_t)(R * mpb);
        }
        CK(cudaMemsetAsync(h->d_perm_thin, 0xff, (size_t)(nbt * mpb) * sizeof(int32_t), st));
        spread_placement<<<(unsigned)((n_left + 255) / 256), 256, 0, st>>>(h->d_perm, n_left, (int)per_block, (int)mpb, h->d_perm_thin);
        CK(cudaGetLastError());
        d.perm = h->d_perm_thin;
        d.n_slots = nbt * mpb;
      }
    } else if (ramp) {
      d.iter_budget = q;
    } else {
      // K front blocks keep their slots for the whole launch, the other R - K slots take the other blocks, one quantum
      // per block
      int64_t K = R * prio_pct / 100;
      if (K > R - 1) K = R - 1;
      if (K < 0) K = 0;
      const int64_t W = (nb - K + (R - K) - 1) / (R - K);
      if (!stop_flag && K > 0 && W > 1) K += (W * (R - K) - (nb - K)) / (W - 1);   // the last of the W rounds is full
      if (K > R - 1) K = R - 1;
      d.iter_budget = quantum;
      d.prio_blocks = K;
      d.prio_budget = stop_flag ? (1LL << 60) : quantum * ((nb - K + (R - K) - 1) / (R - K));
      if (thin_front > 0 && thin_front < wpb && h->perm_valid && K > 0 && n_left > K * thin_front * mpw) {
        // the front blocks carry thin_front warps of members each: their passes run at the latency of few warps
        const int64_t per_front = thin_front * mpw;
        const size_t need = (size_t)(K * mpb + n_left);
        if (h->perm_thin_len < need) {
          cudaFree(h->d_perm_thin); h->d_perm_thin = nullptr; h->perm_thin_len = 0;
          CK(cudaMalloc(&h->d_perm_thin, (size_t)(R * mpb + h->M) * sizeof(int32_t)));
          h->perm_thin_len = (size_t)(R * mpb + h->M);
        }
        CK(cudaMemsetAsync(h->d_perm_thin, 0xff, (size_t)(K * mpb) * sizeof(int32_t), st));
        spread_front<<<(unsigned)((n_left + 255) / 256), 256, 0, st>>>(h->d_perm, n_left, K, (int)per_front, (int)mpb, h->d_perm_thin);
        CK(cudaGetLastError());
        d.perm = h->d_perm_thin;
        d.n_slots = K * mpb + (n_left - K * per_front);
      }
    }
    if (stop_flag) {
      // the launch ends when the first slot runs out of work (see DevData::stop)
      CK(cudaMemsetAsync(h->d_stop, 0, 2 * sizeof(int), st));
      d.stop = h->d_stop;
      d.resident = R;
    }
    if (trace_path) {
      TraceRow r{n_left, nb, (int64_t)d.iter_budget, (int64_t)d.prio_blocks, (int64_t)d.n_slots, nullptr, nullptr};
      cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
      cudaEventRecord(r.e0, st);
      trace.push_back(r);
    }
    int rc = launch_integrator(h, d, st, latency);
    if (rc) return rc;
    if (trace_path) cudaEventRecord(trace.back().e1, st);
    if ((rc = regroup_by_progress(h, st, &n_left))) return rc;
  }
  CK(cudaEventRecord(h->ev1, st));
  if (trace_path) {
    cudaEventSynchronize(h->ev1);
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "# run: M=%lld N=%d days=%d; launch n_left blocks budget front_blocks slots kernel_ms\n", (long long)h->M, h->o.n_layers, h->o.n_days);
      for (size_t k = 0; k < trace.size(); k++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, trace[k].e0, trace[k].e1);
        fprintf(f, "%zu %lld %lld %lld %lld %lld %.3f\n", k, (long long)trace[k].n_left, (long long)trace[k].nb, (long long)trace[k].budget,
                (long long)trace[k].prio, (long long)trace[k].slots, ms);
      }
      fclose(f);
    }
    for (auto& r : trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  }
  h->perm_valid = false;       // d_perm now holds a partial order of the last launch
  h->ran = true;
  h->cost_valid = true;
  return lwfb_sync(h);
}

extern "C" int lwfb_water_partitioning(lwfb_handle* h, int32_t n_groups, const int32_t* group_of_record, double* out) {
  if (!h || !out || !group_of_record || n_groups < 1 || !h->ran) return LWFB_ERR_ARG;
  if (!h->d_rec) { h->err = "handle was created with record_daily = 0"; return LWFB_ERR_ARG; }
  CK(cudaSetDevice(h->o.device));
  { int rc = ensure_done(h); if (rc) return rc; }
  const int n_rec = h->o.n_days + 1;
  std::vector<int32_t> n_in((size_t)n_groups, 0);
  for (int r = 0; r < n_rec; r++)
    if (group_of_record[r] >= 0 && group_of_record[r] < n_groups) n_in[(size_t)group_of_record[r]]++;
  int32_t* d_grp = nullptr;
  double* d_out = nullptr;
  const size_t out_bytes = (size_t)h->M * n_groups * LWFB_NWP * sizeof(double);
  CK(cudaMalloc(&d_grp, (size_t)(n_rec + n_groups) * sizeof(int32_t)));
  cudaError_t e = cudaMalloc(&d_out, out_bytes);
  if (e != cudaSuccess) { cudaFree(d_grp); h->err = "cannot allocate the aggregate block"; return LWFB_ERR_NOMEM; }
  e = cudaMemcpy(d_grp, group_of_record, (size_t)n_rec * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_grp + n_rec, n_in.data(), (size_t)n_groups * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d_grp); cudaFree(d_out); h->err = std::string("lwfb_water_partitioning: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  const SIdx ix(h->o.n_layers, h->o.simulate_isotopes);
  water_partitioning_kernel<<<(unsigned)((h->M + 3) / 4), 128>>>(h->d_rec, h->M, h->L, ix.ACC, n_rec, d_grp, n_groups, d_out);
  const int64_t n_cells = h->M * (int64_t)n_groups;
  water_partitioning_mean<<<(unsigned)((n_cells + 255) / 256), 256>>>(d_out, h->M, n_groups, d_grp + n_rec);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_grp); cudaFree(d_out);
  if (e != cudaSuccess) { h->err = std::string("lwfb_water_partitioning: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  return LWFB_OK;
}

extern "C" int lwfb_soil_state(lwfb_handle* h, int32_t n_rec_sel, const int32_t* rec_idx, int32_t n_lay_sel,
                               const int32_t* layer_idx, double* out) {
  if (!h || !out || !rec_idx || !layer_idx || n_rec_sel < 1 || n_lay_sel < 1 || !h->ran) return LWFB_ERR_ARG;
  if (!h->d_rec) { h->err = "handle was created with record_daily = 0"; return LWFB_ERR_ARG; }
  for (int k = 0; k < n_rec_sel; k++)
    if (rec_idx[k] < 0 || rec_idx[k] > h->o.n_days) { h->err = "record index out of range"; return LWFB_ERR_ARG; }
  for (int q = 0; q < n_lay_sel; q++)
    if (layer_idx[q] < 0 || layer_idx[q] >= h->o.n_layers) { h->err = "layer index out of range"; return LWFB_ERR_ARG; }
  CK(cudaSetDevice(h->o.device));
  { int rc = ensure_done(h); if (rc) return rc; }
  int32_t* d_idx = nullptr;
  double* d_out = nullptr;
  const size_t out_bytes = (size_t)h->M * n_rec_sel * n_lay_sel * LWFB_NSV * sizeof(double);
  CK(cudaMalloc(&d_idx, (size_t)(n_rec_sel + n_lay_sel) * sizeof(int32_t)));
  cudaError_t e = cudaMalloc(&d_out, out_bytes);
  if (e != cudaSuccess) { cudaFree(d_idx); h->err = "cannot allocate the soil-state block"; return LWFB_ERR_NOMEM; }
  e = cudaMemcpy(d_idx, rec_idx, (size_t)n_rec_sel * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_idx + n_rec_sel, layer_idx, (size_t)n_lay_sel * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d_idx); cudaFree(d_out); h->err = std::string("lwfb_soil_state: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  soil_state_kernel<<<(unsigned)((h->M + 127) / 128), 128>>>(h->d_rec, h->d_lay, h->M, h->o.n_layers, h->L,
                                                             h->o.simulate_isotopes, h->o.n_days + 1, n_rec_sel, d_idx, n_lay_sel,
                                                             d_idx + n_rec_sel, d_out);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_idx); cudaFree(d_out);
  if (e != cudaSuccess) { h->err = std::string("lwfb_soil_state: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  return LWFB_OK;
}

extern "C" int lwfb_last_run_stats(lwfb_handle* h, int32_t* n_launches, float* device_ms) {
  if (!h) return LWFB_ERR_ARG;
  if (n_launches) *n_launches = h->n_launches;
  if (device_ms) *device_ms = h->last_ms;
  return LWFB_OK;
}

// every getter first waits for an asynchronous run (lwfb_run_async) that has not been synchronised yet: the getters copy
// on the legacy default stream, which does not order against a caller's non-blocking stream
static int ensure_done(lwfb_handle* h) {
  if (!h->pending) return LWFB_OK;
  return lwfb_sync(h);
}

// D2H through a pinned bounce buffer in chunks, two chunks in flight (pageable destinations would serialise on the
// driver's own staging); the device record is already in the layout the caller gets
static int download(lwfb_handle* h, void* dst, const void* src, size_t bytes) {
  const size_t chunk = (size_t)32 << 20;
  cudaPointerAttributes pa;
  const bool dst_pinned = cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  cudaGetLastError();                                 // (an unregistered host pointer may leave an error state behind)
  // page-locked destination: one DMA at link speed, nothing to stage
  if (dst_pinned || bytes <= chunk / 4) { CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return LWFB_OK; }
  if (!h->pin[0]) {
    for (int k = 0; k < 2; k++) {
      CK(cudaMallocHost(&h->pin[k], chunk));
      CK(cudaEventCreateWithFlags(&h->pin_ev[k], cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  }
  CK(cudaDeviceSynchronize());
  size_t off[2] = {0, 0}, len[2] = {0, 0};
  int k = 0;
  for (size_t o = 0; o < bytes || len[0] || len[1]; k ^= 1) {
    if (len[k]) {                                     // the chunk that went out two rounds ago has arrived
      CK(cudaEventSynchronize(h->pin_ev[k]));
      memcpy((char*)dst + off[k], h->pin[k], len[k]);
      len[k] = 0;
    }
    if (o < bytes) {
      const size_t n = bytes - o < chunk ? bytes - o : chunk;
      CK(cudaMemcpyAsync(h->pin[k], (const char*)src + o, n, cudaMemcpyDeviceToHost, h->copy_stream));
      CK(cudaEventRecord(h->pin_ev[k], h->copy_stream));
      off[k] = o; len[k] = n; o += n;
    }
  }
  return LWFB_OK;
}

extern "C" int lwfb_get_daily(lwfb_handle* h, int64_t member_first, int64_t member_count, double* out) {
  if (!h || !out || !h->ran) return LWFB_ERR_ARG;
  if (!h->d_rec) { h->err = "handle was created with record_daily = 0"; return LWFB_ERR_ARG; }
  if (member_first < 0 || member_count < 1 || member_first + member_count > h->M) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  int rc = ensure_done(h);
  if (rc) return rc;
  const size_t K = (size_t)(h->o.n_days + 1) * h->L;  // doubles per member: [n_days+1][L], contiguous on the device too
  return download(h, out, h->d_rec + (size_t)member_first * K, (size_t)member_count * K * sizeof(double));
}

extern "C" int lwfb_get_final(lwfb_handle* h, int64_t member_first, int64_t member_count, double* out) {
  if (!h || !out || !h->ran) return LWFB_ERR_ARG;
  if (member_first < 0 || member_count < 1 || member_first + member_count > h->M) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  int rc = ensure_done(h);
  if (rc) return rc;
  return download(h, out, h->d_fin + (size_t)member_first * h->L, (size_t)member_count * h->L * sizeof(double));
}

extern "C" int lwfb_get_status(lwfb_handle* h, int32_t* status, int64_t* counters) {
  if (!h || !h->ran) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  int rc = ensure_done(h);
  if (rc) return rc;
  if (status) CK(cudaMemcpy(status, h->d_status, (size_t)h->M * sizeof(int32_t), cudaMemcpyDeviceToHost));
  static_assert(sizeof(long long) == sizeof(int64_t), "counter width");
  if (counters) CK(cudaMemcpy(counters, h->d_counters, (size_t)h->M * LWFB_NCOUNTERS * sizeof(int64_t), cudaMemcpyDeviceToHost));
  return LWFB_OK;
}

extern "C" int lwfb_set_observations(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                     const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o) {
  if (!h || n_obs_days < 0) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  cudaFree(h->d_obs_of_rec); cudaFree(h->d_obs_layer); cudaFree(h->d_obs_theta); cudaFree(h->d_obs_d18o);
  h->d_obs_of_rec = h->d_obs_layer = nullptr; h->d_obs_theta = h->d_obs_d18o = nullptr;
  h->n_obs_days = h->n_obs_layers = 0; h->sums_valid = false;
  if (n_obs_days == 0) return LWFB_OK;                 // observations cleared
  if (n_obs_layers < 1 || !obs_day || !obs_layer || (!obs_theta && !obs_d18o)) return LWFB_ERR_ARG;
  std::vector<int32_t> of_rec((size_t)h->o.n_days + 1, -1);
  for (int k = 0; k < n_obs_days; k++) {
    if (obs_day[k] < 0 || obs_day[k] > h->o.n_days) { h->err = "obs_day out of range"; return LWFB_ERR_ARG; }
    if (k > 0 && obs_day[k] <= obs_day[k - 1]) { h->err = "obs_day must be strictly increasing"; return LWFB_ERR_ARG; }
    of_rec[(size_t)obs_day[k]] = k;
  }
  for (int l = 0; l < n_obs_layers; l++)
    if (obs_layer[l] < 0 || obs_layer[l] >= h->o.n_layers) { h->err = "obs_layer out of range"; return LWFB_ERR_ARG; }
  const size_t nobs = (size_t)n_obs_days * n_obs_layers;
  CK(cudaMalloc(&h->d_obs_of_rec, of_rec.size() * sizeof(int32_t)));
  CK(cudaMalloc(&h->d_obs_layer, (size_t)n_obs_layers * sizeof(int32_t)));
  CK(cudaMemcpy(h->d_obs_of_rec, of_rec.data(), of_rec.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_obs_layer, obs_layer, (size_t)n_obs_layers * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (obs_theta) {
    CK(cudaMalloc(&h->d_obs_theta, nobs * sizeof(double)));
    CK(cudaMemcpy(h->d_obs_theta, obs_theta, nobs * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (obs_d18o) {
    CK(cudaMalloc(&h->d_obs_d18o, nobs * sizeof(double)));
    CK(cudaMemcpy(h->d_obs_d18o, obs_d18o, nobs * sizeof(double), cudaMemcpyHostToDevice));
  }
  h->n_obs_days = n_obs_days; h->n_obs_layers = n_obs_layers;
  return LWFB_OK;
}

extern "C" int lwfb_scores_device(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                  const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o,
                                  double** d_scores) {
  if (!h || !h->ran || n_obs_days < 0) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  { int rc = ensure_done(h); if (rc) return rc; }
  const unsigned grid = (unsigned)((h->M + 255) / 256);
  if (n_obs_days == 0) {
    // scores of the observations registered BEFORE the run: the integrator accumulated the sums itself
    if (!h->sums_valid) { h->err = "no observations were registered (lwfb_set_observations) before the run"; return LWFB_ERR_ARG; }
    score_finalize_kernel<<<grid, 256>>>(h->d_sacc, h->M, h->d_scores);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (d_scores) *d_scores = h->d_scores;
    return LWFB_OK;
  }
  if (n_obs_layers < 1 || !obs_day || !obs_layer) return LWFB_ERR_ARG;
  if (!h->d_rec) { h->err = "without the daily record (record_daily = 0) register the observations before the run: lwfb_set_observations, then lwfb_scores with n_obs_days = 0"; return LWFB_ERR_ARG; }
  for (int k = 0; k < n_obs_days; k++)
    if (obs_day[k] < 0 || obs_day[k] > h->o.n_days) { h->err = "obs_day out of range"; return LWFB_ERR_ARG; }
  for (int l = 0; l < n_obs_layers; l++)
    if (obs_layer[l] < 0 || obs_layer[l] >= h->o.n_layers) { h->err = "obs_layer out of range"; return LWFB_ERR_ARG; }
  size_t nobs = (size_t)n_obs_days * n_obs_layers;
  size_t bytes = sizeof(int32_t) * (n_obs_days + n_obs_layers) + 2 * nobs * sizeof(double) + 64;
  int rc = ensure_stage(h, bytes);
  if (rc) return rc;
  if (!h->d_sacc_rec) CK(cudaMalloc(&h->d_sacc_rec, (size_t)h->M * 2 * LWFB_SA * sizeof(double)));
  char* p = (char*)h->d_stage;
  double* d_th = (double*)p; p += nobs * sizeof(double);
  double* d_18 = (double*)p; p += nobs * sizeof(double);
  int32_t* d_day = (int32_t*)p; p += n_obs_days * sizeof(int32_t);
  int32_t* d_layer = (int32_t*)p;
  if (obs_theta) CK(cudaMemcpy(d_th, obs_theta, nobs * sizeof(double), cudaMemcpyHostToDevice));
  if (obs_d18o) CK(cudaMemcpy(d_18, obs_d18o, nobs * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_day, obs_day, n_obs_days * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_layer, obs_layer, n_obs_layers * sizeof(int32_t), cudaMemcpyHostToDevice));
  DevData d = make_dev(h);
  score_sums_kernel<<<grid, 256>>>(d, n_obs_days, d_day, n_obs_layers, d_layer, obs_theta ? d_th : nullptr,
                                   obs_d18o ? d_18 : nullptr, h->d_sacc_rec);
  score_finalize_kernel<<<grid, 256>>>(h->d_sacc_rec, h->M, h->d_scores);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (d_scores) *d_scores = h->d_scores;
  return LWFB_OK;
}

extern "C" int lwfb_scores(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                           const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o, double* scores) {
  if (!scores) return LWFB_ERR_ARG;
  double* dptr = nullptr;
  int rc = lwfb_scores_device(h, n_obs_days, obs_day, n_obs_layers, obs_layer, obs_theta, obs_d18o, &dptr);
  if (rc) return rc;
  CK(cudaMemcpy(scores, dptr, (size_t)h->M * LWFB_NSCORES * sizeof(double), cudaMemcpyDeviceToHost));
  return LWFB_OK;
}

// The reference keeps du_NTFLI in `p` across simulate! calls on the same problem, so a second run starts its first ITER
// from the net flows the first run left behind instead of NaN (golden G5 was produced that way,
// test/01-unit-tests.jl:875, :904).  ntfli0 [n_members][N], or NULL for a fresh problem; applies to the following runs
// until the members are set again.
extern "C" int lwfb_set_ntfli0(lwfb_handle* h, const double* ntfli0) {
  if (!h || h->M < 1) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  const size_t bytes = (size_t)h->M * h->o.n_layers * sizeof(double);
  if (!ntfli0) { cudaFree(h->d_ntfli0); h->d_ntfli0 = nullptr; return LWFB_OK; }
  if (!h->d_ntfli0) CK(cudaMalloc(&h->d_ntfli0, bytes));
  CK(cudaMemcpy(h->d_ntfli0, ntfli0, bytes, cudaMemcpyHostToDevice));
  return LWFB_OK;
}

// du_NTFLI as the last run left it: out [n_members][N]
extern "C" int lwfb_get_ntfli(lwfb_handle* h, double* out) {
  if (!h || !out || !h->ran) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  { int rc = ensure_done(h); if (rc) return rc; }
  CK(cudaMemcpy(out, h->d_ntfli_out, (size_t)h->M * h->o.n_layers * sizeof(double), cudaMemcpyDeviceToHost));
  return LWFB_OK;
}

extern "C" int lwfb_eval_rhs(lwfb_handle* h, double t, const double* swati, const double* gwat, const double* frozen,
                             const double* ntfli_prev, double* du) {
  if (!h || h->M < 1 || !swati || !gwat || !frozen || !ntfli_prev || !du) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  const int N = h->o.n_layers;
  const int64_t M = h->M;
  size_t n_sw = (size_t)M * N, n_fr = (size_t)M * (N + 3), n_du = (size_t)M * (N + 9);
  int rc = ensure_stage(h, (2 * n_sw + M + n_fr + n_du) * sizeof(double));
  if (rc) return rc;
  double* p = h->d_stage;
  double* d_sw = p; p += n_sw;
  double* d_gw = p; p += M;
  double* d_fr = p; p += n_fr;
  double* d_nt = p; p += n_sw;
  double* d_du = p;
  CK(cudaMemcpy(d_sw, swati, n_sw * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_gw, gwat, M * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_fr, frozen, n_fr * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_nt, ntfli_prev, n_sw * sizeof(double), cudaMemcpyHostToDevice));
  DevData d = make_dev(h);
  if (use_lanes(h) || use_wide(h)) {
    // the production kernel evaluates f once through its own stage-0 code (DevData::probe_*): state = u0 with SWATI and
    // GWAT replaced by the arguments
    const int L = h->L;
    const SIdx ix(N, h->o.simulate_isotopes);
    double* d_u0p = nullptr;
    CK(cudaMalloc(&d_u0p, (size_t)M * L * sizeof(double)));
    cudaError_t e = cudaMemcpy(d_u0p, h->d_u0, (size_t)M * L * sizeof(double), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy2D(d_u0p + ix.SWATI, (size_t)L * sizeof(double), swati, (size_t)N * sizeof(double), (size_t)N * sizeof(double), (size_t)M, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy2D(d_u0p + ix.GWAT, (size_t)L * sizeof(double), gwat, sizeof(double), sizeof(double), (size_t)M, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d_u0p); h->err = std::string("lwfb_eval_rhs: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
    d.u0 = d_u0p; d.rec = nullptr; d.perm = nullptr; d.n_obs_layers = 0; d.ntfli0 = nullptr; d.ntfli_out = nullptr;
    d.iter_budget = 1;
    d.probe_frozen = d_fr; d.probe_ntfli = d_nt; d.probe_t = t; d.probe_out = d_du;
    int rc2 = launch_integrator(h, d, 0);
    e = cudaDeviceSynchronize();
    cudaFree(d_u0p);
    h->ran = false; h->sums_valid = false; h->cost_valid = false;      // (fin / status / counters were used as scratch)
    if (rc2) return rc2;
    if (e != cudaSuccess) { h->err = std::string("lwfb_eval_rhs: ") + cudaGetErrorString(e); return LWFB_ERR_CUDA; }
  } else {
    launch_member_rhs(d, t, d_sw, d_gw, d_fr, d_nt, d_du);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(du, d_du, n_du * sizeof(double), cudaMemcpyDeviceToHost));
  return LWFB_OK;
}

extern "C" int lwfb_elem_eval(int32_t device, int32_t fn, int64_t n, const double* x, const double* y, double* out) {
  if (n < 1 || !x || !y || !out) return LWFB_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return LWFB_ERR_CUDA;
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  size_t b = (size_t)n * sizeof(double);
  if (cudaMalloc(&dx, b) != cudaSuccess || cudaMalloc(&dy, b) != cudaSuccess || cudaMalloc(&dout, b) != cudaSuccess) {
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return LWFB_ERR_NOMEM;
  }
  cudaMemcpy(dx, x, b, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, y, b, cudaMemcpyHostToDevice);
  elem_kernel<<<(unsigned)((n + 255) / 256), 256>>>(fn, n, dx, dy, dout);
  cudaError_t e = cudaMemcpy(out, dout, b, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  return e == cudaSuccess ? LWFB_OK : LWFB_ERR_CUDA;
}

// debug/testing: raw device-layout copies of the member blocks ([field][member])
extern "C" int lwfb_debug_get_block(lwfb_handle* h, int32_t which, double* out) {
  if (!h || !out || h->M < 1) return LWFB_ERR_ARG;
  CK(cudaSetDevice(h->o.device));
  const double* src = which == 0 ? h->d_scal : (which == 1 ? h->d_lay : h->d_u0);
  size_t n = (size_t)h->M * (which == 0 ? LWFB_PS : (which == 1 ? LWFB_PL * h->o.n_layers : h->L));
  CK(cudaMemcpy(out, src, n * sizeof(double), cudaMemcpyDeviceToHost));
  return LWFB_OK;
}

// measured FP64 (DFMA) peak of the device in TFLOP/s (2 flop per DFMA); best of `reps` launches
extern "C" int lwfb_fp64_peak(int32_t device, int32_t reps, double* tflops) {
  if (!tflops) return LWFB_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return LWFB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LWFB_ERR_CUDA;
  double* d_out = nullptr;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  if (cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return LWFB_ERR_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 16.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? LWFB_OK : LWFB_ERR_CUDA;
}

// testing aid: select the thread-per-member kernel (the second, independent CUDA implementation)
extern "C" int lwfb_debug_force_legacy(lwfb_handle* h, int32_t on) {
  if (!h) return LWFB_ERR_ARG;
  if (on == 1 && h->o.n_layers > LWFB_MEMBER_MAX_LAYERS) { h->err = "the thread-per-member kernel holds at most LWFB_MEMBER_MAX_LAYERS layers"; return LWFB_ERR_UNSUPPORTED; }
  h->force_legacy = on == 1;        // 1: thread-per-member kernel
  h->force_general = on == 2;       // 2: lane kernel, but never its lean / fixed-N instantiations
  return LWFB_OK;
}

// placement of the members in the lane kernel's slots: 0 = as given, 1 = adaptive (default)
extern "C" int lwfb_set_placement(lwfb_handle* h, int32_t mode) {
  if (!h || mode < 0 || mode > 2) return LWFB_ERR_ARG;
  h->placement = mode;
  h->perm_valid = false; h->cost_valid = false;      // the next run starts cold
  return LWFB_OK;
}
