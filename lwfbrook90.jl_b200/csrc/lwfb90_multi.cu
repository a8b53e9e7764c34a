// lwfb90_multi.cu — several GPUs of one host behind one handle (8b: "multi-GPU inside one call").
//
// The ensemble shards naturally: members are independent initial-value problems.  lwfb_multi owns one lwfb_handle per
// device, deals the members to the devices - contiguous blocks, or, when the caller supplies a cost estimate per member
// (e.g. the RHS-evaluation counts of a short probe run), sorted by cost and dealt in a snake so that every device gets
// the same share of expensive members - runs the shards concurrently (one host thread per device; there is no
// collective during the integration), and all-gathers the per-member score blocks with peer-to-peer copies over
// NVLink (cudaMemcpyPeerAsync; no NCCL, no torch), so that every device holds the scores of the whole ensemble for the
// parameter-estimation update.  Results come back in the caller's member order.
//
// The same device may be listed more than once (several shards on one GPU): that is how the single-GPU test suite
// exercises the sharding, the permutation and the gather.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lwfb90.h"
#include "../../include/lwfb90_setup.h"

struct lwfb_multi {
  lwfb_opts o;
  int L = 0;
  std::vector<int> dev;                  // device of shard s
  std::vector<lwfb_handle*> h;           // handle of shard s
  int64_t M = 0;
  std::vector<int64_t> first;            // shard s holds the members order[first[s] .. first[s+1])
  std::vector<int64_t> order;            // position in shard order -> caller's member index
  std::vector<double*> d_all;            // per shard: [M][LWFB_NSCORES] scores of the whole ensemble (caller's order)
  std::string err;
};

#define MCK(call)                                                                 \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      mh->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
      return LWFB_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

static int fail(lwfb_multi* mh, int s, int rc) {
  mh->err = "shard " + std::to_string(s) + " (device " + std::to_string(mh->dev[(size_t)s]) + "): " + lwfb_last_error(mh->h[(size_t)s]);
  return rc;
}

// run fn(s) for every shard concurrently; first non-zero return code wins
template <typename F>
static int for_shards(lwfb_multi* mh, F fn) {
  const int S = (int)mh->h.size();
  std::vector<int> rc((size_t)S, 0);
  std::vector<std::thread> th;
  for (int s = 0; s < S; s++) th.emplace_back([&, s]() { rc[(size_t)s] = fn(s); });
  for (auto& t : th) t.join();
  for (int s = 0; s < S; s++)
    if (rc[(size_t)s]) return fail(mh, s, rc[(size_t)s]);
  return LWFB_OK;
}

extern "C" int lwfb_multi_create(lwfb_multi** out, const lwfb_opts* opts, int32_t n_shards, const int32_t* devices) {
  if (!out || !opts || n_shards < 1 || !devices) return LWFB_ERR_ARG;
  *out = nullptr;
  lwfb_multi* mh = new lwfb_multi();
  mh->o = *opts;
  mh->L = LWFB_STATE_LEN(opts->n_layers, opts->simulate_isotopes);
  for (int s = 0; s < n_shards; s++) {
    lwfb_opts o = *opts;
    o.device = devices[s];
    lwfb_handle* h = nullptr;
    int rc = lwfb_create(&h, &o);
    if (rc) {
      for (auto* p : mh->h) lwfb_destroy(p);
      delete mh;
      return rc;
    }
    mh->h.push_back(h);
    mh->dev.push_back(devices[s]);
  }
  // peer access between the distinct devices (NVLink / NVSwitch: every pair)
  for (int a : mh->dev)
    for (int b : mh->dev)
      if (a != b) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a, b);
        if (can) {
          cudaSetDevice(a);
          cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
      }
  mh->d_all.assign((size_t)n_shards, nullptr);
  *out = mh;
  return LWFB_OK;
}

extern "C" void lwfb_multi_destroy(lwfb_multi* mh) {
  if (!mh) return;
  for (size_t s = 0; s < mh->h.size(); s++) {
    cudaSetDevice(mh->dev[s]);
    cudaFree(mh->d_all[s]);
    lwfb_destroy(mh->h[s]);
  }
  delete mh;
}

extern "C" const char* lwfb_multi_last_error(const lwfb_multi* mh) { return mh ? mh->err.c_str() : "null handle"; }
extern "C" int32_t lwfb_multi_n_shards(const lwfb_multi* mh) { return mh ? (int32_t)mh->h.size() : 0; }
extern "C" lwfb_handle* lwfb_multi_shard(lwfb_multi* mh, int32_t s) {
  return (mh && s >= 0 && s < (int32_t)mh->h.size()) ? mh->h[(size_t)s] : nullptr;
}

extern "C" int lwfb_multi_set_sites(lwfb_multi* mh, int32_t n_sites, int32_t n_rows, const double* forcing) {
  if (!mh) return LWFB_ERR_ARG;
  return for_shards(mh, [&](int s) { return lwfb_set_sites(mh->h[(size_t)s], n_sites, n_rows, forcing); });
}

// members -> shards.  cost == NULL: contiguous blocks (sizes differ by at most one).  Otherwise: sorted by cost, most
// expensive first, and dealt in a snake (0 1 2 .. S-1 S-1 .. 1 0 0 1 ..): every shard gets the same number of members
// (+-1) and nearly the same total cost whatever the tail of the cost distribution looks like.
static void deal(lwfb_multi* mh, int64_t M, const double* cost) {
  const int S = (int)mh->h.size();
  mh->M = M;
  mh->first.assign((size_t)S + 1, 0);
  mh->order.resize((size_t)M);
  if (!cost) {
    const int64_t base = M / S, rem = M % S;
    for (int s = 0; s < S; s++) mh->first[(size_t)s + 1] = mh->first[(size_t)s] + base + (s < rem ? 1 : 0);
    std::iota(mh->order.begin(), mh->order.end(), (int64_t)0);
    return;
  }
  std::vector<int64_t> idx((size_t)M);
  std::iota(idx.begin(), idx.end(), (int64_t)0);
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
  std::vector<std::vector<int64_t>> sh((size_t)S);
  for (int64_t k = 0; k < M; k++) {
    const int64_t round = k / S, pos = k % S;
    const int s = (int)((round & 1) ? S - 1 - pos : pos);
    sh[(size_t)s].push_back(idx[(size_t)k]);
  }
  int64_t o = 0;
  for (int s = 0; s < S; s++) {
    std::sort(sh[(size_t)s].begin(), sh[(size_t)s].end());          // (caller's order within a shard)
    for (int64_t g : sh[(size_t)s]) mh->order[(size_t)o++] = g;
    mh->first[(size_t)s + 1] = o;
  }
}

template <typename T>
static std::vector<T> gather_rows(const lwfb_multi* mh, int s, const T* src, size_t row) {
  const int64_t a = mh->first[(size_t)s], b = mh->first[(size_t)s + 1];
  std::vector<T> out((size_t)(b - a) * row);
  for (int64_t k = a; k < b; k++) memcpy(out.data() + (size_t)(k - a) * row, src + (size_t)mh->order[(size_t)k] * row, row * sizeof(T));
  return out;
}

extern "C" int lwfb_multi_set_members(lwfb_multi* mh, int64_t n_members, const int32_t* site_of_member, const double* scal,
                                      const double* lay, const double* u0, const double* cost) {
  if (!mh || n_members < (int64_t)mh->h.size() || !site_of_member || !scal || !lay || !u0) return LWFB_ERR_ARG;
  deal(mh, n_members, cost);
  const int N = mh->o.n_layers;
  return for_shards(mh, [&](int s) {
    auto si = gather_rows(mh, s, site_of_member, 1);
    auto sc = gather_rows(mh, s, scal, LWFB_PS);
    auto la = gather_rows(mh, s, lay, (size_t)LWFB_PL * N);
    auto uu = gather_rows(mh, s, u0, (size_t)mh->L);
    return lwfb_set_members(mh->h[(size_t)s], (int64_t)si.size(), si.data(), sc.data(), la.data(), uu.data());
  });
}

extern "C" int lwfb_multi_setup_members(lwfb_multi* mh, int64_t n_members, const double* base_scal, const double* base_lay,
                                        const double* base_u0, const double* knobs, int32_t n_over, const int32_t* over_idx,
                                        int32_t n_profiles, const double* profiles, const double* psim_base, int32_t ilayer,
                                        const double* cost) {
  if (!mh || n_members < (int64_t)mh->h.size() || !knobs) return LWFB_ERR_ARG;
  deal(mh, n_members, cost);
  return for_shards(mh, [&](int s) {
    auto kn = gather_rows(mh, s, knobs, (size_t)(LWFB_K + n_over));
    return lwfb_setup_members(mh->h[(size_t)s], mh->first[(size_t)s + 1] - mh->first[(size_t)s], nullptr, base_scal, base_lay,
                              base_u0, kn.data(), n_over, over_idx, n_profiles, profiles, psim_base, ilayer, nullptr);
  });
}

extern "C" int lwfb_multi_set_observations(lwfb_multi* mh, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                           const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o) {
  if (!mh) return LWFB_ERR_ARG;
  return for_shards(mh, [&](int s) {
    return lwfb_set_observations(mh->h[(size_t)s], n_obs_days, obs_day, n_obs_layers, obs_layer, obs_theta, obs_d18o);
  });
}

// every shard integrates its members on its device, all at the same time
extern "C" int lwfb_multi_run(lwfb_multi* mh) {
  if (!mh || mh->M < 1) return LWFB_ERR_ARG;
  return for_shards(mh, [&](int s) { return lwfb_run(mh->h[(size_t)s], nullptr); });
}

// Scores of the whole ensemble on every device: each shard finalises its block (lwfb_scores_device: from the sums of
// the run with n_obs_days = 0, or from its daily record), the blocks are exchanged peer to peer and laid out in the
// caller's member order.  d_scores_of_shard [n_shards] (optional) receives the device pointers of the full
// [n_members][LWFB_NSCORES] blocks, scores (optional, host) one copy of it.
__global__ void scatter_scores(const double* __restrict__ src, const int64_t* __restrict__ order, int64_t n, double* __restrict__ dst) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * LWFB_NSCORES) return;
  dst[order[k / LWFB_NSCORES] * LWFB_NSCORES + k % LWFB_NSCORES] = src[k];
}

extern "C" int lwfb_multi_allgather_scores(lwfb_multi* mh, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                           const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o,
                                           double** d_scores_of_shard, double* scores) {
  if (!mh || mh->M < 1) return LWFB_ERR_ARG;
  const int S = (int)mh->h.size();
  std::vector<double*> d_local((size_t)S, nullptr);
  int rc = for_shards(mh, [&](int s) {
    return lwfb_scores_device(mh->h[(size_t)s], n_obs_days, obs_day, n_obs_layers, obs_layer, obs_theta, obs_d18o, &d_local[(size_t)s]);
  });
  if (rc) return rc;
  const size_t all_bytes = (size_t)mh->M * LWFB_NSCORES * sizeof(double);
  for (int s = 0; s < S; s++) {
    MCK(cudaSetDevice(mh->dev[(size_t)s]));
    if (!mh->d_all[(size_t)s]) MCK(cudaMalloc(&mh->d_all[(size_t)s], 2 * all_bytes + (size_t)mh->M * sizeof(int64_t)));
  }
  // every device receives every block (shard order) into the second half of its buffer, then un-permutes into the first
  for (int s = 0; s < S; s++) {
    MCK(cudaSetDevice(mh->dev[(size_t)s]));
    double* stage = mh->d_all[(size_t)s] + (size_t)mh->M * LWFB_NSCORES;
    for (int p = 0; p < S; p++) {
      const size_t off = (size_t)mh->first[(size_t)p] * LWFB_NSCORES;
      const size_t bytes = (size_t)(mh->first[(size_t)p + 1] - mh->first[(size_t)p]) * LWFB_NSCORES * sizeof(double);
      MCK(cudaMemcpyPeerAsync(stage + off, mh->dev[(size_t)s], d_local[(size_t)p], mh->dev[(size_t)p], bytes, 0));
    }
    int64_t* d_order = (int64_t*)(mh->d_all[(size_t)s] + 2 * (size_t)mh->M * LWFB_NSCORES);
    MCK(cudaMemcpyAsync(d_order, mh->order.data(), (size_t)mh->M * sizeof(int64_t), cudaMemcpyHostToDevice, 0));
    const int64_t n = mh->M * LWFB_NSCORES;
    scatter_scores<<<(unsigned)((n + 255) / 256), 256>>>(stage, d_order, mh->M, mh->d_all[(size_t)s]);
    MCK(cudaGetLastError());
  }
  for (int s = 0; s < S; s++) {
    MCK(cudaSetDevice(mh->dev[(size_t)s]));
    MCK(cudaDeviceSynchronize());
    if (d_scores_of_shard) d_scores_of_shard[s] = mh->d_all[(size_t)s];
  }
  if (scores) {
    MCK(cudaSetDevice(mh->dev[0]));
    MCK(cudaMemcpy(scores, mh->d_all[0], all_bytes, cudaMemcpyDeviceToHost));
  }
  return LWFB_OK;
}

// status [n_members], counters [n_members][LWFB_NCOUNTERS] in the caller's member order; shard_of_member (optional)
extern "C" int lwfb_multi_get_status(lwfb_multi* mh, int32_t* status, int64_t* counters, int32_t* shard_of_member) {
  if (!mh || mh->M < 1) return LWFB_ERR_ARG;
  const int S = (int)mh->h.size();
  for (int s = 0; s < S; s++) {
    const int64_t a = mh->first[(size_t)s], n = mh->first[(size_t)s + 1] - a;
    std::vector<int32_t> st((size_t)n);
    std::vector<int64_t> cn((size_t)n * LWFB_NCOUNTERS);
    int rc = lwfb_get_status(mh->h[(size_t)s], st.data(), cn.data());
    if (rc) return fail(mh, s, rc);
    for (int64_t k = 0; k < n; k++) {
      const int64_t g = mh->order[(size_t)(a + k)];
      if (status) status[g] = st[(size_t)k];
      if (counters) memcpy(counters + g * LWFB_NCOUNTERS, cn.data() + (size_t)k * LWFB_NCOUNTERS, LWFB_NCOUNTERS * sizeof(int64_t));
      if (shard_of_member) shard_of_member[g] = s;
    }
  }
  return LWFB_OK;
}

// final states [n_members][LWFB_STATE_LEN] in the caller's member order
extern "C" int lwfb_multi_get_final(lwfb_multi* mh, double* out) {
  if (!mh || mh->M < 1 || !out) return LWFB_ERR_ARG;
  const int S = (int)mh->h.size();
  for (int s = 0; s < S; s++) {
    const int64_t a = mh->first[(size_t)s], n = mh->first[(size_t)s + 1] - a;
    std::vector<double> buf((size_t)n * mh->L);
    int rc = lwfb_get_final(mh->h[(size_t)s], 0, n, buf.data());
    if (rc) return fail(mh, s, rc);
    for (int64_t k = 0; k < n; k++)
      memcpy(out + (size_t)mh->order[(size_t)(a + k)] * mh->L, buf.data() + (size_t)k * mh->L, (size_t)mh->L * sizeof(double));
  }
  return LWFB_OK;
}

// device time of the last lwfb_multi_run per shard (ms) and the slowest one
extern "C" int lwfb_multi_last_run_ms(lwfb_multi* mh, float* ms_of_shard, float* ms_max) {
  if (!mh) return LWFB_ERR_ARG;
  float mx = 0;
  for (size_t s = 0; s < mh->h.size(); s++) {
    float ms = 0;
    int32_t n = 0;
    lwfb_last_run_stats(mh->h[s], &n, &ms);
    if (ms_of_shard) ms_of_shard[s] = ms;
    mx = ms > mx ? ms : mx;
  }
  if (ms_max) *ms_max = mx;
  return LWFB_OK;
}
