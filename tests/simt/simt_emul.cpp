// simt_emul.cpp — lock-step warp emulator for the liblwfb90 kernels.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernel headers of lwfbrook90.jl_b200/csrc with a host compiler (-DLWFB_SIMT_EMUL, see
// csrc/lwfb90_simt.h) and runs them with the semantics the hardware gives warp-synchronous, block-synchronous code:
// the threads of a thread block (1..8 warps of 32 lanes) are coroutines (ucontext) on one host thread that advance
// round-robin from one primitive to the next; a warp primitive returns when all lanes of the warp have reached it, a
// block barrier (__syncthreads, __syncthreads_or) when all threads of the block have.  Every primitive carries its
// source line; lanes meeting at different lines, a thread leaving before a barrier, or warps making different numbers
// of barrier calls abort the run (undefined behaviour or a hang on the GPU).
//
// Purpose: the container that builds this repository has no GPU.  The emulator lets the non-GPU test
// suite check the kernels' arithmetic, shuffle choreography and control flow bit-for-bit against the
// CPU oracle (tests/test_simt_emulation.py).  It is never linked into liblwfb90.so and is not a
// fallback: the product path needs the CUDA library and a GPU.
#define LWFB_SIMT_EMUL 1
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "../../lwfbrook90.jl_b200/csrc/lwfb90_simt.h"
#ifdef LWFB_EMUL_WIDE
// second build of this file (libsimt_emul_wide.so): the warp-per-member kernel for N > 32, whose translation unit is
// compiled with the workspace form of the daily callback (as lwfb90_wide.cu is); kernel id 2, nothing else
#define LWFB_CB_SCRATCH 1
#include "../../lwfbrook90.jl_b200/csrc/lwfb90_device.cuh"
#include "../../lwfbrook90.jl_b200/csrc/lwfb90_wide.cuh"
#define LWFB_LANES_BLOCK 256
namespace lwfb { enum { KC_COUNT = 1, KM_COUNT = 1 }; }
#else
#include "../../lwfbrook90.jl_b200/csrc/lwfb90_device.cuh"
#include "../../lwfbrook90.jl_b200/csrc/lwfb90_lanes.cuh"
#include "../../lwfbrook90.jl_b200/csrc/lwfb90_member.cuh"
#endif

namespace {
constexpr int WARP = 32;
constexpr int MAX_WARPS = LWFB_LANES_BLOCK / 32;     // warps of an emulated thread block (run-time choice 1..MAX_WARPS)
constexpr int MAX_LANES = WARP * MAX_WARPS;
constexpr size_t STACK_BYTES = 1u << 20;

// One emulated thread block: n_lanes = 32 * warps coroutines on one host thread.  A lane runs until it reaches a
// primitive, deposits its operand and WAITS: it hands control to the next live lane of the block (round robin) and, each
// time control comes back, looks whether its primitive is complete - a warp primitive when all lanes of the warp have
// deposited, a block barrier (__syncthreads*) when all threads of the block have, a counted named barrier
// (bar.sync id, n) when n threads have arrived.  Warps may therefore be in different code at the same time (the lane
// kernel runs daily callbacks in some warps while the others step), as on the hardware.  Every primitive carries its
// source line; lanes meeting at different lines, a thread leaving before a barrier, warps making different numbers of
// barrier calls or a block in which nobody can make progress abort the run (undefined behaviour or a hang on the GPU).
struct Block {
  ucontext_t main_ctx, ctx[MAX_LANES];
  char* stacks = nullptr;
  int n_lanes = WARP;
  bool done[MAX_LANES];
  uint64_t slot[2][MAX_LANES];          // warp primitives, double-buffered by the parity of the lane's primitive count
  int opid[2][MAX_LANES];
  unsigned long counter[MAX_LANES];
  uint64_t bslot[2][MAX_LANES];         // block primitives
  int bopid[2][MAX_LANES];
  unsigned long bcounter[MAX_LANES];
  // counted named barrier (one id is enough for the kernels): generation, arrivals, what the first arrival asked for
  unsigned long named_gen = 0;
  int named_arrived = 0, named_expect = 0, named_op = 0;
  unsigned long progress = 0;           // primitives completed / lanes finished (deadlock detection)
  int cur = 0;
  void (*body)(void* arg, int lane) = nullptr;
  void* arg = nullptr;
  long n_ops = 0;
};
thread_local Block* W = nullptr;
thread_local int g_warps_per_block = 1;

int next_alive_in_block(int l) {
  for (int s = 1; s <= W->n_lanes; s++) {
    int n = (l + s) % W->n_lanes;
    if (!W->done[n]) return n;
  }
  return -1;
}

void trampoline(int lane) {
  W->body(W->arg, lane);
  W->done[lane] = true;
  W->progress++;
  int n = next_alive_in_block(lane);
  if (n < 0) setcontext(&W->main_ctx);
  W->cur = n;
  setcontext(&W->ctx[n]);
}

// hand control to the next live lane of the block; returns when it is this lane's turn again
struct Waiter {
  int lane, op_id;
  const char* what;
  unsigned long seen;
  int idle_rounds = 0;
  Waiter(int l, int op, const char* w) : lane(l), op_id(op), what(w), seen(W->progress) {}
  void yield() {
    if (W->progress != seen) { seen = W->progress; idle_rounds = 0; }
    else if (++idle_rounds > 4) {
      fprintf(stderr, "simt_emul: deadlock: lane %d waits at the %s of line %d and no thread of the block can make progress\n",
              lane, what, op_id);
      abort();
    }
    const int n = next_alive_in_block(lane);
    if (n < 0 || n == lane) { if (idle_rounds > 1) { fprintf(stderr, "simt_emul: lane %d waits alone at line %d\n", lane, op_id); abort(); } return; }
    W->cur = n;
    swapcontext(&W->ctx[lane], &W->ctx[n]);
  }
};

// warp scope: deposit a value, wait until every lane of the warp has deposited at the same primitive, return the slot parity
int rendezvous(uint64_t v, int op_id) {
  const int l = W->cur;
  const int base = l & ~(WARP - 1);
  const unsigned long my_op = W->counter[l]++;
  const int par = (int)(my_op & 1ul);
  W->slot[par][l] = v;
  W->opid[par][l] = op_id;
  W->progress++;
  Waiter w(l, op_id, "warp primitive");
  for (;;) {
    bool all = true;
    for (int i = base; i < base + WARP; i++) {
      if (W->done[i] && W->counter[i] <= my_op) { fprintf(stderr, "simt_emul: lane %d exited before the warp primitive at line %d\n", i, op_id); abort(); }
      if (W->counter[i] <= my_op) all = false;
    }
    if (all) break;
    w.yield();
  }
  for (int i = base; i < base + WARP; i++)
    if (W->opid[par][i] != op_id) {
      fprintf(stderr, "simt_emul: divergent warp primitive: lane %d at line %d, lane %d at line %d\n", l, op_id, i, W->opid[par][i]);
      abort();
    }
  W->progress++;
  if (l == 0) W->n_ops++;
  return par;
}

// block scope (barrier 0: every thread of the block)
int block_rendezvous(uint64_t v, int op_id) {
  const int l = W->cur;
  const unsigned long my_op = W->bcounter[l]++;
  const int par = (int)(my_op & 1ul);
  W->bslot[par][l] = v;
  W->bopid[par][l] = op_id;
  W->progress++;
  Waiter w(l, op_id, "block barrier");
  for (;;) {
    bool all = true;
    for (int i = 0; i < W->n_lanes; i++) {
      if (W->done[i] && W->bcounter[i] <= my_op) { fprintf(stderr, "simt_emul: lane %d exited before the block barrier at line %d\n", i, op_id); abort(); }
      if (W->bcounter[i] <= my_op) all = false;
    }
    if (all) break;
    w.yield();
  }
  for (int i = 0; i < W->n_lanes; i++)
    if (W->bopid[par][i] != op_id) {
      fprintf(stderr, "simt_emul: block barrier mismatch: lane %d at line %d, lane %d at line %d\n", l, op_id, i, W->bopid[par][i]);
      abort();
    }
  W->progress++;
  return par;
}

// counted named barrier (bar.sync 1, n_threads): released when n_threads threads have arrived
void named_rendezvous(int n_threads, int op_id) {
  const int l = W->cur;
  if (n_threads <= 0 || n_threads > W->n_lanes || (n_threads & (WARP - 1))) {
    fprintf(stderr, "simt_emul: bar.sync with thread count %d at line %d\n", n_threads, op_id); abort();
  }
  const unsigned long g = W->named_gen;
  if (W->named_arrived == 0) { W->named_expect = n_threads; W->named_op = op_id; }
  else if (W->named_expect != n_threads || W->named_op != op_id) {
    fprintf(stderr, "simt_emul: named barrier mismatch: lane %d arrives with (line %d, %d threads), the barrier is open for (line %d, %d threads)\n",
            l, op_id, n_threads, W->named_op, W->named_expect);
    abort();
  }
  W->progress++;
  if (++W->named_arrived == n_threads) { W->named_arrived = 0; W->named_gen++; return; }
  Waiter w(l, op_id, "named barrier");
  while (W->named_gen == g) w.yield();
}

void run_block(void (*body)(void*, int), void* arg, int warps) {
  Block* w = new Block;
  w->n_lanes = WARP * warps;
  w->stacks = (char*)malloc(STACK_BYTES * (size_t)w->n_lanes);
  w->body = body; w->arg = arg;
  W = w;
  g_warps_per_block = warps;
  for (int l = 0; l < w->n_lanes; l++) {
    w->done[l] = false; w->counter[l] = 0; w->bcounter[l] = 0;
    w->opid[0][l] = w->opid[1][l] = -1;
    w->bopid[0][l] = w->bopid[1][l] = -1;
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = w->stacks + (size_t)l * STACK_BYTES;
    w->ctx[l].uc_stack.ss_size = STACK_BYTES;
    w->ctx[l].uc_link = nullptr;
    makecontext(&w->ctx[l], (void (*)())trampoline, 1, l);
  }
  w->cur = 0;
  swapcontext(&w->main_ctx, &w->ctx[0]);
  free(w->stacks);
  delete w;
  W = nullptr;
}
}  // namespace

namespace simt {
uint64_t exchange(uint64_t v, int src_lane, int op_id) {
  const int l = W->cur;
  const int base = l & ~(WARP - 1);
  const int par = rendezvous(v, op_id);
  return (src_lane >= 0 && src_lane < WARP) ? W->slot[par][base + src_lane] : W->slot[par][l];
}
unsigned vote(bool p, int op_id) {
  const int base = W->cur & ~(WARP - 1);
  const int par = rendezvous(p ? 1u : 0u, op_id);
  unsigned m = 0;
  for (int i = 0; i < WARP; i++) m |= (W->slot[par][base + i] ? 1u : 0u) << i;
  return m;
}
std::atomic<long> g_stage_syncs{0};      // stage barriers executed by thread 0 of every block (profiling: passes of the blocks)
std::atomic<long> g_votes{0};            // block votes executed by thread 0 of every block (3 per callback pass, 4 per step pass)
bool block_or(bool p, int op_id) {
  if (W->cur == 0) (op_id == 1000000 ? g_stage_syncs : g_votes).fetch_add(1, std::memory_order_relaxed);
  const int par = block_rendezvous(p ? 1u : 0u, op_id);
  bool any = false;
  for (int i = 0; i < W->n_lanes; i++) any = any || W->bslot[par][i] != 0;
  return any;
}
int block_count(bool p, int op_id) {
  if (W->cur == 0) g_votes.fetch_add(1, std::memory_order_relaxed);
  const int par = block_rendezvous(p ? 1u : 0u, op_id);
  int n = 0;
  for (int i = 0; i < W->n_lanes; i++) n += W->bslot[par][i] != 0;
  return n;
}
void named_sync(int n_threads, int op_id) {
  if ((W->cur & (WARP - 1)) == 0) g_stage_syncs.fetch_add(1, std::memory_order_relaxed);
  named_rendezvous(n_threads, op_id);
}
int lane_id() { return W->cur & (WARP - 1); }
int warps_per_block() { return g_warps_per_block; }
}  // namespace simt

// ----------------------------------------------------------------------------------------------
// harness: same argument meaning as lwfb_oracle_run (member-major blocks in, member-major record out)
// ----------------------------------------------------------------------------------------------
namespace {
struct Job {
  lwfb::DevData d;
  int kernel;                 // 0 lanes, 1 thread-per-member
  std::atomic<int64_t> next{0};
  int64_t n_units = 0;        // thread blocks (lanes kernel) or members
  int warps = 1;              // warps per emulated thread block (lanes kernel)
  std::atomic<long> ops{0};
  // stop flag of the work-quantum execution, as the __global__ wrapper of integrate_lanes drives it: the first block that
  // uses up its budget when no block is left to take its slot ends the launch for every block
  bool use_stop = false;
  int stop[2] = {0, 0};
  std::atomic<int64_t> exited{0};
  int resident = 1;           // blocks "resident" at a time = worker threads
  bool lean = false;          // no member uses an optional flow path: the lean instantiations apply (as in the library)
};

struct WarpArg { const lwfb::DevData* d; int64_t block; int warps; double* cst; bool by_budget; bool lean; double* msm; };

#ifdef LWFB_EMUL_WIDE
void lanes_entry(void* a, int tid) {             // one warp = one member; cst: the member's shared-memory block
  WarpArg* wa = (WarpArg*)a;
  lwfb::wide_body(*wa->d, tid & 31, wa->block, wa->cst);
}
#else
void lanes_entry(void* a, int tid) {             // tid = thread index within the emulated block
  WarpArg* wa = (WarpArg*)a;
  const lwfb::DevData& d = *wa->d;
  double* cst = wa->cst + tid;                   // this thread's column of the "shared memory" constant block
  const int lane = tid & 31;
  const int64_t warp_g = wa->block * wa->warps + (tid >> 5);
  bool b;
  if (wa->lean && d.N == 5) b = lwfb::lanes_body<8, 2, 5, true>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  else if (wa->lean && d.N == 7) b = lwfb::lanes_body<8, 2, 7, true>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  else if (d.N == 3) b = lwfb::lanes_body<8, 3>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  else if (d.N < 8) b = lwfb::lanes_body<8, 2>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  else if (d.N == 8) b = lwfb::lanes_body<8, 1>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  else b = lwfb::lanes_body<32, 1>(d, lane, warp_g, cst, lwfb_log_tab_h, lwfb_exp_tab_h, wa->msm);
  if (tid == 0) wa->by_budget = b;
}
#endif

void* worker(void* p) {
  Job* j = (Job*)p;
  for (;;) {
    int64_t u = j->next.fetch_add(1);
    if (u >= j->n_units) break;
#ifdef LWFB_EMUL_WIDE
    {
      std::vector<double> blk(LWFB_WIDE_DOUBLES_PER_MEMBER(j->d.N));
      WarpArg wa{&j->d, u, 1, blk.data(), false, false, nullptr};
      run_block(lanes_entry, &wa, 1);
    }
#else
    if (j->kernel == 0) {
      std::vector<double> cst((size_t)lwfb::KC_COUNT * LWFB_LANES_BLOCK);
      std::vector<double> msm((size_t)(lwfb::KM_COUNT + 2 * LWFB_F) * j->warps * (32 / j->d.N));
      WarpArg wa{&j->d, u, j->warps, cst.data(), false, j->lean, msm.data()};
      run_block(lanes_entry, &wa, j->warps);
      if (j->use_stop) {
        const int64_t n_exited = j->exited.fetch_add(1) + 1;
        if (wa.by_budget && n_exited + j->resident > j->n_units) ((volatile int*)j->stop)[1] = 1;
      }
    } else {
      lwfb::integrate_member<0>(j->d, u);
    }
#endif
  }
  return nullptr;
}

}  // namespace

// profiling aid: stage barriers executed by thread 0 of all blocks since the last call (6 or 7 per step pass of a block)
extern "C" long lwfb_emul_stage_syncs() { return simt::g_stage_syncs.exchange(0); }
extern "C" long lwfb_emul_votes() { return simt::g_votes.exchange(0); }
// calls of the cold paths of exp / log (per lane) since the last call
extern "C" long lwfb_emul_cold_calls() { return __atomic_exchange_n(&lwfb_cold_calls, 0, __ATOMIC_RELAXED); }

// observations scored during the following runs (the emulator's lwfb_set_observations) and the sums of the last run
static std::vector<int32_t> g_obs_of_rec, g_obs_layer;
static std::vector<double> g_obs_theta, g_obs_d18o, g_sacc;
extern "C" void lwfb_emul_set_obs(int n_days, int n_obs_days, const int32_t* obs_day, int n_obs_layers, const int32_t* obs_layer,
                                  const double* obs_theta, const double* obs_d18o) {
  g_obs_of_rec.assign((size_t)n_days + 1, -1);
  for (int k = 0; k < n_obs_days; k++) g_obs_of_rec[(size_t)obs_day[k]] = k;
  g_obs_layer.assign(obs_layer, obs_layer + n_obs_layers);
  const size_t n = (size_t)n_obs_days * n_obs_layers;
  g_obs_theta.clear(); g_obs_d18o.clear();
  if (n_obs_days > 0 && obs_theta) g_obs_theta.assign(obs_theta, obs_theta + n);
  if (n_obs_days > 0 && obs_d18o) g_obs_d18o.assign(obs_d18o, obs_d18o + n);
  if (n_obs_days == 0) g_obs_layer.clear();
}
// du_NTFLI cache handed to the following runs / left by the last run (the emulator's lwfb_set_ntfli0 / lwfb_get_ntfli)
static std::vector<double> g_ntfli0, g_ntfli_out;
extern "C" void lwfb_emul_set_ntfli0(const double* p, int64_t n) { if (p && n > 0) g_ntfli0.assign(p, p + n); else g_ntfli0.clear(); }
extern "C" void lwfb_emul_get_ntfli(double* out, int64_t n) { for (int64_t k = 0; k < n && k < (int64_t)g_ntfli_out.size(); k++) out[k] = g_ntfli_out[(size_t)k]; }
// lane-kernel RHS probe of the next lwfb_emul_run (the emulator's lwfb_eval_rhs for 3 <= N <= 32): one launch, one pass
static const double *g_probe_frozen = nullptr, *g_probe_ntfli = nullptr;
static double g_probe_t = 0;
static double* g_probe_out = nullptr;
extern "C" void lwfb_emul_set_probe(const double* frozen, const double* ntfli, double t, double* out) {
  g_probe_frozen = frozen; g_probe_ntfli = ntfli; g_probe_t = t; g_probe_out = out;
}
extern "C" void lwfb_emul_get_sums(double* out, int64_t n) { for (int64_t k = 0; k < n && k < (int64_t)g_sacc.size(); k++) out[k] = g_sacc[(size_t)k]; }

extern "C" int lwfb_emul_run(int kernel, int N, int iso, int t0_day, int n_days, int n_sites, int n_rows,
                             const double* forcing, int64_t M, const int32_t* site_of_member, const double* scal,
                             const double* lay, const double* u0, double reltol, double abstol, double dt_init,
                             double dtmax, int64_t maxiters, int n_threads, double* out_daily, int32_t* status,
                             int64_t* counters, const int32_t* perm, int seg_len, int resort,
                             int64_t perm_len /* slots described by perm (may exceed M: empty slots = -1); 0: M */,
                             int warps_per_block /* 1 .. LWFB_LANES_BLOCK / 32 */) {
  (void)n_sites;
  const int L = LWFB_STATE_LEN(N, iso);
  if (kernel == 0 && (N < 3 || N > 32)) return -1;
#ifdef LWFB_EMUL_WIDE
  if (kernel != 2) return -1;
#else
  if (kernel == 2) return -1;
  if (kernel == 1 && N > LWFB_MEMBER_MAX_LAYERS) return -1;
#endif
  // the device layouts are member-major = the layouts of the arguments: the kernels read and write them in place
  std::vector<double> fin((size_t)M * L);
  static_assert(sizeof(long long) == sizeof(int64_t), "counter width");
  Job job;
  lwfb::DevData& d = job.d;
  d.M = M; d.N = N; d.iso = iso; d.t0_day = t0_day; d.n_days = n_days; d.n_rows = n_rows; d.L = L;
  d.scal = scal; d.lay = lay; d.forc = forcing; d.site = site_of_member; d.perm = perm; d.u0 = u0;
  d.rec = out_daily; d.fin = fin.data(); d.status = status; d.counters = (long long*)counters;
  d.reltol = reltol; d.abstol = abstol; d.dt_init = dt_init; d.dtmax = dtmax; d.maxiters = maxiters;
  g_sacc.assign((size_t)M * 2 * lwfb::LWFB_SA, 0.0);
  const bool with_obs = !g_obs_layer.empty() && (int)g_obs_of_rec.size() == n_days + 1;
  d.n_obs_layers = with_obs ? (int)g_obs_layer.size() : 0;
  d.obs_of_rec = g_obs_of_rec.data(); d.obs_layer = g_obs_layer.data();
  d.obs_theta = g_obs_theta.empty() ? nullptr : g_obs_theta.data();
  d.obs_d18o = g_obs_d18o.empty() ? nullptr : g_obs_d18o.data();
  d.sacc = g_sacc.data();
  g_ntfli_out.assign((size_t)M * N, 0.0);
  d.ntfli0 = (g_ntfli0.size() == (size_t)M * N) ? g_ntfli0.data() : nullptr;
  d.ntfli_out = g_ntfli_out.data();
  d.probe_frozen = g_probe_frozen; d.probe_ntfli = g_probe_ntfli; d.probe_t = g_probe_t; d.probe_out = g_probe_out;
  {   // the library's classification (optional_paths_kernel); LWFB_EMUL_GENERAL=1 keeps the general instantiation
    bool used = false;
    for (int64_t m = 0; m < M; m++) {
      const double* sc = scal + m * LWFB_PS;
      const bool slope = !(sc[LWFB_S_LENGTH_SLOPE] == 0 || sc[LWFB_S_DSLOPE] == 0);
      used = used || !(sc[LWFB_S_GSC] < 1.0e-8) || sc[LWFB_S_BYPAR] == 1.0 || (int)sc[LWFB_S_QLAYER] > 0 || slope;
    }
    const char* g = getenv("LWFB_EMUL_GENERAL");
    job.lean = !used && !(g && atoi(g) != 0);
  }
  job.kernel = kernel;
  if (warps_per_block < 1 || warps_per_block > MAX_WARPS) return 4;
  job.warps = warps_per_block;
  if (kernel == 0) { const int mpw = 32 / N; job.n_units = ((M + mpw - 1) / mpw + job.warps - 1) / job.warps; }
  else job.n_units = M;
  if (n_threads < 1) n_threads = 1;
  // work-quantum execution as in lwfb_run: a launch ends after `seg_len` passes of the block loop (here: of the warp
  // loop), the unfinished members are regrouped by progress (resort) and the next launch resumes them from fin + ck
  const size_t ck_rows = (size_t)(lwfb::LWFB_CK + lwfb::LWFB_CKL * N);
  std::vector<double> ck(ck_rows * M);
  std::vector<int32_t> perm2((size_t)M);
  const bool quanta = kernel == 0 && seg_len > 0;
  d.ck = quanta ? ck.data() : nullptr;
  d.iter_budget = quanta ? seg_len : 0;
  d.n_slots = (perm && perm_len > 0) ? perm_len : 0;
  d.resume = 0;
  job.use_stop = quanta && resort == 2;              // resort == 2: regrouping + stop flag (front block without a budget)
  job.resident = n_threads;
  d.stop = job.stop; d.resident = n_threads; d.cb_wait = 0;
  if (const char* e = getenv("LWFB_CB_WAIT")) d.cb_wait = atoi(e);     // profiling knob, as in the library
  d.prio_blocks = quanta && resort ? 1 : 0;          // the first block runs three times as long as the others ...
  d.prio_budget = job.use_stop ? (1LL << 60) : 3 * (long long)seg_len;   // ... or until the stop flag is raised
  const double tend = (double)(t0_day + n_days);
  for (int launch = 0;; launch++) {
    if (launch > 100000) return 3;
    job.next = 0;
    job.exited = 0; job.stop[0] = job.stop[1] = 0;
    if (kernel == 0) { const int mpw = 32 / N; const int64_t ns = d.n_slots > 0 ? d.n_slots : M; job.n_units = ((ns + mpw - 1) / mpw + job.warps - 1) / job.warps; }
    std::vector<pthread_t> th(n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], nullptr, worker, &job);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], nullptr);
    if (!quanta || d.probe_out) break;
    int64_t left = 0;
    for (int64_t i = 0; i < M; i++)
      if (status[i] == LWFB_ST_OK && ck[(size_t)i * ck_rows + lwfb::LWFB_CK_T] < tend) perm2[(size_t)left++] = (int32_t)i;
    if (left == 0) break;
    d.resume = 1;
    if (resort) {
      // experiment knob LWFB_EMUL_KEY=1: (day, passes to the day boundary at the current step size, most first) instead of t
      static const int key_mode = getenv("LWFB_EMUL_KEY") ? atoi(getenv("LWFB_EMUL_KEY")) : 0;
      auto keyf = [&](int32_t a) {
        const double* c = ck.data() + (size_t)a * ck_rows;
        const double ta = c[lwfb::LWFB_CK_T];
        if (key_mode == 0) return ta;
        const double left_p = std::min(65535.0, std::max(0.0, (c[lwfb::LWFB_CK_TSTOP] - ta) / c[lwfb::LWFB_CK_DT]));
        const int fl = (int)c[lwfb::LWFB_CK_FLAGS];
        const double dayk = floor(ta) + ((fl & 2) ? -0.0 : 0.0);
        return dayk * 65536.0 + (65535.0 - left_p);
      };
      std::stable_sort(perm2.begin(), perm2.begin() + left, [&](int32_t a, int32_t b) { return keyf(a) < keyf(b); });
      d.perm = perm2.data();
      d.n_slots = left;
    }
  }
  return 0;
}
