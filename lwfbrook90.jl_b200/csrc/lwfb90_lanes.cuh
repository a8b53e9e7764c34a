// lwfb90_lanes.cuh — the production integrator: one LANE per soil layer.
//
// Mapping: a member with N layers (3 <= N <= 32) occupies N consecutive lanes of a warp ("group");
// a warp holds floor(32/N) members, the remaining lanes shadow lane 0 and never commit.  Lane i of a
// group owns layer i: its hydraulic constants (one column of shared memory per thread, ~40 doubles),
// its SWATI component with the Runge-Kutta bookkeeping, its delta18O / delta2H, and up to EQ of the 8
// integrated flow accumulators (pure quadratures, dealt round-robin to the lanes of the group).  GWAT
// and the step-size controller are member-level scalars replicated in the lanes of the group; they
// evolve identically because every input is a broadcast value.
//
// Everything that changes only once per day (interception / snow states, the 23 non-integrated
// accumulator entries, RWU, delta of INTS/INTR/SNOW) lives in global memory in `fin`, the member's full
// state vector in the reference's order; the daily record is a snapshot copy of it.
//
// Why this mapping (see profiles/r01_*): with one member per THREAD the per-layer arrays only stay in
// registers if the layer loops are unrolled; one RHS is then ~9000 instructions, the kernel needs 255
// registers (8 warps per SM) and ncu showed the warps waiting ~10 cycles per issued instruction on
// instruction fetch, FP64 pipe 6 % busy.  One lane per layer keeps the per-layer body once in the code,
// divides the register need by N, and turns cross-layer couplings into warp shuffles:
//   * Darcy flux needs psi_t and log K of the layer below           -> __shfl_down
//   * ITER's min-chain, bypass / downslope sums, the error norm      -> broadcast chains evaluated in
//     exactly the oracle's order so that results stay bit-identical; ITER takes a plain min when no
//     clamp / NaN is in play (warp vote), the sums are skipped when every term is zero
//   * INFLOW's bottom-up oversaturation limiter                      -> skipped by a warp vote when no layer
//     limits (the common case), otherwise an N-step chain
// The daily callback runs warp-synchronously (lanes wait at the day boundary for the slowest member of
// the warp).  The elementary functions (include/lwfb90_elem.h) are branch-free so that the independent
// pow/exp/log of one layer interleave in the FP64 pipe.
#pragma once
#include "lwfb90_device.cuh"

namespace lwfb {

// Tsit5 tableau for the rolled stage loop.  The rows are indexed with the (warp-uniform) stage number at run time:
// constant memory on the device (a __device__ constexpr array would be read from global memory).
// CQ[ks][p] = a(ks+1+p, ks): contribution of k_ks to the input of stage ks+1+p (stage 7 = the new state)
#define LWFB_TAB_CQ { {0, 0, 0, 0, 0}, \
    {-0.008480655492356989, 2.8971530571054935, 5.325864828439257, 5.86145544294642, 0.09646076681806523}, \
    {0.335480655492357, -6.359448489975075, -11.748883564062828, -12.92096931784711, 0.01}, \
    {4.3622954328695815, 7.4955393428898365, 8.159367898576159, 0.4798896504144996, 0}, \
    {-0.09249506636175525, -0.071584973281401, 1.379008574103742, 0, 0}, \
    {-0.028269050394068383, -3.290069515436081, 0, 0, 0}, \
    {2.324710524099774, 0, 0, 0, 0} }
#define LWFB_TAB_B7 {0, 0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774, 0}
#define LWFB_TAB_BT {0, -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, \
                     0.5823571654525552, -0.45808210592918697, 0.015151515151515152}
#define LWFB_TAB_C {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0}
#ifdef __CUDACC__
#define LWFB_TABQ static __constant__
#else
#define LWFB_TABQ static const
#endif
LWFB_TABQ double lCQ[7][5] = LWFB_TAB_CQ;
LWFB_TABQ double lB7[8] = LWFB_TAB_B7;        // a(7, ks)
LWFB_TABQ double lBT[8] = LWFB_TAB_BT;
LWFB_TABQ double lC[7] = LWFB_TAB_C;          // time of the stage evaluated at s
__device__ constexpr int lAccPos[8] = {LWFB_A_FLOW, LWFB_A_SEEP, LWFB_A_SRFL, LWFB_A_SLFL,
                                       LWFB_A_BYFL, LWFB_A_DSFL, LWFB_A_GWFL, LWFB_A_VRFLN};

#define LFULL 0xffffffffu
#ifndef LWFB_LANES_BLOCK
#define LWFB_LANES_BLOCK 256        // threads per block = stride of the shared-memory constant columns
#endif
#ifndef LWFB_LANES_MINBLOCKS
#define LWFB_LANES_MINBLOCKS (512 / LWFB_LANES_BLOCK)   // resident blocks per SM the register allocation is sized for
#endif
// Block-level control votes: the warps of a block advance in LOCK STEP.  "Does anybody step / need the callback /
// need the FSAL re-evaluation" are block-uniform decisions taken with __syncthreads_or, and every Runge-Kutta stage
// starts with a barrier.  Reason (profiles/r01_*): the stepping loop is ~40 KB of straight-line code; free-running
// warps each stream it through the instruction caches on their own and spent 7.8 cycles per issued instruction
// waiting for instruction fetch (IPC 1.0 per SM).  In lock step the warps that share a scheduler fetch the same
// lines at the same time: fetch stalls drop to 0.2-0.6 cycles per instruction and IPC doubles.  The price is that a
// block waits for its slowest member at every step (a few % more executed instructions).  Results do not depend on
// the grouping: a member that waits does not change state.  -DLWFB_LANES_FREE_RUNNING restores warp-level votes.
// (Round 2 tried WARP-level phases - a warp runs its daily callbacks while the other warps of the block step, stage
// barriers counted among the stepping warps only: executed / needed member-steps 1.03 instead of 1.16 in the emulator,
// but on the B200 the callback's ~160 KB of straight-line code running beside the stepping warps evicts their
// instructions: fetch stalls 0.5 -> 3.6 cycles per instruction, 97 -> 158 ms on the profile workload.  Any second
// instruction stream next to the stepping loop costs more than the waiting it saves; profiles/r02_notes.md.)
#if !defined(LWFB_LANES_FREE_RUNNING) && (defined(__CUDACC__) || defined(LWFB_SIMT_EMUL))
#define LWFB_VOTE_ANY(p) (__syncthreads_or((p) ? 1 : 0) != 0)
#define LWFB_VOTE_COUNT(p) __syncthreads_count((p) ? 1 : 0)
#ifdef LWFB_SIMT_EMUL
// the stage barrier is reached from two places (stage loop / warps that idle through a pass): the hardware matches
// barriers by number, the emulator by this id
#define LWFB_STAGE_SYNC() ((void)simt::block_or(false, 1000000))
#else
#define LWFB_STAGE_SYNC() __syncthreads()
#endif
#else
#error "LWFB_LANES_FREE_RUNNING: the daily callback is a block-level phase (one member per lane over the whole block); the free-running experiment of rounds 1 / 2 is kept as kernel text under profiles/"
#endif
// When does the block run its (block-uniform) daily callback phase?  A member that has reached its day boundary idles
// until the phase runs; the phase itself costs every member of the block a few step passes, whether it takes part or not,
// and the members it releases leave the block's common rhythm.
//  * One launch over the whole record (lwfb_run_async, small ensembles): nothing ever brings the members of a block back
//    together, so the block waits for everybody - but at most LWFB_CB_WAIT = 32 passes (the oldest waiter's), which bounds
//    the cost of one member's bad day for the others (round 1).  On DAV2020 this keeps a block at one phase per day.
//  * Work quanta (lwfb_run): the members of a block are regrouped by progress every LWFB_QUANTUM passes, they reach the
//    day boundary together and fall apart only when one of them has a bad day.  Here the block weighs the two costs
//    against each other (ski-rental rule): it counts the member-passes idled at the day boundary since the last phase
//    and runs the phase as soon as they reach LWFB_CB_WASTE2 / 2 passes for each member that still steps.  Measured on
//    the B200 against the 32-pass bound (profiles/r02_notes.md): DAV2020 127872 x 365 d 2827 -> 2768 ms, isoBEA2016
//    100000 x 366 d 17.0 -> 16.2 s, BEA2016 10000 6.39 -> 6.26 s; in ONE launch of unsorted DAV2020 members it loses
//    (74 -> 91 ms on the profile workload), hence the two modes.  Emulator (tools/emul_waiting.py, BEA2016 in quanta):
//    executed / needed member-steps 1.242 -> 1.124.
// Results do not depend on any of this: a member that waits does not change state.
#ifndef LWFB_CB_WAIT
#define LWFB_CB_WAIT 32
#endif
#ifndef LWFB_CB_WASTE2
#define LWFB_CB_WASTE2 5
#endif

// per-thread constant column (shared memory): element k of this thread is cst[k * LWFB_LANES_BLOCK]
enum {
  KC_THSAT = 0, KC_THR, KC_DTH, KC_SWATMAX, KC_E1, KC_E2, KC_E3, KC_TORT, KC_CPSI, KC_KSAT, KC_PSIG,
  KC_THICK, KC_OMS, KC_INFRAC, KC_THICK_DN, KC_KSAT_DN, KC_TSUM, KC_THALF, KC_STFAC, KC_CITER,
  KC_RDZ_UP, KC_RDZ_DN, KC_RTHS2, KC_RSWATMAX, KC_RDTH, KC_RTSUM, KC_RTHALF,   // R...: reciprocals (div_c, LWFB_RCP)
  // isotope state of the lane (touched once per accepted step): delta18O / delta2H of SWATI[i]
  KC_DS_18O, KC_DS_2H,
  // frozen uptake of the layer (set by the daily callback, read twice per RHS evaluation)
  KC_TRANI,
  // exchange slot: a value of the lane that the other lanes of its group read (instead of N double shuffles)
  KC_XCH,
  KC_COUNT_LEAN,                  // the lean instantiations stop here
  // bypass flow / downslope flow only
  KC_WETF = KC_COUNT_LEAN, KC_ARATIO,
  KC_COUNT
};
// per-MEMBER values (shared memory, one slot per member of the block: element k of the member is msm[k * mpb], mpb =
// members per block; the lanes of a group read the same word - a broadcast - and the groups of a warp neighbouring words)
enum {
  KM_DPSIMAX = 0, KM_DTIMAX, KM_DRAIN, KM_GSC, KM_GSP, KM_IMPERV, KM_QFFC, KM_QFPAR, KM_SWATQF, KM_SWATQX,
  KM_CS, KM_SNRHO, KM_LEN1000, KM_LAM, KM_VXYLEM,
  // day constants written by the daily callback
  KM_SUMT, KM_SUMT_V, KM_D0_18O, KM_D0_2H,
  // isotope state of the member: delta18O / delta2H of GWAT, XYLEM
  KM_DG_18O, KM_DG_2H, KM_DXY_18O, KM_DXY_2H,
  // RNET + SMLT and SLVP of the day (set by the daily callback)
  KM_RS, KM_SLVP,
  // the rates of the 8 integrated accumulators of the last RHS evaluation (+ a zero for lanes that hold none): every
  // lane picks its own by index instead of walking a chain of selects
  KM_RATE0, KM_RATE_ZERO = KM_RATE0 + 8,
  // step-size control and bookkeeping of the member that only the header / footer of a pass touch (kept out of the
  // registers that the stage loop needs: every spilled register costs a trip to L2 here).  All lanes of the group write
  // the same value; read-modify-write sequences read into locals, __syncwarp, then write (no lane may see a value
  // another lane of its group has already advanced).  Counters are stored as doubles (exact).
  KM_TSTOP, KM_DT, KM_QOLD, KM_ITER, KM_NACCEPT, KM_NREJECT, KM_NF, KM_REC, KM_STATUS,
  // GWAT of the member and its Runge-Kutta bookkeeping in the lean N = 7 kernel (no groundwater feedback: GWAT is a pure
  // quadrature of the bottom flux and no stage needs it).  One lane per member - the top lane - accumulates in these
  // slots; in registers the six values would be replicated in every lane, 12 registers of a kernel that spills at its cap.
  KM_G, KM_KFG, KM_AEG, KM_SBG, KM_KCG,
#ifdef LWFB_CB_POLICY_WAIT
  KM_WAITED,
#endif
  KM_COUNT
};
#define LC(k) (*(const volatile double*)(cst + (k) * LWFB_LANES_BLOCK))
#define LM(k) (*(const volatile double*)(msm + (k) * mpb))
#define SETM(k, v) (msm[(k) * mpb] = (v))
#ifdef LWFB_SIMT_EMUL
#define LWFB_WARPS_PER_BLOCK (simt::warps_per_block())          // the emulator's block size is a run-time choice
#else
#define LWFB_WARPS_PER_BLOCK (LWFB_LANES_BLOCK / 32)
#endif
#define LWFB_BLOCK_OF_WARP(w) ((w) / LWFB_WARPS_PER_BLOCK)
// shared memory of a block (doubles): per-thread columns, log and exp tables, per-member slots, forcing rows of two days per member
#define LWFB_SMEM_TABLES (3 * LWFB_LOG_TAB_N + 2 * LWFB_EXP_TAB_N)
#define LWFB_SMEM_DOUBLES(mpb, lean) ((size_t)((lean) ? KC_COUNT_LEAN : KC_COUNT) * LWFB_LANES_BLOCK + LWFB_SMEM_TABLES + (size_t)(KM_COUNT + 2 * LWFB_F) * (mpb))

// NX > 0: the layer count is the compile-time constant NX (index arithmetic and the broadcast loops resolve at compile
// time); 0: run-time N <= NMAX.  LEAN: the ensemble uses none of the optional flow paths - no bypass flow (BYPAR), no
// source-area flow (QLAYER), no downslope flow, no groundwater storage feedback (GSC) - which is what DevData::lean
// certifies (checked on the device when the members are set); their code, registers and constants disappear from the
// stepping loop.  All of BASELINE.json's ensemble configurations are lean.  Same results as the general kernel.
template <int NMAX, int EQ, int NX = 0, bool LEAN = false>
__device__ __forceinline__ bool lanes_body(const DevData& d, const int lane, const int64_t warp_g, double* cst,
                                           const double* ltab /* log table (shared-memory copy on the device) */,
                                           const double* etab /* exp table (likewise) */,
                                           double* msm_block /* per-member slots + forcing rows of the block */) {
  const int N = NX > 0 ? NX : d.N;
  const int mpw = 32 / N;                            // members per warp
  const int mpb = LWFB_WARPS_PER_BLOCK * mpw;        // members per block
  const int64_t M = d.M;
  int g = lane / N;
  int i = lane - g * N;
  const bool lane_valid = g < mpw;
  if (!lane_valid) { g = 0; i = 0; }                 // surplus lanes shadow lane 0 and never commit
  const int base = g * N;
  const int mib = (int)(warp_g % LWFB_WARPS_PER_BLOCK) * mpw + g;   // member within the block
  double* const msm = msm_block + mib;               // the member's slots: msm[k * mpb]
  const double* const xgrp = cst - (lane - base) + (size_t)KC_XCH * LWFB_LANES_BLOCK;   // exchange slots of the group: xgrp[j] = lane j's
  double* const frow2 = msm_block + (size_t)KM_COUNT * mpb + (size_t)mib * 2 * LWFB_F;   // its two staged forcing rows
  const int64_t m_raw = warp_g * mpw + g;            // slot of the group
  const int64_t n_slots = d.n_slots > 0 ? d.n_slots : M;
  const int64_t slot = (m_raw < n_slots) ? m_raw : n_slots - 1;
  const int64_t placed = d.perm ? (int64_t)__ldg(d.perm + slot) : slot;   // member placed in this slot; < 0: the slot is empty
  const bool live = lane_valid && m_raw < n_slots && placed >= 0;
  const int64_t m = placed >= 0 ? placed : 0;        // (lanes that are not live read some member's data and never write)
  const bool top = (i == 0), bot = (i == N - 1);
  const bool nodn = bot || !lane_valid;              // lanes without a real neighbour below
  const int L = d.L;
  const bool iso = d.iso != 0;
  const SIdx ix(N, d.iso);
  const ParamView pv{N, m, M, d.scal, d.lay};
  const double* __restrict__ forc = d.forc + (int64_t)__ldg(d.site + m) * d.n_rows * LWFB_F;
  double* const fin = d.fin;
  // member-major blocks: the member's state vector / checkpoint rows are contiguous, lanes i, i+1, ... of a group touch
  // neighbouring words
  auto U0 = [&](int j) { return __ldg(d.u0 + m * L + j); };
  auto FIN = [&](int j) -> double& { return fin[m * L + j]; };
  auto bcast = [&](double v, int layer) { return __shfl_sync(LFULL, v, base + layer); };

  // ---- per-lane constants -> shared-memory column
  bool BYPAR, SLOPE;
  int QLAYER;
  {
    auto SETC = [&](int k, double v) { cst[k * LWFB_LANES_BLOCK] = v; };
    const double THSAT = pv.Lp(LWFB_L_THSAT, i), SWATMAX = pv.Lp(LWFB_L_SWATMAX, i), THR = pv.Lp(LWFB_L_THR, i);
    const double NPAR = pv.Lp(LWFB_L_NPAR, i), MPAR = pv.Lp(LWFB_L_MPAR, i);
    const double E2 = 1.0 / NPAR;
    const double THICK = pv.Lp(LWFB_L_THICK, i), STONEF = pv.Lp(LWFB_L_STONEF, i);
    const int idn = bot ? i : i + 1;
    const double THICK_dn = pv.Lp(LWFB_L_THICK, idn), STONEF_dn = pv.Lp(LWFB_L_STONEF, idn);
    SETC(KC_THSAT, THSAT); SETC(KC_THR, THR); SETC(KC_DTH, THSAT - THR); SETC(KC_SWATMAX, SWATMAX);
    SETC(KC_RSWATMAX, LWFB_RCP(SWATMAX)); SETC(KC_RDTH, LWFB_RCP(THSAT - THR));
    SETC(KC_E1, -1.0 / MPAR); SETC(KC_E2, E2); SETC(KC_E3, NPAR / (NPAR - 1.0));
    SETC(KC_TORT, pv.Lp(LWFB_L_TORT, i)); SETC(KC_CPSI, 9.81 * (-1.0 / pv.Lp(LWFB_L_ALPHA, i)));
    SETC(KC_KSAT, pv.Lp(LWFB_L_KSAT, i)); SETC(KC_PSIG, pv.Lp(LWFB_L_PSIG, i));
    SETC(KC_THICK, THICK); SETC(KC_OMS, 1 - STONEF); SETC(KC_INFRAC, pv.Lp(LWFB_L_INFRAC, i));
    SETC(KC_THICK_DN, THICK_dn); SETC(KC_KSAT_DN, pv.Lp(LWFB_L_KSAT, idn));
    SETC(KC_TSUM, THICK + THICK_dn); SETC(KC_THALF, (THICK + THICK_dn) / 2.0);
    SETC(KC_RTSUM, LWFB_RCP(THICK + THICK_dn)); SETC(KC_RTHALF, LWFB_RCP((THICK + THICK_dn) / 2.0));
    SETC(KC_STFAC, 1.0 - (STONEF + STONEF_dn) / 2.0);
    SETC(KC_CITER, 0.01 * pv.S(LWFB_S_DSWMAX) * THICK * THSAT * (1 - STONEF));
    BYPAR = pv.S(LWFB_S_BYPAR) == 1.0;
    if (!LEAN) SETC(KC_WETF, BYPAR ? pv.Lp(LWFB_L_WETF, i) : 0.0);
    SETC(KC_RDZ_UP, top ? 1.0 : LWFB_RCP(pv.Lp(LWFB_L_DZ, i - 1))); SETC(KC_RDZ_DN, bot ? 1.0 : LWFB_RCP(pv.Lp(LWFB_L_DZ, i)));
    SETC(KC_RTHS2, LWFB_RCP(THSAT * THSAT));
    SETM(KM_DPSIMAX, pv.S(LWFB_S_DPSIMAX)); SETM(KM_DTIMAX, pv.S(LWFB_S_DTIMAX)); SETM(KM_DRAIN, pv.S(LWFB_S_DRAIN));
    SETM(KM_GSC, pv.S(LWFB_S_GSC)); SETM(KM_GSP, pv.S(LWFB_S_GSP)); SETM(KM_IMPERV, pv.S(LWFB_S_IMPERV));
    SETM(KM_QFFC, pv.S(LWFB_S_QFFC)); SETM(KM_QFPAR, pv.S(LWFB_S_QFPAR));
    SETM(KM_SWATQF, pv.S(LWFB_S_SWATQF)); SETM(KM_SWATQX, pv.S(LWFB_S_SWATQX));
    QLAYER = (int)pv.S(LWFB_S_QLAYER);
    const double DSLOPE = pv.S(LWFB_S_DSLOPE), LEN = pv.S(LWFB_S_LENGTH_SLOPE);
    SLOPE = !(LEN == 0 || DSLOPE == 0);
    const double cs = SLOPE ? lwfb_cos(DSLOPE) : 1.0, sn = SLOPE ? lwfb_sin(DSLOPE) : 0.0;   // DSLOPE used as radians (reference quirk)
    SETM(KM_CS, cs); SETM(KM_SNRHO, sn * p_RHOWG); SETM(KM_LEN1000, 1000 * LEN);
    if (!LEAN) SETC(KC_ARATIO, SLOPE ? THICK * (1 - STONEF) * cs / (1000 * LEN) : 0.0);
    SETM(KM_LAM, pv.S(LWFB_S_DISPERSIVITY)); SETM(KM_VXYLEM, pv.S(LWFB_S_VXYLEM));
  }
  const unsigned grpmask = (N == 32) ? LFULL : (((1u << N) - 1u) << base);
  // warp-uniform switches, voted once (inside the loops they would be re-derived from global memory at every use)
  const bool any_bypar = LEAN ? false : __any_sync(LFULL, BYPAR), any_qlayer = LEAN ? false : __any_sync(LFULL, QLAYER > 0);
  const bool any_slope = LEAN ? false : __any_sync(LFULL, SLOPE);
  const bool gw_feedback = LEAN ? false : __any_sync(LFULL, !(pv.S(LWFB_S_GSC) < 1.0e-8));   // GWAT enters the RHS only through GSC >= 1e-8

  // ---- live state: fin := u0 (then kept current at day boundaries); registers for what evolves per step
  const bool resume = d.resume != 0;                 // continue from the checkpoint of the previous launch
  const int ck_rows = LWFB_CK + LWFB_CKL * N;
  auto CKP = [&](int r) -> double& { return d.ck[m * ck_rows + r]; };
  auto CKL = [&](int r) -> double& { return d.ck[m * ck_rows + LWFB_CK + r * N + i]; };   // per-layer rows
  if (live && !resume) for (int j = i; j < L; j += N) FIN(j) = U0(j);
  __syncwarp(LFULL);
  double y = FIN(ix.SWATI + i), G = FIN(ix.GWAT);
  double Q[EQ], kFQ[EQ];
  int qe[EQ];                                        // accumulator held in slot j (>= 8: none)
  int qr[EQ];                                        // its rate slot (KM_RATE0 + qr; 8 = the zero)
  SETM(KM_RATE_ZERO, 0.0);
#pragma unroll
  for (int j = 0; j < EQ; j++) {
    qe[j] = i + j * N;
    qr[j] = qe[j] < 8 ? qe[j] : 8;
    Q[j] = qe[j] < 8 ? FIN(ix.ACC + lAccPos[qe[j] & 7]) : 0.0;
    kFQ[j] = resume ? CKL(LWFB_CKL_KFQ + j) : 0.0;
  }
  double kFy = resume ? CKL(LWFB_CKL_KFY) : 0.0, kFG = resume ? CKP(LWFB_CK_KFG) : 0.0;   // FSAL derivative of SWATI / GWAT
  // GWAT bookkeeping in shared slots (KM_G ...): measured on the same box, isoBEA2016 (N = 7) 13.93 -> 13.52 s, DAV2020
  // (N = 5) 2576 -> 2596 ms (profiles/r02_exp_gwat_slots_ab.log) - the seven-layer kernel spills more and gains, the
  // five-layer kernel pays for the extra shared-memory traffic: on for the N = 7 instantiation only
  constexpr bool GWS = LEAN && NX == 7;
  const bool gw_lane = top && lane_valid;            // GWS: the lane that keeps the member's GWAT bookkeeping in shared slots
  if (GWS) { SETM(KM_G, G); SETM(KM_KFG, kFG); }     // (from here on such a kernel does not use the registers G, kFG)
  auto SETK = [&](int k, double v) { cst[k * LWFB_LANES_BLOCK] = v; };
  for (int s = 0; s < 2; s++) {
    SETK(KC_DS_18O + s, iso ? FIN(ix.SWATI + (s + 1) * N + i) : 0.0);
    SETM(KM_DG_18O + s, iso ? FIN(ix.GWAT + 1 + s) : 0.0);
    SETM(KM_DXY_18O + s, iso ? FIN(ix.XYLEM + 1 + s) : 0.0);
  }
  // (the frozen fluxes of the day live in shared memory: read once or twice per RHS evaluation, they were the first
  // victims of the register allocator, and a spilled register costs an L2 round trip here - the L1 left beside
  // 2 x 100 KB of shared memory is far smaller than the blocks' local memory)
  SETK(KC_TRANI, resume ? CKL(LWFB_CKL_TRANI) : NAN);   // frozen uptake of the layer
  // du_NTFLI of the previous RHS call (NaN on a fresh problem, or what an earlier run on the same problem left behind)
  double ntfli = resume ? CKL(LWFB_CKL_NTFLI) : (d.ntfli0 ? __ldg(d.ntfli0 + m * N + i) : NAN);
  SETM(KM_RS, resume ? CKP(LWFB_CK_RS) : NAN); SETM(KM_SLVP, resume ? CKP(LWFB_CK_SLVP) : NAN);   // RNET + SMLT and SLVP of the day
  if (resume) {                                      // day constants of the callback
    SETM(KM_SUMT, CKP(LWFB_CK_SUMT)); SETM(KM_SUMT_V, CKP(LWFB_CK_SUMT_V));
    SETM(KM_D0_18O, CKP(LWFB_CK_D0_18O)); SETM(KM_D0_2H, CKP(LWFB_CK_D0_2H));
  }
  double VV = 0, INFLI = 0, DSFLI = 0;               // flux caches of the last RHS (isotope step)

  // value of accumulator e (held by lane e % N in slot e / N), broadcast to the group
  auto q_bcast = [&](const double* v, int e) {
    const int slot = (e >= N) + (e >= 2 * N);          // e / N for e < 8, N >= 3
    double mine = v[0];
#pragma unroll
    for (int j = 1; j < EQ; j++) mine = (slot == j) ? v[j] : mine;
    return bcast(mine, e - slot * N);
  };
  // registers -> fin for everything that evolves per step
  auto flush_state = [&]() {
    if (!live) return;
    FIN(ix.SWATI + i) = y;
    if (iso) { FIN(ix.SWATI + N + i) = LC(KC_DS_18O); FIN(ix.SWATI + 2 * N + i) = LC(KC_DS_2H); }
#pragma unroll
    for (int j = 0; j < EQ; j++) if (qe[j] < 8) FIN(ix.ACC + lAccPos[qe[j] & 7]) = Q[j];
    if (top) {
      FIN(ix.GWAT) = GWS ? LM(KM_G) : G;
      if (iso) { FIN(ix.GWAT + 1) = LM(KM_DG_18O); FIN(ix.GWAT + 2) = LM(KM_DG_2H); FIN(ix.XYLEM + 1) = LM(KM_DXY_18O); FIN(ix.XYLEM + 2) = LM(KM_DXY_2H); }
    }
  };
  // Forcing: the rows of the current and of the next day of the member's site are STAGED in shared memory (two slots per
  // member, row r in slot r & 1).  The daily callback at day d reads row d, which was fetched a day of simulated time
  // earlier, and starts the asynchronous copy (cp.async, 16 bytes per lane and instruction) of row d + 1 into the other
  // slot; the per-step isotope update reads delta of precipitation from whichever of the two rows applies (the value of
  // row r holds for t in (r - 0.5, r + 0.5]).  No load of forcing data is ever waited for inside the stepping loop.
  auto stage_row = [&](int day) {
    const int r = day < 0 ? 0 : (day > d.n_rows - 1 ? d.n_rows - 1 : day);
    const double* src = forc + (int64_t)r * LWFB_F;
    double* dst = frow2 + (day & 1) * LWFB_F;
    for (int c = i; c < LWFB_F / 2; c += N) lwfb_cp_async16(dst + 2 * c, src + 2 * c);
  };
  auto staged_wait = [&]() { lwfb_cp_async_wait_all(); __syncwarp(LFULL); };   // own copies landed, then visible to the group
  auto prec_delta = [&](double t, double* dp) {
    int r = (int)ceil(t - 0.5);
    r = r < 0 ? 0 : (r > d.n_rows - 1 ? d.n_rows - 1 : r);
    const volatile double* row = frow2 + (r & 1) * LWFB_F;
    dp[0] = row[LWFB_F_D18O];
    dp[1] = row[LWFB_F_D2H];
  };
  // goodness-of-fit sums at the observation days (lwfb_set_observations), from what daily record r holds: theta of the
  // observed layers from SWATI, delta18O.  Called by the whole warp; `mine`: the lane's member saves record r now.
  // Same expressions and the same order (days outer, layers inner) as score_sums_kernel uses on a resident record.
  auto score_rec = [&](int r, bool mine, double sw_i, double d18_i) {
    if (d.n_obs_layers == 0) return;                 // (uniform)
    const int ko = mine ? __ldg(d.obs_of_rec + r) : -1;
    if (!__any_sync(LFULL, ko >= 0)) return;
    const double th_ = score_theta(LC(KC_THSAT), LC(KC_THR), LC(KC_SWATMAX), sw_i);
    double* a = d.sacc + m * (2 * LWFB_SA);
    for (int l = 0; l < d.n_obs_layers; l++) {
      const int il = __ldg(d.obs_layer + l);
      const double s_th = bcast(th_, il), s_d = bcast(d18_i, il);
      if (ko >= 0 && top && live) {
        if (d.obs_theta) score_add(a, s_th, __ldg(d.obs_theta + (int64_t)ko * d.n_obs_layers + l));
        if (d.obs_d18o && iso) score_add(a + LWFB_SA, s_d, __ldg(d.obs_d18o + (int64_t)ko * d.n_obs_layers + l));
      }
    }
  };

  const double qmin = 0.2, qmax = 10.0, gamma = 0.9, beta1 = 0.14, beta2 = 0.08, qoldinit = 1e-4;
  const double t0 = (double)d.t0_day, tend = (double)(d.t0_day + d.n_days);
  double t = resume ? CKP(LWFB_CK_T) : t0;
  SETM(KM_TSTOP, resume ? CKP(LWFB_CK_TSTOP) : t0);
  SETM(KM_DT, resume ? CKP(LWFB_CK_DT) : d.dt_init); SETM(KM_QOLD, resume ? CKP(LWFB_CK_QOLD) : qoldinit);
  const double dtmin_opt = fmax(eps_of(tend), eps_of(t0));
  const double maxiters = (double)(d.maxiters > 0x7fffffffLL ? 0x7fffffffLL : d.maxiters);
  SETM(KM_ITER, resume ? CKP(LWFB_CK_ITER) : 0.0); SETM(KM_NACCEPT, resume ? CKP(LWFB_CK_NACCEPT) : 0.0);
  SETM(KM_NREJECT, resume ? CKP(LWFB_CK_NREJECT) : 0.0); SETM(KM_NF, resume ? CKP(LWFB_CK_NF) : 0.0);
  SETM(KM_REC, resume ? CKP(LWFB_CK_REC) : 1.0);
#ifdef LWFB_CB_POLICY_WAIT
  SETM(KM_WAITED, 0.0);                              // step iterations spent waiting at the day boundary
#endif
  bool reeval = false, at_boundary = true, cand_nan = (y != y);
  const int status_in = resume ? d.status[m] : LWFB_ST_OK;
  SETM(KM_STATUS, (double)status_in);
  if (resume) {
    const int fl = (int)CKP(LWFB_CK_FLAGS);
    reeval = (fl & 1) != 0; at_boundary = (fl & 2) != 0;
    cand_nan = CKL(LWFB_CKL_NAN) != 0.0;
  }
  bool by_budget = false;
#ifdef LWFB_CB_POLICY_WAIT
  const int cb_wait = d.cb_wait > 0 ? d.cb_wait : LWFB_CB_WAIT;
#else
  const bool waste_rule = d.iter_budget > 0;         // work quanta: weigh idling against the phase; one launch: bounded waiting
  const int cb_waste2 = d.cb_wait > 0 ? d.cb_wait : LWFB_CB_WASTE2;
  const int cb_wait = (!waste_rule && d.cb_wait > 0) ? d.cb_wait : LWFB_CB_WAIT;
  int idle_acc = 0;                                  // member-passes idled at the day boundary since the last callback phase (block-uniform)
  int oldest_wait = 0;                               // passes the longest-waiting member has idled (block-uniform: a phase releases all)
#endif
  // block-uniform: every thread of a block makes the same passes
  long long passes_left = d.iter_budget > 0 ? (LWFB_BLOCK_OF_WARP(warp_g) < d.prio_blocks ? d.prio_budget : d.iter_budget) : -1;
#ifdef LWFB_DEBUG_ITERS
  unsigned dbg_iters = 0;          // step-phase iterations of the block (profiling: lock-step waiting overhead)
  unsigned dbg_day_iter0 = 0;
#endif

  double* const rec_m = d.rec ? d.rec + m * (int64_t)(d.n_days + 1) * L : nullptr;   // the member's records [n_days+1][L]
  if (d.rec && live && !resume) for (int j = i; j < L; j += N) __stcs(rec_m + j, U0(j));     // record 0 = raw u0
  __syncwarp(LFULL);
  if (!resume) score_rec(0, true, U0(ix.SWATI + i), iso ? U0(ix.SWATI + N + i) : 0.0);
  { const int d0 = (int)floor(t); stage_row(d0); stage_row(d0 + 1); staged_wait(); }
  const bool probe = d.probe_out != nullptr;         // lwfb_eval_rhs: one evaluation of f(u, t), no callback, no step
  if (probe) {
    const double* fr = d.probe_frozen + m * (N + 3);
    t = d.probe_t; at_boundary = false; reeval = true;
    SETK(KC_TRANI, __ldg(fr + 3 + i)); SETM(KM_RS, __ldg(fr) + __ldg(fr + 1)); SETM(KM_SLVP, __ldg(fr + 2));
    ntfli = __ldg(d.probe_ntfli + m * N + i);
    __syncwarp(LFULL);
  }

  for (;;) {
    const int status = (int)LM(KM_STATUS);
#ifdef LWFB_CB_POLICY_WAIT
    const int waited = (int)LM(KM_WAITED);
#endif
    const bool active = ((status == LWFB_ST_OK) && (t < tend)) || probe;
    if (passes_left == 0) { by_budget = true; break; }   // quantum used up: checkpoint, the host regroups and relaunches
    if (passes_left > 0) passes_left--;
    // (a raised stop flag ends the launch for everybody; threads may see it one pass apart, the vote decides)
    const bool go_on = d.stop ? (*(volatile const int*)(d.stop + 1) == 0) : true;
    if (!LWFB_VOTE_ANY(active && live && go_on)) break;
    const bool can_step = active && !at_boundary;
    const bool at_cb = active && at_boundary;        // waits for (or is about to get) its daily callback
#ifdef LWFB_CB_POLICY_WAIT
    // (two votes: __syncthreads_or returns a truth value, not the OR of the arguments)
    const bool any_step = LWFB_VOTE_ANY(can_step && live);
    const bool any_overdue = LWFB_VOTE_ANY(at_cb && live && waited >= cb_wait);
#else
    // members of the block that step / that idle at their day boundary (one lane per member counts)
    const int n_step = LWFB_VOTE_COUNT(can_step && live && top);
    const int n_wait = LWFB_VOTE_COUNT(at_cb && live && top);
    const bool any_step = n_step > 0;
    idle_acc += n_wait;
    oldest_wait = n_wait > 0 ? oldest_wait + 1 : 0;
    const bool any_overdue = n_wait > 0 && (oldest_wait > cb_wait || (waste_rule && 2 * idle_acc >= cb_waste2 * n_step));
    if (!any_step || any_overdue) { idle_acc = 0; oldest_wait = 0; }
#endif

    if (!any_step || any_overdue) {
      // ============================== block-synchronous daily callback at t = tstop
      // (1) every lane: soil state of its own layer (SWCHEK clamps the state itself), posted in shared memory
      const double SWATMAX = LC(KC_SWATMAX), E2 = LC(KC_E2);
      const double tstop0 = LM(KM_TSTOP);
      double sw = y;
      if (sw <= 0.0) sw = 0.0; else if (sw > SWATMAX) sw = SWATMAX;
      const int day = (int)floor(t + 0.5);
      if (__any_sync(LFULL, at_cb)) {
        const double w = jminc(div_c(div_c(LC(KC_THSAT) * sw, SWATMAX, LC(KC_RSWATMAX)) - LC(KC_THR), LC(KC_DTH), LC(KC_RDTH)), 1.0);
        const double a = jmaxc(w, 1.e-6);
        const double la = lwfb_log(a);
        const double p1 = lwfb_pow_from_log(a, la, LC(KC_E1));
        const double psim = LC(KC_CPSI) * lwfb_pow(p1 - 1.0, E2);
        const double p3 = lwfb_pow_from_log(a, la, LC(KC_E3));
        const double inner = 1.0 - lwfb_pow(1.0 - p3, 1.0 - E2);
        const double kkv = LC(KC_KSAT) * lwfb_pow_from_log(a, la, LC(KC_TORT)) * (inner * inner);
        const double psit = psim + LC(KC_PSIG);
        if (at_cb && lane_valid) {
          // total potential and conductivity of the layer: into the lane's exchange slot and its frozen-uptake slot (the
          // uptake of the day that ends is dead; the callback writes the new one there)
          cst[KC_XCH * LWFB_LANES_BLOCK] = psit;
          cst[KC_TRANI * LWFB_LANES_BLOCK] = kkv;
        }
        if (top && lane_valid) {
          // the member's request, in its (transient) accumulator-rate slots: 0 none, 1 compute only, 2 compute and commit
          SETM(KM_RATE0 + 0, at_cb ? (live ? 2.0 : 1.0) : 0.0);
          SETM(KM_RATE0 + 1, t); SETM(KM_RATE0 + 2, (double)m); SETM(KM_RATE0 + 3, psim);
        }
      } else if (top && lane_valid) SETM(KM_RATE0 + 0, 0.0);
      staged_wait();                                // forcing row `day` was requested at the previous callback (or at the start)
      __syncthreads();
      // (2) The scalar core of the callback (canopy, snow, interception, uptake by layer: ~2 10^4 operations per member that
      // do not parallelise over layers) runs with ONE MEMBER PER LANE: thread k of the block takes member k of the block,
      // so the 48 members of a DAV2020 block fill two warps instead of occupying six lanes in each of eight.  A callback
      // phase is bound by instruction issue (every warp walks the whole core), so its cost falls with the number of warps
      // that walk it: measured 94 us = 3.9 step passes per phase with one lane per member in every warp
      // (profiles/r02_exp_callback_phase_cost.log).  Rounds 1 / 2 ran the core redundantly in every lane of a member.
      {
        const int tib = (int)(warp_g % LWFB_WARPS_PER_BLOCK) * 32 + lane;     // thread within the block = member it serves
        double* const cst_blk = cst - tib;                                     // column 0 of the block
        double* const msm2 = msm_block + tib;
        const double req = tib < mpb ? *(const volatile double*)(msm2 + (KM_RATE0 + 0) * mpb) : 0.0;
        if (req != 0.0) {
          const double t2 = msm2[(KM_RATE0 + 1) * mpb];
          const int64_t m2 = (int64_t)msm2[(KM_RATE0 + 2) * mpb];
          const double psim0 = msm2[(KM_RATE0 + 3) * mpb];
          const bool live2 = req == 2.0;
          double* const col0 = cst_blk + (tib / mpw) * 32 + (tib % mpw) * N;  // column of the member's layer 0
          const double* frow_m = msm_block + (size_t)KM_COUNT * mpb + (size_t)tib * 2 * LWFB_F;
          const int day2 = (int)floor(t2 + 0.5);
          const ParamView pv2{N, m2, M, d.scal, d.lay};
          double* const fin2 = fin + m2 * L;
          double psiti_a[NMAX], kk_a[NMAX], TRANI_a[NMAX], TRANIst_a[NMAX];
          for (int j = 0; j < N; j++) { psiti_a[j] = col0[j + KC_XCH * LWFB_LANES_BLOCK]; kk_a[j] = col0[j + KC_TRANI * LWFB_LANES_BLOCK]; }
          double nRNET, nSMLT, nSLVP, nTA, nRWU;
          int cbst = LWFB_ST_OK;
          DayState nds;
          nds.INTS = fin2[ix.INTS]; nds.INTR = fin2[ix.INTR]; nds.SNOW = fin2[ix.SNOW]; nds.CC = fin2[ix.CC]; nds.SNOWLQ = fin2[ix.SNOWLQ];
          cb_core<NMAX>(pv2, frow_m + (day2 & 1) * LWFB_F, t2, psim0, psiti_a, kk_a, nds, TRANI_a, nRNET, nSMLT, nSLVP, nTA,
                        TRANIst_a, nRWU, cbst);
          double sT = 0;                         // sum of the layer uptakes and its ratio to the xylem volume (isotope step)
          for (int j = 0; j < N; j++) {
            col0[j + KC_TRANI * LWFB_LANES_BLOCK] = TRANI_a[j];      // the frozen uptake of layer j, into that lane's column
            sT += TRANI_a[j];
          }
          // the surplus lanes of a warp shadow layer 0 of its first member: they must see a finite uptake as well, or their
          // (never committed) state turns NaN and drags the whole warp through the cold paths of exp / log at every call
          if (tib % mpw == 0) for (int l = mpw * N; l < 32; l++) (cst_blk + (tib / mpw) * 32 + l)[KC_TRANI * LWFB_LANES_BLOCK] = TRANI_a[0];
          msm2[KM_SUMT * mpb] = sT;
          msm2[KM_SUMT_V * mpb] = sT / msm2[KM_VXYLEM * mpb];
          msm2[KM_RS * mpb] = nRNET + nSMLT; msm2[KM_SLVP * mpb] = nSLVP;
          {
            // molecular diffusivities of the day (soil temperature = daily mean air temperature)
            const double Tk = nTA + 273.15;
            const double ee = lwfb_exp(-535400 / (Tk * Tk) + 1393.3 / Tk + 2.1876);
            msm2[KM_D0_18O * mpb] = 0.96691 * 1e-9 * ee * 3600 * 24;
            msm2[KM_D0_2H * mpb] = 0.98331 * 1e-9 * ee * 3600 * 24;
          }
          if (cbst != LWFB_ST_OK) msm2[KM_STATUS * mpb] = (double)cbst;
          if (live2) {
            for (int j = 0; j < N; j++) fin2[ix.TRANI + j] = TRANIst_a[j];
            fin2[ix.INTS] = nds.INTS; fin2[ix.INTR] = nds.INTR; fin2[ix.SNOW] = nds.SNOW; fin2[ix.CC] = nds.CC;
            fin2[ix.SNOWLQ] = nds.SNOWLQ; fin2[ix.RWU] = nRWU;
            for (int a2 = 0; a2 < LWFB_A_STORAGESWAT; a2++) fin2[ix.ACC + a2] = nds.acc[a2];
            if (iso) {
              // delta of precipitation at t2 (the value of forcing row r holds for t in (r - 0.5, r + 0.5])
              int r = (int)ceil(t2 - 0.5);
              r = r < 0 ? 0 : (r > d.n_rows - 1 ? d.n_rows - 1 : r);
              const double* row = frow_m + (r & 1) * LWFB_F;
              const double dp[2] = {row[LWFB_F_D18O], row[LWFB_F_D2H]};
#pragma unroll
              for (int s = 0; s < 2; s++) { fin2[ix.INTS + 1 + s] = dp[s]; fin2[ix.INTR + 1 + s] = dp[s]; fin2[ix.SNOW + 1 + s] = dp[s]; }
            }
          }
        }
      }
      __syncthreads();
      // (3) the lanes of the members that had their callback start the new day
      if (at_cb) {
#ifdef LWFB_CB_POLICY_WAIT
        SETM(KM_WAITED, 0.0);
#endif
        y = sw;
#pragma unroll
        for (int j = 0; j < EQ; j++) Q[j] = 0.0;
        reeval = true;
        at_boundary = false;
        SETM(KM_TSTOP, fmin(tstop0 + 1.0, tend));
        stage_row(day + 1);                          // asynchronous: needed from t = day + 0.5 on (delta of precipitation)
      }
      __syncwarp(LFULL);
      continue;
    }

#ifdef LWFB_DEBUG_ITERS
    dbg_iters++;
#endif
#ifdef LWFB_CB_POLICY_WAIT
    if (at_cb) SETM(KM_WAITED, (double)(waited + 1));
#endif
    // ============================== one attempted step (all lanes execute, only `do_step` lanes commit)
    // loopheader! (dt already holds the controller's proposal / the shrunk step after a rejection)
    double dt_s = LM(KM_DT);
    const double tstop = LM(KM_TSTOP);
    int st_s = LWFB_ST_OK;
    if (LM(KM_ITER) + 1.0 > maxiters) st_s = LWFB_ST_MAXITERS;
    {
      // unstable_check: NaN in any SWATI of the member (the state last written by the solver)
      const unsigned nanmask = __ballot_sync(LFULL, cand_nan);
      if (st_s == LWFB_ST_OK && (nanmask & grpmask)) st_s = LWFB_ST_UNSTABLE;
    }
    dt_s = dt_s < d.dtmax ? dt_s : d.dtmax;
    { const double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dt_s < dtmin_t) dt_s = dtmin_t; }
    if (dt_s > tstop - t) dt_s = tstop - t;
    if (st_s == LWFB_ST_OK && !(fabs(dt_s) > dtmin_opt)) st_s = LWFB_ST_DTMIN;
    const bool do_step = can_step && (st_s == LWFB_ST_OK);   // member-level: identical in all lanes of the group

    // perform_step!: rolled stage loop (s = 0 is the FSAL re-evaluation after a callback)
    double xy = y, xG = G;                           // stage inputs
    double qy[5] = {0, 0, 0, 0, 0}, qG[4] = {0, 0, 0, 0};    // qG: inputs of stages 3..6 (only if GWAT feeds back)
    double aEy = 0, aEG = 0, kcy = kFy, kcG = kFG, SbG = 0;
    double Sb[EQ], Sbt[EQ], kcQ[EQ];
#pragma unroll
    for (int j = 0; j < EQ; j++) { Sb[j] = 0; Sbt[j] = 0; kcQ[j] = kFQ[j]; }
    double nt = ntfli;
    unsigned nf_s = 0;                               // RHS evaluations of this pass
    const int s_first = LWFB_VOTE_ANY(reeval && can_step) ? 0 : 1;
    if (!__any_sync(LFULL, can_step && live)) {
      // no member of THIS warp steps (all wait at the day boundary for slower members of the block, or are done):
      // keep the block's barrier choreography and issue nothing else
#ifdef LWFB_STAGE_SYNC_MASK
      for (int s = s_first; s <= (probe ? 0 : 6); s++) if ((LWFB_STAGE_SYNC_MASK >> s) & 1) LWFB_STAGE_SYNC();
#else
      for (int s = s_first; s <= (probe ? 0 : 6); s++) LWFB_STAGE_SYNC();   // (a probe launch ends after stage 0)
#endif
      if (probe) break;
      continue;
    }
#pragma unroll 1
    for (int s = s_first; s <= 6; s++) {
#ifdef LWFB_STAGE_SYNC_MASK
      if ((LWFB_STAGE_SYNC_MASK >> s) & 1) LWFB_STAGE_SYNC();
#else
      LWFB_STAGE_SYNC();
#endif
      // ---- stage input
      if (s == 1) {
        const double a = dt_s * 0.161;
        xy = y + a * kFy;
        if (!GWS) xG = G + a * kFG;
        aEy = lBT[1] * kFy;
        if (GWS) {
          if (gw_lane) { const double kfg = LM(KM_KFG); SETM(KM_AEG, lBT[1] * kfg); SETM(KM_SBG, lB7[1] * kfg); }
        } else { aEG = lBT[1] * kFG; SbG = lB7[1] * kFG; }
#pragma unroll
        for (int p = 0; p < 5; p++) qy[p] = lCQ[1][p] * kFy;
        if (gw_feedback) {
#pragma unroll
          for (int p = 0; p < 4; p++) qG[p] = lCQ[1][p] * kFG;
        }
#pragma unroll
        for (int j = 0; j < EQ; j++) { Sb[j] = lB7[1] * kFQ[j]; Sbt[j] = lBT[1] * kFQ[j]; }
      } else if (s >= 2) {
        xy = y + dt_s * qy[0];
#pragma unroll
        for (int p = 0; p < 4; p++) qy[p] = qy[p + 1];
        qy[4] = 0;
        if (s == 6) { if (!GWS) xG = G + dt_s * SbG; }   // the new GWAT: uprev + dt * sum a7j kj (same summation order)
        else if (gw_feedback) {
          xG = G + dt_s * qG[0];
#pragma unroll
          for (int p = 0; p < 3; p++) qG[p] = qG[p + 1];
          qG[3] = 0;
        }
      } else {
        xy = y;
        if (!GWS) xG = G;
      }
      const double ts = (s == 0) ? t : ((s < 5) ? t + lC[s] * dt_s : t + dt_s);
      const bool skip0 = (s == 0) && !reeval;     // groups that do not need the re-evaluation keep their k1

      // ---- RHS f(x, ts)
      nf_s += skip0 ? 0u : 1u;
      const double SWATMAX = LC(KC_SWATMAX), THR = LC(KC_THR), DTH = LC(KC_DTH), E2 = LC(KC_E2);
      double sw = xy;
      if (sw <= 0.0) sw = 0.0; else if (sw > SWATMAX) sw = SWATMAX;
      const double w = jminc(div_c(div_c(LC(KC_THSAT) * sw, SWATMAX, LC(KC_RSWATMAX)) - THR, DTH, LC(KC_RDTH)), 1.0);   // run constants: div_c
      const double a = jmaxc(w, 1.e-6);
      // a in [1e-6, 1]: log on its main path.  The exponents -1/m, n/(n-1), 1/n, 1-1/n are non-zero for every valid
      // van Genuchten n (finite, > 1), so x^y = exp(y log x) needs no y == 0 case; only the tortuosity l may be 0.
      const double la = lwfb_log_main_t(a, ltab);
      const double TORT = LC(KC_TORT);
      double p1, p3, pT;                                      // a^(-1/m), a^(n/(n-1)), a^l: three chains interleaved
      lwfb_exp3_t(LC(KC_E1) * la, LC(KC_E3) * la, TORT * la, &p1, &p3, &pT, etab);
      pT = (TORT == 0.0) ? 1.0 : pT;                          // x^0 = 1 also where exp(0 * log x) would not say so
      const double x1 = p1 - 1.0, x3 = 1.0 - p3;             // both are 0 at saturation (log -> -inf, power -> 0)
      double l1, l3;
      lwfb_log_main2_t(x1, x3, &l1, &l3, ltab);               // two interleaved chains, then the rare fix-ups
      if (lwfb_log_isspecial(x1) | lwfb_log_isspecial(x3)) {                // one test for both (rare: saturation)
        if (lwfb_log_isspecial(x1)) l1 = lwfb_log_rare(x1);
        if (lwfb_log_isspecial(x3)) l3 = lwfb_log_rare(x3);
      }
      double e1, e3;
      lwfb_exp2_t(E2 * l1, (1.0 - E2) * l3, &e1, &e3, etab);
      const double psim = LC(KC_CPSI) * e1;
      const double theta = w * DTH + THR;
      const double inner = 1.0 - e3;
      const double KSAT = LC(KC_KSAT);
      const double kkv = KSAT * pT * (inner * inner);
      double lk = lwfb_log_main_t(kkv, ltab);
      if (lwfb_log_isspecial(kkv)) lk = lwfb_log_rare(kkv);
      const double psit = psim + LC(KC_PSIG);
      double BYFRAC = 0;
      if (any_bypar) {
        const double WETF = LC(KC_WETF);
        double bf = jminc(lwfb_pow(LM(KM_QFFC), 1 - (1 / LM(KM_QFPAR)) * (w - WETF) / (1 - WETF)), 1.0);
        if (w > 0.99) bf = 1;
        BYFRAC = BYPAR ? bf : 0.0;
      }
      double SAFRAC = 0.;
      if (any_qlayer) {
        double sq = 0;
        for (int j = 0; j < N; j++) { const double v = bcast(sw, j); if (j < QLAYER) sq += v; }
        if (QLAYER > 0) {
          const double SWATQF = LM(KM_SWATQF);
          SAFRAC = jminc(lwfb_pow(LM(KM_QFFC), 1 - (1 / LM(KM_QFPAR)) * (sq - SWATQF) / (LM(KM_SWATQX) - SWATQF)), 1.0);
        }
      }
      const double RS = LM(KM_RS);
      const double SRFL = jminc(LM(KM_IMPERV) + SAFRAC, 1.) * RS;
      const double SLFL = RS - SRFL;
      DSFLI = 0;
      if (any_slope) {
        const double GRAD = LM(KM_CS) * (psim * 2 / LM(KM_LEN1000)) + LM(KM_SNRHO);
        const double dd = kkv * LC(KC_ARATIO) * GRAD / p_RHOWG;
        DSFLI = (SLOPE && !(dd < 0)) ? dd : 0.0;
      }
      // first approximation of the flux to the layer below (bottom layer: drainage)
      const double psit_dn = __shfl_down_sync(LFULL, psit, 1), lk_dn = __shfl_down_sync(LFULL, lk, 1);
      double V0;
      {
        const double THICK = LC(KC_THICK), DRAIN = LM(KM_DRAIN);
        double kkmean = lwfb_exp_t(div_c(LC(KC_THICK_DN) * lk + THICK * lk_dn, LC(KC_TSUM), LC(KC_RTSUM)), etab);
        kkmean = jminc(jminc(kkmean, KSAT), LC(KC_KSAT_DN));   // (a NaN conductivity has made kkmean NaN already)
        const double dpsi = nodn ? 1.0 : (psit - psit_dn);     // no layer below: value unused, keep the divisions on their fast path
        const double grad = div_c(dpsi, LC(KC_THALF), LC(KC_RTHALF));
        const double vd = div_c(grad * kkmean, p_RHOWG, LWFB_RCP(p_RHOWG)) * LC(KC_STFAC);
        const double vb = (DRAIN > 0.00001) ? DRAIN * kkv * LC(KC_OMS) : 0.0;
        V0 = bot ? vb : ((fabs(psit - psit_dn) < LM(KM_DPSIMAX)) ? 0.0 : vd);
      }
      const double V0_up = __shfl_up_sync(LFULL, V0, 1);        // prior flux from the layer above (unused at the top)
      // ITER: min-chain over the layers in index order (the clamp to DTIMIN makes the order matter)
      const double DTRI = 1.0 - (ts - floor(ts));                // = 1 - fmod(ts, 1) for ts >= 0
      double DTI = jmin(DTRI, LM(KM_DTIMAX));
      {
        const double c1 = LC(KC_CITER) / jmaxc(fabs(nt), 0.000001);
        const double c2 = div_c(((THR - theta) * LC(KC_THICK)) / nt, 1.30, LWFB_RCP(1.30));
        const bool neg = nt < 0;
        const double v = neg ? (c2 < c1 ? c2 : c1) : c1;
        // fast path: no NaN and nothing below DTIMIN anywhere => the chain is a plain minimum
        const bool slow = (c1 != c1) || (neg && (c2 != c2)) || !(v >= p_DTIMIN) || !(DTI >= p_DTIMIN);
        if (__any_sync(LFULL, slow)) {
          const unsigned negm = __ballot_sync(LFULL, neg);
          double DTINEW = DTI;
          for (int j = 0; j < N; j++) {
            DTINEW = jmin(DTINEW, bcast(c1, j));
            const double c2j = bcast(c2, j);
            if ((negm >> (base + j)) & 1u) {
              DTINEW = jmin(DTINEW, c2j);
              if (DTINEW < p_DTIMIN) DTINEW = p_DTIMIN;
            }
          }
          if (DTINEW < DTI) DTI = DTINEW;
        } else {
          // every lane posts its bound in its exchange slot and reads the N of its group (same comparison order as the
          // broadcast loop of the slow path)
          cst[KC_XCH * LWFB_LANES_BLOCK] = v;
          __syncwarp(LFULL);
          double gm = v;
          if (NX > 0) {
#pragma unroll
            for (int j = 0; j < NX; j++) { const double vj = *(const volatile double*)(xgrp + j); gm = vj < gm ? vj : gm; }
          } else {
            for (int j = 0; j < N; j++) { const double vj = *(const volatile double*)(xgrp + j); gm = vj < gm ? vj : gm; }
          }
          if (gm < DTI) DTI = gm;
        }
      }
      // INFLOW
      const double INFRAC = LC(KC_INFRAC);
      double BYFLI = LEAN ? 0.0 : SLFL * INFRAC * BYFRAC;
      INFLI = LEAN ? SLFL * INFRAC : SLFL * INFRAC * (1.0 - BYFRAC);
      VV = V0;
      {
        const double room = div_z(SWATMAX - sw, DTI);          // (a saturated layer: 0 / DTI without the division's slow path)
        // no layer limits (DAV2020: always) <=> every layer passes its test with the unmodified fluxes.  (Surplus lanes do
        // not vote: their garbage entered the chain in 18 % of the BEA2016 evaluations.  BEA2016 keeps one layer saturated
        // and limits in 79 % of the evaluations, a cascade of one layer in 98 % of them; sweeping all layers at once until
        // nothing changes - 2 sweeps instead of N chain steps - was measured 5 % SLOWER there: a sweep costs every lane
        // what a chain step costs one, profiles/r02_notes.md.)
        const double TRANI = LC(KC_TRANI), SLVP = LM(KM_SLVP);
        double MAXIN = room + V0 + DSFLI + TRANI;
        if (top) MAXIN = MAXIN + SLVP;
        const bool limited = top ? (INFLI > MAXIN) : (V0_up + INFLI > MAXIN);
        if (__any_sync(LFULL, limited && lane_valid)) {
          // bottom-up chain: layer j fixes the posterior flux from layer j-1.  (With the layer count fixed at compile time
          // the chain unrolls into ~12 predicated instructions per layer.  Two attempts to do less were measured SLOWER on
          // BEA2016 by 5 % each, same box: sweeping all layers at once until nothing changes, and starting the chain at
          // the highest limiting layer with an early exit - votes and a run-time loop cost more than the steps they save.)
          for (int j = N - 1; j >= 0; j--) {
            double MX = room + VV + DSFLI + TRANI;
            double VVup = V0_up;
            if (j == 0) {
              MX = MX + SLVP;
              if (i == 0 && INFLI > MX) {
                const double too = INFLI - MX;
                INFLI = INFLI - too;
                BYFLI = BYFLI + too;
              }
            } else {
              if (i == j && V0_up + INFLI > MX) {
                if (BYFRAC > 0) {
                  if (V0_up < MX) {
                    const double too = INFLI + V0_up - MX;
                    INFLI = INFLI - too;
                    BYFLI = BYFLI + too;
                  } else {
                    VVup = MX;
                    BYFLI = BYFLI + INFLI;
                    INFLI = 0;
                  }
                } else {
                  VVup = MX - INFLI;
                }
              }
            }
            const double recv = __shfl_down_sync(LFULL, VVup, 1);   // lane j-1 receives its posterior outflow
            if (j >= 1 && i == j - 1) VV = recv;
          }
        }
      }
      const double VV_up = __shfl_up_sync(LFULL, VV, 1);
      double du0;
      if (top) du0 = INFLI - VV - DSFLI - LC(KC_TRANI) - LM(KM_SLVP);
      else du0 = VV_up + INFLI - VV - DSFLI - LC(KC_TRANI);
      const double VN = bcast(VV, N - 1);
      double GWFL, SEEP;
      {
        const double GSC = LM(KM_GSC), GSP = LM(KM_GSP);
        if (LEAN || GSC < 1.0e-8) { SEEP = GSP * VN; GWFL = VN - SEEP; }
        else {
          SEEP = xG * GSC * GSP;
          GWFL = xG * GSC * (1. - GSP);
          if (xG - (GWFL + SEEP) * 1 < 0) { SEEP = xG / 1 * GSP; GWFL = xG / 1 * (1. - GSP); }
        }
      }
      double sby = 0, sds = 0;
      if (__any_sync(LFULL, BYFLI != 0.0)) for (int j = 0; j < N; j++) sby += bcast(BYFLI, j);
      if (any_slope) for (int j = 0; j < N; j++) sds += bcast(DSFLI, j);
      const double kny = du0, knG = VN - GWFL - SEEP;
      // rates of the 8 integrated accumulators (member-level values, identical in the lanes of the group): posted in the
      // member's slots, every lane then reads the (up to EQ) ones it integrates
      SETM(KM_RATE0 + 0, SRFL + sby + sds + GWFL);      // flow
      SETM(KM_RATE0 + 1, SEEP); SETM(KM_RATE0 + 2, SRFL); SETM(KM_RATE0 + 3, SLFL); SETM(KM_RATE0 + 4, sby);
      SETM(KM_RATE0 + 5, sds); SETM(KM_RATE0 + 6, GWFL); SETM(KM_RATE0 + 7, VN);
      __syncwarp(LFULL);
      double knQ[EQ];
#pragma unroll
      for (int j = 0; j < EQ; j++) knQ[j] = LM(KM_RATE0 + qr[j]);
      if (!skip0) nt = du0;
      // ---- end of RHS

      if (s == 0) {
        // the re-evaluated k1 is committed at once: a member that needed it either steps now or has ended
        if (!skip0 && can_step) {
          kFy = kny;
          if (GWS) { if (gw_lane) SETM(KM_KFG, knG); } else kFG = knG;
#pragma unroll
          for (int j = 0; j < EQ; j++) kFQ[j] = knQ[j];
        }
        if (probe) {                                 // (uniform over the launch)
          if (live) {
            double* o = d.probe_out + m * (N + 9);
            o[i] = kny;
            if (top) { o[N] = knG; for (int e = 0; e < 8; e++) o[N + 1 + e] = LM(KM_RATE0 + e); }
          }
          break;
        }
      } else {
        const int ks = s + 1;
        aEy = aEy + lBT[ks] * kny;
        if (GWS) {
          if (gw_lane) {
            SETM(KM_AEG, LM(KM_AEG) + lBT[ks] * knG);
            if (ks <= 6) SETM(KM_SBG, LM(KM_SBG) + lB7[ks] * knG);
            else SETM(KM_KCG, knG);                  // k7 = the FSAL derivative of the next step
          }
        } else {
          aEG = aEG + lBT[ks] * knG;
          if (ks <= 6) SbG = SbG + lB7[ks] * knG;
        }
#pragma unroll
        for (int p = 0; p < 5; p++)
          if (ks + 1 + p <= 7) qy[p] = qy[p] + lCQ[ks][p] * kny;
        if (gw_feedback) {
#pragma unroll
          for (int p = 0; p < 4; p++)
            if (ks + 1 + p <= 6) qG[p] = qG[p] + lCQ[ks][p] * knG;
        }
#pragma unroll
        for (int j = 0; j < EQ; j++) {
          if (ks <= 6) Sb[j] = Sb[j] + lB7[ks] * knQ[j];
          Sbt[j] = Sbt[j] + lBT[ks] * knQ[j];
          kcQ[j] = knQ[j];
        }
        kcy = kny;
        if (!GWS) kcG = knG;
        if (s == 6) xy = sw;          // the new state u: SWCHEK clamped it in place
      }
    }
    if (probe) break;
    // residuals + norm_to_use: sequential chain SWATI[0..N-1], GWAT, 8 accumulators (oracle order)
    double Qn[EQ], r2Q[EQ];
    double ss = 0;
    {
      const double abstol = d.abstol, reltol = d.reltol;
      double r = div_z(dt_s * aEy, abstol + jmax_abs(y, xy) * reltol);
      const double r2y = r * r;
      if (GWS) {                                     // the top lane's sums, the old GWAT; the new one as stage 7 would have formed it
        __syncwarp(LFULL);
        G = LM(KM_G); aEG = LM(KM_AEG);
        xG = G + dt_s * LM(KM_SBG);
      }
      r = div_z(dt_s * aEG, abstol + jmax_abs(G, xG) * reltol);
      const double r2G = r * r;
#pragma unroll
      for (int j = 0; j < EQ; j++) {
        Qn[j] = Q[j] + dt_s * Sb[j];
        r = div_z(dt_s * Sbt[j], abstol + jmax_abs(Q[j], Qn[j]) * reltol);
        r2Q[j] = r * r;
      }
      if (NMAX <= 8) {
        double rj[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; j++) rj[j] = bcast(r2y, j < N ? j : 0);
#pragma unroll
        for (int j = 0; j < NMAX; j++) if (j < N) ss += rj[j];
      } else {
        for (int j = 0; j < N; j++) ss += bcast(r2y, j);
      }
      ss += r2G;
#pragma unroll 1
      for (int e = 0; e < 8; e++) ss += q_bcast(r2Q, e);
    }
    const double EEst = sqrt(ss / (double)(3 * N + 39));
    // loopfooter!: PI controller.  (the member's bookkeeping: read into locals, then - all lanes have read - written)
    const double qold = LM(KM_QOLD), iter0 = LM(KM_ITER), nacc0 = LM(KM_NACCEPT), nrej0 = LM(KM_NREJECT), nf0 = LM(KM_NF);
    __syncwarp(LFULL);
    double q, q11 = 1.0;
    if (EEst == 0.0) q = 1 / qmax;
    else {
      q11 = fastpow(EEst, beta1);
      q = q11 / fastpow(qold, beta2);
      q = jmax(1 / qmax, jmin(1 / qmin, q / gamma));
    }
    const bool acc_s = (EEst <= 1.0);

    // ---- commit
    if (can_step) {
      SETM(KM_ITER, iter0 + 1.0);
      if (st_s != LWFB_ST_OK) {
        SETM(KM_STATUS, (double)st_s);
        // the reference re-evaluates f right after a modifying callback, i.e. before the check that stopped the member;
        // here that evaluation is stage 0 of the step that is not taken: count it
        if (reeval) { SETM(KM_NF, nf0 + 1.0); reeval = false; }
      }
    }
    const bool accepted = do_step && acc_s;
    double dtt = 0;
    if (do_step) {
      SETM(KM_NF, nf0 + (double)nf_s); ntfli = nt;
      reeval = false;
      cand_nan = (xy != xy);
      if (acc_s) {
        SETM(KM_NACCEPT, nacc0 + 1.0);
        y = xy; G = xG;
        kFy = kcy;                                  // FSAL
        if (GWS) { if (gw_lane) { SETM(KM_G, xG); SETM(KM_KFG, LM(KM_KCG)); } } else kFG = kcG;
#pragma unroll
        for (int j = 0; j < EQ; j++) { Q[j] = Qn[j]; kFQ[j] = kcQ[j]; }
        SETM(KM_QOLD, jmax(EEst, qoldinit));
        const double dtnew = dt_s / q;
        const double ttmp = t + dt_s;
        const double tn = (fabs(ttmp - tstop) < 100 * eps_of(fmax(t, tstop))) ? tstop : ttmp;
        dtt = tn - t;
        t = tn;
        double dt = d.dtmax < dtnew ? d.dtmax : dtnew;
        { const double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dt < dtmin_t) dt = dtmin_t; }
        SETM(KM_DT, dt);
      } else {
        SETM(KM_NREJECT, nrej0 + 1.0);
        SETM(KM_DT, dt_s / jmin(1 / qmin, q11 / gamma));
      }
    }
    const bool at_stop = accepted && (t == tstop);
    // ---- record + balance at the day boundary (saved before the callbacks and before this step's isotope update)
    if (__any_sync(LFULL, at_stop)) {
      const int rec = (int)LM(KM_REC);              // (read before the broadcasts below, written after them)
      double Ssum = 0;
      for (int j = 0; j < N; j++) Ssum += bcast(y, j);
      double ev[8];
#pragma unroll 1
      for (int e = 0; e < 8; e++) ev[e] = q_bcast(Q, e);
      if (at_stop) flush_state();
#ifdef LWFB_DEBUG_ITERS
      if (at_stop && live && top) { const unsigned it_ = (unsigned)LM(KM_ITER); FIN(ix.AUX) = (double)(it_ - dbg_day_iter0); dbg_day_iter0 = it_; }   // attempted steps of this day
#endif
      __syncwarp(LFULL);
      double bs = 0, bt = 0, Wt = 0;
      score_rec(rec, at_stop, y, LC(KC_DS_18O));
      if (at_stop) {
        if (d.rec && live) {
          // 8 L contiguous bytes per member-day; the lanes of the group write neighbouring words, L2 merges the sectors
          double* __restrict__ dst = rec_m + (int64_t)rec * L;
          // streaming stores (evict-first): the record is written once and must not push the blocks' local memory and
          // the members' live state out of L2
          for (int j = i; j < L; j += N) __stcs(dst + j, FIN(j));
        }
        SETM(KM_REC, (double)(rec + 1));
        if (t < tend) {
          // LWFBrook90R_check_balance_errors!
          Wt = FIN(ix.INTR) + FIN(ix.INTS) + FIN(ix.SNOW) + Ssum + G;
          const double qin = 0 + (ev[3] - ev[4]);
          const double qout = ev[7] + ev[5] + FIN(ix.ACC + LWFB_A_CUM_D_TRAN) + FIN(ix.ACC + LWFB_A_CUM_D_SLVP);
          const double qin_t = FIN(ix.ACC + LWFB_A_CUM_D_PREC);
          const double qout_t = FIN(ix.ACC + LWFB_A_CUM_D_EVAP) + ev[0] + ev[1];
          bs = (qin - qout) - (Ssum - FIN(ix.ACC + LWFB_A_STORAGESWAT));
          bt = (qin_t - qout_t) - (Wt - FIN(ix.ACC + LWFB_A_STORAGEWATER));
          at_boundary = true;
        }
      }
      __syncwarp(LFULL);                            // all reads of the old balance entries done
      if (at_stop && t < tend && top && live) {
        FIN(ix.ACC + LWFB_A_BALERD_SWAT) = bs; FIN(ix.ACC + LWFB_A_BALERD_TOTAL) = bt;
        FIN(ix.ACC + LWFB_A_STORAGESWAT) = Ssum; FIN(ix.ACC + LWFB_A_STORAGEWATER) = Wt;
      }
      __syncwarp(LFULL);
    }
    // ---- isotope advection-dispersion step with the flux caches of stage 7
    const bool do_ade = accepted && iso && (t != tstop || t < tend);
    if (__any_sync(LFULL, do_ade)) {
      double dp[2];
      staged_wait();
      prec_delta(t, dp);
      const double THR = LC(KC_THR), DTH = LC(KC_DTH);
      const double S1 = y;
      const double w1 = jminc(div_c(div_c(LC(KC_THSAT) * S1, LC(KC_SWATMAX), LC(KC_RSWATMAX)) - THR, DTH, LC(KC_RDTH)), 1.0);
      const double th1 = w1 * DTH + THR;
      const double sumT = LM(KM_SUMT), sumT_V = LM(KM_SUMT_V);
      const double TRANI = LC(KC_TRANI), SLVP = LM(KM_SLVP);
      const double tauth = lwfb_pow_t(th1, 7.0 / 3, ltab, etab) * LC(KC_RTHS2);
      const double aq = LM(KM_LAM) * fabs(VV * LWFB_RCP(1000.0));
      const double VN = bcast(VV, N - 1);
      const double VV_up = __shfl_up_sync(LFULL, VV, 1);
      const double RDZ_up = LC(KC_RDZ_UP), RDZ_dn = LC(KC_RDZ_DN);
      const double rS1 = 1 / S1, VN_G = div_z(VN, G);
      // the two species share one copy of the code (rolled loop, state picked by select)
#ifdef LWFB_ADE_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int s = 0; s < 2; s++) {
        const double R = s == 0 ? R18 : R2H;
        const double D0 = LM(KM_D0_18O + s);
        const double dSs = LC(KC_DS_18O + s), dGs = LM(KM_DG_18O + s), dXYs = LM(KM_DXY_18O + s);
        const double dps = s == 0 ? dp[0] : dp[1];
        // the three member-level mixing ratios (xylem, precipitation, groundwater) are converted by lanes 0, 1, 2 of the
        // group in ONE evaluation of d_to_x and broadcast
        const double X = d_to_x(i == 1 ? dps : (i == 2 ? dGs : dXYs), R);
        const double Cxy = bcast(X, 0), Cinf = bcast(X, 1), CG = bcast(X, 2);
        const double Dd = D0 * tauth * th1 + aq;
        const double C = d_to_x(dSs, R);
        const double Ctr = (TRANI < 0) ? Cxy : C;
        const double C_up = __shfl_up_sync(LFULL, C, 1), C_dn = __shfl_down_sync(LFULL, C, 1);
        const double Dd_up = __shfl_up_sync(LFULL, Dd, 1), Dd_dn = __shfl_down_sync(LFULL, Dd, 1);
        const double diff_upp = top ? 0.0 : ((Dd_up + Dd) / 2) * ((C - C_up) * RDZ_up);
        const double diff_low = bot ? 0.0 : ((Dd + Dd_dn) / 2) * ((C_dn - C) * RDZ_dn);
        const double qC_upp = top ? 0.0 : VV_up * C_up;
        const double qC_low = VV * C;
        const double slv = top ? C : 0.0;
        double innr = -diff_upp * 1000 + diff_low * 1000 + qC_upp;
        innr = innr - qC_low;
        innr = innr + INFLI * Cinf;
        innr = innr - DSFLI * C;
        innr = innr - SLVP * slv;
        innr = innr - TRANI * Ctr;
        const double dC = -C * rS1 * ntfli + rS1 * innr;
        const double dd = dxdt_to_dddt(dC, C, R);
        const double CN = bcast(C, N - 1);
        double ddG = dxdt_to_dddt(VN_G * (CN - CG), CG, R);
        if (G == 0) ddG = (VN == 0) ? NAN : ((VN > 0) ? 0.0 : ddG);
        double swt = 0;
        const double ctt = Ctr * TRANI;
        for (int j = 0; j < N; j++) swt += bcast(ctt, j);
        // 0/0 (no uptake at all) is NaN by definition; answered without sending the divisions down their slow path
        const bool nouptake = (swt == 0.0) && (sumT == 0.0);
        const double dRWc = x_to_d((nouptake ? 1.0 : swt) / (nouptake ? 2.0 : sumT), R);
        const double dRW = nouptake ? NAN : dRWc;
        const double ddxy = (dRW != dRW) ? 0.0 : sumT_V * (dRW - dXYs);
        if (do_ade) {
          SETM(KM_DG_18O + s, dGs + dtt * ddG); SETK(KC_DS_18O + s, dSs + dtt * dd); SETM(KM_DXY_18O + s, dXYs + dtt * ddxy);
          if (top && live) FIN(ix.RWU + 1 + s) = dRW;
        }
      }
    }
  }
  __syncwarp(LFULL);
  const int status = (int)LM(KM_STATUS);
  int rec = (int)LM(KM_REC);
  if (live) {
    if (d.ck) {
      // checkpoint for the next launch: everything the loop keeps in registers / shared memory and flush_state does
      // not put into fin (controller, counters, clock, FSAL derivatives, day constants)
      CKL(LWFB_CKL_NTFLI) = ntfli; CKL(LWFB_CKL_TRANI) = LC(KC_TRANI); CKL(LWFB_CKL_KFY) = kFy; CKL(LWFB_CKL_NAN) = cand_nan ? 1.0 : 0.0;
#pragma unroll
      for (int j = 0; j < EQ; j++) CKL(LWFB_CKL_KFQ + j) = kFQ[j];
      if (top) {
        CKP(LWFB_CK_DT) = LM(KM_DT); CKP(LWFB_CK_QOLD) = LM(KM_QOLD); CKP(LWFB_CK_ITER) = LM(KM_ITER);
        CKP(LWFB_CK_NACCEPT) = LM(KM_NACCEPT); CKP(LWFB_CK_NREJECT) = LM(KM_NREJECT); CKP(LWFB_CK_NF) = LM(KM_NF);
        CKP(LWFB_CK_REC) = (double)rec; CKP(LWFB_CK_T) = t; CKP(LWFB_CK_TSTOP) = LM(KM_TSTOP);
        CKP(LWFB_CK_FLAGS) = (double)((reeval ? 1 : 0) | (at_boundary ? 2 : 0));
        CKP(LWFB_CK_RS) = LM(KM_RS); CKP(LWFB_CK_SLVP) = LM(KM_SLVP); CKP(LWFB_CK_KFG) = GWS ? LM(KM_KFG) : kFG;
        CKP(LWFB_CK_SUMT) = LM(KM_SUMT); CKP(LWFB_CK_SUMT_V) = LM(KM_SUMT_V);
        CKP(LWFB_CK_D0_18O) = LM(KM_D0_18O); CKP(LWFB_CK_D0_2H) = LM(KM_D0_2H);
      }
    }
    if (d.n_obs_layers > 0 && top && status != LWFB_ST_OK && status_in == LWFB_ST_OK) {
      // ... and so are its simulated values at the observation days it did not reach
      double* a = d.sacc + m * (2 * LWFB_SA);
      for (int r = rec; r <= d.n_days; r++) {
        const int ko = __ldg(d.obs_of_rec + r);
        if (ko < 0) continue;
        for (int l = 0; l < d.n_obs_layers; l++) {
          if (d.obs_theta) score_add(a, NAN, __ldg(d.obs_theta + (int64_t)ko * d.n_obs_layers + l));
          if (d.obs_d18o && iso) score_add(a + LWFB_SA, NAN, __ldg(d.obs_d18o + (int64_t)ko * d.n_obs_layers + l));
        }
      }
    }
    if (d.rec && status != LWFB_ST_OK && status_in == LWFB_ST_OK) {
      // the member ended in this launch before the end of the record: the records it did not reach are NaN
      for (; rec <= d.n_days; rec++) {
        double* dst = rec_m + (int64_t)rec * L;
        for (int j = i; j < L; j += N) __stcs(dst + j, (double)NAN);
      }
    }
    flush_state();
    if (d.ntfli_out) d.ntfli_out[m * N + i] = ntfli;
    if (top) {
      d.status[m] = status;
      d.counters[m * LWFB_NCOUNTERS + LWFB_C_NACCEPT] = (long long)LM(KM_NACCEPT);
#ifdef LWFB_DEBUG_ITERS
      d.counters[m * LWFB_NCOUNTERS + LWFB_C_NREJECT] = dbg_iters;
#else
      d.counters[m * LWFB_NCOUNTERS + LWFB_C_NREJECT] = (long long)LM(KM_NREJECT);
#endif
      d.counters[m * LWFB_NCOUNTERS + LWFB_C_NF] = (long long)LM(KM_NF);
    }
  }
  return by_budget;
}

#ifdef __CUDACC__
// MINB = resident blocks per SM the register allocation is sized for.  LWFB_LANES_MINBLOCKS (2: 128 registers, 16 warps
// per SM) is the throughput build.  MINB = 1 (255 registers, hardly any spills) is the LATENCY build for what is left
// when the unfinished members no longer fill the device - the critical path of the slowest members, where a pass of a
// block is a chain of dependent instructions and every spill reload is waited for: BEA2016 10000 members 6.36 -> 5.23 s
// (profiles/r02_notes.md), while a full device runs 26 % slower with it.
template <int NMAX, int EQ, int NX = 0, bool LEAN = false, int MINB = LWFB_LANES_MINBLOCKS>
__global__ void __launch_bounds__(LWFB_LANES_BLOCK, MINB) integrate_lanes(DevData d) {
  extern __shared__ double lanes_smem[];
  double* ltab = lanes_smem + (size_t)(LEAN ? KC_COUNT_LEAN : KC_COUNT) * LWFB_LANES_BLOCK;     // shared-memory copy of the log table
  double* etab = ltab + 3 * LWFB_LOG_TAB_N;                            // ... and of the exp table
  double* msm_block = etab + 2 * LWFB_EXP_TAB_N;                       // per-member slots, then the staged forcing rows
  for (int k = threadIdx.x; k < 3 * LWFB_LOG_TAB_N; k += LWFB_LANES_BLOCK) ltab[k] = lwfb_log_tab_d[k];
  for (int k = threadIdx.x; k < 2 * LWFB_EXP_TAB_N; k += LWFB_LANES_BLOCK) etab[k] = lwfb_exp_tab_d[k];
  __syncthreads();
  unsigned tid = threadIdx.x;
  asm volatile("" : "+r"(tid));     // opaque copy of %tid: the column address stays in a register (it would be re-derived
                                    // from %tid at every use otherwise) and is still known to point to shared memory
  const bool by_budget =
      lanes_body<NMAX, EQ, NX, LEAN>(d, threadIdx.x & 31, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, lanes_smem + tid, ltab,
                                     etab, msm_block);
  if (d.stop) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long n_exited = atomicAdd(d.stop, 1) + 1;
      // this block used up its quantum and no block is waiting for its slot: end the launch for everybody
      if (by_budget && n_exited + d.resident > (long long)gridDim.x) atomicExch(d.stop + 1, 1);
    }
  }
}
#endif

}  // namespace lwfb
