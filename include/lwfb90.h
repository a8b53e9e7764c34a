/*
 * lwfb90.h — C ABI of liblwfb90 (B200-native batched integrator for the LWFBrook90.jl
 * soil-water / isotope ODE).
 *
 * This is the drop-in boundary for ONE path of the reference: what happens inside
 *     simulate!(s::DiscretizedSPAC)  ->  solve(s.ODEProblem; ...)
 * (reference: src/LWFBrook90.jl:628-655 calling the problem built at src/LWFBrook90.jl:552-567).
 * The reference hands OrdinaryDiffEq an (f!, u0, tspan, p, callbacks) bundle and gets an
 * ODESolution back.  Here the model is compiled into the library, so no function pointer
 * crosses the ABI: the caller flattens `p` and `u0` of each ensemble member into the blocks
 * described below, and receives the daily records (`saveat = t0:1:t1`) in the reference's own
 * state-vector order (src/func_DiffEq_definition_u0.jl:42-73), so that a ComponentVector view can
 * be laid over every record.
 *
 * All pointers are caller-owned host memory, row-major, IEEE double unless noted.
 * No function throws; every entry point returns LWFB_OK (0) or a negative error code and
 * lwfb_last_error() gives a message.  A handle is not thread-safe.
 *
 * The enums below ARE the layout contract.  python (lwfbrook90.jl_b200/layout.py) parses this
 * file; the CPU oracle (oracle/lwfb90_oracle.c) and the CUDA kernels include it.
 */
#ifndef LWFB90_H
#define LWFB90_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- error codes */
enum {
  LWFB_OK = 0,
  LWFB_ERR_ARG = -1,      /* invalid argument / call order                       */
  LWFB_ERR_CUDA = -2,     /* CUDA runtime error, see lwfb_last_error             */
  LWFB_ERR_NOMEM = -3,    /* host or device allocation failed                    */
  LWFB_ERR_UNSUPPORTED = -4 /* e.g. n_layers above the compiled maximum          */
};

/* per-member status word (replaces the reference's retcode / exceptions,
 * src/LWFBrook90.jl:648-650, src/func_DiffEq_definition_ode.jl:7-12,
 * src/func_MSB_functions.jl:114-119) */
enum {
  LWFB_ST_OK = 0,
  LWFB_ST_UNSTABLE = 1,   /* NaN in SWATI.mm  (unstable_check)                    */
  LWFB_ST_DTMIN = 2,      /* dt fell below dtmin                                  */
  LWFB_ST_MAXITERS = 3,   /* more than opts.maxiters attempted steps              */
  LWFB_ST_RSS = 4,        /* soil evaporation resistance < 1e-6 (reference error) */
  LWFB_ST_INPUT = 5       /* invalid member input                                 */
};

/* ---------------------------------------------------------------- forcing table
 * One table per site, one row per day d = 0..n_rows-1 (n_rows >= t0_day + n_days + 1;
 * piecewise constant in time as the reference's Constant{Previous} interpolants,
 * src/func_DiffEq_definition_p.jl:715-724, :850-855).  Row d holds what the reference
 * evaluates at integer t = d. */
enum {
  LWFB_F_GLOBRAD = 0,   /* MJ/m2/d                                               */
  LWFB_F_TMAX,          /* degC                                                  */
  LWFB_F_TMIN,          /* degC                                                  */
  LWFB_F_VAPPRES,       /* kPa                                                   */
  LWFB_F_WIND,          /* m/s                                                   */
  LWFB_F_PREC,          /* mm/d                                                  */
  LWFB_F_DENSEF_REL,    /* DENSEF_rel/100 (multiplied by member DENSEF_BASELINE) */
  LWFB_F_HEIGHT_REL,    /* HEIGHT_rel/100                                        */
  LWFB_F_LAI_REL,       /* LAI_rel/100                                           */
  LWFB_F_SAI_REL,       /* SAI_rel/100                                           */
  LWFB_F_DOY,           /* day of year (1..366) of reference_date + d            */
  LWFB_F_IHD,           /* floor((DURATN[month]+0.1)/2), src/module_EVP.jl:645   */
  LWFB_F_D18O,          /* delta18O of precipitation at t = d  (permil)          */
  LWFB_F_D2H,           /* delta2H  of precipitation at t = d  (permil)          */
  LWFB_F_PAD0,
  LWFB_F_PAD1,
  LWFB_F                /* = 16 doubles = 128 B per site-day                     */
};

/* ---------------------------------------------------------------- per-member scalars
 * Flattened `p` NamedTuple, src/func_DiffEq_definition_p.jl:421-498. */
enum {
  LWFB_S_DPSIMAX = 0, LWFB_S_DSWMAX, LWFB_S_DTIMAX,
  LWFB_S_VXYLEM, LWFB_S_DISPERSIVITY,            /* mm, m                          */
  LWFB_S_LAT, LWFB_S_ESLOPE, LWFB_S_L1, LWFB_S_L2, /* radians                      */
  LWFB_S_MELFAC, LWFB_S_RSTEMP,
  LWFB_S_FETCH, LWFB_S_WNDRAT, LWFB_S_Z0G, LWFB_S_Z0S, LWFB_S_Z0W, LWFB_S_ZMINH, LWFB_S_ZW,
  LWFB_S_HS, LWFB_S_HR, LWFB_S_CZS, LWFB_S_CZR, LWFB_S_C1, LWFB_S_C2, LWFB_S_C3,
  LWFB_S_SNODEN, LWFB_S_CCFAC, LWFB_S_GRDMLT, LWFB_S_LAIMLT, LWFB_S_SAIMLT, LWFB_S_MAXLQF,
  LWFB_S_CR, LWFB_S_LWIDTH, LWFB_S_RHOTP, LWFB_S_LPC, LWFB_S_KSNVP, LWFB_S_ALBSN, LWFB_S_ALB,
  LWFB_S_NN, LWFB_S_RM,
  LWFB_S_CINTRL, LWFB_S_CINTRS, LWFB_S_CINTSL, LWFB_S_CINTSS,
  LWFB_S_FRINTL, LWFB_S_FRINTS, LWFB_S_FSINTL, LWFB_S_FSINTS,
  LWFB_S_MXKPL, LWFB_S_FXYLEM, LWFB_S_GLMAX, LWFB_S_GLMIN,
  LWFB_S_TH, LWFB_S_T1, LWFB_S_T2, LWFB_S_TL, LWFB_S_PSICR, LWFB_S_CVPD, LWFB_S_R5,
  LWFB_S_NOOUTF, LWFB_S_RTRAD, LWFB_S_MXRTLN,
  LWFB_S_RSSA, LWFB_S_RSSB,
  LWFB_S_IMPERV, LWFB_S_QLAYER, LWFB_S_BYPAR, LWFB_S_DRAIN, LWFB_S_DSLOPE,
  LWFB_S_GSC, LWFB_S_GSP, LWFB_S_LENGTH_SLOPE, LWFB_S_QFFC, LWFB_S_QFPAR,
  LWFB_S_SWATQX, LWFB_S_SWATQF,
  /* canopy baselines, src/func_DiffEq_definition_p.jl:816-833 */
  LWFB_S_MAXLAI, LWFB_S_SAI_BASELINE, LWFB_S_DENSEF_BASELINE, LWFB_S_HEIGHT_BASELINE,
  /* root growth in time, src/func_discretize_soil_domain.jl:550-645,
   * src/func_DiffEq_definition_p.jl:606-628 */
  LWFB_S_AGE0, LWFB_S_RGROPER, LWFB_S_RL0, LWFB_S_RELDEN_CONST,
  LWFB_S_PAD0,
  LWFB_PS               /* number of scalars per member                          */
};

/* ---------------------------------------------------------------- per-member per-layer vectors
 * KPT_SOILPAR_Mvg1d (src/module_KPT.jl:310-330) + INFRAC + root density. */
enum {
  LWFB_L_THICK = 0,     /* mm                                                    */
  LWFB_L_STONEF, LWFB_L_THSAT, LWFB_L_KSAT, LWFB_L_ALPHA, LWFB_L_NPAR, LWFB_L_MPAR,
  LWFB_L_TORT,          /* MvG l                                                 */
  LWFB_L_THR,           /* residual water content                                */
  LWFB_L_PSIF, LWFB_L_THETAF, LWFB_L_PSIG, LWFB_L_SWATMAX, LWFB_L_WETF,
  LWFB_L_INFRAC,
  LWFB_L_RELDEN,        /* final relative root density                           */
  LWFB_L_TINI,          /* stand age at which roots reach the layer (1e20 never) */
  LWFB_L_DZ,            /* centre-to-centre distance to next layer, m (N-1 used) */
  LWFB_PL
};

/* ---------------------------------------------------------------- accumulator names
 * order of `accum` in the state vector, src/func_DiffEq_definition_u0.jl:10-15 */
enum {
  LWFB_A_CUM_D_PREC = 0, LWFB_A_CUM_D_RFAL, LWFB_A_CUM_D_SFAL, LWFB_A_CUM_D_RINT,
  LWFB_A_CUM_D_SINT, LWFB_A_CUM_D_RSNO, LWFB_A_CUM_D_RNET, LWFB_A_CUM_D_SMLT,
  LWFB_A_CUM_D_EVAP, LWFB_A_CUM_D_TRAN, LWFB_A_CUM_D_IRVP, LWFB_A_CUM_D_ISVP,
  LWFB_A_CUM_D_SLVP, LWFB_A_CUM_D_SNVP, LWFB_A_CUM_D_PINT, LWFB_A_CUM_D_PTRAN,
  LWFB_A_CUM_D_PSLVP,
  LWFB_A_FLOW, LWFB_A_SEEP, LWFB_A_SRFL, LWFB_A_SLFL, LWFB_A_BYFL, LWFB_A_DSFL,
  LWFB_A_GWFL, LWFB_A_VRFLN,
  LWFB_A_CUM_D_RTHR, LWFB_A_CUM_D_STHR,
  LWFB_A_STORAGESWAT, LWFB_A_STORAGEWATER, LWFB_A_BALERD_SWAT, LWFB_A_BALERD_TOTAL,
  LWFB_NACCUM           /* = 31                                                  */
};

/* State vector length and offsets (memory order = ComponentArray order,
 * src/func_DiffEq_definition_u0.jl:42-73).  With isotopes every (mm,d18O,d2H) triple is
 * stored as three consecutive blocks; without isotopes only the first block exists.
 *   iso:    SWATI[3N] GWAT[3] INTS[3] INTR[3] SNOW[3] RWU[3] XYLEM[3] CC SNOWLQ TRANI[3N] aux[3N] accum[31]
 *   no iso: SWATI[N]  GWAT    INTS    INTR    SNOW    RWU    XYLEM    CC SNOWLQ TRANI[N]  aux[3N] accum[31] */
#define LWFB_STATE_LEN(N, iso) ((iso) ? (9 * (N) + 51) : (5 * (N) + 39))

/* counters returned per member */
enum { LWFB_C_NACCEPT = 0, LWFB_C_NREJECT, LWFB_C_NF, LWFB_NCOUNTERS };

/* per-member goodness-of-fit scores (lwfb_scores) */
enum { LWFB_SC_RMSE_THETA = 0, LWFB_SC_KGE_THETA, LWFB_SC_RMSE_D18O, LWFB_SC_KGE_D18O, LWFB_NSCORES };

/* columns of the water partitioning, get_water_partitioning (src/func_postprocess.jl:1310-1364): sums per period of
 * ETa, Esoil, Esnow, Einterception, Ta, Precip, Td, D, R (mm) and the period mean of Swat (mm) */
enum { LWFB_WP_ETA = 0, LWFB_WP_ESOIL, LWFB_WP_ESNOW, LWFB_WP_EINTERCEPTION, LWFB_WP_TA, LWFB_WP_PRECIP, LWFB_WP_TD,
       LWFB_WP_D, LWFB_WP_R, LWFB_WP_SWAT, LWFB_NWP };

/* ---------------------------------------------------------------- options
 * Defaults mirror the solve() keyword arguments fixed in setup(),
 * src/LWFBrook90.jl:552-567: Tsit5, reltol 1e-5, abstol 1e-6 (default), dt 1e-3, dtmax 1/24 d. */
typedef struct lwfb_opts {
  int32_t n_layers;           /* N, 1..LWFB_MAX_LAYERS                               */
  int32_t simulate_isotopes;  /* 0/1                                                  */
  int32_t t0_day;             /* first simulated day = forcing row of tspan[1]        */
  int32_t n_days;             /* tspan = (t0_day, t0_day + n_days)                    */
  double reltol;              /* 1e-5                                                 */
  double abstol;              /* 1e-6                                                 */
  double dt_init;             /* 1e-3 d                                               */
  double dtmax;               /* 60/60/24 d                                           */
  int64_t maxiters;           /* attempted steps per member before LWFB_ST_MAXITERS   */
  int32_t device;             /* CUDA device ordinal                                  */
  int32_t record_daily;       /* 1: keep the full daily record in HBM; 0: scores only */
} lwfb_opts;

/* Layer counts: 3..32 run in the lane-per-layer kernel, 33..LWFB_MAX_LAYERS in the warp-per-member kernel (a member's
 * per-layer arrays live in shared memory; the reference's integration tests go up to N = 400,
 * test/02-integration-tests.jl:383), 1..2 (and, for testing, up to LWFB_MEMBER_MAX_LAYERS) in the thread-per-member kernel. */
#define LWFB_MAX_LAYERS 512
#define LWFB_MEMBER_MAX_LAYERS 64

typedef struct lwfb_handle lwfb_handle;

/* fills *o with the reference defaults (src/LWFBrook90.jl:552-567) */
void lwfb_default_opts(lwfb_opts* o);

/* replaces ODEProblem(...) construction, src/LWFBrook90.jl:552 */
int lwfb_create(lwfb_handle** h, const lwfb_opts* opts);
void lwfb_destroy(lwfb_handle* h);
const char* lwfb_last_error(const lwfb_handle* h);

/* replaces the forcing closures inside `p` (p_GLOBRAD ... p_δ2H_PREC, p_DOY, p_MONTHN),
 * src/func_DiffEq_definition_p.jl:482-497.  forcing: [n_sites][n_rows][LWFB_F]. */
int lwfb_set_sites(lwfb_handle* h, int32_t n_sites, int32_t n_rows, const double* forcing);

/* replaces (u0, p) of each ODEProblem, src/LWFBrook90.jl:489-503.
 *   site_of_member [n_members]                    index into the site tables
 *   scal           [n_members][LWFB_PS]
 *   lay            [n_members][LWFB_PL][N]
 *   u0             [n_members][LWFB_STATE_LEN(N,iso)]   reference state order */
int lwfb_set_members(lwfb_handle* h, int64_t n_members, const int32_t* site_of_member,
                     const double* scal, const double* lay, const double* u0);

/* replaces remakeSPAC -> setup of every member (src/LWFBrook90.jl:214-345, :365-587: define_LWFB90_p, the
 * KPT_SOILPAR_Mvg1d constructor with its bisection, init_LWFB90_u0!) ON THE DEVICE: the member blocks are built in HBM
 * from ONE base member (the blocks lwfb_set_members would take for the unmodified site definition) and a row of knob
 * values per member, so only the knobs cross the bus.  Layouts and arithmetic: include/lwfb90_setup.h.
 *   site_of_member [n_members] or NULL (all members on site 0)
 *   knobs          [n_members][LWFB_K + n_over]   NaN = keep the base value
 *   over_idx       [n_over]   LWFB_S_* indices of the scalar columns that follow the LWFB_K fixed columns
 *   profiles       [n_profiles][2 N + 1]   root profiles (RELDEN[N], TINI[N], RELDEN_CONST) selected by LWFB_K_ROOTPROFILE
 *   psim_base      [N]   initial matric potential of the soil discretisation (kPa)
 *   ilayer         number of layers that receive infiltration (IDEPTH)
 *   n_invalid      (optional) members for which no field-capacity wetness exists (the reference throws for those)
 * Equivalent to lwfb_set_members with the blocks that lwfb_setup_member() produces on the host. */
int lwfb_setup_members(lwfb_handle* h, int64_t n_members, const int32_t* site_of_member,
                       const double* base_scal, const double* base_lay, const double* base_u0,
                       const double* knobs, int32_t n_over, const int32_t* over_idx,
                       int32_t n_profiles, const double* profiles, const double* psim_base, int32_t ilayer,
                       int32_t* n_invalid);

/* replaces solve(prob), src/LWFBrook90.jl:643.  Integrates every member over tspan with
 * Tsit5 + PI controller + daily / per-step callbacks (src/func_DiffEq_definition_cb.jl:11-57).
 * Synchronous; if cuda_stream != NULL it is the cudaStream_t to launch on (still synchronised
 * before returning unless lwfb_run_async is used). */
int lwfb_run(lwfb_handle* h, void* cuda_stream);
int lwfb_run_async(lwfb_handle* h, void* cuda_stream); /* no sync; pair with lwfb_sync */
int lwfb_sync(lwfb_handle* h);

/* replaces ODESolution.u at the saveat points (src/LWFBrook90.jl:561):
 * out [member_count][n_days+1][LWFB_STATE_LEN]; record k = state saved at t = t0_day + k. */
int lwfb_get_daily(lwfb_handle* h, int64_t member_first, int64_t member_count, double* out);

/* final state only (always available, also with record_daily = 0): out [member_count][LWFB_STATE_LEN] */
int lwfb_get_final(lwfb_handle* h, int64_t member_first, int64_t member_count, double* out);

/* replaces ODESolution.retcode / .destats (src/LWFBrook90.jl:645-650):
 * status [n_members] (LWFB_ST_*), counters [n_members][LWFB_NCOUNTERS] */
int lwfb_get_status(lwfb_handle* h, int32_t* status, int64_t* counters);

/* per-member goodness of fit against observations (the only reference analogue is
 * RMS_differences, test/fct-helpers-for-integration-tests.jl:6-18):
 *   obs_theta [n_obs_days][n_obs_layers], obs_d18o likewise; NaN = missing; either may be NULL
 *   obs_day   [n_obs_days]     record index k (0..n_days)
 *   obs_layer [n_obs_layers]   0-based layer index
 *   scores    [n_members][LWFB_NSCORES]
 * Two ways, bit-identical results:
 *   (a) from the resident daily record (record_daily = 1): pass the observations here, after the run;
 *   (b) without a record (record_daily = 0, e.g. 1 M members on one GPU): register the observations BEFORE the run
 *       with lwfb_set_observations - the integrator then accumulates the sums itself at the observation days -
 *       and call lwfb_scores / lwfb_scores_device with n_obs_days = 0 and NULL tables afterwards. */
int lwfb_scores(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day,
                int32_t n_obs_layers, const int32_t* obs_layer,
                const double* obs_theta, const double* obs_d18o, double* scores);

/* registers (n_obs_days > 0) or clears (n_obs_days = 0) the observations scored during the following runs;
 * obs_day must be strictly increasing.  See lwfb_scores. */
int lwfb_set_observations(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day,
                          int32_t n_obs_layers, const int32_t* obs_layer,
                          const double* obs_theta, const double* obs_d18o);

/* device-resident views for callers that stay on the GPU (bench, NCCL all-gather):
 * pointer to scores [n_members][LWFB_NSCORES] after lwfb_scores_device. */
int lwfb_scores_device(lwfb_handle* h, int32_t n_obs_days, const int32_t* obs_day,
                       int32_t n_obs_layers, const int32_t* obs_layer,
                       const double* obs_theta, const double* obs_d18o, double** d_scores);

/* replaces get_water_partitioning (src/func_postprocess.jl:1310-1364) for the whole ensemble, on the device: the daily
 * record stays in HBM, only the aggregates come back.
 *   group_of_record [n_days+1]   period index (0..n_groups-1) of each daily record, < 0 = record not used; the
 *                                records of a period are expected to be contiguous (months, years)
 *   out [n_members][n_groups][LWFB_NWP]   sums in record order (Swat: mean over the records of the period) */
int lwfb_water_partitioning(lwfb_handle* h, int32_t n_groups, const int32_t* group_of_record, double* out);

/* soil variables of get_soil_ (src/func_postprocess.jl:1025-1045; derive_auxiliary_SOILVAR, src/module_KPT.jl:432-443) */
enum { LWFB_SV_THETA = 0, LWFB_SV_PSI /* kPa */, LWFB_SV_W, LWFB_SV_SWATI /* mm */, LWFB_SV_K /* mm/day */,
       LWFB_SV_D18O, LWFB_SV_D2H /* permil, NaN without isotopes */, LWFB_NSV };

/* replaces get_soil_ (src/func_postprocess.jl:1025-1045) for the whole ensemble, on the device:
 *   rec_idx   [n_rec_sel]   daily records to read (0..n_days)
 *   layer_idx [n_lay_sel]   0-based layers (the depth -> layer lookup, get_soil_idx, stays on the host)
 *   out [n_members][n_rec_sel][n_lay_sel][LWFB_NSV] */
int lwfb_soil_state(lwfb_handle* h, int32_t n_rec_sel, const int32_t* rec_idx, int32_t n_lay_sel,
                    const int32_t* layer_idx, double* out);

/* du_NTFLI across runs.  The reference keeps the net layer flows of the last RHS evaluation in `p` (du_NTFLI), and `p`
 * survives simulate!: a second run on the same problem starts its first ITER from them instead of NaN (SURVEY A.5 #19;
 * test/01-unit-tests.jl:875, :904 produce golden G5 that way).  lwfb_get_ntfli returns them after a run
 * (out [n_members][N]); lwfb_set_ntfli0 hands them to the following runs (NULL = fresh problem; setting the members
 * anew also clears them). */
int lwfb_set_ntfli0(lwfb_handle* h, const double* ntfli0);
int lwfb_get_ntfli(lwfb_handle* h, double* out);

/* one evaluation of the right-hand side f(u,p,t) for every member, for unit parity tests
 * (replaces f_LWFBrook90!, src/func_DiffEq_definition_f.jl:7-243):
 *   swati [n_members][N], gwat [n_members], t scalar, daily-frozen fluxes as set by the
 *   daily callback are taken from `frozen` [n_members][N+3] = (RNET, SMLT, SLVP, TRANI[N]);
 *   ntfli_prev [n_members][N] is the stale du_NTFLI consumed by ITER (NaN on first call).
 *   du [n_members][N+9] = (dSWATI[N], dGWAT, flow, seep, srfl, slfl, byfl, dsfl, gwfl, vrfln)
 * Evaluated by the kernel that would integrate the members: for 3 <= N <= 32 the lane-per-layer production kernel
 * (through its own stage code), else the thread-per-member kernel (lwfb_debug_force_legacy selects explicitly). */
int lwfb_eval_rhs(lwfb_handle* h, double t, const double* swati, const double* gwat,
                  const double* frozen, const double* ntfli_prev, double* du);

/* elementary functions of include/lwfb90_elem.h evaluated on the GPU (fn: 0 exp, 1 log, 2 pow, 3 sin,
 * 4 cos, 5 tan, 6 asin, 7 acos, 8 DiffEqBase-style fastpow) - lets a host check device/host bit-equality */
int lwfb_elem_eval(int32_t device, int32_t fn, int64_t n, const double* x, const double* y, double* out);

/* measured FP64 (DFMA) peak of the device in TFLOP/s: the roofline denominator of this path (the
 * driver-written MEASURED_PEAKS.json has HBM and bf16 tensor figures only) */
int lwfb_fp64_peak(int32_t device, int32_t reps, double* tflops);

/* testing aid: copy of the device-resident member blocks (member-major like the arguments of lwfb_set_members);
 * which = 0 scal, 1 lay, 2 u0 */
int lwfb_debug_get_block(lwfb_handle* h, int32_t which, double* out);

/* testing aid: on = 1 runs the thread-per-member kernel (a second, independently written CUDA implementation that
 * is otherwise used only for N < 3 or N > 32 layers) instead of the lane-per-layer production kernel; on = 2 keeps the
 * lane kernel but only its general instantiation (run-time layer count, every optional flow path compiled in) where
 * the library would pick a lean one (ensembles without bypass / source-area / downslope flow and groundwater feedback,
 * N = 5 or 7); 0 restores the defaults */
int lwfb_debug_force_legacy(lwfb_handle* h, int32_t on);

/* Placement of the members in the kernel's execution slots (no reference analogue; results do not depend on it).
 * The members that share a thread block advance in lock step and wait for each other at every step, so a block
 * is as slow as its member with the most steps.  mode 1 (default, "adaptive"): lwfb_run integrates in WORK QUANTA - a
 * launch ends after LWFB_QUANTUM (environment, default 64) passes of a block's loop, every member checkpoints where
 * it is, the unfinished members are regrouped by progress (least first) and resumed; the front blocks of a launch keep
 * their slots for the whole launch (LWFB_PRIO_PCT % of the resident slots, default 15).  Nothing needs to be known
 * about the members in advance; if the handle has run before (same member count), the first launch is sorted by the
 * RHS evaluations of that run.  When the unfinished members no longer fill the device they are dealt thinly to all
 * resident slots (LWFB_SPREAD, default 1).  Ensembles that do not fill the device once (LWFB_MIN_WAVES, default 1) run
 * in one launch, and lwfb_run_async is always a single launch (cost-sorted if the handle has run before).
 * mode 0: one launch, members as given.  mode 2: like 1 but also for small launches (testing).
 * Every call also forgets what earlier runs of the handle learned about the members' cost (the next run is "cold");
 * lwfb_set_members does the same. */
int lwfb_set_placement(lwfb_handle* h, int32_t mode);

/* kernel launch statistics of the last lwfb_run: number of kernel launches, device ms */
int lwfb_last_run_stats(lwfb_handle* h, int32_t* n_launches, float* device_ms);

/* ---------------------------------------------------------------- several GPUs of one host behind one handle
 * (SURVEY 8e, 8b "multi-GPU inside one call"; no reference analogue: the reference is single-threaded per problem).
 * lwfb_multi owns one lwfb_handle per shard, deals the members to the shards, runs them concurrently (one host thread
 * per shard, no collective during the integration) and all-gathers the per-member scores with peer-to-peer copies over
 * NVLink, so that every device holds the scores of the whole ensemble.  No NCCL, no torch: callable from Julia with
 * plain ccall.  devices [n_shards]: CUDA ordinals; one ordinal may appear more than once (several shards on one GPU).
 * All per-member arrays are in the CALLER'S member order on both sides of every call. */
typedef struct lwfb_multi lwfb_multi;
int lwfb_multi_create(lwfb_multi** mh, const lwfb_opts* opts, int32_t n_shards, const int32_t* devices);
void lwfb_multi_destroy(lwfb_multi* mh);
const char* lwfb_multi_last_error(const lwfb_multi* mh);
int32_t lwfb_multi_n_shards(const lwfb_multi* mh);
lwfb_handle* lwfb_multi_shard(lwfb_multi* mh, int32_t shard);        /* the per-device handle (getters, records) */
int lwfb_multi_set_sites(lwfb_multi* mh, int32_t n_sites, int32_t n_rows, const double* forcing);
/* cost [n_members] or NULL.  NULL: contiguous blocks of members per shard.  Otherwise (e.g. the RHS-evaluation counts of
 * a short probe run, lwfb_multi_get_status): members sorted by cost and dealt to the shards in a snake, so that every
 * shard gets the same number of members and the same share of the expensive ones. */
int lwfb_multi_set_members(lwfb_multi* mh, int64_t n_members, const int32_t* site_of_member, const double* scal,
                           const double* lay, const double* u0, const double* cost);
int lwfb_multi_setup_members(lwfb_multi* mh, int64_t n_members, const double* base_scal, const double* base_lay,
                             const double* base_u0, const double* knobs, int32_t n_over, const int32_t* over_idx,
                             int32_t n_profiles, const double* profiles, const double* psim_base, int32_t ilayer,
                             const double* cost);
int lwfb_multi_set_observations(lwfb_multi* mh, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o);
int lwfb_multi_run(lwfb_multi* mh);
/* arguments as lwfb_scores_device (n_obs_days = 0: the sums accumulated during the run); d_scores_of_shard [n_shards]
 * (optional) receives, per shard, the device pointer of the [n_members][LWFB_NSCORES] block of the WHOLE ensemble;
 * scores (optional) a host copy */
int lwfb_multi_allgather_scores(lwfb_multi* mh, int32_t n_obs_days, const int32_t* obs_day, int32_t n_obs_layers,
                                const int32_t* obs_layer, const double* obs_theta, const double* obs_d18o,
                                double** d_scores_of_shard, double* scores);
int lwfb_multi_get_status(lwfb_multi* mh, int32_t* status, int64_t* counters, int32_t* shard_of_member);
int lwfb_multi_get_final(lwfb_multi* mh, double* out);
int lwfb_multi_last_run_ms(lwfb_multi* mh, float* ms_of_shard, float* ms_max);

#ifdef __cplusplus
}
#endif
#endif /* LWFB90_H */
