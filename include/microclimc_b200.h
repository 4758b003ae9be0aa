/* microclimc_b200 — C ABI of the B200-native batched canopy energy-balance hot path.
 *
 * The reference (ilyamaclean/microclimc) is a pure-R package with no native/FFI layer (NAMESPACE:1-22), so
 * the drop-in boundary is created at the level of its exported R functions.  Each entry point below names
 * the reference function (file:line under /root/reference) whose batched evaluation it provides; the R-side
 * `.Call` shim that binds them is shown in INTEGRATION.md (r-package/src/shim.c).
 *
 * Conventions
 *  - plain C, FP64 everywhere, caller-owned HOST memory; the library copies and never retains caller pointers;
 *  - every per-column structure is a 2-D array [row][column] (column index fastest, "column-interleaved SoA");
 *    row maps are the MC_V_*, MC_S_*, MC_P_* enums below;
 *  - functions return 0 on success or a negative MC_ERR_* code; mc_last_error() gives the message
 *    (the R shim turns it into Rf_error, mirroring the reference's stop() guards);
 *  - columns are split statically and contiguously over the context's GPUs (no collective; one host thread
 *    and one stream per GPU, joined before return); no R API is touched off the calling thread;
 *  - there is NO CPU fallback: without a CUDA device mc_create fails with MC_ERR_NO_DEVICE.
 */
#ifndef MICROCLIMC_B200_H
#define MICROCLIMC_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 1

/* vegp rows: 17 scalars then PAI[m], pLAI[m], iw[m], thickw[m]  (21 fields, running-microclimc.md:367-370) */
enum { MC_V_HGT, MC_V_X, MC_V_LW, MC_V_CD, MC_V_HGTG, MC_V_ZM0, MC_V_REFLS, MC_V_REFG, MC_V_REFW, MC_V_REFLP, MC_V_VEGEM,
       MC_V_GSMAX, MC_V_Q50, MC_V_CPW, MC_V_PHW, MC_V_KWOOD, MC_V_CLUMP, MC_V_NSCAL };
/* soilp rows: 12 scalars then z[sm]  (R/code.R:619-627) */
enum { MC_S_SMAX, MC_S_SMIN, MC_S_ALPHA, MC_S_N, MC_S_KSAT, MC_S_VQ, MC_S_VM, MC_S_VO, MC_S_MC, MC_S_RHO, MC_S_B, MC_S_PSIE,
       MC_S_NSCAL };
/* state (`previn`) rows: 9 scalars then tc, tleaf, uz, z, rh, Rabs, gv, gha, L, Rswin, Rlwin [m each], gt[m+1],
 * soiltc[sm], sz[sm]  (R/code.R:581-584, 898-901) */
enum { MC_P_TABOVE, MC_P_ZABOVE, MC_P_RELHUM, MC_P_TAIR, MC_P_TSOIL, MC_P_PK, MC_P_H, MC_P_G, MC_P_PSIM, MC_P_NSCAL };
/* forcing row per time step; rows 0..7 accept the per-column affine perturbation v*scale+offset, after which
 * relhum is limited to [1,100], u and Rsw to >= 0, skyem to [0,1], theta to >= 0.001.
 * dp = difrad/swrad with non-finite -> 0.5 (R/code.R:1166-1167); jd = microctools::jday(tme); lt = decimal hour. */
enum { MC_F_TAIR, MC_F_RELHUM, MC_F_PK, MC_F_U, MC_F_SKYEM, MC_F_RSW, MC_F_DP, MC_F_THETA, MC_F_JD, MC_F_LT, MC_F_N };
#define MC_NPERT 8
/* run configuration = scalar arguments of runonestep/runmodel (R/code.R:687-689, 1125-1129).
 * reqhgt NaN == NA; harvest_char != 0 == reqhgt "all"/"allclim". */
enum { MC_C_TIMESTEP, MC_C_EDGEDIST, MC_C_REQHGT, MC_C_ZU, MC_C_MERID, MC_C_DST, MC_C_N, MC_C_METOPEN, MC_C_WINDHGT,
       MC_C_ZLAFACT, MC_C_SURFWET, MC_C_SPINSTEPS, MC_C_HARVEST_CHAR, MC_C_N_ };
/* per-step outputs of runmodel (R/code.R:1222-1265, 1324-1325) */
enum { MC_O_TOUT, MC_O_TLEAF, MC_O_RELHUM, MC_O_SWIN, MC_O_LWIN, MC_O_L, MC_O_H, MC_O_G, MC_O_N };
/* steady path (runmodelS, R/NicheMapRcode.R:717-885) */
enum { MC_SC_TEMP, MC_SC_RELHUM, MC_SC_PRES, MC_SC_WINDSPEED, MC_SC_SWRAD, MC_SC_DIFRAD, MC_SC_SKYEM, MC_SC_JD, MC_SC_LT, MC_SC_N };
enum { MC_SN_D0CM, MC_SN_WC0CM, MC_SN_SNOWDEP, MC_SN_SN1, MC_SN_N = 11 };          /* nmrout inputs (NMR.R:795-796, 872, 518) */
enum { MC_SG_REQHGT, MC_SG_METOPEN, MC_SG_WINDHGT, MC_SG_SURFWET, MC_SG_GROUNDEM, MC_SG_N };
enum { MC_SO_TLOC, MC_SO_TLEAF, MC_SO_RHLOC, MC_SO_RSWLOC, MC_SO_RLWLOC, MC_SO_WINDSPEED, MC_SO_DP, MC_SO_N };   /* metout, NMR.R:864-865 */

enum { MC_OK = 0, MC_ERR_ARG = -1, MC_ERR_LAYERS = -2 /* "At least three canopy layers" code.R:699 */,
       MC_ERR_NO_DEVICE = -3, MC_ERR_CUDA = -4, MC_ERR_STATE = -5, MC_ERR_WINDHGT = -6 /* code.R:1141-1143, NMR.R:719-721 */,
       MC_ERR_SNOWLAYER = -7 /* .snowtempf column lookup out of range, NMR.R:508 */ };

typedef struct mc_ctx mc_ctx;

int mc_version(void);
int mc_device_count(void);                       /* number of visible CUDA devices (0 => nothing can run) */
int mc_nveg(int m);   int mc_nsoil(int sm);   int mc_nstate(int m, int sm);
const char* mc_last_error(const mc_ctx* ctx);    /* ctx may be NULL: last mc_create failure */

/* Context for columns with m canopy and sm soil nodes on n_gpus devices (gpu_ids NULL => 0..n_gpus-1). */
mc_ctx* mc_create(int m, int sm, int n_gpus, const int* gpu_ids);
void mc_destroy(mc_ctx* ctx);

/* vegp [NVEG][ncol], soilp [NSOIL][ncol], lat/lon [ncol].  (vegp/soilp arguments of every reference entry point) */
int mc_set_columns(mc_ctx* ctx, int64_t ncol, const double* vegp, const double* soilp, const double* lat, const double* lon);
/* forcing [T][MC_F_N] shared base series; fscale/foffset [MC_NPERT][ncol] or NULL; pai_season [T] or NULL
 * (PAI(t) = PAI * pai_season[t], the batched form of .vegpsort, R/code.R:904-916). */
int mc_set_forcing(mc_ctx* ctx, int64_t T, const double* forcing, const double* fscale, const double* foffset,
                   const double* pai_season);
/* General time-varying vegetation (.vegpsort, R/code.R:904-916: vegp$PAI / vegp$pLAI as m x T matrices, vegp$zm0 of length T, what
 * microctools::habitatvars returns): vegt [T][2m+1][nvt] = PAI[m], pLAI[m], zm0 of every forcing row, one series shared by all
 * columns (nvt == 1) or one per column (nvt == ncol); T must equal the rows of the run that uses it (the forcing rows of
 * mc_run_transient, the T of mc_run_steady: checked there, MC_ERR_ARG); NULL clears.  Transient runs with it take the general kernel
 * (the streaming kernel hoists the geometry once per launch); mc_run_steady reads PAI and pLAI from it (NMR.R:730-735). */
int mc_set_vegt(mc_ctx* ctx, int64_t T, int64_t nvt, const double* vegt);
/* Soil moisture by depth (theta matrix of runmodel, R/code.R:1146-1154, consumed by soilk at code.R:790): thetaz [T][sm], shared by
 * all columns (the per-column perturbation of the theta forcing row applies to every depth).  The forcing's theta column is then
 * ignored.  R recycles the soil vector over the canopy layers inside leaftemp (code.R:416-437; canopy layer i sees soil node
 * (i-1) mod sm), which is reproduced; sm > m is rejected.  T must equal the forcing rows of the run; NULL clears.  General kernel. */
int mc_set_thetaz(mc_ctx* ctx, int64_t T, const double* thetaz);
/* Parameter classes (grids with K << ncol distinct vegetation / soil parameter sets, e.g. habitat maps): class_of [ncol], the caller's
 * promise that columns with the same id >= 0 have identical vegp AND soilp columns (lat / lon / forcing perturbation may differ; a
 * negative id marks a column of its own).  Tiles of 128 consecutive columns that belong to one class then share their per-layer geometry
 * rows, which are read from L2 instead of DRAM (the streamed geometry is 5.3 of the 8.7 KB the streaming kernel moves per column-step);
 * results are bit-identical to a run without classes.  Order the columns by class for it to take effect.  Call after mc_set_columns
 * (which clears it); NULL clears. */
int mc_set_classes(mc_ctx* ctx, const int32_t* class_of);
/* Statistics across calls: on != 0 starts (or restarts) an accumulation — the mean / min / max (and the flag bits) that mc_run_transient
 * returns then cover every step run since this call, carried on the device between calls, so a year run week by week needs no host-side
 * combining and can skip the read-back (stats = NULL) until its last call; on == 0 returns to per-call statistics. */
int mc_set_stats_accumulate(mc_ctx* ctx, int on);
int mc_set_config(mc_ctx* ctx, const double* cfg /* [MC_C_N_] */);
int mc_set_state(mc_ctx* ctx, const double* state /* [NSTATE][ncol] */);
int mc_get_state(mc_ctx* ctx, double* state);

/* paraminit (R/code.R:566-585), batched: args [6][ncol] = hgt, tair, u, relhum, tsoil, Rsw -> device state. */
int mc_paraminit(mc_ctx* ctx, const double* args);
/* runonestep (R/code.R:687-902): climvars [8][ncol] = tair, relhum, pk, u, tsoil, skyem, Rsw, dp (NaN dp => difprop);
 * tme as (jd, lt); theta/thetap [ncol]; nmerged [ncol] or NULL receives the number of merged air layers. */
int mc_run_onestep(mc_ctx* ctx, const double* climvars, double jd, double lt, const double* theta, const double* thetap,
                   int32_t* nmerged);
/* runmodel hot loop (R/code.R:1193-1266) over steps [t0,t1) of the forcing, preceded by paraminit + spinup
 * (R/code.R:982-1013, 1155-1159) when cfg[MC_C_SPINSTEPS] > 0 and t0 == 0.
 *  tsoil [ncol] or NULL (NaN => mean air temperature of the column, code.R:1165);
 *  samp_idx [nsamp] columns whose hourly outputs are returned: series [t1-t0][MC_O_N][nsamp],
 *  fullseries [t1-t0][NSTATE][nsamp] (reqhgt "all"/"allclim" tables); either may be NULL;
 *  stats [6][ncol] = mean,min,max of tout then of tleaf (NULL to skip); flags [ncol] bit0 non-finite output,
 *  bit1 merged-layer branch taken, bit3 the column left the streaming kernel and was redone by the general kernel inside the
 *  same call (NULL to skip).  samp_idx must not hold a column twice (MC_ERR_ARG). */
int mc_run_transient(mc_ctx* ctx, int64_t t0, int64_t t1, const double* tsoil, const int64_t* samp_idx, int64_t nsamp,
                     double* series, double* fullseries, double* stats, int32_t* flags);
/* spinup (R/code.R:982-1013) as a call of its own: paraminit + `steps` iterations on forcing row 0. */
int mc_spinup(mc_ctx* ctx, int steps);

/* runmodelS + .runmodelsnow (R/NicheMapRcode.R:717-885, 514-677) for ncol columns x T hours.
 *  clim [T][MC_SC_N] shared; nmr [T][MC_SN_N] shared (nmr_percol == 0) or [T][MC_SN_N][ncol];
 *  scfg [MC_SG_N]; pai_season [T] or NULL; out [T][MC_SO_N][ncol] or NULL; stats [6][ncol] (Tloc, tleaf) or NULL. */
int mc_run_steady(mc_ctx* ctx, int64_t T, const double* clim, const double* nmr, int nmr_percol, const double* scfg,
                  const double* pai_season, double* out, double* stats);

/* Page-locked host memory for callers that want true asynchronous copies (the library accepts any host pointer; pinned
 * input / result buffers remove the driver's staging copy: bench.py's e2e legs, microclimc_b200.gridrun).  NULL on failure. */
void* mc_host_alloc(size_t bytes);
void mc_host_free(void* p);

/* Device time (ms, CUDA events on the launching streams, max over GPUs) of the kernels of the last mc_run_* /
 * mc_spinup call, and the number of kernel launches it made. */
double mc_last_kernel_ms(const mc_ctx* ctx);
int64_t mc_last_kernel_launches(const mc_ctx* ctx);
/* Re-run the last mc_run_transient / mc_run_steady kernels on the device-resident inputs `reps` times without any
 * host<->device copy (state is restored before each repetition); returns the mean device ms per repetition.
 * The state copy a transient replay needs is kept only once this entry has been called on the context: the first call after a
 * transient run without spin-up arms it and returns -1 (MC_ERR_STATE in mc_last_error); repeat the run, then call again. */
double mc_bench_resident(mc_ctx* ctx, int reps);
/* Measured FP64 (non-tensor DFMA) peak of the context's first GPU in TFLOP/s: the roofline denominator of this
 * FP64-bound path (the driver's MEASURED_PEAKS.json carries no FP64 figure). */
double mc_fp64_peak_tflops(mc_ctx* ctx);

/* Diagnostic: evaluates the device's fast FP64 primitives (microclimc_b200/csrc/mc_phys.cuh) on n host values.
 * op: 0 exp, 1 log, 2 1/x, 3 sqrt, 4 x/y (y = second operand array, may be null for the other ops).
 * The kernels of this library replace IEEE division and the libm exp/log/sqrt by these; tests bound their error in ulp.
 * Returns 0 or a negative MC_ERR_* code. */
int mc_eval_primitive(mc_ctx* ctx, int op, int64_t n, const double* x, const double* y, double* out);

#ifdef __cplusplus
}
#endif
#endif
