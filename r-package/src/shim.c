/* .Call shim between R and the C-ABI of include/microclimc_b200.h.
 *
 * The reference package (ilyamaclean/microclimc) is pure R with no native layer (NAMESPACE:1-22), so this file is the
 * whole FFI surface a maintainer adds: every entry converts R vectors / matrices (column-major doubles, which already
 * are the library's [row][column] layout when the matrix is stored as ncol x nrow, see R/microclimc_b200.R) to plain
 * pointers, calls one mc_* function and turns a negative return code into an R error carrying mc_last_error()
 * (mirroring the reference's stop() guards, R/code.R:699, 1141-1143; R/NicheMapRcode.R:719-721).
 * Nothing numerical happens here.  R was not available in the build container, so this file is compiled there only
 * against stub headers in the test-suite's syntax check; tests exercise the same ABI from Python (ctypes). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "microclimc_b200.h"

static void ctx_finalizer(SEXP ptr) {
  mc_ctx* ctx = (mc_ctx*)R_ExternalPtrAddr(ptr);
  if (ctx) { mc_destroy(ctx); R_ClearExternalPtr(ptr); }
}
static mc_ctx* get_ctx(SEXP ptr) {
  mc_ctx* ctx = (mc_ctx*)R_ExternalPtrAddr(ptr);
  if (!ctx) Rf_error("microclimc_b200: context already destroyed");
  return ctx;
}
static void check(mc_ctx* ctx, int rc) {
  if (rc != MC_OK) Rf_error("%s", mc_last_error(ctx));
}
static const double* opt(SEXP x) { return Rf_isNull(x) ? NULL : REAL(x); }

SEXP mcb_create(SEXP m, SEXP sm, SEXP ngpu) {
  mc_ctx* ctx = mc_create(Rf_asInteger(m), Rf_asInteger(sm), Rf_asInteger(ngpu), NULL);
  if (!ctx) Rf_error("%s", mc_last_error(NULL));      /* "At least three canopy layers ...", or no CUDA device */
  SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}
/* vegp: matrix ncol x NVEG (R column-major == [NVEG][ncol]); soilp: ncol x NSOIL; lat, lon: length ncol */
SEXP mcb_set_columns(SEXP ptr, SEXP vegp, SEXP soilp, SEXP lat, SEXP lon) {
  mc_ctx* ctx = get_ctx(ptr);
  check(ctx, mc_set_columns(ctx, (int64_t)XLENGTH(lat), REAL(vegp), REAL(soilp), REAL(lat), REAL(lon)));
  return R_NilValue;
}
/* forcing: matrix MC_F_N x T (so that R's column-major storage is [T][MC_F_N]); fscale/foffset: ncol x 8 or NULL */
SEXP mcb_set_forcing(SEXP ptr, SEXP forcing, SEXP fscale, SEXP foffset, SEXP season) {
  mc_ctx* ctx = get_ctx(ptr);
  check(ctx, mc_set_forcing(ctx, (int64_t)(XLENGTH(forcing) / MC_F_N), REAL(forcing), opt(fscale), opt(foffset), opt(season)));
  return R_NilValue;
}
SEXP mcb_set_config(SEXP ptr, SEXP cfg) {
  mc_ctx* ctx = get_ctx(ptr);
  if (XLENGTH(cfg) != MC_C_N_) Rf_error("microclimc_b200: cfg must have %d entries", (int)MC_C_N_);
  check(ctx, mc_set_config(ctx, REAL(cfg)));
  return R_NilValue;
}
SEXP mcb_set_state(SEXP ptr, SEXP state) { mc_ctx* ctx = get_ctx(ptr); check(ctx, mc_set_state(ctx, REAL(state))); return R_NilValue; }
SEXP mcb_get_state(SEXP ptr, SEXP ncol, SEXP nstate) {
  mc_ctx* ctx = get_ctx(ptr);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(ncol), Rf_asInteger(nstate)));
  check(ctx, mc_get_state(ctx, REAL(out)));
  UNPROTECT(1);
  return out;
}
SEXP mcb_paraminit(SEXP ptr, SEXP args6) { mc_ctx* ctx = get_ctx(ptr); check(ctx, mc_paraminit(ctx, REAL(args6))); return R_NilValue; }
SEXP mcb_spinup(SEXP ptr, SEXP steps) { mc_ctx* ctx = get_ctx(ptr); check(ctx, mc_spinup(ctx, Rf_asInteger(steps))); return R_NilValue; }
SEXP mcb_run_onestep(SEXP ptr, SEXP climvars, SEXP jd, SEXP lt, SEXP theta, SEXP thetap) {
  mc_ctx* ctx = get_ctx(ptr);
  check(ctx, mc_run_onestep(ctx, REAL(climvars), Rf_asReal(jd), Rf_asReal(lt), REAL(theta), REAL(thetap), NULL));
  return R_NilValue;
}
/* vegt: array c(nvt, 2m+1, T) (== [T][2m+1][nvt]); .vegpsort for every time step (R/code.R:904-916) */
SEXP mcb_set_vegt(SEXP ptr, SEXP vegt, SEXP T, SEXP nvt) {
  mc_ctx* ctx = get_ctx(ptr);
  check(ctx, mc_set_vegt(ctx, (int64_t)Rf_asReal(T), (int64_t)Rf_asReal(nvt), opt(vegt)));
  return R_NilValue;
}
/* thetaz: matrix sm x T (== [T][sm]): runmodel's theta matrix as the user passes it (R/code.R:1146-1154) */
SEXP mcb_set_thetaz(SEXP ptr, SEXP thetaz, SEXP T) {
  mc_ctx* ctx = get_ctx(ptr);
  check(ctx, mc_set_thetaz(ctx, (int64_t)Rf_asReal(T), opt(thetaz)));
  return R_NilValue;
}
/* returns list(series = [nsamp x 8 x nt], stats = [ncol x 6], flags = int[ncol], full = [nsamp x NSTATE x nt] or NULL);
 * `full` holds the whole state of the sampled columns at every step: the reqhgt "all" / "allclim" tables (R/code.R:1267-1322) */
SEXP mcb_run_transient(SEXP ptr, SEXP t0, SEXP t1, SEXP tsoil, SEXP samp, SEXP ncol_, SEXP want_full, SEXP nstate) {
  mc_ctx* ctx = get_ctx(ptr);
  const int64_t a = (int64_t)Rf_asReal(t0), b = (int64_t)Rf_asReal(t1), ncol = (int64_t)Rf_asReal(ncol_);
  const int64_t ns = XLENGTH(samp), nt = b - a;
  const int wf = Rf_asInteger(want_full) != 0 && ns > 0;
  int64_t* idx = (int64_t*)R_alloc(ns > 0 ? ns : 1, sizeof(int64_t));
  for (int64_t j = 0; j < ns; ++j) idx[j] = (int64_t)REAL(samp)[j] - 1;          /* R is 1-based */
  SEXP series = PROTECT(Rf_allocVector(REALSXP, nt * MC_O_N * ns));
  SEXP stats = PROTECT(Rf_allocMatrix(REALSXP, (int)ncol, 6));
  SEXP flags = PROTECT(Rf_allocVector(INTSXP, ncol));
  SEXP full = PROTECT(wf ? Rf_allocVector(REALSXP, nt * (int64_t)Rf_asInteger(nstate) * ns) : R_NilValue);
  check(ctx, mc_run_transient(ctx, a, b, opt(tsoil), idx, ns, ns ? REAL(series) : NULL, wf ? REAL(full) : NULL, REAL(stats),
                              (int32_t*)INTEGER(flags)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, series); SET_VECTOR_ELT(out, 1, stats); SET_VECTOR_ELT(out, 2, flags); SET_VECTOR_ELT(out, 3, full);
  UNPROTECT(5);
  return out;
}
/* clim: MC_SC_N x T; nmr: MC_SN_N x T (shared); returns [ncol x MC_SO_N x T] */
SEXP mcb_run_steady(SEXP ptr, SEXP clim, SEXP nmr, SEXP scfg, SEXP season, SEXP ncol_) {
  mc_ctx* ctx = get_ctx(ptr);
  const int64_t T = XLENGTH(clim) / MC_SC_N, ncol = (int64_t)Rf_asReal(ncol_);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, T * MC_SO_N * ncol));
  check(ctx, mc_run_steady(ctx, T, REAL(clim), REAL(nmr), 0, REAL(scfg), opt(season), REAL(out), NULL));
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
  {"mcb_create", (DL_FUNC)&mcb_create, 3}, {"mcb_set_columns", (DL_FUNC)&mcb_set_columns, 5},
  {"mcb_set_forcing", (DL_FUNC)&mcb_set_forcing, 5}, {"mcb_set_config", (DL_FUNC)&mcb_set_config, 2},
  {"mcb_set_state", (DL_FUNC)&mcb_set_state, 2}, {"mcb_get_state", (DL_FUNC)&mcb_get_state, 3},
  {"mcb_paraminit", (DL_FUNC)&mcb_paraminit, 2}, {"mcb_spinup", (DL_FUNC)&mcb_spinup, 2},
  {"mcb_run_onestep", (DL_FUNC)&mcb_run_onestep, 6}, {"mcb_run_transient", (DL_FUNC)&mcb_run_transient, 8},
  {"mcb_set_vegt", (DL_FUNC)&mcb_set_vegt, 4}, {"mcb_set_thetaz", (DL_FUNC)&mcb_set_thetaz, 3},
  {"mcb_run_steady", (DL_FUNC)&mcb_run_steady, 6}, {NULL, NULL, 0}};
void R_init_microclimcB200(DllInfo* dll) { R_registerRoutines(dll, NULL, call_methods, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }
