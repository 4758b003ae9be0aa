/* TEST-ONLY miniature of the R C API: just enough of a runtime behind tests/rstub/*.h for r-package/src/shim.c to be COMPILED AND
 * EXECUTED by the test-suite without R (tests/test_r_shim.py drives its mcb_* entry points through ctypes with vectors laid out as
 * R lays them out, column-major).  Never shipped, never part of the product library. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"

struct SEXPREC { unsigned type; R_xlen_t len; void* data; SEXP* elts; R_CFinalizer_t fin; };
static struct SEXPREC nil_rec = {0, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec;
static jmp_buf g_jmp; static int g_jmp_set = 0; static char g_err[1024];

static SEXP mk(unsigned type, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type; s->len = n; s->data = calloc(n > 0 ? (size_t)n : 1, elt);
  if (type == VECSXP) s->elts = (SEXP*)s->data;
  return s;
}
double* REAL(SEXP s) { return (double*)s->data; }
int* INTEGER(SEXP s) { return (int*)s->data; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int Rf_isNull(SEXP s) { return s == R_NilValue; }
int Rf_asInteger(SEXP s) { return s->type == INTSXP ? INTEGER(s)[0] : (int)REAL(s)[0]; }
double Rf_asReal(SEXP s) { return s->type == INTSXP ? (double)INTEGER(s)[0] : REAL(s)[0]; }
SEXP Rf_allocVector(unsigned type, R_xlen_t n) { return mk(type, n, type == REALSXP ? sizeof(double) : type == INTSXP ? sizeof(int) : sizeof(SEXP)); }
SEXP Rf_allocMatrix(unsigned type, int nr, int nc) { return Rf_allocVector(type, (R_xlen_t)nr * nc); }
SEXP SET_VECTOR_ELT(SEXP v, R_xlen_t i, SEXP x) { v->elts[i] = x; return x; }
SEXP PROTECT(SEXP s) { return s; }
void UNPROTECT(int n) { (void)n; }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  if (g_jmp_set) longjmp(g_jmp, 1);
  fprintf(stderr, "Rf_error outside rstub_call: %s\n", g_err); abort();
}
SEXP R_MakeExternalPtr(void* p, SEXP a, SEXP b) { (void)a; (void)b; SEXP s = mk(22, 0, 1); free(s->data); s->data = p; return s; }
void* R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }     /* (leaks in the test process, as R's would until the call returns) */
int R_registerRoutines(void* d, const void* a, const void* b, const void* c, const void* e) { (void)d; (void)a; (void)b; (void)c; (void)e; return 1; }
Rboolean R_useDynamicSymbols(void* d, Rboolean v) { (void)d; (void)v; return FALSE; }

/* ---- helpers for the Python side ---- */
SEXP rstub_real(const double* v, R_xlen_t n) { SEXP s = Rf_allocVector(REALSXP, n); if (n > 0) memcpy(s->data, v, sizeof(double) * (size_t)n); return s; }
SEXP rstub_int(const int* v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); if (n > 0) memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_elt(SEXP v, R_xlen_t i) { return v->elts[i]; }
unsigned rstub_type(SEXP s) { return s->type; }
const char* rstub_last_error(void) { return g_err; }
void rstub_finalize(SEXP s) { if (s->fin) s->fin(s); }
typedef SEXP (*fn_t)();
/* calls fn(args...) under a setjmp that stands in for R's condition system; returns NULL and sets *err when Rf_error fired */
SEXP rstub_call(fn_t fn, int n, SEXP* a, int* err) {
  *err = 0; g_err[0] = 0; g_jmp_set = 1;
  if (setjmp(g_jmp)) { g_jmp_set = 0; *err = 1; return NULL; }
  SEXP r = NULL;
  switch (n) {
    case 1: r = fn(a[0]); break;
    case 2: r = fn(a[0], a[1]); break;
    case 3: r = fn(a[0], a[1], a[2]); break;
    case 4: r = fn(a[0], a[1], a[2], a[3]); break;
    case 5: r = fn(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = fn(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 8: r = fn(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    default: snprintf(g_err, sizeof g_err, "rstub_call: unsupported arity %d", n); *err = 1;
  }
  g_jmp_set = 0;
  return r;
}
