#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include "R.h"
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
extern SEXP R_NilValue;
double* REAL(SEXP); int* INTEGER(SEXP); R_xlen_t XLENGTH(SEXP);
int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned, R_xlen_t); SEXP Rf_allocMatrix(unsigned, int, int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP PROTECT(SEXP); void UNPROTECT(int);
void Rf_error(const char*, ...);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
char* R_alloc(size_t, int);
#endif
