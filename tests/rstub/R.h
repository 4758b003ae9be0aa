/* Minimal stand-ins for the R C API, ONLY so that the test-suite can syntax-check r-package/src/shim.c in a container
 * without R (tests/test_abi.py).  Never shipped, never linked. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#endif
