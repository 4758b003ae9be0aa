"""TEST-ONLY harness: compiles r-package/src/shim.c together with the miniature R runtime of tests/rstub/ and lets Python drive the
shim's mcb_* entry points with vectors laid out as R lays them out (column-major), i.e. exactly what r-package/R/microclimc_b200.R
hands to .Call.  R itself is not installed in the build image; this is how the shim gets executed anyway."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
REALSXP, INTSXP, VECSXP = 14, 13, 19


class RError(RuntimeError):
    """An Rf_error raised inside the shim (R would signal a condition; stop() semantics)."""


def build():
    out = os.path.join(ROOT, "tests", "rstub", "libshim_test.so")
    srcs = [os.path.join(ROOT, "r-package", "src", "shim.c"), os.path.join(ROOT, "tests", "rstub", "rstub_runtime.c")]
    deps = srcs + [os.path.join(ROOT, "include", "microclimc_b200.h"), os.path.join(ROOT, "microclimc_b200", "libmicroclimc_b200.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        libdir = os.path.join(ROOT, "microclimc_b200")
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "tests", "rstub"),
                               "-I" + os.path.join(ROOT, "include"), "-o", out] + srcs +
                              ["-L" + libdir, "-lmicroclimc_b200", "-Wl,-rpath," + libdir])
    return out


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.rstub_real.restype = C.c_void_p; L.rstub_real.argtypes = [C.POINTER(C.c_double), C.c_ssize_t]
        L.rstub_int.restype = C.c_void_p; L.rstub_int.argtypes = [C.POINTER(C.c_int), C.c_ssize_t]
        L.rstub_nil.restype = C.c_void_p
        L.rstub_elt.restype = C.c_void_p; L.rstub_elt.argtypes = [C.c_void_p, C.c_ssize_t]
        L.rstub_type.restype = C.c_uint; L.rstub_type.argtypes = [C.c_void_p]
        L.rstub_last_error.restype = C.c_char_p
        L.rstub_call.restype = C.c_void_p; L.rstub_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.rstub_finalize.argtypes = [C.c_void_p]
        L.REAL.restype = C.POINTER(C.c_double); L.REAL.argtypes = [C.c_void_p]
        L.INTEGER.restype = C.POINTER(C.c_int); L.INTEGER.argtypes = [C.c_void_p]
        L.XLENGTH.restype = C.c_ssize_t; L.XLENGTH.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def r_real(a):
    """numeric vector / matrix / array as R stores it: the elements in column-major (Fortran) order."""
    a = np.asarray(a, dtype=np.float64)
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return lib().rstub_real(flat.ctypes.data_as(C.POINTER(C.c_double)), flat.size)


def r_int(a):
    flat = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
    return lib().rstub_int(flat.ctypes.data_as(C.POINTER(C.c_int)), flat.size)


def r_null():
    return lib().rstub_nil()


def to_numpy(sexp, dim=None):
    """numeric / integer SEXP -> numpy array; `dim` as R's dim attribute (column-major)."""
    L = lib()
    n = L.XLENGTH(sexp)
    t = L.rstub_type(sexp)
    if t == REALSXP:
        v = np.ctypeslib.as_array(L.REAL(sexp), shape=(max(n, 1),))[:n].copy()
    elif t == INTSXP:
        v = np.ctypeslib.as_array(L.INTEGER(sexp), shape=(max(n, 1),))[:n].copy()
    else:
        return None
    return v if dim is None else v.reshape(dim, order="F")


def call(name, *args):
    """.Call(name, ...): args are SEXPs; raises RError with the message when the shim called Rf_error."""
    L = lib()
    fn = C.cast(getattr(L, name), C.c_void_p)
    arr = (C.c_void_p * len(args))(*args)
    err = C.c_int(0)
    r = L.rstub_call(fn, len(args), arr, C.byref(err))
    if err.value:
        raise RError(L.rstub_last_error().decode())
    return r


def elt(sexp, i):
    return lib().rstub_elt(sexp, i)


def routines():
    """name -> arity as registered in shim.c's R_CallMethodDef table."""
    import re
    src = open(os.path.join(ROOT, "r-package", "src", "shim.c")).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(mcb_\w+)", \(DL_FUNC\)&mcb_\w+, (\d+)\}', src)}
