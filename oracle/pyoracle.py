"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/libmc_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
D = C.c_double
PD = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmc_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        for name in ["satvap", "cpair", "orc_leafem"]:
            pass
        def sig(n, res, args):
            f = getattr(L, n); f.restype = res; f.argtypes = args
        sig("orc_satvap", D, [D]); sig("orc_dewpoint", D, [D, D]); sig("orc_phair", D, [D, D]); sig("orc_cpair", D, [D])
        sig("orc_zeroplanedis", D, [D, D]); sig("orc_roughlength", D, [D, D, D]); sig("orc_mixinglength", D, [D, D])
        sig("orc_attencoef", D, [D] * 5); sig("orc_solalt", D, [D] * 6); sig("orc_jday", D, [C.c_int] * 3)
        sig("orc_gcanopy", D, [D] * 11); sig("orc_gturb", D, [D] * 11); sig("orc_gforcedfree", D, [D] * 6)
        sig("orc_windprofile", D, [D] * 9); sig("orc_windcanopy", D, [D] * 10); sig("orc_abovecanopytemp", D, [D] * 10)
        sig("orc_leafem", D, [D, D]); sig("orc_windprofile2", D, [D] * 7)
        sig("orc_thomas", None, [C.c_int, PD, D, D, PD, PD, C.c_int, D, PD, C.c_int, PD])
        sig("orc_paraminit", None, [C.c_int, C.c_int, D, D, D, D, D, D, PD])
        sig("orc_soilinit_z", None, [C.c_int, D, D, PD])
        sig("orc_runonestep", C.c_int, [C.c_int64, C.c_int, C.c_int, PD, PD, PD, PD, PD, PD, D, D, PD, PD, PD,
                                       C.POINTER(C.c_int)])
        sig("orc_run_transient", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, PD, PD, PD, PD, PD, PD, PD, PD, PD, PD,
                                          PD, PD, C.POINTER(C.c_int64), C.c_int64, PD, PD, C.POINTER(C.c_int32), C.c_int])
        sig("orc_run_transient2", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, PD, PD, PD, PD, PD, PD, PD, PD, PD, PD,
                                           PD, PD, C.POINTER(C.c_int64), C.c_int64, PD, PD, C.POINTER(C.c_int32), C.c_int, PD, C.c_int64, PD])
        sig("orc_run_steady", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, PD, PD, PD, PD, C.c_int, PD, PD, PD, PD,
                                       PD, PD, C.c_int])
        sig("orc_run_steady2", C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, PD, PD, PD, PD, C.c_int, PD, PD, PD, PD,
                                        PD, PD, C.c_int, PD, C.c_int64])
        sig("orc_tleafS", None, [PD, PD]); sig("orc_snowtempf", C.c_int, [D, PD, D, PD])
        for n in ["orc_nveg", "orc_nsoil", "orc_nforc", "orc_ncfg"]:
            pass
        L.orc_nstate.restype = C.c_int; L.orc_nstate.argtypes = [C.c_int, C.c_int]
        L.orc_nveg.restype = C.c_int; L.orc_nveg.argtypes = [C.c_int]
        L.orc_nsoil.restype = C.c_int; L.orc_nsoil.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(PD)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def thomas(tc, tmsoil, tair, k, cd, f=0.6, X=0.0):
    tc = _f(tc); k = _f(k); cd = _f(np.atleast_1d(cd)); X = _f(np.atleast_1d(X))
    N = tc.size - 2
    out = np.empty(N + 2)
    lib().orc_thomas(N, _p(tc), tmsoil, tair, _p(k), _p(cd), cd.size, f, _p(X), X.size, _p(out))
    return out


def paraminit(m, sm, hgt, tair, u, relhum, tsoil, Rsw):
    out = np.empty((lib().orc_nstate(m, sm), 1))
    lib().orc_paraminit(m, sm, hgt, tair, u, relhum, tsoil, Rsw, _p(out))
    return out


def soilinit_z(m=10, sdepth=2.0, reqdepth=float("nan")):
    out = np.empty(m)
    lib().orc_soilinit_z(m, sdepth, reqdepth, _p(out))
    return out


def runonestep(vegp, soilp, climvars, lat, lon, cfg, jd, lt, theta, thetap, state, m, sm):
    """state [NSTATE][ncol] is copied; returns (new_state, nmerged)."""
    vegp, soilp, climvars, lat, lon, cfg, theta, thetap = map(_f, (vegp, soilp, climvars, lat, lon, cfg, theta, thetap))
    st = np.array(state, dtype=np.float64, order="C", copy=True)
    ncol = st.shape[1]
    nm = np.zeros(ncol, dtype=np.int32)
    rc = lib().orc_runonestep(ncol, m, sm, _p(vegp), _p(soilp), _p(climvars), _p(lat), _p(lon), _p(cfg), jd, lt, _p(theta),
                              _p(thetap), _p(st), nm.ctypes.data_as(C.POINTER(C.c_int)))
    if rc:
        raise ValueError("At least three canopy layers must be provided")
    return st, nm


def run_transient(vegp, soilp, forcing, lat, lon, cfg, m, sm, state=None, fscale=None, foffset=None, tsoil=None,
                  pai_season=None, samp_idx=None, want_full=False, nthreads=0, vegt=None, thetaz=None):
    """vegt [T][2m+1][1 or ncol]: PAI, pLAI, zm0 per forcing row (.vegpsort); thetaz [T][sm]: soil moisture by depth."""
    vegp, soilp, forcing, lat, lon, cfg = map(_f, (vegp, soilp, forcing, lat, lon, cfg))
    fscale, foffset, tsoil, pai_season, vegt, thetaz = map(_f, (fscale, foffset, tsoil, pai_season, vegt, thetaz))
    ncol = vegp.shape[1]; T = forcing.shape[0]
    NS = lib().orc_nstate(m, sm)
    st = np.zeros((NS, ncol)) if state is None else np.array(state, dtype=np.float64, order="C", copy=True)
    samp = np.arange(ncol, dtype=np.int64) if samp_idx is None else np.ascontiguousarray(samp_idx, dtype=np.int64)
    ns = samp.size
    series = np.zeros((T, 8, ns)); stats = np.zeros((6, ncol)); flags = np.zeros(ncol, dtype=np.int32)
    full = np.zeros((T, NS, ns)) if want_full else None
    nvt = 0
    if vegt is not None:
        assert vegt.ndim == 3 and vegt.shape[0] == T and vegt.shape[1] == 2 * m + 1 and vegt.shape[2] in (1, ncol), vegt.shape
        nvt = vegt.shape[2]
    if thetaz is not None:
        assert thetaz.shape == (T, sm), thetaz.shape
    rc = lib().orc_run_transient2(ncol, m, sm, T, _p(vegp), _p(soilp), _p(forcing), _p(fscale), _p(foffset), _p(tsoil),
                                  _p(lat), _p(lon), _p(cfg), _p(pai_season), _p(st), _p(series),
                                  samp.ctypes.data_as(C.POINTER(C.c_int64)), ns, _p(stats), _p(full),
                                  flags.ctypes.data_as(C.POINTER(C.c_int32)), nthreads, _p(vegt), nvt, _p(thetaz))
    if rc:
        raise ValueError("oracle run_transient rc=%d" % rc)
    return dict(state=st, series=series, stats=stats, flags=flags, full=full)


def run_steady(vegp, soilp, clim, nmr, lat, lon, scfg, m, sm, pai_season=None, want_out=True, nthreads=0, vegt=None):
    vegp, soilp, clim, nmr, lat, lon, scfg, pai_season, vegt = map(_f, (vegp, soilp, clim, nmr, lat, lon, scfg, pai_season, vegt))
    ncol = vegp.shape[1]; T = clim.shape[0]
    percol = 1 if nmr.ndim == 3 else 0
    out = np.zeros((T, 7, ncol)) if want_out else None
    stats = np.zeros((6, ncol))
    rc = lib().orc_run_steady2(ncol, m, sm, T, _p(vegp), _p(soilp), _p(clim), _p(nmr), percol, _p(lat), _p(lon), _p(scfg),
                               _p(pai_season), _p(out), _p(stats), nthreads, _p(vegt), 0 if vegt is None else vegt.shape[2])
    return dict(out=out, stats=stats, rc=rc)
