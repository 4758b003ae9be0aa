"""ctypes binding of the C-ABI library (include/microclimc_b200.h).  No CPU fallback: if the CUDA library is
missing or no device is visible every compute entry point raises."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MC_B200_LIB: developer knob to A/B another build of the same CUDA library (never a CPU path)
LIB_PATH = os.environ.get("MC_B200_LIB") or os.path.join(_HERE, "libmicroclimc_b200.so")
_LIB = None
D, PD, I64 = C.c_double, C.POINTER(C.c_double), C.c_int64
PI32, PI64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)

# every symbol the public header declares (tests/test_abi.py checks the list against the header)
SYMBOLS = {
    "mc_version": (C.c_int, []),
    "mc_device_count": (C.c_int, []),
    "mc_nveg": (C.c_int, [C.c_int]), "mc_nsoil": (C.c_int, [C.c_int]), "mc_nstate": (C.c_int, [C.c_int, C.c_int]),
    "mc_last_error": (C.c_char_p, [C.c_void_p]),
    "mc_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "mc_destroy": (None, [C.c_void_p]),
    "mc_set_columns": (C.c_int, [C.c_void_p, I64, PD, PD, PD, PD]),
    "mc_set_forcing": (C.c_int, [C.c_void_p, I64, PD, PD, PD, PD]),
    "mc_set_vegt": (C.c_int, [C.c_void_p, I64, I64, PD]),
    "mc_set_thetaz": (C.c_int, [C.c_void_p, I64, PD]),
    "mc_set_classes": (C.c_int, [C.c_void_p, PI32]),
    "mc_set_stats_accumulate": (C.c_int, [C.c_void_p, C.c_int]),
    "mc_set_config": (C.c_int, [C.c_void_p, PD]),
    "mc_set_state": (C.c_int, [C.c_void_p, PD]),
    "mc_get_state": (C.c_int, [C.c_void_p, PD]),
    "mc_paraminit": (C.c_int, [C.c_void_p, PD]),
    "mc_run_onestep": (C.c_int, [C.c_void_p, PD, D, D, PD, PD, PI32]),
    "mc_run_transient": (C.c_int, [C.c_void_p, I64, I64, PD, PI64, I64, PD, PD, PD, PI32]),
    "mc_spinup": (C.c_int, [C.c_void_p, C.c_int]),
    "mc_run_steady": (C.c_int, [C.c_void_p, I64, PD, PD, C.c_int, PD, PD, PD, PD]),
    "mc_last_kernel_ms": (D, [C.c_void_p]),
    "mc_last_kernel_launches": (I64, [C.c_void_p]),
    "mc_bench_resident": (D, [C.c_void_p, C.c_int]),
    "mc_fp64_peak_tflops": (D, [C.c_void_p]),
    "mc_eval_primitive": (C.c_int, [C.c_void_p, C.c_int, I64, PD, PD, PD]),
    "mc_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mc_host_free": (None, [C.c_void_p]),
}


class MicroclimcError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MicroclimcError("CUDA library %s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                  "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("MC_B200_LIB") and not hasattr(L, name):
                continue                      # A/B of an older build: newer diagnostic entries may be absent
            f = getattr(L, name)
            f.restype = res; f.argtypes = args
        _LIB = L
    return _LIB


class _Pinned:
    """Owner of one mc_host_alloc block; numpy views keep it alive through their .base chain."""

    def __init__(self, nbytes):
        self.ptr = lib().mc_host_alloc(max(int(nbytes), 8))
        if not self.ptr:
            raise MicroclimcError("mc_host_alloc(%d) failed" % nbytes)
        self.buf = (C.c_byte * max(int(nbytes), 8)).from_address(self.ptr)

    def __del__(self):
        if getattr(self, "ptr", None) and _LIB is not None:
            _LIB.mc_host_free(self.ptr); self.ptr = None


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (mc_host_alloc): copies to / from it are true asynchronous DMA."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    own = _Pinned(n)
    a = np.frombuffer(own.buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    a.flags.writeable = True
    _PINNED_OWNERS[id(own)] = own          # (np.frombuffer keeps `own.buf`, not `own`: hold the owner until the module goes away)
    return a


_PINNED_OWNERS = {}


def pinned_copy(a, dtype=np.float64):
    out = pinned_empty(np.shape(a), dtype)
    out[...] = a
    return out


def _p(a, typ=PD):
    return None if a is None else a.ctypes.data_as(typ)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """Owns an mc_ctx: columns with m canopy / sm soil nodes on `n_gpus` devices (static contiguous split)."""

    def __init__(self, m, sm, n_gpus=1, gpu_ids=None):
        L = lib()
        ids = None
        if gpu_ids is not None:
            ids = (C.c_int * len(gpu_ids))(*gpu_ids); n_gpus = len(gpu_ids)
        self._h = L.mc_create(m, sm, n_gpus, ids)
        if not self._h:
            msg = L.mc_last_error(None).decode()
            if "three canopy layers" in msg:
                raise ValueError(msg)
            raise MicroclimcError(msg)
        self.m, self.sm, self.ncol, self.T = m, sm, 0, 0
        self.nstate = L.mc_nstate(m, sm)

    def close(self):
        if getattr(self, "_h", None):
            lib().mc_destroy(self._h); self._h = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            msg = lib().mc_last_error(self._h).decode()
            if rc == -2:
                raise ValueError(msg)
            raise MicroclimcError("rc=%d: %s" % (rc, msg))

    def set_columns(self, vegp, soilp, lat, lon):
        vegp, soilp, lat, lon = map(_f, (vegp, soilp, lat, lon))
        assert vegp.shape[0] == lib().mc_nveg(self.m) and soilp.shape[0] == lib().mc_nsoil(self.sm), (vegp.shape, soilp.shape)
        self.ncol = vegp.shape[1]
        assert soilp.shape[1] == self.ncol and lat.size == self.ncol and lon.size == self.ncol
        self._chk(lib().mc_set_columns(self._h, self.ncol, _p(vegp), _p(soilp), _p(lat), _p(lon)))

    def set_forcing(self, forcing, fscale=None, foffset=None, pai_season=None):
        forcing, fscale, foffset, pai_season = map(_f, (forcing, fscale, foffset, pai_season))
        self.T = forcing.shape[0]
        self._chk(lib().mc_set_forcing(self._h, self.T, _p(forcing), _p(fscale), _p(foffset), _p(pai_season)))

    def set_vegt(self, vegt):
        """vegt [T][2m+1][1 or ncol]: PAI[m], pLAI[m], zm0 per forcing row (.vegpsort); None clears."""
        if vegt is None:
            self._chk(lib().mc_set_vegt(self._h, 0, 0, None)); return
        vegt = _f(vegt)
        assert vegt.ndim == 3 and vegt.shape[1] == 2 * self.m + 1, vegt.shape
        self._chk(lib().mc_set_vegt(self._h, vegt.shape[0], vegt.shape[2], _p(vegt)))

    def set_thetaz(self, thetaz):
        """thetaz [T][sm]: soil moisture by depth per forcing row; None clears."""
        if thetaz is None:
            self._chk(lib().mc_set_thetaz(self._h, 0, None)); return
        thetaz = _f(thetaz)
        assert thetaz.ndim == 2 and thetaz.shape[1] == self.sm, thetaz.shape
        self._chk(lib().mc_set_thetaz(self._h, thetaz.shape[0], _p(thetaz)))

    def set_classes(self, class_of):
        """class_of [ncol] int32: columns with the same id >= 0 have identical vegp and soilp (the caller's promise); None clears."""
        if class_of is None:
            self._chk(lib().mc_set_classes(self._h, None)); return
        c = np.ascontiguousarray(class_of, dtype=np.int32); assert c.shape == (self.ncol,)
        self._chk(lib().mc_set_classes(self._h, _p(c, PI32)))

    def set_stats_accumulate(self, on=True):
        """Carry the per-column statistics and flags across run_transient calls (on the device); (re)starts the accumulation."""
        self._chk(lib().mc_set_stats_accumulate(self._h, 1 if on else 0))

    def set_config(self, cfg):
        cfg = _f(cfg); assert cfg.size == 13
        self._chk(lib().mc_set_config(self._h, _p(cfg)))

    def set_state(self, state):
        state = _f(state); assert state.shape == (self.nstate, self.ncol)
        self._chk(lib().mc_set_state(self._h, _p(state)))

    def get_state(self):
        out = np.empty((self.nstate, self.ncol))
        self._chk(lib().mc_get_state(self._h, _p(out)))
        return out

    def paraminit(self, args):
        args = _f(args); assert args.shape == (6, self.ncol)
        self._chk(lib().mc_paraminit(self._h, _p(args)))

    def run_onestep(self, climvars, jd, lt, theta, thetap):
        climvars, theta, thetap = map(_f, (climvars, theta, thetap))
        nm = np.zeros(self.ncol, dtype=np.int32)
        self._chk(lib().mc_run_onestep(self._h, _p(climvars), float(jd), float(lt), _p(theta), _p(thetap), _p(nm, PI32)))
        return nm

    def spinup(self, steps):
        self._chk(lib().mc_spinup(self._h, int(steps)))

    def run_transient(self, t0=0, t1=None, tsoil=None, samp_idx=None, want_series=True, want_full=False, want_stats=True,
                      want_flags=True, out_stats=None, out_flags=None):
        """Steps [t0, t1) of the runmodel loop.  `out_stats` ([6][ncol] float64) / `out_flags` ([ncol] int32): caller-owned result
        buffers to fill instead of fresh arrays (e.g. pinned memory reused across calls)."""
        t1 = self.T if t1 is None else t1
        nt = t1 - t0
        tsoil = _f(tsoil)
        samp = None; ns = 0
        if samp_idx is not None and (want_series or want_full):
            samp = np.ascontiguousarray(samp_idx, dtype=np.int64); ns = samp.size
        series = np.zeros((nt, 8, ns)) if (want_series and ns) else None
        full = np.zeros((nt, self.nstate, ns)) if (want_full and ns) else None
        stats = (out_stats if out_stats is not None else np.zeros((6, self.ncol))) if want_stats else None
        flags = (out_flags if out_flags is not None else np.zeros(self.ncol, dtype=np.int32)) if want_flags else None
        if stats is not None:
            assert stats.dtype == np.float64 and stats.shape == (6, self.ncol) and stats.flags["C_CONTIGUOUS"]
        if flags is not None:
            assert flags.dtype == np.int32 and flags.shape == (self.ncol,) and flags.flags["C_CONTIGUOUS"]
        self._chk(lib().mc_run_transient(self._h, t0, t1, _p(tsoil), _p(samp, PI64), ns, _p(series), _p(full), _p(stats),
                                         _p(flags, PI32)))
        return dict(series=series, full=full, stats=stats, flags=flags)

    def run_steady(self, clim, nmr, scfg, pai_season=None, want_out=True, want_stats=True, out_stats=None, out_out=None):
        """runmodelS over T hours.  `out_stats` ([6][ncol]) / `out_out` ([T][7][ncol]): caller-owned (e.g. pinned) result buffers."""
        clim, nmr, scfg, pai_season = map(_f, (clim, nmr, scfg, pai_season))
        T = clim.shape[0]
        percol = 1 if nmr.ndim == 3 else 0
        out = (out_out if out_out is not None else np.zeros((T, 7, self.ncol))) if want_out else None
        stats = (out_stats if out_stats is not None else np.zeros((6, self.ncol))) if want_stats else None
        for a_, shp in ((out, (T, 7, self.ncol)), (stats, (6, self.ncol))):
            if a_ is not None:
                assert a_.dtype == np.float64 and a_.shape == shp and a_.flags["C_CONTIGUOUS"]
        self._chk(lib().mc_run_steady(self._h, T, _p(clim), _p(nmr), percol, _p(scfg), _p(pai_season), _p(out), _p(stats)))
        return dict(out=out, stats=stats)

    def last_kernel_ms(self):
        return lib().mc_last_kernel_ms(self._h)

    def last_kernel_launches(self):
        return lib().mc_last_kernel_launches(self._h)

    def fp64_peak_tflops(self):
        return lib().mc_fp64_peak_tflops(self._h)

    def eval_primitive(self, op, x, y=None):
        """Device fast-math primitive `op` ("exp", "log", "rcp", "sqrt", "div") on host arrays (diagnostic)."""
        code = {"exp": 0, "log": 1, "rcp": 2, "sqrt": 3, "div": 4}[op]
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty_like(x)
        self._chk(lib().mc_eval_primitive(self._h, code, x.size, _p(x), _p(y), _p(out)))
        return out

    def bench_resident(self, reps=1):
        ms = lib().mc_bench_resident(self._h, reps)
        if ms < 0:
            raise MicroclimcError(lib().mc_last_error(self._h).decode())
        return ms
