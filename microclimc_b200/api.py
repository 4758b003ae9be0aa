"""Host-side mirror of the reference's exported R functions for the hot path (same names, argument meaning,
defaults and error behaviour), dispatching to the CUDA library through the C-ABI (microclimc_b200._lib).

Reference signatures (under /root/reference): paraminit R/code.R:566, soilinit R/code.R:619, runonestep
R/code.R:687-689, spinup R/code.R:982-985, runmodel R/code.R:1125-1129, runmodelS R/NicheMapRcode.R:717-718,
.vegpsort / .snowsort R/code.R:904-933.  R is not installed in the build container, so this Python layer stands
where the R wrappers of r-package/R/microclimc_b200.R stand; both are thin and do only input shaping.
Everything numerical runs on the GPU — there is no CPU path here.

Batched entry points: `runmodel_batch`, `runmodelS_batch` take stacked vegp/soilp ([ncol] lists of the reference's
lists, or pre-packed [row][column] arrays) and one shared climdata (optionally a per-column affine perturbation).
"""
import datetime as _dt
import warnings

import numpy as np

from . import layout as L
from . import synth as S
from ._lib import Context

__all__ = ["paraminit", "soilinit", "runonestep", "spinup", "runmodel", "runmodelS", "runmodel_batch", "runmodelS_batch",
           "leafem", "weather", "dailyprecip", "soilparams", "climvars"]


# ---- bundled datasets (R/data.R:1-65) ------------------------------------------------------------------------
def weather():
    return S.load_weather()


def dailyprecip():
    return S.load_dailyprecip()


def soilparams():
    return S.load_soilparams()


def climvars():
    return S.load_climvars()


def leafem(tc, vegem):
    """R/code.R:59-62 (exported, not on the hot path; pure formula)."""
    return vegem * 5.67 * 10 ** -8 * (np.asarray(tc, float) + 273.15) ** 4


# ---- small helpers ---------------------------------------------------------------------------------------------
def _isna(x):
    return x is None or (isinstance(x, float) and np.isnan(x))


def _tme_jd_lt(tme):
    """POSIXlt-like -> (jday, decimal hour). Accepts datetime, 'YYYY-mm-dd HH:MM[:SS]' or (jd, lt)."""
    if isinstance(tme, (tuple, list)) and len(tme) == 2:
        return float(tme[0]), float(tme[1])
    if isinstance(tme, str):
        jd, lt, _ = S.parse_obs_time([tme])
        return float(jd[0]), float(lt[0])
    if isinstance(tme, _dt.datetime):
        return float(S.jday(tme.year, tme.month, tme.day)), tme.hour + tme.minute / 60 + tme.second / 3600
    raise TypeError("tme must be a datetime, a time string or (jd, lt)")


def _vegpsort(vegp, i):
    """.vegpsort (R/code.R:904-916): column i of time-varying PAI / pLAI, element i of zm0."""
    v = dict(vegp)
    for k in ("PAI", "pLAI"):
        a = np.asarray(vegp[k], float)
        if a.ndim == 2:
            v[k] = a[:, i].copy()
    z = np.atleast_1d(np.asarray(vegp["zm0"], float))
    v["zm0"] = float(z[i] if z.size > 1 else z[0])
    return v


def _snowsort(vegp, snow, i=0):
    """.snowsort (R/code.R:918-933); `snow` None/NA -> unchanged."""
    v = dict(vegp)
    if snow is not None and not (np.ndim(snow) == 0 and _isna(snow)):
        if np.atleast_1d(snow)[i] == 1:
            v["lw"] = v["lw"] * 1.5
            v.update(zm0=0.024, refls=0.85, refg=0.8, refw=0.7, reflp=0.8, gsmax=10.0, q50=1.0)
    return v


def _season_of(vegp):
    """Time-varying vegp (PAI / pLAI as m x T matrices, zm0 of length T; what .vegpsort indexes, R/code.R:904-916) ->
    (static vegp, season[T] or None, vegt [T][2m+1] or None).

    When only PAI varies and PAI[:, t] = PAI[:, 0] * s[t] exactly, the cheap separable form (pai_season) is used; any other
    time dependence is passed through in full as `vegt` rows (PAI[m], pLAI[m], zm0 per time step)."""
    P = np.asarray(vegp["PAI"], float)
    pl = np.asarray(vegp["pLAI"], float)
    z0 = np.atleast_1d(np.asarray(vegp["zm0"], float))
    if P.ndim == 1 and pl.ndim == 1 and z0.size == 1:
        return vegp, None, None
    T = P.shape[1] if P.ndim == 2 else (pl.shape[1] if pl.ndim == 2 else z0.size)
    v = dict(vegp)
    v["PAI"] = P[:, 0].copy() if P.ndim == 2 else P
    v["pLAI"] = pl[:, 0].copy() if pl.ndim == 2 else pl
    v["zm0"] = float(z0[0])
    pl_static = pl.ndim == 1 or np.array_equal(pl, np.repeat(pl[:, :1], pl.shape[1], axis=1))
    z_static = z0.size == 1 or np.all(z0 == z0[0])
    if P.ndim == 2 and pl_static and z_static:
        base = P[:, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.nanmean(np.where(base[:, None] > 0, P / base[:, None], np.nan), axis=0)
        if np.all(np.isfinite(s)) and np.allclose(base[:, None] * s[None, :], P, rtol=1e-12, atol=0):
            return v, np.ascontiguousarray(s), None
    m = P.shape[0]
    vt = np.empty((T, 2 * m + 1))
    vt[:, :m] = P.T if P.ndim == 2 else P[None, :]
    vt[:, m:2 * m] = pl.T if pl.ndim == 2 else pl[None, :]
    vt[:, 2 * m] = z0 if z0.size > 1 else z0[0]
    return v, None, vt


def _veg_time(vl, snow=None):
    """Per-column vegp lists -> (static lists, shared season or None, vegt [T][2m+1][1 or ncol] or None)."""
    out, seasons, vts = [], [], []
    for v in vl:
        v2, s, vt = _season_of(v)
        out.append(v2); seasons.append(s); vts.append(vt)
    if all(s is None for s in seasons) and all(vt is None for vt in vts):
        return out, None, None
    same_season = all(s is not None for s in seasons) and all(np.array_equal(s, seasons[0]) for s in seasons)
    if same_season:
        return out, seasons[0], None
    # general form: every column gets its rows (columns given in separable form are expanded)
    T = next((vt.shape[0] for vt in vts if vt is not None), None) or next(s.size for s in seasons if s is not None)
    m = len(np.asarray(out[0]["PAI"]))
    full = np.empty((T, 2 * m + 1, len(vl)))
    for c, (v2, s, vt) in enumerate(zip(out, seasons, vts)):
        if vt is None:
            f = np.ones(T) if s is None else s
            full[:, :m, c] = np.asarray(v2["PAI"], float)[None, :] * f[:, None]
            full[:, m:2 * m, c] = np.asarray(v2["pLAI"], float)[None, :]
            full[:, 2 * m, c] = v2["zm0"]
        else:
            full[:, :, c] = vt
    if len(vl) > 1 and all(np.array_equal(full[:, :, c], full[:, :, 0]) for c in range(1, len(vl))):
        full = full[:, :, :1]
    return out, None, np.ascontiguousarray(full)


def _state_to_list(st, m, sm, c=0):
    rows = L.state_rows(m, sm)
    out = {}
    for k, sl in rows.items():
        v = st[sl, c]
        out[k] = float(v[0]) if k in L.STATE_SCALARS else v.copy()
    return out


def _list_to_state(previn, m, sm):
    st = np.zeros((L.nstate(m, sm), 1))
    for k, sl in L.state_rows(m, sm).items():
        st[sl, 0] = np.asarray(previn[k], float)
    return st


def _pack(vegp, soilp):
    if isinstance(vegp, np.ndarray):
        veg = np.ascontiguousarray(vegp, float)
        m = (veg.shape[0] - len(L.VEG_SCALARS)) // 4
    else:
        vl = vegp if isinstance(vegp, (list, tuple)) else [vegp]
        m = len(np.asarray(vl[0]["thickw"]))
        veg = S.pack_vegp(vl, m)
    if isinstance(soilp, np.ndarray):
        soil = np.ascontiguousarray(soilp, float)
        sm = soil.shape[0] - len(L.SOIL_SCALARS)
    else:
        sl = soilp if isinstance(soilp, (list, tuple)) else [soilp]
        sm = len(np.asarray(sl[0]["z"]))
        soil = S.pack_soilp(sl, sm)
    if soil.shape[1] == 1 and veg.shape[1] > 1:
        soil = np.repeat(soil, veg.shape[1], axis=1)
    return veg, soil, m, sm


def _cfg(timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, spin, char=0):
    rq = np.nan if _isna(reqhgt) else float(reqhgt)
    ed = np.nan if _isna(edgedist) else float(edgedist)
    return np.array([timestep, ed, rq, zu, merid, dst, n, 1.0 if metopen else 0.0, windhgt, zlafact, surfwet, spin, char], float)


def _check_pai(veg, m):
    r = L.veg_rows(m)
    if np.min(veg[r["PAI"]]) == 0:
        warnings.warn("Some PAI values are 0. Zero PAI values set to 0.0001 to enable model to run\n")   # code.R:695


# ---- constructors ------------------------------------------------------------------------------------------------
def soilinit(soiltype, m=10, sdepth=2, reqdepth=None):
    """R/code.R:619-627."""
    sp = S.load_soilparams()
    if soiltype not in sp["Soil.type"]:
        raise ValueError("unknown soil type %r" % (soiltype,))
    i = sp["Soil.type"].index(soiltype)
    out = {k: sp[k][i] for k in sp}
    z = (sdepth / m ** 1.2) * np.arange(1, m + 1, dtype=float) ** 1.2
    if not _isna(reqdepth):
        z[int(np.argmin(np.abs(z - reqdepth)))] = reqdepth
    out["z"] = z
    return out


def paraminit(m, sm, hgt, tair, u, relhum, tsoil, Rsw):
    """R/code.R:566-585 — evaluated by the batched device kernel (mc_paraminit)."""
    ctx = Context(m, sm)
    veg = np.zeros((L.nveg(m), 1)); soil = np.zeros((L.nsoil(sm), 1))
    ctx.set_columns(veg, soil, np.zeros(1), np.zeros(1))
    ctx.paraminit(np.array([[hgt], [tair], [u], [relhum], [tsoil], [Rsw]], float))
    return _state_to_list(ctx.get_state(), m, sm)


def runonestep(climvars, previn, vegp, soilp, timestep, tme, lat, long, edgedist=1000, sdepth=2, reqhgt=None, zu=2, theta=0.3,
               thetap=0.3, merid=0, dst=0, n=0.6, metopen=True, windhgt=2, zlafact=1, surfwet=1):
    """R/code.R:687-902 for one column."""
    m = len(previn["tc"]); sm = len(previn["soiltc"])
    if m < 3:
        raise ValueError("At least three canopy layers must be provided")
    veg, soil, m2, sm2 = _pack(vegp, soilp)
    _check_pai(veg, m)
    ctx = Context(m, sm)
    ctx.set_columns(veg, soil, np.array([lat], float), np.array([long], float))
    ctx.set_config(_cfg(timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, 0))
    ctx.set_state(_list_to_state(previn, m, sm))
    u = climvars["u"] if "u" in climvars else climvars["u2"]            # partial `$` matching, quirk Q4
    dp = climvars.get("dp", None)
    cv = np.array([[climvars["tair"]], [climvars["relhum"]], [climvars["pk"]], [u], [climvars["tsoil"]], [climvars["skyem"]],
                   [climvars["Rsw"]], [np.nan if _isna(dp) else dp]], float)
    jd, lt = _tme_jd_lt(tme)
    ctx.run_onestep(cv, jd, lt, np.array([np.atleast_1d(theta)[0]], float), np.array([np.atleast_1d(thetap)[0]], float))
    return _state_to_list(ctx.get_state(), m, sm)


def _transient_ctx(climdata, veg, soil, m, sm, lat, long, cfg, theta, fscale, foffset, season, n_gpus):
    ncol = veg.shape[1]
    f, ts = S.forcing_from_climdata(climdata, theta=theta)
    ctx = Context(m, sm, n_gpus)
    ctx.set_columns(veg, soil, np.broadcast_to(np.asarray(lat, float), (ncol,)).copy(),
                    np.broadcast_to(np.asarray(long, float), (ncol,)).copy())
    ctx.set_forcing(f, fscale, foffset, season)
    return ctx, f, ts


def _theta_series(theta, T, sm):
    """runmodel's theta argument (R/code.R:1093-1100, 1146-1154) -> (theta for the forcing column, thetaz [T][sm] or None)."""
    th = np.asarray(theta, float)
    if th.ndim == 0:
        return float(th), None
    if th.ndim == 2:
        if th.shape[0] == 1:
            return th[0], None
        if th.shape != (sm, T):
            raise ValueError("theta matrix must have one row per soil layer and one column per time step")
        return th[0], np.ascontiguousarray(th.T)
    if th.size == 1:
        return float(th[0]), None
    if th.size == sm:                       # code.R:1151-1153 (tested before the time-series reading, as R does when sm == T)
        return float(th[0]), np.ascontiguousarray(np.repeat(th[None, :], T, axis=0))
    if th.size == T:
        return th, None
    raise ValueError("theta must be a scalar, a vector over time or soil layers, or a matrix [soil layers x time]")


def spinup(climdata, vegp, soilp, lat, long, edgedist=100, reqhgt=None, sdepth=2, zu=2, theta=0.3, thetap=0.3, merid=0, dst=0,
           n=0.6, plotout=False, steps=200, metopen=True, windhgt=2, zlafact=1, surfwet=1):
    """R/code.R:982-1013 (plotout is accepted and ignored: graphics are out of scope)."""
    if not np.array_equal(np.atleast_1d(theta), np.atleast_1d(thetap)):
        raise NotImplementedError("spinup with thetap != theta")
    vegp1 = _vegpsort(vegp, 0)
    veg, soil, m, sm = _pack(vegp1, soilp)
    _check_pai(veg, m)
    f, ts = S.forcing_from_climdata(climdata, theta=float(np.atleast_1d(theta)[0]))
    ctx = Context(m, sm)
    ctx.set_columns(veg, soil, np.array([lat], float), np.array([long], float))
    ctx.set_forcing(f)
    if np.size(theta) > 1:                                     # theta by soil layer (runmodel passes theta[, 1], code.R:1158)
        if np.size(theta) != sm:
            raise ValueError("theta must be a scalar or one value per soil layer")
        ctx.set_thetaz(np.repeat(np.asarray(theta, float)[None, :], f.shape[0], axis=0))
    ctx.set_config(_cfg(ts, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, steps))
    ctx.spinup(steps)
    return _state_to_list(ctx.get_state(), m, sm)


def _prepare_transient(climdata, vegp, soilp, lat, long, edgedist, reqhgt, zu, theta, merid, dst, n, steps, metopen, windhgt, zlafact,
                       surfwet, previn, snow, fscale, foffset, n_gpus):
    """Argument checks and input shaping of `runmodel` (R/code.R:1130-1192) for stacked columns; returns the loaded context."""
    char = isinstance(reqhgt, str)
    if char and reqhgt not in ("all", "allclim"):
        raise ValueError("input reqhgt no recognised\n")                                     # code.R:1136-1139
    season = vegt = None
    if not isinstance(vegp, np.ndarray):
        vl = list(vegp) if isinstance(vegp, (list, tuple)) else [vegp]
        vl = [_snowsort(v, snow, 0) for v in vl]                                             # Q3: snow[1] decides for the whole run
        vegp, season, vegt = _veg_time(vl)
        if vegt is not None and snow is not None and not (np.ndim(snow) == 0 and _isna(snow)) and np.atleast_1d(snow)[0] == 1:
            vegt[:, -1, :] = 0.024                                                          # .snowsort after .vegpsort (code.R:1196-1197)
    veg, soil, m, sm = _pack(vegp, soilp)
    ncol = veg.shape[1]
    _check_pai(veg, m)
    rq = None if char else reqhgt
    if not _isna(rq) and rq < 0:                                                            # code.R:1130-1135
        r = L.soil_rows(sm)["z"]
        for c in range(ncol):
            z = soil[r, c]
            dif = z + rq
            sel = int(np.argmin(np.abs(dif)))
            z[sel] = -rq
            if dif[sel] < 0 and sel != sm - 1:
                z[sel + 1] = z[sel + 1] - dif[sel]
    hg = veg[L.veg_rows(m)["hgt"]][0]
    if (not metopen) and np.any(windhgt < hg):
        raise ValueError("When metopen is FALSE, windhgt must be above canopy\n")            # code.R:1141-1143
    T = len(climdata["temp"])
    th, thetaz = _theta_series(theta, T, sm)
    spin = steps if previn is None else 0
    ctx, f, ts = _transient_ctx(climdata, veg, soil, m, sm, lat, long, None, th, fscale, foffset, season, n_gpus)
    if vegt is not None:
        if vegt.shape[0] != T:
            raise ValueError("time-varying vegp must have one column per time step of climdata")
        ctx.set_vegt(vegt)
    if thetaz is not None:
        ctx.set_thetaz(thetaz)
    ctx.set_config(_cfg(ts, edgedist, rq, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, spin, 1 if char else 0))
    if previn is not None:
        st = previn if isinstance(previn, np.ndarray) else np.concatenate([_list_to_state(p, m, sm) for p in
                                                                           (previn if isinstance(previn, (list, tuple)) else [previn])], axis=1)
        ctx.set_state(st)
    return ctx, dict(m=m, sm=sm, ncol=ncol, T=T, char=char, timestep=ts)


def _runmodel_batch_by_class(classes, climdata, vegp, soilp, lat, long, kw):
    """Columns ordered by parameter class for the run (tiles of one class share their geometry rows, mc_set_classes) and the results
    put back in the caller's order.  vegp / soilp must be packed arrays here; columns of a class must be identical (checked)."""
    classes = np.asarray(classes)
    veg = np.ascontiguousarray(vegp, float); soil = np.ascontiguousarray(soilp, float)
    ncol = veg.shape[1]
    if soil.shape[1] == 1:
        soil = np.repeat(soil, ncol, axis=1)
    order = np.argsort(classes, kind="stable"); inv = np.empty(ncol, np.int64); inv[order] = np.arange(ncol)
    ids = classes[order].astype(np.int32)
    first = np.r_[0, np.flatnonzero(np.diff(ids)) + 1]
    rep = np.repeat(first, np.diff(np.r_[first, ncol]))
    if not (np.array_equal(veg[:, order], veg[:, order][:, rep]) and np.array_equal(soil[:, order], soil[:, order][:, rep])):
        raise ValueError("columns of one class must have identical vegp and soilp")
    per = lambda a: None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, float), (ncol,))[order])
    per2 = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, float)[:, order])
    kw = dict(kw)
    sample = kw.pop("sample", None)
    samp = np.arange(ncol) if sample is None else np.asarray(sample, dtype=np.int64)
    kw["fscale"], kw["foffset"] = per2(kw.get("fscale")), per2(kw.get("foffset"))
    if kw.get("previn") is not None:
        kw["previn"] = per2(kw["previn"])
    if kw.get("tsoil") is not None and np.ndim(kw["tsoil"]) > 0:
        kw["tsoil"] = per(kw["tsoil"])
    r = runmodel_batch(climdata, veg[:, order].copy(), soil[:, order].copy(), per(lat), per(long), sample=inv[samp], _class_ids=ids, **kw)
    r["stats"] = r["stats"][:, inv]; r["flags"] = r["flags"][inv]; r["state"] = r["state"][:, inv]; r["sample"] = samp
    return r


def runmodel_batch(climdata, vegp, soilp, lat, long, edgedist=100, reqhgt=None, sdepth=2, zu=2, theta=0.3, merid=0, dst=0, n=0.6,
                   steps=200, tsoil=None, metopen=True, windhgt=2, zlafact=1, surfwet=1, previn=None, snow=None, fscale=None,
                   foffset=None, sample=None, full=False, n_gpus=1, classes=None, _class_ids=None):
    """Batched `runmodel` (R/code.R:1125-1328) over stacked columns.

    vegp / soilp: lists of the reference's lists (one per column) or packed [row][column] arrays.
    classes: optional parameter-class id per column (habitat maps: K << ncol distinct vegp / soilp sets; packed arrays only) — the
    columns are run ordered by class so that the streamed geometry is shared (mc_set_classes), results come back in the given order.
    Returns dict(series [T][8][nsample] (layout.SERIES), stats [6][ncol], flags, state [NSTATE][ncol], full, sample, obs_time)."""
    if classes is not None:
        return _runmodel_batch_by_class(classes, climdata, vegp, soilp, lat, long, dict(
            edgedist=edgedist, reqhgt=reqhgt, sdepth=sdepth, zu=zu, theta=theta, merid=merid, dst=dst, n=n, steps=steps, tsoil=tsoil,
            metopen=metopen, windhgt=windhgt, zlafact=zlafact, surfwet=surfwet, previn=previn, snow=snow, fscale=fscale, foffset=foffset,
            sample=sample, full=full, n_gpus=n_gpus))
    ctx, info = _prepare_transient(climdata, vegp, soilp, lat, long, edgedist, reqhgt, zu, theta, merid, dst, n, steps, metopen, windhgt,
                                   zlafact, surfwet, previn, snow, fscale, foffset, n_gpus)
    m, sm, ncol, char = info["m"], info["sm"], info["ncol"], info["char"]
    ts_arr = None
    if not _isna(tsoil):
        ts_arr = np.broadcast_to(np.asarray(tsoil, float), (ncol,)).copy()
    samp = np.arange(ncol) if sample is None else np.asarray(sample, dtype=np.int64)
    if _class_ids is not None:
        ctx.set_classes(_class_ids)
    r = ctx.run_transient(tsoil=ts_arr, samp_idx=samp, want_series=True, want_full=bool(full or char))
    r["state"] = ctx.get_state(); r["sample"] = samp; r["obs_time"] = np.asarray(climdata["obs_time"])
    r["m"], r["sm"], r["kernel_ms"] = m, sm, ctx.last_kernel_ms()
    return r


def _all_tables(full, obs, m, sm, with_fluxes):
    """The data.frames `runmodel` returns for reqhgt "allclim" / "all" (R/code.R:1267-1322) from the full-state rows [T][NSTATE]
    of one column."""
    import pandas as pd
    rows = L.state_rows(m, sm)
    z = full[-1, rows["z"]]; sz = full[-1, rows["sz"]]
    zab = full[-1, rows["zabove"]][0]

    def df(cols, names):
        d = {"obs_time": obs}
        for nm_, c in zip(names, cols.T):
            d[nm_] = c
        return pd.DataFrame(d)
    zn = ["%g" % round(v, 3) for v in z]
    out = dict(
        Airtemp=df(np.column_stack([full[:, rows["tc"]], full[:, rows["tabove"]], full[:, rows["psi_m"]]]),
                   ["Temp_%sm" % v for v in zn] + ["Temp_%gm" % round(zab, 3), "psi_m"]),
        Leaftemp=df(full[:, rows["tleaf"]], ["Temp_%sm" % v for v in zn]),
        Soiltemp=df(full[:, rows["soiltc"]], ["Temp_%gm" % round(v, 3) for v in sz]),
        Windspeed=df(full[:, rows["uz"]], ["u_%s" % v for v in zn]),
        Relhum=df(full[:, rows["rh"]], ["relh_%s" % v for v in zn]))
    if with_fluxes:
        # code.R:1302-1310 binds gha, gt (m+1 columns), gv in that order; its name vector is one short, so plain names here
        out["Conductivity"] = df(np.column_stack([full[:, rows["gha"]], full[:, rows["gt"]], full[:, rows["gv"]]]),
                                 ["Leafbound_%s" % v for v in zn] + ["Air_%d" % i for i in range(m + 1)] + ["Leafvap_%s" % v for v in zn])
        out["Fluxes"] = df(np.column_stack([full[:, rows["Rswin"]], full[:, rows["Rlwin"]], full[:, rows["L"]], full[:, rows["H"]],
                                            full[:, rows["G"]]]),
                           ["SWin_%s" % v for v in zn] + ["Lwin_%s" % v for v in zn] + ["Latent_%sW/m^2" % v for v in zn] + ["Sensible", "Ground"])
    return out


def runmodel(climdata, vegp, soilp, lat, long, edgedist=100, reqhgt=None, sdepth=2, zu=2, theta=0.3, merid=0, dst=0, n=0.6,
             steps=200, plotout=False, plotsteps=100, tsoil=None, metopen=True, windhgt=2, zlafact=1, surfwet=1, previn=None,
             snow=None):
    """R/code.R:1125-1328 for one column. Returns a pandas DataFrame (obs_time, reftemp, tout, tleaf, relhum, SWin, LWin, H, L, G)
    or, for reqhgt "all"/"allclim", a dict of DataFrames (Airtemp, Leaftemp, Soiltemp, Windspeed, Relhum[, Conductivity, Fluxes])."""
    import pandas as pd
    r = runmodel_batch(climdata, [vegp], [soilp], lat, long, edgedist, reqhgt, sdepth, zu, theta, merid, dst, n, steps, tsoil, metopen,
                       windhgt, zlafact, surfwet, previn=None if previn is None else [previn], snow=snow)
    m, sm = r["m"], r["sm"]
    obs = r["obs_time"]
    if not isinstance(reqhgt, str):
        s = r["series"][:, :, 0]
        return pd.DataFrame(dict(obs_time=obs, reftemp=np.asarray(climdata["temp"], float), tout=s[:, 0], tleaf=s[:, 1], relhum=s[:, 2],
                                 SWin=s[:, 3], LWin=s[:, 4], H=s[:, 6], L=s[:, 5], G=s[:, 7]))
    return _all_tables(r["full"][:, :, 0], obs, m, sm, reqhgt == "all")


# ---- steady state --------------------------------------------------------------------------------------------------
def _nmr_arrays(nmrout, T):
    nmr = np.zeros((T, 11))
    nmr[:, 0] = np.asarray(nmrout["soiltemps"]["D0cm"], float)
    nmr[:, 1] = np.asarray(nmrout["soilmoist"]["WC0cm"], float)
    sd = np.asarray(nmrout["metout"]["SNOWDEP"], float)
    nmr[:, 2] = sd
    snow = nmrout.get("snowtemp", nmrout.get("snow", None))                                # partial matching, quirk Q17
    if np.max(sd) > 0:
        if snow is None:
            raise ValueError("object 'L' not found (snow present but nmrout$snowtemp is not a data.frame; NMR.R:519)")
        for q in range(8):
            nmr[:, 3 + q] = np.asarray(snow["SN%d" % (q + 1)], float)
    return nmr


def _q16_remap(tleaf, tloc, snowdep_cm, reqhgt):
    """Quirk Q16 (NMR.R:590-594, 668): under snow `tleaf2 <- tz2[sbs]` double-indexes the already subset vector."""
    sbs = np.nonzero(reqhgt < snowdep_cm / 100.0)[0]
    if sbs.size:
        tz2 = tloc[sbs]
        idx = sbs                                            # 0-based positions used as 1-based-1 indices into tz2
        vals = np.where(idx < tz2.size, tz2[np.minimum(idx, tz2.size - 1)], np.nan)
        tleaf = tleaf.copy(); tleaf[sbs] = vals
    return tleaf


def runmodelS_batch(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen=True, windhgt=2, surfwet=1, groundem=0.95, n_gpus=1,
                    want_out=True):
    """Batched `runmodelS` (R/NicheMapRcode.R:717-885, incl. the .runmodelsnow overlay). nmrout: the reference's list
    (shared by all columns) or a packed array [T][11] / [T][11][ncol] (layout.STEADY_NMR)."""
    season = vegt = None
    if not isinstance(vegp, np.ndarray):
        vl = list(vegp) if isinstance(vegp, (list, tuple)) else [vegp]
        vegp, season, vegt = _veg_time(vl)
    veg, soil, m, sm = _pack(vegp, soilp)
    ncol = veg.shape[1]
    hg = veg[L.veg_rows(m)["hgt"]][0]
    if (not metopen) and np.any(windhgt < hg):
        raise ValueError("When metopen is FALSE, windhgt must be above canopy\n")            # NMR.R:719-721
    clim = S.steady_clim_from_climdata(climdata)
    T = clim.shape[0]
    nmr = nmrout if isinstance(nmrout, np.ndarray) else _nmr_arrays(nmrout, T)
    ctx = Context(m, sm, n_gpus)
    ctx.set_columns(veg, soil, np.broadcast_to(np.asarray(lat, float), (ncol,)).copy(),
                    np.broadcast_to(np.asarray(long, float), (ncol,)).copy())
    if vegt is not None:
        if vegt.shape[0] != T:
            raise ValueError("time-varying vegp must have one column per time step of climdata")
        ctx.set_vegt(vegt)
    r = ctx.run_steady(clim, nmr, np.array([reqhgt, 1.0 if metopen else 0.0, windhgt, surfwet, groundem], float), season, want_out=want_out)
    r["nmr"] = nmr; r["kernel_ms"] = ctx.last_kernel_ms()
    return r


def runmodelS(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen=True, windhgt=2, surfwet=1, groundem=0.95):
    """R/NicheMapRcode.R:717-885 for one column -> DataFrame(obs_time, Tref, Tloc, tleaf, RHref, RHloc, RSWloc, RLWloc, windspeed, dp)."""
    import pandas as pd
    r = runmodelS_batch(climdata, [vegp], [soilp], nmrout, reqhgt, lat, long, metopen, windhgt, surfwet, groundem)
    o = r["out"][:, :, 0]
    tleaf = _q16_remap(o[:, 1], o[:, 0], r["nmr"][:, 2], reqhgt)
    return pd.DataFrame(dict(obs_time=np.asarray(climdata["obs_time"]), Tref=np.asarray(climdata["temp"], float), Tloc=o[:, 0], tleaf=tleaf,
                             RHref=np.asarray(climdata["relhum"], float), RHloc=o[:, 2], RSWloc=o[:, 3], RLWloc=o[:, 4], windspeed=o[:, 5],
                             dp=o[:, 6]))
