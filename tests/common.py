"""Shared input builders and comparison helpers for the test-suite."""
import numpy as np
from microclimc_b200 import synth as S, layout as L

M, SM = 20, 10
# north_star tolerances: 1e-6 K absolute on temperatures, 1e-8 relative on vapour pressure / fluxes
TOL_T = 1e-6
TOL_REL = 1e-8


def loam(oracle):
    sp = S.load_soilparams(); i = sp["Soil.type"].index("Loam")
    soil = {k: sp[k][i] for k in L.SOIL_SCALARS}
    soil["z"] = oracle.soilinit_z()
    return soil


def make_cfg(timestep=3600.0, edgedist=100.0, reqhgt=np.nan, zu=2.0, merid=0.0, dst=0.0, n=0.6, metopen=1, windhgt=2.0,
             zlafact=1.0, surfwet=1.0, spin_steps=200, harvest_char=0):
    return np.array([timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, spin_steps, harvest_char],
                    dtype=np.float64)


def transient_inputs(oracle, ncol=16, T=240, seed=1234, perturb=True, weather=True):
    clim = S.load_weather() if weather else S.synthetic_forcing(T)
    f, ts = S.forcing_from_climdata(clim)
    f = np.ascontiguousarray(f[:T])
    veg, soil = S.ensemble(S.vegp_deciduous(), loam(oracle), ncol, M, SM, seed=seed)
    lat = np.full(ncol, 50.0); lon = np.full(ncol, -5.0)
    fs = fo = None
    if perturb:
        fs, fo = S.forcing_perturbation(ncol, seed=seed + 1)
    return dict(veg=veg, soil=soil, f=f, ts=ts, lat=lat, lon=lon, fs=fs, fo=fo)


def assert_state_close(a, b, m=M, sm=SM, what=""):
    """a, b: [NSTATE][ncol]. Temperatures to TOL_T absolute, everything else to TOL_REL relative (+ tiny abs floor)."""
    rows = L.state_rows(m, sm)
    temp_keys = ["tabove", "tair", "tsoil", "tc", "tleaf", "soiltc"]
    # flux densities are differences of O(100 W m-2) terms (H = Rnet - G - L): relative to max(|ref|, 1 W m-2)
    flux_keys = ["H", "G", "L", "Rabs", "Rswin", "Rlwin"]
    for k, sl in rows.items():
        x, y = a[sl], b[sl]
        assert np.array_equal(np.isnan(x), np.isnan(y)), "%s NaN pattern differs in %s" % (what, k)
        x = np.nan_to_num(x); y = np.nan_to_num(y)
        if k in temp_keys:
            err = np.max(np.abs(x - y)) if x.size else 0
            assert err <= TOL_T, "%s %s abs err %g" % (what, k, err)
        else:
            floor = 1.0 if k in flux_keys else 1e-6
            err = np.max(np.abs(x - y) / np.maximum(np.abs(y), floor)) if x.size else 0
            assert err <= TOL_REL, "%s %s rel err %g" % (what, k, err)


def assert_series_close(a, b, what=""):
    """[T][8][n] runmodel series: tout, tleaf temperatures; rest relative."""
    assert np.array_equal(np.isnan(a), np.isnan(b)), what + " NaN pattern"
    a = np.nan_to_num(a); b = np.nan_to_num(b)
    for k in (0, 1):
        err = np.max(np.abs(a[:, k] - b[:, k]))
        assert err <= TOL_T, "%s series %s abs err %g" % (what, L.SERIES[k], err)
    for k in range(2, 8):
        floor = 1e-3 if k == 2 else 1.0              # relhum (%) ; flux densities relative to max(|ref|, 1 W m-2)
        err = np.max(np.abs(a[:, k] - b[:, k]) / np.maximum(np.abs(b[:, k]), floor))
        assert err <= TOL_REL, "%s series %s rel err %g" % (what, L.SERIES[k], err)
