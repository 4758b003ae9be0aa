"""GPU: the R-API mirror (microclimc_b200.api) end to end, the golden vectors through the CUDA path, multi-GPU split,
and size-independent properties at BASELINE sizes."""
import json
import os

import numpy as np
import pytest

from tests.common import M, SM, make_cfg, transient_inputs, assert_state_close, assert_series_close, loam
from microclimc_b200 import api, layout as L, synth as S

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intree_known_answers.json")))


def test_paraminit_golden_through_cuda():
    for c in GOLD["paraminit"]:
        m, sm = int(c["args"][0]), int(c["args"][1])
        got = api.paraminit(m, sm, *c["args"][2:])
        assert len(got) == 23
        for k, v in c["out"].items():
            np.testing.assert_allclose(np.atleast_1d(got[k]), np.atleast_1d(v), rtol=1e-13, atol=1e-13, err_msg=k)


def test_runmodel_config1_single_column_bundled_weather(oracle):
    """BASELINE config 1 (first fortnight): runmodel on `weather` + soilinit("Loam"), 20 layers, hourly, spin-up 200."""
    w = {k: v[:336] for k, v in S.load_weather().items()}
    vegp = S.vegp_deciduous(); soilp = api.soilinit("Loam")
    df = api.runmodel(w, vegp, soilp, 50.2178, -5.32656)
    assert list(df.columns) == ["obs_time", "reftemp", "tout", "tleaf", "relhum", "SWin", "LWin", "H", "L", "G"] and len(df) == 336
    veg, soil, m, sm = api._pack([vegp], [soilp])
    f, ts = S.forcing_from_climdata(w)
    want = oracle.run_transient(veg, soil, f, np.array([50.2178]), np.array([-5.32656]), make_cfg(timestep=ts, spin_steps=200), m, sm)
    s = want["series"][:, :, 0]
    assert np.max(np.abs(df["tout"].values - s[:, 0])) <= 1e-6 and np.max(np.abs(df["tleaf"].values - s[:, 1])) <= 1e-6
    for col, k in [("relhum", 2), ("SWin", 3), ("LWin", 4), ("L", 5), ("H", 6), ("G", 7)]:
        assert np.max(np.abs(df[col].values - s[:, k]) / np.maximum(np.abs(s[:, k]), 1.0)) <= 1e-8, col
    assert np.all(np.isfinite(df["tout"].values)) and abs(df["tout"].mean() - np.mean(w["temp"])) < 6.0      # tracks the forcing


def test_runmodel_all_tables_and_previn_restart(oracle):
    w = {k: v[2000:2048] for k, v in S.load_weather().items()}
    vegp = S.vegp_deciduous(); soilp = api.soilinit("Loam")
    out = api.runmodel(w, vegp, soilp, 50, -5, reqhgt="all", steps=50)
    assert sorted(out) == ["Airtemp", "Conductivity", "Fluxes", "Leaftemp", "Relhum", "Soiltemp", "Windspeed"]
    assert out["Airtemp"].shape == (48, 1 + 20 + 2) and out["Soiltemp"].shape == (48, 1 + 10) and out["Conductivity"].shape == (48, 1 + 61)
    # previn: continue from a spin-up state instead of spinning up again (code.R:1155-1159)
    pv = api.spinup(w, vegp, soilp, 50, -5, steps=50)
    a = api.runmodel(w, vegp, soilp, 50, -5, previn=pv)
    b = api.runmodel(w, vegp, soilp, 50, -5, steps=50)
    np.testing.assert_allclose(a["tout"].values, b["tout"].values, rtol=0, atol=1e-9)


def test_runonestep_and_spinup_api(oracle):
    vegp = S.vegp_deciduous(); soilp = api.soilinit("Loam")
    pv = api.paraminit(20, 10, 15.0, 16, 2.1, 90, 11, 500)
    cv = dict(tair=16, relhum=90, pk=101.3, u2=2.1, tsoil=11, skyem=0.9, Rsw=500, dp=None)     # field named u2: quirk Q4
    new = api.runonestep(cv, pv, vegp, soilp, 60, "2021-06-21 13:00:00", 50, -5, edgedist=100)
    veg, soil, m, sm = api._pack([vegp], [soilp])
    clim = np.array([[16], [90], [101.3], [2.1], [11], [0.9], [500], [np.nan]], float)
    cfg = make_cfg(timestep=60, edgedist=100, spin_steps=0)
    want, _ = oracle.runonestep(veg, soil, clim, np.array([50.]), np.array([-5.]), cfg, float(S.jday(2021, 6, 21)), 13.0, np.array([0.3]),
                                np.array([0.3]), api._list_to_state(pv, 20, 10), 20, 10)
    assert_state_close(api._list_to_state(new, 20, 10), want, what="runonestep api")


def test_runmodelS_api_with_snow(oracle):
    w = {k: v[:240] for k, v in S.load_weather().items()}
    vegp = S.vegp_deciduous(); soilp = api.soilinit("Loam")
    T = 240
    vegp_t = dict(vegp, PAI=np.repeat(vegp["PAI"][:, None], T, axis=1), pLAI=np.repeat(vegp["pLAI"][:, None], T, axis=1))
    rng = np.random.default_rng(0)
    sd = np.where(w["temp"] < 4, 30.0, 0.0)
    nmrout = dict(soiltemps=dict(D0cm=w["temp"] + 0.5), soilmoist=dict(WC0cm=np.full(T, 0.3)), metout=dict(SNOWDEP=sd),
                  snowtemp={"SN%d" % q: rng.uniform(-8, 0, T) for q in range(1, 9)})
    df = api.runmodelS(w, vegp_t, soilp, nmrout, 1.0, 50, -5)
    assert list(df.columns) == ["obs_time", "Tref", "Tloc", "tleaf", "RHref", "RHloc", "RSWloc", "RLWloc", "windspeed", "dp"]
    veg, soil, m, sm = api._pack([vegp], [soilp])
    want = oracle.run_steady(veg, soil, S.steady_clim_from_climdata(w), api._nmr_arrays(nmrout, T), np.array([50.]), np.array([-5.]),
                             np.array([1.0, 1, 2, 1, 0.95]), m, sm)["out"][:, :, 0]
    assert np.max(np.abs(df["Tloc"].values - want[:, 0])) <= 1e-6
    assert np.max(np.abs(df["RLWloc"].values - want[:, 4]) / np.maximum(np.abs(want[:, 4]), 1e-3)) <= 1e-8
    assert sd.max() > 0 and np.all(df["windspeed"].values[1.0 < sd / 100] == 0) if np.any(1.0 < sd / 100) else True


def test_multi_gpu_split_equals_single_gpu(oracle):
    from microclimc_b200._lib import Context, lib
    if lib().mc_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    inp = transient_inputs(oracle, ncol=1001, T=24)
    cfg = make_cfg(spin_steps=20)
    res = []
    for g in (1, 2):
        ctx = Context(M, SM, g)
        ctx.set_columns(inp["veg"], inp["soil"], inp["lat"], inp["lon"]); ctx.set_forcing(inp["f"], inp["fs"], inp["fo"]); ctx.set_config(cfg)
        r = ctx.run_transient(samp_idx=np.array([0, 500, 1000]))
        res.append((ctx.get_state(), r["series"], r["stats"]))
    for a, b in zip(res[0], res[1]):
        np.testing.assert_array_equal(a, b)


def test_full_size_properties_one_million_columns(oracle):
    """BASELINE config 3 size (10^6 columns, 20 layers) for a day: every column finite, replicated columns agree bit for
    bit wherever they sit in the grid, and a sample of columns matches the oracle."""
    ncol, T = 1_000_000, 24
    base = transient_inputs(oracle, ncol=1000, T=T)
    rep = ncol // 1000
    big = {k: np.ascontiguousarray(np.tile(base[k], (1, rep))) for k in ("veg", "soil", "fs", "fo")}
    from microclimc_b200._lib import Context
    ctx = Context(M, SM)
    ctx.set_columns(big["veg"], big["soil"], np.full(ncol, 50.0), np.full(ncol, -5.0))
    ctx.set_forcing(base["f"], big["fs"], big["fo"]); ctx.set_config(make_cfg(spin_steps=30))
    samp = np.r_[0:16, 999_000:999_016]
    r = ctx.run_transient(samp_idx=samp)
    assert np.count_nonzero(r["flags"] & 1) == 0 and np.all(np.isfinite(r["stats"]))
    st = r["stats"].reshape(6, rep, 1000)
    assert np.all(st == st[:, :1, :])                                    # position in the grid does not matter
    want = oracle.run_transient(base["veg"][:, :16], base["soil"][:, :16], base["f"], base["lat"][:16], base["lon"][:16],
                                make_cfg(spin_steps=30), M, SM, fscale=base["fs"][:, :16], foffset=base["fo"][:, :16])
    assert_series_close(r["series"][:, :, :16], want["series"], "1M grid sample")
    assert_series_close(r["series"][:, :, 16:], want["series"], "1M grid sample (far end)")


def _ulp_err(got, want):
    """|got - want| in units of the spacing of `want` (want computed in extended precision where available)."""
    want = np.asarray(want, dtype=np.longdouble)
    sp = np.spacing(np.abs(want.astype(np.float64))).astype(np.longdouble)
    return np.max(np.abs(np.asarray(got, dtype=np.longdouble) - want) / sp)


def test_fast_fp64_primitives_are_within_two_ulp():
    """The kernels replace IEEE division and libm exp/log/sqrt by branch-free sequences (mc_phys.cuh); bound them in ulp
    over the ranges this path uses, and check the special values the code relies on.  Measured on B200 (tools/ulp_probe.py, round 2):
    exp 1.26, log 1.45, 1/x 0.505, sqrt 0.75, x/y 1.45 ulp (the quotient is a * (1/b) without the exact-residual step since r02n)."""
    from microclimc_b200._lib import Context
    ctx = Context(20, 10, 1, [0])
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-5, 5, 200000), rng.uniform(-1e-3, 1e-3, 50000)])
    assert _ulp_err(ctx.eval_primitive("exp", x), np.exp(x.astype(np.longdouble))) <= 2.0
    sp = ctx.eval_primitive("exp", np.array([-800.0, 800.0, np.nan, 0.0, -708.5]))
    assert 0.0 <= sp[0] < 4e-308 and np.isinf(sp[1]) and np.isnan(sp[2]) and sp[3] == 1.0 and 0.0 <= sp[4] < 4e-308
    p = np.concatenate([rng.uniform(1e-6, 1e6, 200000), np.exp(rng.uniform(-300, 300, 100000)), rng.uniform(0.5, 2.0, 100000)])
    assert _ulp_err(ctx.eval_primitive("log", p), np.log(p.astype(np.longdouble))) <= 2.0
    assert _ulp_err(ctx.eval_primitive("rcp", p), 1 / p.astype(np.longdouble)) <= 1.0
    assert _ulp_err(ctx.eval_primitive("sqrt", p), np.sqrt(p.astype(np.longdouble))) <= 1.0
    q = np.concatenate([rng.uniform(-1e3, 1e3, 300000), rng.uniform(-1, 1, 100000)])
    assert _ulp_err(ctx.eval_primitive("div", q, p), q.astype(np.longdouble) / p.astype(np.longdouble)) <= 1.5
    assert ctx.eval_primitive("sqrt", np.array([0.0]))[0] == 0.0
