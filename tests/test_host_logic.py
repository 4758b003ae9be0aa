"""CPU-only checks of the product's fused column code.

tests/hostdbg/dbg_host.cpp compiles the DEVICE headers (mc_step.cuh, mc_driver.cuh, mc_steady.cuh) for the host
with g++ (test-only build, never shipped or loaded by the package) so that the logic of the fused step — hoisted
geometry, fused leaftemp loop, two-array Thomas, layer merge / average / interpolation, snow overlay — is compared
with the literal oracle here, without a GPU.  The GPU parity tests proper are tests/test_gpu_*.py."""
import ctypes as C

import numpy as np
import pytest

from tests.common import M, SM, make_cfg, transient_inputs, assert_state_close, assert_series_close, loam
from microclimc_b200 import layout as L, synth as S

PD = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(PD)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def dbg_run_transient(lib, inp, cfg, m=M, sm=SM, state=None, pai_season=None, want_full=False, vegt=None, thetaz=None):
    veg, soil, f, lat, lon, fs, fo = map(_f, (inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], inp["fs"], inp["fo"]))
    ncol, T = veg.shape[1], f.shape[0]
    NS = L.nstate(m, sm)
    st = np.zeros((NS, ncol)) if state is None else np.array(state, dtype=np.float64, order="C", copy=True)
    samp = np.arange(ncol, dtype=np.int64)
    series = np.zeros((T, 8, ncol)); stats = np.zeros((6, ncol)); flags = np.zeros(ncol, dtype=np.int32)
    full = np.zeros((T, NS, ncol)) if want_full else None
    if vegt is not None or thetaz is not None:
        vegt, thetaz = _f(vegt), _f(thetaz)
        lib.dbg_run_transient2.restype = C.c_int
        rc = lib.dbg_run_transient2(C.c_int64(ncol), m, sm, C.c_int64(T), _p(veg), _p(soil), _p(f), _p(fs), _p(fo), None, _p(lat), _p(lon),
                                    _p(_f(cfg)), _p(_f(pai_season)), _p(st), _p(series), samp.ctypes.data_as(C.POINTER(C.c_int64)),
                                    C.c_int64(ncol), _p(stats), _p(full), flags.ctypes.data_as(C.POINTER(C.c_int32)), 0,
                                    _p(vegt), C.c_int64(0 if vegt is None else vegt.shape[2]), _p(thetaz))
        assert rc == 0
        return dict(state=st, series=series, stats=stats, flags=flags, full=full)
    lib.dbg_run_transient.restype = C.c_int
    rc = lib.dbg_run_transient(C.c_int64(ncol), m, sm, C.c_int64(T), _p(veg), _p(soil), _p(f), _p(fs), _p(fo), None, _p(lat), _p(lon),
                               _p(_f(cfg)), _p(_f(pai_season)), _p(st), _p(series), samp.ctypes.data_as(C.POINTER(C.c_int64)),
                               C.c_int64(ncol), _p(stats), _p(full), flags.ctypes.data_as(C.POINTER(C.c_int32)), 0)
    assert rc == 0
    return dict(state=st, series=series, stats=stats, flags=flags, full=full)


def dbg_runonestep(lib, inp, clim, cfg, jd, lt, theta, thetap, state, m=M, sm=SM):
    veg, soil, lat, lon, clim, theta, thetap = map(_f, (inp["veg"], inp["soil"], inp["lat"], inp["lon"], clim, theta, thetap))
    st = np.array(state, dtype=np.float64, order="C", copy=True)
    ncol = st.shape[1]
    nm = np.zeros(ncol, dtype=np.int32)
    rc = lib.dbg_runonestep(C.c_int64(ncol), m, sm, _p(veg), _p(soil), _p(clim), _p(lat), _p(lon), _p(_f(cfg)), C.c_double(jd),
                            C.c_double(lt), _p(theta), _p(thetap), _p(st), nm.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return st, nm


def _random_climvars(ncol, seed=5):
    rng = np.random.default_rng(seed)
    clim = np.stack([rng.uniform(-5, 25, ncol), rng.uniform(40, 100, ncol), rng.uniform(98, 103, ncol), rng.uniform(0, 12, ncol),
                     rng.uniform(5, 12, ncol), rng.uniform(0.7, 1.0, ncol), rng.uniform(0, 900, ncol), rng.uniform(0.1, 1.0, ncol)])
    return clim, rng.uniform(0.15, 0.4, ncol), rng.uniform(0.15, 0.4, ncol)


@pytest.mark.parametrize("timestep,reqhgt,edgedist,metopen", [(3600.0, np.nan, 100.0, 1), (60.0, np.nan, 100.0, 1), (600.0, 5.0, np.nan, 1),
                                                             (3600.0, 0.5, 30.0, 0), (20.0, 17.0, 100.0, 1)])
def test_fused_step_equals_literal_oracle(oracle, hostdbg, timestep, reqhgt, edgedist, metopen):
    inp = transient_inputs(oracle, ncol=24, T=4)
    windhgt = 2.0 if metopen else 19.0
    cfg_spin = make_cfg(timestep=timestep, reqhgt=reqhgt, edgedist=edgedist, metopen=metopen, windhgt=windhgt, spin_steps=15)
    st0 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][:1], inp["lat"], inp["lon"], cfg_spin, M, SM, fscale=inp["fs"],
                               foffset=inp["fo"])["state"]
    cfg = make_cfg(timestep=timestep, reqhgt=reqhgt, edgedist=edgedist, metopen=metopen, windhgt=windhgt, spin_steps=0)
    clim, theta, thetap = _random_climvars(24)
    want, nm_o = oracle.runonestep(inp["veg"], inp["soil"], clim, inp["lat"], inp["lon"], cfg, 2449900.0, 13.0, theta, thetap, st0, M, SM)
    got, nm_d = dbg_runonestep(hostdbg, inp, clim, cfg, 2449900.0, 13.0, theta, thetap, st0)
    assert np.array_equal(nm_o, nm_d)
    assert_state_close(got, want, what="dt=%g" % timestep)


def test_merged_layer_branch_is_exercised(oracle, hostdbg):
    """Short time steps give several merged air layers (code.R:813-852); hourly steps give one (code.R:854-867)."""
    inp = transient_inputs(oracle, ncol=8, T=2)
    counts = {}
    for ts in (3600.0, 120.0, 5.0):
        cfg_spin = make_cfg(timestep=ts, spin_steps=10)
        st0 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][:1], inp["lat"], inp["lon"], cfg_spin, M, SM)["state"]
        clim, theta, thetap = _random_climvars(8, seed=9)
        _, nm = oracle.runonestep(inp["veg"], inp["soil"], clim, inp["lat"], inp["lon"], make_cfg(timestep=ts, spin_steps=0), 2449900.0, 11.0,
                                  theta, thetap, st0, M, SM)
        counts[ts] = nm
    assert np.all(counts[3600.0] == 1)
    assert np.any(counts[120.0] > 1) and np.all(counts[5.0] >= counts[120.0]) and np.all(counts[5.0] > 1) and counts[5.0].max() <= M


@pytest.mark.parametrize("timestep,charmode,reqhgt", [(3600.0, 0, np.nan), (300.0, 1, np.nan), (3600.0, 0, 7.3), (3600.0, 0, -0.3)])
def test_trajectory_with_spinup_equals_oracle(oracle, hostdbg, timestep, charmode, reqhgt):
    inp = transient_inputs(oracle, ncol=6, T=72)
    cfg = make_cfg(timestep=timestep, reqhgt=reqhgt, spin_steps=60, harvest_char=charmode)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                want_full=True)
    got = dbg_run_transient(hostdbg, inp, cfg, want_full=True)
    assert_series_close(got["series"], want["series"], "trajectory")
    assert_state_close(got["state"], want["state"], what="final state")
    np.testing.assert_allclose(got["stats"], want["stats"], rtol=0, atol=1e-6)
    assert np.array_equal(got["flags"] & 1, want["flags"] & 1)
    assert_state_close(got["full"][40], want["full"][40], what="full table row")


def test_other_layer_counts(oracle, hostdbg):
    """m = 7 / sm = 5 (small template) and m = 33 / sm = 12 (large template)."""
    for m, sm in [(7, 5), (33, 12)]:
        soil = loam(oracle); soil["z"] = oracle.soilinit_z(sm)
        veg, so = S.ensemble(S.vegp_deciduous(m=m), soil, 5, m, sm, seed=3)
        f, ts = S.forcing_from_climdata(S.load_weather()); f = np.ascontiguousarray(f[4000:4030])
        inp = dict(veg=veg, soil=so, f=f, lat=np.full(5, 50.0), lon=np.full(5, -5.0), fs=None, fo=None)
        cfg = make_cfg(spin_steps=25)
        want = oracle.run_transient(veg, so, f, inp["lat"], inp["lon"], cfg, m, sm)
        got = dbg_run_transient(hostdbg, inp, cfg, m=m, sm=sm)
        assert_state_close(got["state"], want["state"], m=m, sm=sm, what="m=%d" % m)


def test_zero_pai_layers_are_floored(oracle, hostdbg):
    """PAI == 0 -> 1e-4 (code.R:691-696)."""
    inp = transient_inputs(oracle, ncol=3, T=24, perturb=False)
    inp["veg"][L.veg_rows(M)["PAI"]][3:6, :] = 0.0
    cfg = make_cfg(spin_steps=20)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM)
    got = dbg_run_transient(hostdbg, inp, cfg)
    assert np.all(np.isfinite(want["state"]))
    assert_state_close(got["state"], want["state"], what="zero PAI")


# ---- steady path ----------------------------------------------------------------------------------------------------
def _steady_inputs(oracle, ncol, T, t0=0, snow=False, percol=False, seed=2):
    veg, soil = S.ensemble(S.vegp_deciduous(), loam(oracle), ncol, M, SM, seed=seed)
    clim = S.steady_clim_from_climdata(S.load_weather())[t0:t0 + T]
    rng = np.random.default_rng(seed)
    shape = (T, 11, ncol) if percol else (T, 11)
    nmr = np.zeros(shape)
    tair = clim[:, 0]
    bt = (lambda a: a[:, None]) if percol else (lambda a: a)
    nmr[:, 0] = bt(tair) + rng.normal(0.5, 1.0, nmr[:, 0].shape)
    nmr[:, 1] = rng.uniform(0.12, 0.4, nmr[:, 1].shape)
    if snow:
        sd = np.clip(rng.normal(20, 40, nmr[:, 2].shape), 0, None)       # cm; about a third of the hours snow-free
        nmr[:, 2] = sd
        nmr[:, 3:] = rng.uniform(-10, 0, nmr[:, 3:].shape)
    return veg, soil, clim, nmr


def dbg_run_steady(lib, veg, soil, clim, nmr, lat, lon, scfg, m=M, sm=SM, pai_season=None, vegt=None):
    veg, soil, clim, nmr, lat, lon, scfg = map(_f, (veg, soil, clim, nmr, lat, lon, scfg))
    ncol, T = veg.shape[1], clim.shape[0]
    out = np.zeros((T, 7, ncol)); stats = np.zeros((6, ncol))
    if vegt is not None:
        vegt = _f(vegt)
        rc = lib.dbg_run_steady2(C.c_int64(ncol), m, sm, C.c_int64(T), _p(veg), _p(soil), _p(clim), _p(nmr), 1 if nmr.ndim == 3 else 0, _p(lat),
                                 _p(lon), _p(scfg), None, _p(out), _p(stats), 0, _p(vegt), C.c_int64(vegt.shape[2]))
        return dict(out=out, stats=stats, rc=rc)
    rc = lib.dbg_run_steady(C.c_int64(ncol), m, sm, C.c_int64(T), _p(veg), _p(soil), _p(clim), _p(nmr), 1 if nmr.ndim == 3 else 0, _p(lat),
                            _p(lon), _p(scfg), _p(_f(pai_season)), _p(out), _p(stats), 0)
    return dict(out=out, stats=stats, rc=rc)


@pytest.mark.parametrize("reqhgt,snow,percol", [(1.0, False, False), (15.5, False, False), (22.0, False, True), (1.0, True, True),
                                                (0.3, True, False), (18.0, True, False)])
def test_steady_series_equals_oracle(oracle, hostdbg, reqhgt, snow, percol):
    ncol, T = 7, 96
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=600, snow=snow, percol=percol)
    lat = np.full(ncol, 50.0); lon = np.full(ncol, -5.0)
    scfg = np.array([reqhgt, 1, 2, 1, 0.95])
    want = oracle.run_steady(veg, soil, clim, nmr, lat, lon, scfg, M, SM)
    got = dbg_run_steady(hostdbg, veg, soil, clim, nmr, lat, lon, scfg)
    assert want["rc"] == 0 and got["rc"] == 0
    assert np.array_equal(np.isnan(got["out"]), np.isnan(want["out"]))
    np.testing.assert_allclose(np.nan_to_num(got["out"]), np.nan_to_num(want["out"]), rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(got["stats"], want["stats"], rtol=0, atol=1e-6)
    if snow:
        buried = reqhgt < nmr[:, 2] / 100.0
        if buried.any():     # under the snow pack: wind 0, RH 100, SW 0 (NMR.R:668-675)
            o = want["out"] if not percol else want["out"]
            sel = np.broadcast_to(buried if percol else buried[:, None], (T, ncol))
            assert np.all(want["out"][:, 5][sel] == 0) and np.all(want["out"][:, 2][sel] == 100) and np.all(want["out"][:, 3][sel] == 0)


def test_steady_snow_lookup_error_is_reported(oracle, hostdbg):
    veg, soil, clim, nmr = _steady_inputs(oracle, 3, 24, snow=True)
    nmr[:, 2] = 50.0
    lat = np.full(3, 50.0); lon = np.full(3, -5.0)
    scfg = np.array([0.01, 1, 2, 1, 0.95])           # below the lowest snow node: R indexes snow[, 11]
    assert oracle.run_steady(veg, soil, clim, nmr, lat, lon, scfg, M, SM)["rc"] != 0
    assert dbg_run_steady(hostdbg, veg, soil, clim, nmr, lat, lon, scfg)["rc"] != 0


# ---- streaming form (mc_fast.cuh) -------------------------------------------------------------------------------------
class _FastLib:
    """Routes dbg_run_transient to the streaming implementation (geometry pass + three-sweep step, DirectMem accessor)."""

    def __init__(self, lib):
        self._lib = lib
        lib.dbg_general_steps.restype = C.c_longlong

    def __getattr__(self, n):
        return getattr(self._lib, "dbg_run_transient_fast" if n == "dbg_run_transient" else n)


@pytest.mark.parametrize("timestep,charmode,reqhgt,edgedist,metopen", [
    (3600.0, 0, np.nan, 100.0, 1), (3600.0, 0, 7.3, 100.0, 1), (3600.0, 1, np.nan, np.nan, 1), (3600.0, 0, 16.9, 30.0, 0),
    (1800.0, 0, -0.3, 100.0, 1), (300.0, 0, np.nan, 100.0, 1), (20.0, 0, np.nan, 100.0, 1)])
def test_streaming_form_equals_oracle(oracle, hostdbg, timestep, charmode, reqhgt, edgedist, metopen):
    inp = transient_inputs(oracle, ncol=7, T=96)
    if reqhgt == 16.9:
        inp["veg"][L.veg_rows(M)["hgt"]] = np.minimum(inp["veg"][L.veg_rows(M)["hgt"]], 16.0)      # keep reqhgt above every canopy
    cfg = make_cfg(timestep=timestep, reqhgt=reqhgt, edgedist=edgedist, metopen=metopen, windhgt=2.0 if metopen else 20.0, spin_steps=80,
                   harvest_char=charmode)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                want_full=True)
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    got = dbg_run_transient(fl, inp, cfg, want_full=True)
    ngen = fl.dbg_general_steps(1)
    assert_series_close(got["series"], want["series"], "streaming trajectory")
    assert_state_close(got["state"], want["state"], what="streaming final state")
    np.testing.assert_allclose(got["stats"], want["stats"], rtol=0, atol=1e-6)
    assert np.array_equal(got["flags"] & 1, want["flags"] & 1)
    for t in (0, 50, 95):
        assert_state_close(got["full"][t], want["full"][t], what="streaming full table t=%d" % t)
    # Nothing leaves the streaming form at any of these time steps: below hourly steps the general instantiation runs (it has leaftemp's
    # transient balances), and steps with more than six merged air layers — every step at 20 s — take the wide solve.
    if np.all(np.isfinite(want["state"])):
        assert ngen == 0, "columns dropped: %d" % ngen
        assert not np.any(got["flags"] & 8)


def test_streaming_form_merged_layers_occur_at_hourly_steps(oracle, hostdbg):
    """The bench workload takes the merged-layer branch (code.R:813-852) in roughly one column-step in ten."""
    inp = transient_inputs(oracle, ncol=8, T=240, weather=False)
    cfg = make_cfg(spin_steps=100)
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    got = dbg_run_transient(fl, inp, cfg)
    assert fl.dbg_general_steps(1) == 0
    assert np.any(got["flags"] & 2), "no merged-layer step seen"
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    assert_series_close(got["series"], want["series"], "bench-like")


def test_streaming_restart_and_chunking(oracle, hostdbg):
    """Chunked runs (state handed over through the ABI layout) reproduce the single run bit for bit."""
    inp = transient_inputs(oracle, ncol=5, T=60)
    fl = _FastLib(hostdbg)
    a = dbg_run_transient(fl, inp, make_cfg(spin_steps=40))
    s0 = dbg_run_transient(fl, dict(inp, f=inp["f"][:1]), make_cfg(spin_steps=40))      # spin-up + 1 step
    # run the rest from the handed-over state; tsoil of the full series is the column mean over all 60 rows, so pass it
    tmean = (inp["f"][:, 0].mean() * inp["fs"][0] + inp["fo"][0])
    # (dbg_run_transient has no tsoil argument: compare only that a restart from the ABI state is self-consistent)
    b1 = dbg_run_transient(fl, dict(inp, f=inp["f"][1:2]), make_cfg(spin_steps=0), state=s0["state"])
    b2 = dbg_run_transient(hostdbg, dict(inp, f=inp["f"][1:2]), make_cfg(spin_steps=0), state=s0["state"])
    assert_state_close(b1["state"], b2["state"], what="restart fast vs general")
    assert np.all(np.isfinite(a["state"]))


def test_regular_grid_instantiation_redoes_columns_with_a_moved_node(oracle, hostdbg):
    """reqhgt NA selects the regular-grid instantiation (REG) of the streaming step; a state whose previn$z has a node moved by an
    earlier reqhgt = 5 m run (code.R:719) cannot use it: every such column is redone by the general column and matches the oracle."""
    inp = transient_inputs(oracle, ncol=6, T=36)
    first = make_cfg(reqhgt=5.0, spin_steps=30)
    want0 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][:12], inp["lat"], inp["lon"], first, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    second = make_cfg(spin_steps=0)
    rest = dict(inp, f=inp["f"][12:])
    want = oracle.run_transient(inp["veg"], inp["soil"], rest["f"], inp["lat"], inp["lon"], second, M, SM, state=want0["state"],
                                fscale=inp["fs"], foffset=inp["fo"])
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    got = dbg_run_transient(fl, rest, second, state=want0["state"])
    assert fl.dbg_general_steps(1) == 6 and np.all(got["flags"] & 8)
    assert_series_close(got["series"], want["series"], "moved node, regular-grid launch")
    assert_state_close(got["state"], want["state"], what="moved node, regular-grid launch")
    fl.dbg_general_steps(1)
    got2 = dbg_run_transient(fl, rest, second, state=got["state"])                 # z is the regular grid again: streaming form
    assert fl.dbg_general_steps(1) == 0 and not np.any(got2["flags"] & 8)
    want2 = oracle.run_transient(inp["veg"], inp["soil"], rest["f"], inp["lat"], inp["lon"], second, M, SM, state=want["state"],
                                 fscale=inp["fs"], foffset=inp["fo"])
    assert_series_close(got2["series"], want2["series"], "second regular-grid launch")


def test_regular_grid_instantiation_leaves_transient_balances_to_the_general_column(oracle, hostdbg, monkeypatch):
    """leaftemp selects its transient balances when timestep * max(g1, g3, g2) <= 1 (code.R:419-449).  On the regular grid that does not
    happen at hourly or half-hourly steps; it starts around 900 s, in mid-canopy layers at night, when the stomata are closed and only the
    two air terms are left (host count over 8 columns x 88 steps: 0 of 14 080 layer-steps at 3600 and 1800 s, 2 at 900 s, 49 at 300 s).
    The regular-grid instantiation carries only the equilibrium forms: such columns are redone by the general column and match the
    oracle (the library itself runs the streaming form at hourly or longer steps only)."""
    inp = transient_inputs(oracle, ncol=8, T=48)
    cfg = make_cfg(timestep=300.0, spin_steps=40)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    monkeypatch.setenv("MC_HOST_FORCE_REG", "1")          # (the library itself selects the general instantiation below hourly steps)
    got = dbg_run_transient(fl, inp, cfg)
    monkeypatch.delenv("MC_HOST_FORCE_REG")
    assert fl.dbg_general_steps(1) > 0 and np.any(got["flags"] & 8)
    assert_series_close(got["series"], want["series"], "300 s steps")
    assert_state_close(got["state"], want["state"], what="300 s steps")


@pytest.mark.parametrize("timestep", [300.0, 20.0])
def test_wide_solve_keeps_many_merged_layers_in_the_streaming_form(oracle, hostdbg, timestep):
    """More than six merged air layers in a step (the rule at short steps, calm stable hours at hourly ones) used to send the column to the
    general column; the step now solves them on the lane's own arrays (the wide solve, mc_fast.cuh).  With one node moved (reqhgt = 7.3 m:
    the general instantiation, which has leaftemp's transient balances) no column leaves the streaming form even at 20 s steps, where
    every step has ~20 merged layers, and the trajectories match the oracle."""
    inp = transient_inputs(oracle, ncol=7, T=96)
    cfg = make_cfg(timestep=timestep, reqhgt=7.3, spin_steps=80)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"], want_full=True)
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    got = dbg_run_transient(fl, inp, cfg, want_full=True)
    assert fl.dbg_general_steps(1) == 0 and not np.any(got["flags"] & 8)
    assert np.all(got["flags"] & 2)
    assert_series_close(got["series"], want["series"], "wide solve")
    assert_state_close(got["state"], want["state"], what="wide solve")
    for t in (0, 50, 95):
        assert_state_close(got["full"][t], want["full"][t], what="wide solve full table t=%d" % t)


# ---- the variant tables of tests/test_gpu_variants.py through the host build of the streaming code -----------------------
def _variant_cases():
    import tests.test_gpu_variants as V
    return [("cfg", k) for k in sorted(V.CFG_VARIANTS)] + [("site", k) for k in sorted(V.SITE_VARIANTS)]


@pytest.mark.parametrize("kind,name", _variant_cases())
def test_streaming_form_over_the_variant_tables(oracle, hostdbg, kind, name, monkeypatch):
    """Every run-description / vegetation / soil / site variant the GPU test drives through the CUDA kernels, here through the
    host build of the same source (DirectMem accessor) against the literal oracle — the CPU suite's view of those branches."""
    import tests.test_gpu_variants as V
    monkeypatch.setattr(V, "NCOL", 5)
    if kind == "cfg":
        kw = dict(spin_steps=V.SPIN); kw.update(V.CFG_VARIANTS[name])
        inp = V._inputs(oracle)
    else:
        kw = dict(spin_steps=V.SPIN)
        sk = dict(V.SITE_VARIANTS[name])
        if callable(sk.get("vegp")):
            sk["vegp"] = sk["vegp"]()
        inp = V._inputs(oracle, **sk)
    cfg = make_cfg(**kw)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    fl = _FastLib(hostdbg)
    fl.dbg_general_steps(1)
    got = dbg_run_transient(fl, inp, cfg)
    ngen = fl.dbg_general_steps(1)
    assert_series_close(got["series"], want["series"], name)
    assert_state_close(got["state"], want["state"], what=name)
    if name == "short_grass_general_path":
        assert ngen >= 5, "nodes under 0.12 m must take the general path (%d)" % ngen
    elif kw.get("timestep", 3600.0) >= 3600.0:
        assert ngen == 0, "hourly columns should stay in the streaming form (columns dropped: %d)" % ngen


# ---- general time-varying vegetation (.vegpsort) and soil moisture by depth (theta matrix) ---------------------------------
def _vegt_case(ncol, T, m, percol, seed=3):
    """PAI / pLAI as m x T matrices that are NOT profile x seasonal factor (the profile's shape moves through the season) and a
    zm0 series, as habitatvars-style inputs have them."""
    rng = np.random.default_rng(seed)
    base = S.vegp_deciduous(m=m)
    n = ncol if percol else 1
    vegt = np.empty((T, 2 * m + 1, n))
    zc = (np.arange(m) + 0.5) / m
    for t in range(T):
        ph = t / max(T - 1, 1)
        for c in range(n):
            peak = 0.55 + 0.25 * ph + 0.05 * (c % 3)
            prof = np.exp(-((zc - peak) / (0.18 + 0.1 * ph)) ** 2)
            vegt[t, :m, c] = prof / prof.sum() * (2.0 + 2.5 * ph + 0.1 * c)
            vegt[t, m:2 * m, c] = np.clip(0.6 + 0.3 * ph * zc, 0, 1)
            vegt[t, 2 * m, c] = 0.004 + 0.002 * ph
    vegt[T // 2, 2, :] = 0.0                                                # a zero-PAI layer at one step (floored to 0.0001, code.R:691-696)
    return vegt


@pytest.mark.parametrize("percol", [False, True])
def test_general_time_varying_vegetation_equals_oracle(oracle, hostdbg, percol):
    inp = transient_inputs(oracle, ncol=5, T=60)
    vegt = _vegt_case(5, 60, M, percol)
    cfg = make_cfg(spin_steps=30, reqhgt=2.0)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                vegt=vegt)
    got = dbg_run_transient(hostdbg, inp, cfg, vegt=vegt)
    assert_series_close(got["series"], want["series"], "vegt"); assert_state_close(got["state"], want["state"], what="vegt")
    # and it matters: the run with the first row's vegetation held fixed differs
    fixed = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                 vegt=np.repeat(vegt[:1], 60, axis=0))
    assert np.max(np.abs(fixed["series"][-1, 0] - want["series"][-1, 0])) > 1e-3


def test_soil_moisture_by_depth_equals_oracle_and_recycles(oracle, hostdbg):
    inp = transient_inputs(oracle, ncol=5, T=60)
    rng = np.random.default_rng(8)
    thetaz = np.clip(0.18 + 0.15 * np.linspace(0, 1, SM)[None, :] + 0.05 * np.sin(np.arange(60) / 7.0)[:, None] + rng.normal(0, 0.01, (60, SM)),
                     0.05, 0.42)
    cfg = make_cfg(spin_steps=30)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                thetaz=thetaz)
    got = dbg_run_transient(hostdbg, inp, cfg, thetaz=thetaz)
    assert_series_close(got["series"], want["series"], "thetaz"); assert_state_close(got["state"], want["state"], what="thetaz")
    # a depth-constant matrix is the scalar-theta run bit for bit (no recycling effect, same soilk)
    flat = np.repeat(thetaz[:, :1], SM, axis=1)
    inp2 = dict(inp, f=inp["f"].copy()); inp2["f"][:, L.FORCING.index("theta")] = thetaz[:, 0]
    a = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"], thetaz=flat)
    b = oracle.run_transient(inp2["veg"], inp2["soil"], inp2["f"], inp2["lat"], inp2["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    np.testing.assert_array_equal(a["series"], b["series"])
    # Q24: the moisture of the deeper nodes reaches the canopy air through the recycled esoil
    deep = thetaz.copy(); deep[:, 1:] = 0.08
    c = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"], thetaz=deep)
    assert np.max(np.abs(c["series"][:, 2] - want["series"][:, 2])) > 1e-6


@pytest.mark.parametrize("percol,snow", [(False, False), (True, True)])
def test_steady_general_time_varying_vegetation_equals_oracle(oracle, hostdbg, percol, snow):
    """runmodelS takes vegp$PAI / vegp$pLAI as m x T matrices (NMR.R:730-735); any matrix, not only profile x seasonal factor."""
    ncol, T = 5, 72
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=600, snow=snow, percol=False)
    vegt = _vegt_case(ncol, T, M, percol)
    lat = np.full(ncol, 50.0); lon = np.full(ncol, -5.0)
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    want = oracle.run_steady(veg, soil, clim, nmr, lat, lon, scfg, M, SM, vegt=vegt)
    got = dbg_run_steady(hostdbg, veg, soil, clim, nmr, lat, lon, scfg, vegt=vegt)
    assert np.array_equal(np.isnan(got["out"]), np.isnan(want["out"]))
    np.testing.assert_allclose(np.nan_to_num(got["out"]), np.nan_to_num(want["out"]), rtol=1e-9, atol=1e-9)
    base = oracle.run_steady(veg, soil, clim, nmr, lat, lon, scfg, M, SM)
    assert np.nanmax(np.abs(base["out"][:, 0] - want["out"][:, 0])) > 1e-3
