"""GPU parity cases added in round 2 ("untested on the GPU"): transient output above the canopy (code.R:722-725,
quirk Q13 at 1235-1240), `runmodel(snow = 1)` end to end (code.R:918-933: a very different operating point of the step), the
bench workload itself (BASELINE config 3 shapes) over a fortnight with sampled columns against the oracle, columns that leave
the streaming form, and the in-process multi-GPU split."""
import numpy as np
import pytest

from tests.common import M, SM, make_cfg, transient_inputs, assert_state_close, assert_series_close, loam
from microclimc_b200 import layout as L, synth as S

pytestmark = pytest.mark.gpu


def _ctx(inp, cfg, n_gpus=1):
    from microclimc_b200._lib import Context
    ctx = Context(M, SM, n_gpus)
    ctx.set_columns(inp["veg"], inp["soil"], inp["lat"], inp["lon"])
    ctx.set_forcing(inp["f"], inp["fs"], inp["fo"])
    ctx.set_config(cfg)
    return ctx


@pytest.mark.parametrize("reqhgt", [17.0, 25.0])
def test_transient_output_above_canopy(oracle, reqhgt):
    """reqhgt > hgt: zabove <- reqhgt (code.R:722-725) changes the canopy-top temperature's reference height, and the harvest takes
    tabove / the first row's relhum, Rsw, skyem (quirk Q13, code.R:1235-1240)."""
    inp = transient_inputs(oracle, ncol=48, T=120)
    hg = L.veg_rows(M)["hgt"]
    inp["veg"][hg] = np.minimum(inp["veg"][hg], 16.0)                      # every canopy below the output height
    cfg = make_cfg(reqhgt=reqhgt, spin_steps=60)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    ctx = _ctx(inp, cfg)
    got = ctx.run_transient(samp_idx=np.arange(48))
    st = ctx.get_state()
    assert np.all(st[L.state_rows(M, SM)["zabove"]] == reqhgt)
    assert_series_close(got["series"], want["series"], "reqhgt %g above canopy" % reqhgt)
    assert_state_close(st, want["state"], what="reqhgt above canopy")
    assert np.count_nonzero(got["flags"] & 8) <= 2, "hourly columns stay in the streaming form (a node snapped above a short canopy may not)"


def test_transient_runmodel_with_snow_on_the_canopy(oracle):
    """runmodel(snow = 1): .snowsort (code.R:918-933) swaps in lw*1.5, zm0 0.024, refls 0.85, refg 0.8, refw 0.7, gsmax 10, q50 1 for the
    whole run (quirk Q3) — through the public host mirror (api.runmodel_batch) against the oracle fed the swapped parameters."""
    from microclimc_b200 import api
    w = S.load_weather()
    T = 96
    clim = {k: (np.asarray(v)[300:300 + T]) for k, v in w.items()}          # midwinter
    soil = loam(oracle)
    base = S.vegp_deciduous()
    vl = []
    rng = np.random.default_rng(4)
    for c in range(24):
        v = dict(base); v["PAI"] = np.asarray(base["PAI"]) * rng.uniform(0.7, 1.3); v["hgt"] = base["hgt"] * rng.uniform(0.8, 1.2)
        vl.append(v)
    r = api.runmodel_batch(clim, vl, [soil] * 24, 50.0, -5.0, reqhgt=1.0, steps=40, snow=1, sample=np.arange(24))
    snowed = [api._snowsort(v, 1, 0) for v in vl]
    veg = S.pack_vegp(snowed, M); so = S.pack_soilp([soil] * 24, SM)
    f, ts = S.forcing_from_climdata(clim, theta=0.3)
    cfg = make_cfg(timestep=ts, reqhgt=1.0, spin_steps=40)
    want = oracle.run_transient(veg, so, f, np.full(24, 50.0), np.full(24, -5.0), cfg, M, SM)
    assert veg[L.veg_rows(M)["gsmax"]][0, 0] == 10.0 and veg[L.veg_rows(M)["refls"]][0, 0] == 0.85
    assert_series_close(r["series"], want["series"], "snow = 1")
    assert_state_close(r["state"], want["state"], what="snow = 1")


def test_bench_workload_fortnight_sampled_columns(oracle):
    """BASELINE config 3 as bench.py / tools/bench_year.py run it (same generator, same seeds), 2 x 10^5 columns x (200 spin-up +
    336 hourly steps) in two weekly calls; 32 columns spread over the grid against the oracle, statistics combined as the year run does."""
    import bench as B
    ncol, hours = 200_000, 336
    w = B.workload(ncol, hours, seed=1234)
    cfg = B.make_cfg(w["ts"], 200)
    samp = np.unique(np.linspace(0, ncol - 1, 32).astype(np.int64))
    from microclimc_b200._lib import Context
    ctx = Context(M, SM)
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"]); ctx.set_forcing(w["f"], w["fs"], w["fo"]); ctx.set_config(cfg)
    ser = []; dropped = 0
    for t0 in (0, 168):
        r = ctx.run_transient(t0, t0 + 168, samp_idx=samp)
        ser.append(r["series"]); dropped += int(np.count_nonzero(r["flags"] & 8))
        assert not np.any(r["flags"] & 1)
    got = np.concatenate(ser)
    want = oracle.run_transient(w["veg"][:, samp], w["soil"][:, samp], w["f"], w["lat"][samp], w["lon"][samp], cfg, M, SM,
                                fscale=w["fs"][:, samp], foffset=w["fo"][:, samp])
    assert_series_close(got, want["series"], "C3 fortnight")
    assert_state_close(ctx.get_state()[:, samp], want["state"], what="C3 fortnight state")
    assert dropped <= ncol // 1000, "the bench workload should stay in the streaming form (%d columns dropped)" % dropped


def test_columns_leaving_the_streaming_form_are_redone_by_the_general_kernel(oracle):
    """A mixed grid: forest columns stay in the streaming form, short-grass columns (nodes under 0.12 m need the generic edge-wind
    profile) leave it at once; those are flagged (bit 3), redone by the general kernel inside the same call — spin-up launch and run
    launch each with its own list — and agree with the oracle like the rest.  A ragged count, so padding lanes are in play."""
    ncol = 300 + 21
    inp = transient_inputs(oracle, ncol=ncol, T=72)
    g = S.vegp_deciduous(hgt=0.4, PAIt=1.5); g.update(x=0.4, lw=0.01, gsmax=0.4, hgtg=0.02)
    vg, _ = S.ensemble(g, loam(oracle), ncol, M, SM, seed=77)
    grass = np.zeros(ncol, bool); grass[::3] = True; grass[-5:] = True
    inp["veg"][:, grass] = vg[:, grass]
    cfg = make_cfg(spin_steps=40)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    ctx = _ctx(inp, cfg)
    got = ctx.run_transient(samp_idx=np.arange(ncol))
    assert np.array_equal((got["flags"] & 8) != 0, grass)
    assert_series_close(got["series"], want["series"], "mixed grid")
    assert_state_close(ctx.get_state(), want["state"], what="mixed grid")
    np.testing.assert_allclose(got["stats"], want["stats"], rtol=0, atol=1e-6)
    assert np.array_equal(got["flags"] & 1, want["flags"] & 1)
    # a second call continues from the device state: the run launch's list alone
    r2 = ctx.run_transient(40, 72, samp_idx=np.arange(ncol))
    assert np.array_equal((r2["flags"] & 8) != 0, grass)


def test_duplicate_sample_columns_are_rejected(oracle):
    from microclimc_b200._lib import MicroclimcError
    inp = transient_inputs(oracle, ncol=9, T=6)
    ctx = _ctx(inp, make_cfg(spin_steps=5))
    with pytest.raises((ValueError, MicroclimcError), match="duplicate"):
        ctx.run_transient(samp_idx=np.array([1, 3, 3]))


def test_in_process_split_over_all_visible_gpus_equals_one_gpu(oracle):
    """The product's own multi-GPU path (mc_create(n_gpus): static contiguous column split, one worker thread per GPU, gathered into
    the caller's arrays) must reproduce the single-GPU run bit for bit — on one GPU the split degenerates to one shard and the
    test still drives the gather path with a ragged column count."""
    from microclimc_b200._lib import lib
    ng = max(1, min(8, lib().mc_device_count()))
    inp = transient_inputs(oracle, ncol=1000 + 37, T=48)
    cfg = make_cfg(spin_steps=30)
    samp = np.array([0, 5, 517, 518, 1036])
    a = _ctx(inp, cfg, 1); ra = a.run_transient(samp_idx=samp)
    b = _ctx(inp, cfg, ng); rb = b.run_transient(samp_idx=samp)
    np.testing.assert_array_equal(a.get_state(), b.get_state())
    for k in ("series", "stats", "flags"):
        np.testing.assert_array_equal(ra[k], rb[k])


# ---- general time-varying vegetation (.vegpsort) and soil moisture by depth, through the C-ABI and the host mirror ----------------------
def _vegt(ncol, T, percol):
    from tests.test_host_logic import _vegt_case
    return _vegt_case(ncol, T, M, percol)


@pytest.mark.parametrize("percol", [False, True])
def test_general_time_varying_vegetation_on_the_gpu(oracle, percol):
    """vegp$PAI / vegp$pLAI as m x T matrices that are not profile x seasonal factor, vegp$zm0 of length T (code.R:904-916): mc_set_vegt."""
    ncol, T = 70, 96
    inp = transient_inputs(oracle, ncol=ncol, T=T)
    vegt = _vegt(ncol, T, percol)
    cfg = make_cfg(spin_steps=40, reqhgt=2.0)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"], vegt=vegt)
    ctx = _ctx(inp, cfg)
    ctx.set_vegt(vegt)
    got = ctx.run_transient(samp_idx=np.arange(ncol))
    assert_series_close(got["series"], want["series"], "vegt"); assert_state_close(ctx.get_state(), want["state"], what="vegt")
    # chunked continuation picks the right rows
    ctx2 = _ctx(inp, cfg); ctx2.set_vegt(vegt)
    a = ctx2.run_transient(0, 50, samp_idx=np.arange(ncol)); b = ctx2.run_transient(50, T, samp_idx=np.arange(ncol))
    np.testing.assert_array_equal(np.concatenate([a["series"], b["series"]]), got["series"])


def test_soil_moisture_by_depth_on_the_gpu(oracle):
    """theta as a [soil layers x time] matrix (code.R:1146-1154): soilk per node, esoil recycled over the canopy layers (quirk Q24)."""
    from microclimc_b200._lib import MicroclimcError
    ncol, T = 70, 96
    inp = transient_inputs(oracle, ncol=ncol, T=T)
    rng = np.random.default_rng(8)
    thetaz = np.clip(0.18 + 0.15 * np.linspace(0, 1, SM)[None, :] + 0.05 * np.sin(np.arange(T) / 7.0)[:, None] + rng.normal(0, 0.01, (T, SM)), 0.05, 0.42)
    cfg = make_cfg(spin_steps=40)
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"], thetaz=thetaz)
    ctx = _ctx(inp, cfg)
    ctx.set_thetaz(thetaz)
    got = ctx.run_transient(samp_idx=np.arange(ncol))
    assert_series_close(got["series"], want["series"], "thetaz"); assert_state_close(ctx.get_state(), want["state"], what="thetaz")
    with pytest.raises((ValueError, MicroclimcError)):
        ctx.set_thetaz(thetaz[:10]); ctx.run_transient()                      # rows must match the forcing


def test_runmodel_with_matrix_vegp_and_theta_matrix_through_the_host_mirror(oracle):
    """The reference's own call shape: runmodel(climdata, vegp with PAI m x T, soilp, ..., theta = matrix) for one column."""
    from microclimc_b200 import api
    w = S.load_weather(); T = 72
    clim = {k: np.asarray(v)[2000:2000 + T] for k, v in w.items()}
    vt = _vegt(1, T, False)[:, :, 0]
    v = dict(S.vegp_deciduous(), PAI=vt[:, :M].T.copy(), pLAI=vt[:, M:2 * M].T.copy(), zm0=vt[:, 2 * M].copy())
    soil = loam(oracle)
    theta = np.clip(0.2 + 0.1 * np.linspace(0, 1, SM)[:, None] + 0.03 * np.cos(np.arange(T) / 5.0)[None, :], 0.05, 0.4)
    df = api.runmodel(clim, v, soil, 50.0, -5.0, reqhgt=1.0, steps=30, theta=theta)
    v0 = dict(v, PAI=v["PAI"][:, 0], pLAI=v["pLAI"][:, 0], zm0=float(v["zm0"][0]))
    veg = S.pack_vegp([v0], M); so = S.pack_soilp([soil], SM)
    f, ts = S.forcing_from_climdata(clim, theta=theta[0])
    want = oracle.run_transient(veg, so, f, np.array([50.0]), np.array([-5.0]), make_cfg(timestep=ts, reqhgt=1.0, spin_steps=30), M, SM,
                                vegt=vt[:, :, None], thetaz=theta.T.copy())
    assert np.max(np.abs(df["tout"].to_numpy() - want["series"][:, 0, 0])) <= 1e-6
    assert np.max(np.abs(df["tleaf"].to_numpy() - want["series"][:, 1, 0])) <= 1e-6


def test_steady_path_with_matrix_vegp_on_the_gpu(oracle):
    from microclimc_b200._lib import Context
    ncol, T = 40, 120
    veg, so = S.ensemble(S.vegp_deciduous(), loam(oracle), ncol, M, SM, seed=5)
    clim = S.steady_clim_from_climdata(S.load_weather())[600:600 + T]
    rng = np.random.default_rng(2)
    nmr = np.zeros((T, 11)); nmr[:, 0] = clim[:, 0] + rng.normal(0, 1.5, T); nmr[:, 1] = rng.uniform(0.1, 0.4, T)
    nmr[:, 2] = np.where(clim[:, 0] < 6, rng.uniform(1, 60, T), 0.0); nmr[:, 3:] = rng.uniform(-9, 0, (T, 8))
    lat = np.full(ncol, 50.0); lon = np.full(ncol, -5.0); scfg = np.array([1.0, 1, 2, 1, 0.95])
    for percol in (False, True):
        vegt = _vegt(ncol, T, percol)
        want = oracle.run_steady(veg, so, clim, nmr, lat, lon, scfg, M, SM, vegt=vegt)["out"]
        ctx = Context(M, SM); ctx.set_columns(veg, so, lat, lon); ctx.set_vegt(vegt)
        got = ctx.run_steady(clim, nmr, scfg)["out"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        g = np.nan_to_num(got); w_ = np.nan_to_num(want)
        assert np.max(np.abs(g[:, :2] - w_[:, :2])) <= 1e-6
        assert np.max(np.abs(g[:, 2:] - w_[:, 2:]) / np.maximum(np.abs(w_[:, 2:]), 1.0)) <= 1e-8


def test_statistics_accumulated_across_calls_equal_one_call(oracle):
    """mc_set_stats_accumulate: mean / min / max and flags carried on the device over chunked calls are bit-identical to one call over the
    whole range (same summation order), including for columns redone by the general kernel; turning it off restores per-call statistics."""
    ncol = 200
    inp = transient_inputs(oracle, ncol=ncol, T=96)
    g = S.vegp_deciduous(hgt=0.4, PAIt=1.5); g.update(x=0.4, lw=0.01, gsmax=0.4, hgtg=0.02)
    vg, _ = S.ensemble(g, loam(oracle), ncol, M, SM, seed=7)
    inp["veg"][:, ::7] = vg[:, ::7]                                        # some columns take the general kernel
    cfg = make_cfg(spin_steps=30)
    one = _ctx(inp, cfg).run_transient(want_series=False)
    ctx = _ctx(inp, cfg)
    ctx.set_stats_accumulate(True)
    ctx.run_transient(0, 40, want_series=False, want_stats=False, want_flags=False)
    ctx.run_transient(40, 41, want_series=False, want_stats=False, want_flags=False)
    r = ctx.run_transient(41, 96, want_series=False)
    np.testing.assert_array_equal(r["stats"], one["stats"])
    np.testing.assert_array_equal(r["flags"], one["flags"])
    ctx.set_stats_accumulate(False)
    r2 = ctx.run_transient(90, 96, want_series=False)
    assert np.all(r2["stats"][1] >= one["stats"][1]) and not np.array_equal(r2["stats"], one["stats"])


def _classed_grid(oracle, ncol, K=4, seed=3):
    """A grid of K habitat classes (identical vegp / soilp within a class), per-column lat / forcing perturbation."""
    inp = transient_inputs(oracle, ncol=K, T=60, seed=seed)
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, K, ncol)
    veg = inp["veg"][:, cls].copy(); soil = inp["soil"][:, cls].copy()
    fs, fo = S.forcing_perturbation(ncol, seed=seed + 1)
    return dict(veg=veg, soil=soil, f=inp["f"], ts=inp["ts"], lat=rng.uniform(45, 55, ncol), lon=np.full(ncol, -5.0), fs=fs, fo=fo), cls


def test_parameter_classes_share_geometry_rows_bit_identically(oracle):
    """mc_set_classes: tiles of one class read their per-layer geometry rows from the class's first tile; the results must not change by
    a single bit, whether the columns are ordered by class (sharing takes effect) or not (mixed tiles keep their own rows)."""
    ncol = 4 * 256 + 77
    inp, cls = _classed_grid(oracle, ncol)
    order = np.argsort(cls, kind="stable")
    srt = {k: (v[..., order] if isinstance(v, np.ndarray) and v.shape[-1] == ncol else v) for k, v in inp.items()}
    for case, ids in ((srt, cls[order]), (inp, cls)):
        for cfg in (make_cfg(spin_steps=30), make_cfg(spin_steps=30, reqhgt=2.0)):           # (reqhgt: the spin-up launch cannot share)
            a = _ctx(case, cfg); ra = a.run_transient(want_series=False)
            b = _ctx(case, cfg); b.set_classes(ids); rb = b.run_transient(want_series=False)
            np.testing.assert_array_equal(a.get_state(), b.get_state())
            np.testing.assert_array_equal(ra["stats"], rb["stats"]); np.testing.assert_array_equal(ra["flags"], rb["flags"])
    want = oracle.run_transient(srt["veg"], srt["soil"], srt["f"], srt["lat"], srt["lon"], make_cfg(spin_steps=30), M, SM, fscale=srt["fs"], foffset=srt["fo"])
    b = _ctx(srt, make_cfg(spin_steps=30)); b.set_classes(cls[order]); b.run_transient(want_series=False)
    assert_state_close(b.get_state(), want["state"], what="classes vs oracle")


def test_runmodel_batch_by_class_returns_the_callers_order(oracle):
    from microclimc_b200 import api
    ncol = 300
    inp, cls = _classed_grid(oracle, ncol, K=3, seed=9)
    w = S.load_weather(); clim = {k: np.asarray(v)[1000:1048] for k, v in w.items()}
    kw = dict(steps=20, fscale=inp["fs"], foffset=inp["fo"], sample=np.array([5, 250, 17]))
    a = api.runmodel_batch(clim, inp["veg"], inp["soil"], inp["lat"], inp["lon"], **kw)
    b = api.runmodel_batch(clim, inp["veg"], inp["soil"], inp["lat"], inp["lon"], classes=cls, **kw)
    for k in ("series", "stats", "flags", "state"):
        np.testing.assert_array_equal(a[k], b[k])
    bad = inp["veg"].copy(); bad[3, 7] *= 1.01
    with pytest.raises(ValueError, match="identical"):
        api.runmodel_batch(clim, bad, inp["soil"], inp["lat"], inp["lon"], classes=cls, **kw)


def test_regular_grid_launch_from_a_state_with_a_moved_node(oracle):
    """A launch with reqhgt NA runs the regular-grid instantiation of the streaming kernel, which assumes paraminit's
    node heights in previn$z as well.  A state handed over from a run with reqhgt = 5 m has one node moved (code.R:719): those columns
    must be redone by the general kernel (flags bit 3) and still match the oracle continuing from the same state."""
    inp = transient_inputs(oracle, ncol=40, T=48)
    first = make_cfg(reqhgt=5.0, spin_steps=30)
    want0 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][:24], inp["lat"], inp["lon"], first, M, SM, fscale=inp["fs"], foffset=inp["fo"])
    second = make_cfg(spin_steps=0)                                             # reqhgt NA from here on
    tsoil = inp["f"][:, 0].mean() * inp["fs"][0] + inp["fo"][0]
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][24:], inp["lat"], inp["lon"], second, M, SM, state=want0["state"],
                                fscale=inp["fs"], foffset=inp["fo"], tsoil=tsoil)
    ctx = _ctx(dict(inp, f=inp["f"][24:]), second)
    ctx.set_state(want0["state"])
    got = ctx.run_transient(tsoil=tsoil, samp_idx=np.arange(40))
    assert_series_close(got["series"], want["series"], "regular-grid launch from a moved-node state")
    assert_state_close(ctx.get_state(), want["state"], what="regular-grid launch from a moved-node state")
    assert np.all(got["flags"] & 8), "columns whose previn$z is not the regular grid leave the regular-grid instantiation"
    # the next launch starts from the regular grid again (runonestep returned z = the current heights) and stays in the streaming form
    got2 = ctx.run_transient(tsoil=tsoil, samp_idx=np.arange(40))
    assert not np.any(got2["flags"] & 8)
    want2 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"][24:], inp["lat"], inp["lon"], second, M, SM, state=want["state"],
                                 fscale=inp["fs"], foffset=inp["fo"], tsoil=tsoil)
    assert_series_close(got2["series"], want2["series"], "second regular-grid launch")


def test_geometry_workspace_follows_new_inputs_between_launches(oracle):
    """The library re-uses the geometry workspace of the streaming path across run launches (mc_lib.cu, Shard::GeomTag).  Every input it
    depends on must invalidate it: a new forcing perturbation (only those rows are rewritten), a new run description, new columns.
    Each continuation is compared with the oracle continuing from the same state with the same inputs."""
    inp = transient_inputs(oracle, ncol=40, T=72)
    cfg = make_cfg(spin_steps=30)
    nospin = make_cfg(spin_steps=0)
    f = inp["f"]

    def tsoil_of(fs, fo):
        return f[:, 0].mean() * fs[0] + fo[0]

    ctx = _ctx(inp, cfg)
    ts0 = tsoil_of(inp["fs"], inp["fo"])
    ctx.run_transient(0, 24, tsoil=ts0, want_series=False)                       # spin-up + 24 h: fills the workspace
    ctx.run_transient(24, 30, tsoil=ts0, want_series=False)                      # re-uses it
    st = ctx.get_state()
    # (1) a new perturbation between two launches
    fs2, fo2 = S.forcing_perturbation(40, seed=77)
    ctx.set_forcing(f, fs2, fo2)
    got = ctx.run_transient(30, 48, tsoil=ts0, samp_idx=np.arange(40))
    want = oracle.run_transient(inp["veg"], inp["soil"], f[30:48], inp["lat"], inp["lon"], nospin, M, SM, state=st, fscale=fs2, foffset=fo2,
                                tsoil=ts0)
    assert_series_close(got["series"], want["series"], "new perturbation")
    assert_state_close(ctx.get_state(), want["state"], what="new perturbation")
    # (2) a new run description (edge distance): the whole workspace is rebuilt
    cfg2 = make_cfg(spin_steps=30, edgedist=25.0)
    ctx.set_config(cfg2)
    st = ctx.get_state()
    got = ctx.run_transient(48, 60, tsoil=ts0, samp_idx=np.arange(40))
    want = oracle.run_transient(inp["veg"], inp["soil"], f[48:60], inp["lat"], inp["lon"], make_cfg(spin_steps=0, edgedist=25.0), M, SM, state=st,
                                fscale=fs2, foffset=fo2, tsoil=ts0)
    assert_series_close(got["series"], want["series"], "new edge distance")
    # (3) new columns (another ensemble) with the state carried over by the caller
    veg2, soil2 = S.ensemble(S.vegp_deciduous(), loam(oracle), 40, M, SM, seed=99)
    st = ctx.get_state()
    ctx.set_columns(veg2, soil2, inp["lat"], inp["lon"]); ctx.set_forcing(f, fs2, fo2); ctx.set_state(st)
    got = ctx.run_transient(60, 72, tsoil=ts0, samp_idx=np.arange(40))
    want = oracle.run_transient(veg2, soil2, f[60:72], inp["lat"], inp["lon"], make_cfg(spin_steps=0, edgedist=25.0), M, SM, state=st,
                                fscale=fs2, foffset=fo2, tsoil=ts0)
    assert_series_close(got["series"], want["series"], "new columns")
    assert_state_close(ctx.get_state(), want["state"], what="new columns")
