"""Host-side input shaping of the R-API mirror (microclimc_b200/api.py) that needs no device: .vegpsort / .snowsort
(R/code.R:904-933), soilinit (R/code.R:619-627), theta normalisation (code.R:1146-1154), the separable seasonal PAI
factorisation, packing into the column-interleaved layout, and the quirk-Q16 re-indexing of the snow series."""
import numpy as np
import pytest

from microclimc_b200 import api, layout as L, synth as S


def test_soilinit_matches_reference_definition():
    s = api.soilinit("Loam")
    assert s["Smax"] == 0.422 and s["b"] == 4.5 and s["psi_e"] == 1.1
    np.testing.assert_allclose(s["z"], (2 / 10 ** 1.2) * np.arange(1, 11) ** 1.2)
    s2 = api.soilinit("Clay", m=6, sdepth=1.5, reqdepth=0.4)
    assert 0.4 in s2["z"] and s2["z"][-1] == 1.5 and len(s2["z"]) == 6
    with pytest.raises(ValueError):
        api.soilinit("Silt")              # listed in the docs (code.R:591-593) but absent from the dataset


def test_vegpsort_and_snowsort():
    v = S.vegp_deciduous()
    T = 5
    v["PAI"] = np.outer(v["PAI"], np.linspace(0.5, 1.0, T)); v["pLAI"] = np.repeat(v["pLAI"][:, None], T, axis=1)
    v["zm0"] = np.linspace(0.004, 0.008, T)
    s = api._vegpsort(v, 3)
    np.testing.assert_array_equal(s["PAI"], v["PAI"][:, 3]); assert s["zm0"] == v["zm0"][3]
    sn = api._snowsort(s, np.array([1, 0, 0]), 0)
    assert sn["lw"] == s["lw"] * 1.5 and (sn["zm0"], sn["refls"], sn["refg"], sn["refw"], sn["reflp"], sn["gsmax"], sn["q50"]) == \
        (0.024, 0.85, 0.8, 0.7, 0.8, 10.0, 1.0)
    assert api._snowsort(s, np.array([0, 1, 1]), 0)["refls"] == s["refls"]       # Q3: only snow[1] is ever tested
    assert api._snowsort(s, None)["refls"] == s["refls"]


def test_seasonal_pai_factorisation():
    v = S.vegp_deciduous()
    season = 0.6 + 0.4 * np.sin(np.linspace(0, np.pi, 12)) ** 2
    v2 = dict(v, PAI=np.outer(v["PAI"], season))
    base, s, vt = api._season_of(v2)
    assert vt is None
    np.testing.assert_allclose(np.outer(base["PAI"], s), v2["PAI"], rtol=1e-13)
    # anything that is not profile x seasonal factor travels in full: PAI[m], pLAI[m], zm0 per time step (.vegpsort, code.R:904-916)
    bad = dict(v, PAI=np.outer(v["PAI"], season)); bad["PAI"][3, 4] *= 1.3
    base, s, vt = api._season_of(bad)
    assert s is None and vt.shape == (12, 41)
    np.testing.assert_array_equal(vt[:, :20], bad["PAI"].T); np.testing.assert_array_equal(vt[4, 20:40], v["pLAI"]); assert np.all(vt[:, 40] == v["zm0"])
    np.testing.assert_array_equal(base["PAI"], bad["PAI"][:, 0])
    tv = dict(v, pLAI=np.outer(v["pLAI"], np.linspace(1, 0.8, 12)), zm0=np.linspace(0.004, 0.006, 12))
    base, s, vt = api._season_of(tv)
    assert s is None and vt[7, 40] == tv["zm0"][7]; np.testing.assert_array_equal(vt[7, 20:40], tv["pLAI"][:, 7]); np.testing.assert_array_equal(vt[7, :20], v["PAI"])
    assert api._season_of(v)[1] is None and api._season_of(v)[2] is None
    # columns: shared season stays a season; differing columns become per-column rows; identical general columns collapse to one series
    _, se, vt = api._veg_time([v2, v2]); assert vt is None and se.shape == (12,)
    _, se, vt = api._veg_time([v2, bad]); assert se is None and vt.shape == (12, 41, 2); np.testing.assert_allclose(vt[:, :20, 0], v2["PAI"].T, rtol=1e-13)
    _, se, vt = api._veg_time([bad, bad]); assert vt.shape == (12, 41, 1)


def test_theta_forms():
    assert api._theta_series(0.3, 10, 4) == (0.3, None)
    th, tz = api._theta_series(np.full((1, 10), 0.25), 10, 4); np.testing.assert_array_equal(th, np.full(10, 0.25)); assert tz is None
    th, tz = api._theta_series(np.linspace(0.2, 0.3, 10), 10, 4); np.testing.assert_array_equal(th, np.linspace(0.2, 0.3, 10)); assert tz is None
    # soil moisture by depth (code.R:1151-1153 and the matrix form): thetaz [T][sm]
    th, tz = api._theta_series(np.array([0.2, 0.25, 0.3, 0.35]), 10, 4)
    assert th == 0.2 and tz.shape == (10, 4) and np.all(tz == np.array([0.2, 0.25, 0.3, 0.35]))
    mat = np.linspace(0.1, 0.4, 40).reshape(4, 10)
    th, tz = api._theta_series(mat, 10, 4); np.testing.assert_array_equal(tz, mat.T); np.testing.assert_array_equal(th, mat[0])
    with pytest.raises(ValueError):
        api._theta_series(np.zeros((3, 10)), 10, 4)


def test_pack_roundtrip_and_state_list():
    v = S.vegp_deciduous(); so = api.soilinit("Loam")
    veg, soil, m, sm = api._pack([v, v], [so])
    assert (m, sm) == (20, 10) and veg.shape == (L.nveg(20), 2) and soil.shape == (L.nsoil(10), 2)
    r = L.veg_rows(20)
    np.testing.assert_array_equal(veg[r["PAI"], 1], v["PAI"]); assert veg[r["hgt"], 0] == 15.0
    st = np.arange(L.nstate(20, 10), dtype=float)[:, None]
    d = api._state_to_list(st, 20, 10)
    assert len(d) == 23 and d["tc"].shape == (20,) and d["gt"].shape == (21,) and isinstance(d["H"], float)
    np.testing.assert_array_equal(api._list_to_state(d, 20, 10), st)


def test_q16_snow_double_subset():
    """tleaf2 <- tz2[sbs] (NMR.R:594): positions of buried hours index the already-subset vector."""
    tloc = np.array([1., 2., 3., 4., 5., 6.]); tleaf = tloc + 10
    sd_cm = np.array([0, 50, 0, 50, 50, 0.])      # buried hours (reqhgt 0.2 m): 1, 3, 4  ->  tz2 = [2, 4, 5]
    out = api._q16_remap(tleaf, tloc, sd_cm, 0.2)
    # R: tz2[c(2,4,5)] = (4, NA, NA)
    assert out[1] == 4.0 and np.isnan(out[3]) and np.isnan(out[4])
    np.testing.assert_array_equal(out[[0, 2, 5]], tleaf[[0, 2, 5]])


def test_guards():
    v = S.vegp_deciduous(); so = api.soilinit("Loam"); w = S.load_weather()
    with pytest.raises(ValueError, match="no recognised"):
        api.runmodel_batch(w, [v], [so], 50, -5, reqhgt="everything")
    with pytest.raises(ValueError, match="windhgt must be above canopy"):
        api.runmodel_batch(w, [v], [so], 50, -5, metopen=False, windhgt=2)
    with pytest.raises(ValueError, match="three canopy layers"):
        api.runonestep(dict(tair=1, relhum=1, pk=1, u=1, tsoil=1, skyem=1, Rsw=1), dict(tc=[1, 2], soiltc=[1, 2]), v, so, 60, (1, 1), 50, -5)


def test_leafem_and_datasets():
    assert abs(api.leafem(11.0, 0.97) - 0.97 * 5.67e-8 * (284.15) ** 4) < 1e-9
    assert len(api.weather()["temp"]) == 8760 and api.dailyprecip().size == 365 and api.climvars()["tair"] == 16


def test_c4_soil_moisture_bucket_and_snow_series():
    """BASELINE config C4 inputs: theta(t) from the documented dailyprecip bucket, snow pack from the same
    precipitation.  Host-side recipes (soil moisture and snow are inputs of the hot path): bounded, reproducible, responsive."""
    w = S.load_weather(); sp = S.load_soilparams(); i = sp["Soil.type"].index("Loam")
    loam = {k: sp[k][i] for k in L.SOIL_SCALARS}
    P = S.load_dailyprecip()
    th = S.theta_from_dailyprecip(P, loam, w["temp"], w["swrad"])
    assert th.shape == (8760,) and th.min() >= loam["Smin"] and th.max() <= loam["Smax"]
    assert np.array_equal(th, S.theta_from_dailyprecip(P, loam, w["temp"], w["swrad"]))
    wet = S.theta_from_dailyprecip(P * 2, loam, w["temp"], w["swrad"]); dry = S.theta_from_dailyprecip(P * 0, loam, w["temp"], w["swrad"])
    assert wet.mean() > th.mean() > dry.mean() and dry[-1] < dry[0]
    dep, sn = S.snow_from_forcing(P * 4.0, w["temp"] - 6.0)
    assert dep.shape == (8760,) and sn.shape == (8760, 8) and dep.min() >= 0 and dep.max() > 100 and np.all(sn <= 0)
    assert np.all(np.diff(dep)[w["temp"][1:] - 6.0 >= 0.5] <= 0)                      # melts, never grows, above the threshold
    d0, _ = S.snow_from_forcing(P * 4.0, w["temp"] + 20.0)
    assert d0.max() == 0.0


def test_gridrun_rejects_bad_chunking_before_touching_a_device(tmp_path):
    from microclimc_b200 import gridrun
    with pytest.raises(ValueError, match="chunk_hours"):
        gridrun.runmodel_grid(S.load_weather(), [S.vegp_deciduous()], [api.soilinit("Loam")], 50, -5, str(tmp_path), chunk_hours=0)
