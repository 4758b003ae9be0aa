"""NicheMapR-free soil / snow provider: physical sanity of the tables it produces and that they drive the steady
path.  There is no reference output to compare with (NicheMapR is absent), so these are property tests, not parity tests."""
import numpy as np
import pytest

from microclimc_b200 import soilsnow as P, synth as S, api


def _case(hours=24 * 120, t0=0, shift=0.0):
    w = S.load_weather()
    clim = {k: np.asarray(v)[t0:t0 + hours].copy() for k, v in w.items()}
    clim["temp"] = clim["temp"] + shift
    return clim, S.load_dailyprecip()


def test_tables_have_the_nmrout_layout_and_sane_physics():
    clim, prec = _case()
    soil = api.soilinit("Loam")
    r = P.soil_snow(clim, prec, S.vegp_deciduous(), soil)
    T = len(clim["temp"])
    assert list(r["soiltemps"])[0] == "D0cm" and list(r["soiltemps"])[-1] == "D200cm" and len(r["soiltemps"]) == 10
    assert "WC0cm" in r["soilmoist"] and "SNOWDEP" in r["metout"] and all("SN%d" % q in r["snowtemp"] for q in range(1, 9))
    d0, d50, d200 = r["soiltemps"]["D0cm"], r["soiltemps"]["D50cm"], r["soiltemps"]["D200cm"]
    assert d0.shape == (T,) and np.all(np.isfinite(d0))
    assert np.std(d0) > np.std(d50) > np.std(d200)                          # the daily / synoptic wave is damped with depth
    assert abs(d0.mean() - clim["temp"].mean()) < 5.0
    lag0 = np.argmax(np.correlate(d0 - d0.mean(), clim["temp"] - clim["temp"].mean(), "full")) - (T - 1)
    assert 0 <= lag0 <= 6                                                   # the surface follows the air within hours, never leads it
    wc = r["soilmoist"]["WC0cm"]
    assert np.all((wc >= soil["Smin"] - 1e-12) & (wc <= soil["Smax"] + 1e-12))
    wet = P.soil_snow(clim, prec, S.vegp_deciduous(), soil, rainmult=3.0)["soilmoist"]["WC0cm"]
    assert wet.mean() > wc.mean()


def test_snow_only_when_cold_and_temperatures_below_zero():
    clim, prec = _case(hours=24 * 60, shift=-8.0)
    r = P.soil_snow(clim, prec * 3, S.vegp_deciduous(), api.soilinit("Loam"))
    dep = r["metout"]["SNOWDEP"]
    assert dep.max() > 1.0
    sn = np.stack([r["snowtemp"]["SN%d" % q] for q in range(1, 9)], axis=1)
    assert np.all(sn[dep > 0] <= 0.0)
    warm, prec2 = _case(hours=24 * 30, t0=24 * 180)
    assert P.soil_snow(warm, prec2, S.vegp_deciduous(), api.soilinit("Loam"))["metout"]["SNOWDEP"].max() == 0.0
    # snow insulates: the soil surface under a pack varies less than the bare one in the same cold spell
    bare = P.soil_snow(clim, prec * 0, S.vegp_deciduous(), api.soilinit("Loam"))["soiltemps"]["D0cm"]
    assert np.std(r["soiltemps"]["D0cm"][dep > 5]) < np.std(bare[dep > 5])


def test_negative_reqhgt_interpolates_between_the_two_nearest_depths():
    clim, prec = _case(hours=48)
    out = P.runwithprovider(clim, prec, S.vegp_deciduous(), api.soilinit("Loam"), -0.07, 50.0, -5.0)
    st = out["nmrout"]["soiltemps"]
    want = st["D5cm"] * (3 / 5) + st["D10cm"] * (2 / 5)                      # 7 cm: 2 cm from D5cm, 3 cm from D10cm (NMR.R:1036-1047)
    np.testing.assert_allclose(out["metout"]["Tloc"].to_numpy(), want, rtol=0, atol=1e-12)
    with pytest.raises(ValueError, match="windhgt must be above canopy"):
        P.runwithprovider(clim, prec, S.vegp_deciduous(), api.soilinit("Loam"), 1.0, 50.0, -5.0, metopen=False, windhgt=2)


@pytest.mark.gpu
def test_provider_drives_the_steady_path_on_the_gpu():
    clim, prec = _case(hours=24 * 20, shift=-6.0)
    out = P.runwithprovider(clim, prec * 3, S.vegp_deciduous(), api.soilinit("Loam"), 1.0, 50.0, -5.0)
    m = out["metout"]
    assert list(m.columns) == ["obs_time", "Tref", "Tloc", "tleaf", "RHref", "RHloc", "RSWloc", "RLWloc", "windspeed", "dp"]
    assert np.all(np.isfinite(m["Tloc"])) and np.all(np.abs(m["Tloc"] - m["Tref"]) < 20)
    assert out["nmrout"]["metout"]["SNOWDEP"].max() > 0                      # the snow overlay (.runmodelsnow) was exercised
