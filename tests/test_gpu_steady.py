"""GPU parity of the steady-state path (runmodelS + .runmodelsnow + .snowtempf + tleafS, R/NicheMapRcode.R:432-885)
through the C-ABI, against the CPU oracle.  Tolerances: 1e-6 K on temperatures, 1e-8 relative on the rest."""
import numpy as np
import pytest

from tests.common import M, SM, loam
from tests.test_host_logic import _steady_inputs
from microclimc_b200 import layout as L, synth as S

pytestmark = pytest.mark.gpu
TEMP = [0, 1]                 # Tloc, tleaf
REST = [2, 3, 4, 5, 6]        # RHloc, RSWloc, RLWloc, windspeed, dp


def _ctx(veg, soil, n_gpus=1):
    from microclimc_b200._lib import Context
    ncol = veg.shape[1]
    ctx = Context(M, SM, n_gpus)
    ctx.set_columns(veg, soil, np.full(ncol, 50.0), np.full(ncol, -5.0))
    return ctx


def _check(got, want):
    assert np.array_equal(np.isnan(got), np.isnan(want))
    g = np.nan_to_num(got); w = np.nan_to_num(want)
    assert np.max(np.abs(g[:, TEMP] - w[:, TEMP])) <= 1e-6
    assert np.max(np.abs(g[:, REST] - w[:, REST]) / np.maximum(np.abs(w[:, REST]), 1e-3)) <= 1e-8


@pytest.mark.parametrize("reqhgt,snow,percol", [(1.0, False, False), (15.5, False, False), (22.0, False, True), (1.0, True, True),
                                                (0.3, True, False), (18.0, True, False), (7.25, True, True)])
def test_steady_series_parity(oracle, reqhgt, snow, percol):
    ncol, T = 96, 240
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=500, snow=snow, percol=percol)
    scfg = np.array([reqhgt, 1, 2, 0.8, 0.95])
    want = oracle.run_steady(veg, soil, clim, nmr, np.full(ncol, 50.0), np.full(ncol, -5.0), scfg, M, SM)
    got = _ctx(veg, soil).run_steady(clim, nmr, scfg)
    _check(got["out"], want["out"])
    np.testing.assert_allclose(got["stats"], want["stats"], rtol=0, atol=1e-6)


def test_steady_full_year_bundled_weather_c2_shape(oracle):
    """BASELINE config 2 in miniature: perturbed-parameter ensemble on the bundled forcing, whole year, reqhgt 1 m."""
    ncol, T = 64, 8760
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, snow=False)
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    want = oracle.run_steady(veg, soil, clim, nmr, np.full(ncol, 50.0), np.full(ncol, -5.0), scfg, M, SM)
    got = _ctx(veg, soil).run_steady(clim, nmr, scfg)
    _check(got["out"], want["out"])


def test_steady_seasonal_pai_and_metopen_false(oracle):
    ncol, T = 32, 120
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=4000)
    season = np.linspace(0.4, 1.0, T)
    scfg = np.array([2.0, 0, 25.0, 1, 0.95])                   # metopen = FALSE, wind measured at 25 m
    want = oracle.run_steady(veg, soil, clim, nmr, np.full(ncol, 50.0), np.full(ncol, -5.0), scfg, M, SM, pai_season=season)
    got = _ctx(veg, soil).run_steady(clim, nmr, scfg, pai_season=season)
    _check(got["out"], want["out"])


def test_snow_layer_lookup_error(oracle):
    from microclimc_b200._lib import MicroclimcError
    veg, soil, clim, nmr = _steady_inputs(oracle, 8, 24, snow=True)
    nmr[:, 2] = 50.0
    with pytest.raises(MicroclimcError, match="snow-layer"):
        _ctx(veg, soil).run_steady(clim, nmr, np.array([0.01, 1, 2, 1, 0.95]))


def test_hours_are_independent_time_reversal(oracle):
    """Size-independent property of the steady path: no time recurrence, so reversing the hours reverses the output."""
    ncol, T = 40, 100
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=3000, snow=True)
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    a = _ctx(veg, soil).run_steady(clim, nmr, scfg)["out"]
    b = _ctx(veg, soil).run_steady(clim[::-1].copy(), nmr[::-1].copy(), scfg)["out"]
    np.testing.assert_array_equal(a, b[::-1])


def test_steady_large_batch_matches_small_batches(oracle):
    """10^4-column ensemble (BASELINE config 2 size): results do not depend on how the batch is tiled over the grid."""
    ncol, T = 10_000, 48
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, t0=4100)
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    big = _ctx(veg, soil).run_steady(clim, nmr, scfg)["out"]
    sel = np.r_[0:50, 5000:5050, 9950:10000]
    small = _ctx(np.ascontiguousarray(veg[:, sel]), np.ascontiguousarray(soil[:, sel])).run_steady(clim, nmr, scfg)["out"]
    np.testing.assert_array_equal(big[:, :, sel], small)
    want = oracle.run_steady(veg[:, sel], soil[:, sel], clim, nmr, np.full(sel.size, 50.0), np.full(sel.size, -5.0), scfg, M, SM)
    _check(small, want["out"])


def test_config2_full_size_statistics_against_oracle(oracle):
    """BASELINE config 2 in full (10^4-column ensemble x the bundled year, 8.76e7 column-hours): per-column statistics of every
    column finite, and the statistics of a 64-column sample equal to the oracle's over the whole year."""
    ncol, T = 10_000, 8760
    veg, soil, clim, nmr = _steady_inputs(oracle, ncol, T, snow=False)
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    st = _ctx(veg, soil).run_steady(clim, nmr, scfg, want_out=False)["stats"]
    assert st.shape == (6, ncol) and np.all(np.isfinite(st))
    sel = np.linspace(0, ncol - 1, 64).astype(int)
    want = oracle.run_steady(veg[:, sel], soil[:, sel], clim, nmr, np.full(64, 50.0), np.full(64, -5.0), scfg, M, SM)["out"]   # [T][7][64]
    for k, col in ((0, 0), (3, 1)):                         # mean / min / max of Tloc, then of tleaf
        np.testing.assert_allclose(st[k, sel], want[:, col].mean(axis=0), rtol=0, atol=1e-6)
        np.testing.assert_allclose(st[k + 1, sel], want[:, col].min(axis=0), rtol=0, atol=1e-6)
        np.testing.assert_allclose(st[k + 2, sel], want[:, col].max(axis=0), rtol=0, atol=1e-6)


def test_config4_full_size_mixed_habitats_with_snow(oracle):
    """BASELINE config 4 size (4 x 10^6 columns, four habitat classes, snow pack and bucket soil moisture from dailyprecip): columns
    of the same class agree bit for bit wherever they sit in the grid, everything is finite, a sample matches the oracle."""
    T, t0, nbase, rep = 36, 516, 4000, 1000
    ncol = nbase * rep
    veg1, soil1 = S.ensemble(S.vegp_deciduous(), loam(oracle), nbase, M, SM, seed=11)
    cls = np.arange(nbase) % 4                               # habitat classes: scale height and plant area of the template
    veg1[L.VEG_SCALARS.index("hgt")] *= (0.4 + 0.4 * cls)
    veg1[len(L.VEG_SCALARS):len(L.VEG_SCALARS) + M] *= (0.5 + 0.35 * cls)
    w = S.load_weather(); sp = S.load_soilparams(); i = sp["Soil.type"].index("Loam")
    lo = {k: sp[k][i] for k in L.SOIL_SCALARS}
    clim = S.steady_clim_from_climdata(w)[t0:t0 + T].copy(); clim[:, 0] -= 6.0
    th = S.theta_from_dailyprecip(S.load_dailyprecip(), lo, w["temp"] - 6.0, w["swrad"])
    dep, sn = S.snow_from_forcing(S.load_dailyprecip() * 4.0, w["temp"] - 6.0)
    nmr = np.zeros((T, 11))
    nmr[:, 1] = th[t0:t0 + T]; nmr[:, 2] = dep[t0:t0 + T]; nmr[:, 3:] = sn[t0:t0 + T]
    nmr[:, 0] = np.minimum(clim[:, 0] + 0.5, 0.0)
    assert nmr[:, 2].min() > 0 and nmr[:, 2].min() < 100 < nmr[:, 2].max()          # the pack grows through the 1 m output height
    veg = np.ascontiguousarray(np.tile(veg1, (1, rep))); soil = np.ascontiguousarray(np.tile(soil1, (1, rep)))
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    st = _ctx(veg, soil).run_steady(clim, nmr, scfg, want_out=False)["stats"]
    assert np.all(np.isfinite(st))
    s3 = st.reshape(6, rep, nbase)
    assert np.all(s3 == s3[:, :1, :])
    sel = np.arange(0, nbase, 125)
    want = oracle.run_steady(veg1[:, sel], soil1[:, sel], clim, nmr, np.full(sel.size, 50.0), np.full(sel.size, -5.0), scfg, M, SM)["out"]
    np.testing.assert_allclose(st[0, sel], want[:, 0].mean(axis=0), rtol=0, atol=1e-6)
    np.testing.assert_allclose(st[5, sel], want[:, 1].max(axis=0), rtol=0, atol=1e-6)
