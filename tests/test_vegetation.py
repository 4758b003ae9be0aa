"""Vegetation generators: documented behaviour of the microctools functions they stand in for
(running-microclimc.md:229-382) — shapes, sums, the two printed values, seasonality — and that their output runs through the model."""
import datetime as dt

import numpy as np
import pytest

from microclimc_b200 import vegetation as V, layout as L, synth as S

FIELDS = ["hgt", "PAI", "x", "lw", "cd", "iw", "hgtg", "zm0", "pLAI", "refls", "refg", "refw", "reflp", "vegem", "gsmax", "q50", "thickw", "cpw",
          "phw", "kwood", "clump"]


def test_profiles():
    p = V.PAIgeometry(100, 10, skew=4, spread=70); q = V.PAIgeometry(100, 10, skew=7, spread=70)
    assert p.shape == (100,) and abs(p.sum() - 10) < 1e-12 and abs(q.sum() - 10) < 1e-12 and np.all(p >= 0)
    assert np.argmax(q) > np.argmax(p)                                      # more skew: foliage higher up
    assert V.PAIgeometry(100, 10, 7, 20).max() > q.max()                    # less spread: narrower
    t = V.thickgeometry(1000, 0.4, 0.7, 0.1)
    assert t[0] == 0.4 and abs(t[-1] - 0.04) < 1e-3 and np.all(np.diff(t) <= 1e-15)
    f = V.LAIfrac(100, 0.8, 5)
    assert abs(f.mean() - 0.8) < 0.02 and np.all((f >= 0) & (f <= 1)) and f[-1] > f[0]
    assert V.LAIfrac(100, 0.8, 10)[10] > f[10]
    iw = V.iwgeometry(100)
    assert iw.shape == (100,) and iw[0] > iw[-1] and np.all((iw > 0.3) & (iw < 1))


def test_pai_from_habitat_and_the_printed_values():
    pxh = V.PAIfromhabitat("Deciduous broadleaf forest", 50, -5.2, 2015)
    assert pxh["height"] == 15 and pxh["x"] == 1.2                          # running-microclimc.md:346-349
    assert len(pxh["lai"]) == 8760 and len(pxh["obs_time"]) == 8760 and pxh["obs_time"][0] == "2015-01-01 00:00:00"
    lai = pxh["lai"]
    assert lai[:24 * 31].mean() < 1.5 and 4 < lai[24 * 190:24 * 220].mean() <= 6     # bare in January, ~5 in July (vignette plot, ylim 0..6)
    south = V.PAIfromhabitat("Deciduous broadleaf forest", -40, 147, 2015)["lai"]
    assert south[:24 * 31].mean() > south[24 * 190:24 * 220].mean()        # seasons flip with the hemisphere
    ever = V.PAIfromhabitat("Evergreen broadleaf forest", 2, 20, 2015)["lai"]
    assert ever.min() > 0.9 * ever.max()
    assert len(V.PAIfromhabitat(4, 50, -5, 2016)["lai"]) == 8784            # habitat by number, leap year
    with pytest.raises(ValueError):
        V.PAIfromhabitat("Moon", 0, 0, 2015)
    assert len(V.habitats) == 16 and V.habitats[3]["descriptor"] == "Deciduous broadleaf forest"


def test_habitatvars_shapes():
    tme = [dt.datetime(2015, 1, 1) + dt.timedelta(days=d) for d in range(365)]
    v = V.habitatvars("Deciduous broadleaf forest", 50, -5.2, tme, 100)
    assert list(v.keys()) == FIELDS
    assert v["PAI"].shape == (100, 8760) and v["pLAI"].shape == (100, 8760)          # daily times become hourly, 00:00 .. 23:00
    prof = v["PAI"].mean(axis=1)
    assert prof[:20].max() > prof[25:40].min()                              # understorey bump near the ground (vignette text)
    one = V.habitatvars("Deciduous broadleaf forest", 50, -5.2, "2015-07-01 12:00:00", 20)
    assert one["PAI"].shape == (20,) and 4 < one["PAI"].sum() <= 6
    fixed = V.habitatvars("Croplands", 50, -5.2, tme, 10, PAIt=2.5)
    assert fixed["PAI"].shape == (10,) and abs(fixed["PAI"].sum() - 2.5) < 1e-12
    node = V.habitatvars("Croplands", 50, -5.2, tme[:3], 10, PAIt=np.linspace(0.1, 0.5, 10))
    assert node["PAI"].shape == (10, 72)
    ser = V.habitatvars("Croplands", 50, -5.2, tme[:3], 10, PAIt=[1.0, 2.0, 3.0, 2.0])
    assert ser["PAI"].shape == (10, 72) and abs(ser["PAI"][:, 0].sum() - 1.0) < 1e-9 and abs(ser["PAI"][:, -1].sum() - 2.0) < 1e-9


def test_habitatvars_feeds_the_time_varying_path():
    """m x T matrices from habitatvars go through the host mirror's .vegpsort handling: a separable season here, so pai_season."""
    from microclimc_b200 import api
    tme = [dt.datetime(2015, 6, 1) + dt.timedelta(hours=h) for h in range(48)]
    v = V.habitatvars("Evergreen needleleaf forest", 60, 10, tme, 20)
    v2, season, vegt = api._season_of(v)
    assert (season is not None) != (vegt is not None)
    d = V.habitatvars("Deciduous broadleaf forest", 50, -5, [dt.datetime(2015, 4, 1) + dt.timedelta(hours=h) for h in range(24 * 40)], 20)
    _, season, vegt = api._season_of(d)                                     # pLAI changes through leaf-out: the general rows are needed
    assert vegt is not None and vegt.shape == (960, 41)
    packed = S.pack_vegp([api._vegpsort(d, 0)], 20)
    assert packed.shape == (L.nveg(20), 1) and np.all(np.isfinite(packed))


@pytest.mark.gpu
def test_runmodel_on_habitatvars_output(oracle):
    """The documented workflow (code.R:1110-1113): vegp <- habitatvars(...); runmodel(weather, vegp, soilp, ...)."""
    from microclimc_b200 import api
    from tests.common import loam
    w = S.load_weather(); T = 96; t0 = 24 * 100
    clim = {k: np.asarray(x)[t0:t0 + T] for k, x in w.items()}
    vegp = V.habitatvars("Deciduous broadleaf forest", 50, -5.2, [str(s) for s in clim["obs_time"]], 20)
    df = api.runmodel(clim, vegp, loam(oracle), 50.0, -5.2, reqhgt=1.0, steps=40)
    assert len(df) == T and np.all(np.isfinite(df["tout"])) and np.all(np.abs(df["tout"] - df["reftemp"]) < 15)
