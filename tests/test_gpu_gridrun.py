"""GPU: grid runs with outputs streamed to disk (microclimc_b200.gridrun) against the in-memory API and the oracle."""
import os

import numpy as np
import pytest

from tests.common import M, SM, make_cfg, assert_series_close, loam
from microclimc_b200 import api, gridrun, layout as L, synth as S

pytestmark = pytest.mark.gpu


def _grid(oracle, ncol, T, t0=3000):
    w = {k: v[t0:t0 + T] for k, v in S.load_weather().items()}
    veg, soil = S.ensemble(S.vegp_deciduous(), loam(oracle), ncol, M, SM, seed=21)
    fs, fo = S.forcing_perturbation(ncol, seed=22)
    return w, veg, soil, fs, fo


def test_chunked_grid_run_on_disk_equals_in_memory_run_and_oracle(oracle, tmp_path):
    ncol, T = 300, 100
    w, veg, soil, fs, fo = _grid(oracle, ncol, T)
    samp = np.array([0, 7, 150, 299])
    meta = gridrun.runmodel_grid(w, veg, soil, 50.0, -5.0, str(tmp_path), chunk_hours=37, sample=samp, steps=40, fscale=fs, foffset=fo)
    assert meta["ncol"] == ncol and meta["T"] == T and meta["sample"] == [0, 7, 150, 299] and gridrun.read_meta(str(tmp_path)) == meta
    mem = api.runmodel_batch(w, veg, soil, 50.0, -5.0, steps=40, fscale=fs, foffset=fo, sample=samp)
    series = np.load(os.path.join(tmp_path, "series.npy"))
    np.testing.assert_array_equal(series, mem["series"])                       # chunking changes nothing, bit for bit
    np.testing.assert_array_equal(np.load(os.path.join(tmp_path, "state.npy")), mem["state"])
    stats = np.load(os.path.join(tmp_path, "stats.npy"))
    np.testing.assert_allclose(stats, mem["stats"], rtol=0, atol=1e-11)        # means are combined per chunk: summation order
    np.testing.assert_array_equal(stats[[1, 2, 4, 5]], mem["stats"][[1, 2, 4, 5]])
    np.testing.assert_array_equal(np.load(os.path.join(tmp_path, "flags.npy")), mem["flags"])
    f, ts = S.forcing_from_climdata(w)
    want = oracle.run_transient(veg[:, samp], soil[:, samp], f, np.full(4, 50.0), np.full(4, -5.0), make_cfg(timestep=ts, spin_steps=40),
                                M, SM, fscale=fs[:, samp], foffset=fo[:, samp])
    assert_series_close(series, want["series"], "grid run on disk")
    df = gridrun.read_tables(str(tmp_path), 2)
    assert list(df.columns) == ["obs_time", "reftemp", "tout", "tleaf", "relhum", "SWin", "LWin", "H", "L", "G"] and len(df) == T
    np.testing.assert_array_equal(df["tout"].values, series[:, 0, 2]); np.testing.assert_array_equal(df["H"].values, series[:, 6, 2])


def test_all_tables_from_disk_equal_runmodel_all(oracle, tmp_path):
    """reqhgt = "all": the tables rebuilt from disk for a sampled column are the ones `runmodel` returns for that column."""
    T = 60
    w = {k: v[2000:2000 + T] for k, v in S.load_weather().items()}
    vegp = S.vegp_deciduous(); soilp = api.soilinit("Loam")
    other = dict(vegp, hgt=9.0, gsmax=0.25)
    gridrun.runmodel_grid(w, [other, vegp, other], [soilp] * 3, 50, -5, str(tmp_path), chunk_hours=25, sample=[1], reqhgt="all", steps=50)
    got = gridrun.read_tables(str(tmp_path), 0)
    want = api.runmodel(w, vegp, soilp, 50, -5, reqhgt="all", steps=50)
    assert sorted(got) == sorted(want) == ["Airtemp", "Conductivity", "Fluxes", "Leaftemp", "Relhum", "Soiltemp", "Windspeed"]
    for k in want:
        assert list(got[k].columns) == list(want[k].columns), k
        np.testing.assert_array_equal(got[k].drop(columns="obs_time").values, want[k].drop(columns="obs_time").values, err_msg=k)
