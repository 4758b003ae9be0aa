"""Real-R parity: compares the in-repo oracle with fixtures produced by the reference itself
(tools/r_parity/dump_reference.R, to be run where R + microctools + microclimc are installed).  The build image has no R, so
the fixtures are absent here and these tests skip; until they exist every parity statement of the repository reads
"CUDA path vs in-repo oracle" (DESIGN.md §2)."""
import csv
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "r_reference")
needs_fixture = pytest.mark.skipif(not os.path.exists(os.path.join(FIX, "step_out.csv")),
                                   reason="no fixtures from a real R run (tools/r_parity/dump_reference.R)")


def read_list(path):
    """'name,v1,v2,...' rows -> dict of float arrays (the wlist() format of dump_reference.R)."""
    out = {}
    with open(path) as f:
        for row in csv.reader(f):
            if row:
                out[row[0]] = np.array([float(x) if x not in ("NA", "NaN") else np.nan for x in row[1:]])
    return out


def test_fixture_reader_roundtrip(tmp_path):
    p = tmp_path / "x.csv"
    p.write_text("hgt,15\nPAI,0.1,0.2,NA\n")
    d = read_list(str(p))
    assert d["hgt"][0] == 15 and np.isnan(d["PAI"][2]) and d["PAI"].shape == (3,)


def _lists():
    one = lambda d: {k: (v if v.size > 1 else float(v[0])) for k, v in d.items()}
    return one(read_list(os.path.join(FIX, "vegp.csv"))), one(read_list(os.path.join(FIX, "soilp.csv")))


def _state_from_list(d, m, sm):
    """previn list of the reference (code.R:581-584, 898-901) -> ABI state column [NSTATE][1]."""
    from microclimc_b200 import layout as L
    st = np.full((L.nstate(m, sm), 1), np.nan)
    rows = L.state_rows(m, sm)
    for k, sl in rows.items():
        if k in d:
            st[sl, 0] = d[k]
    return st


@needs_fixture
def test_oracle_runonestep_equals_real_R(oracle):
    """ONE runonestep from paraminit on the first row of the bundled weather (dump_reference.R: step_in / step_out)."""
    from microclimc_b200 import layout as L, synth as S
    from tests.common import M, SM, make_cfg, assert_state_close
    vegp, soilp = _lists()
    veg = S.pack_vegp([vegp], M); soil = S.pack_soilp([soilp], SM)
    w = S.load_weather()
    f, ts = S.forcing_from_climdata(w)
    clim = np.array([[w["temp"][0]], [w["relhum"][0]], [w["pres"][0]], [w["windspeed"][0]], [np.mean(w["temp"])], [w["skyem"][0]],
                     [w["swrad"][0]], [f[0, 6]]])
    st0 = _state_from_list(read_list(os.path.join(FIX, "step_in.csv")), M, SM)
    want = _state_from_list(read_list(os.path.join(FIX, "step_out.csv")), M, SM)
    cfg = make_cfg(timestep=3600.0, edgedist=100.0, spin_steps=0)
    got, _ = oracle.runonestep(veg, soil, clim, np.array([50.0]), np.array([-5.0]), cfg, f[0, 8], f[0, 9], np.array([0.3]),
                               np.array([0.3]), st0, M, SM)
    assert_state_close(got, want, what="oracle vs R runonestep")


@needs_fixture
def test_oracle_runmodel_equals_real_R(oracle):
    """runmodel on the first 72 hours, 50 spin-up steps, reqhgt = NA (dump_reference.R: runmodel_out.csv)."""
    from microclimc_b200 import synth as S
    from tests.common import M, SM, make_cfg
    vegp, soilp = _lists()
    veg = S.pack_vegp([vegp], M); soil = S.pack_soilp([soilp], SM)
    w = S.load_weather()
    f, ts = S.forcing_from_climdata(w)
    r = oracle.run_transient(veg, soil, np.ascontiguousarray(f[:72]), np.array([50.0]), np.array([-5.0]),
                             make_cfg(timestep=ts, spin_steps=50), M, SM, samp_idx=np.arange(1))
    with open(os.path.join(FIX, "runmodel_out.csv")) as fh:
        rows = list(csv.DictReader(fh))
    for k, col in ((0, "tout"), (1, "tleaf")):
        ref = np.array([float(x[col]) for x in rows])
        assert np.max(np.abs(r["series"][:, k, 0] - ref)) <= 1e-6, col
    for k, col in ((2, "relhum"), (3, "SWin"), (4, "LWin")):
        ref = np.array([float(x[col]) for x in rows])
        assert np.max(np.abs(r["series"][:, k, 0] - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-8, col


def test_harness_mechanics_with_an_oracle_made_fixture(oracle, tmp_path, monkeypatch):
    """Writes a fixture in dump_reference.R's format FROM THE ORACLE and runs the comparison above on it: checks the reader,
    the previn <-> state mapping and the argument plumbing of the harness (it says nothing about R)."""
    import tests.test_r_parity as me
    from microclimc_b200 import layout as L, synth as S
    from tests.common import M, SM, make_cfg

    def wlist(d, path):
        with open(path, "w") as fh:
            for k, v in d.items():
                fh.write(k + "," + ",".join(repr(float(x)) for x in np.atleast_1d(v)) + "\n")

    vegp = S.vegp_deciduous()
    sp = S.load_soilparams(); i = sp["Soil.type"].index("Loam")
    soilp = {k: sp[k][i] for k in L.SOIL_SCALARS}; soilp["z"] = oracle.soilinit_z()
    wlist(vegp, tmp_path / "vegp.csv"); wlist(soilp, tmp_path / "soilp.csv")
    w = S.load_weather()
    f, ts = S.forcing_from_climdata(w)
    st0 = oracle.paraminit(M, SM, 15.0, w["temp"][0], w["windspeed"][0], w["relhum"][0], float(np.mean(w["temp"])), w["swrad"][0])
    clim = np.array([[w["temp"][0]], [w["relhum"][0]], [w["pres"][0]], [w["windspeed"][0]], [np.mean(w["temp"])], [w["skyem"][0]],
                     [w["swrad"][0]], [f[0, 6]]])
    st1, _ = oracle.runonestep(S.pack_vegp([vegp], M), S.pack_soilp([soilp], SM), clim, np.array([50.0]), np.array([-5.0]),
                               make_cfg(timestep=3600.0, edgedist=100.0, spin_steps=0), f[0, 8], f[0, 9], np.array([0.3]), np.array([0.3]),
                               st0, M, SM)
    rows = L.state_rows(M, SM)
    wlist({k: st0[sl, 0] for k, sl in rows.items()}, tmp_path / "step_in.csv")
    wlist({k: st1[sl, 0] for k, sl in rows.items()}, tmp_path / "step_out.csv")
    monkeypatch.setattr(me, "FIX", str(tmp_path))
    me.test_oracle_runonestep_equals_real_R.__wrapped__(oracle) if hasattr(me.test_oracle_runonestep_equals_real_R, "__wrapped__") \
        else me.test_oracle_runonestep_equals_real_R(oracle)
