"""Real-R parity: compares the in-repo oracle with fixtures produced by the reference itself
(tools/r_parity/dump_reference.R, to be run where R + microctools + microclimc are installed).  The build image has no R, so
the fixtures are absent here and the `needs_fixture` tests skip; until they exist every parity statement of the repository reads
"CUDA path vs in-repo oracle" (DESIGN.md §2).  Once tests/golden/r_reference/ exists they RUN, and a mismatch fails the suite.
`test_harness_mechanics_*` exercise every comparison below on fixtures written in the same file formats from the oracle itself:
that checks the readers, the previn <-> state mapping and the argument plumbing of the harness (it says nothing about R)."""
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




@needs_fixture
def test_oracle_merged_layer_step_equals_real_R(oracle):
    """ONE runonestep with timestep = 60 s from the spun-up state (dump_reference.R: step60_*): the merged-layer branch, i.e. the
    layermerge / layeraverage / layerinterp definitions designed in this repository against the real microctools ones."""
    from microclimc_b200 import synth as S
    from tests.common import M, SM, make_cfg, assert_state_close
    vegp, soilp = _lists()
    veg = S.pack_vegp([vegp], M); soil = S.pack_soilp([soilp], SM)
    cvd = read_list(os.path.join(FIX, "step60_climvars.csv"))
    clim = np.array([[cvd[k][0]] for k in ("tair", "relhum", "pk", "u2", "tsoil", "skyem", "Rsw", "dp")])
    st0 = _state_from_list(read_list(os.path.join(FIX, "step60_in.csv")), M, SM)
    want = _state_from_list(read_list(os.path.join(FIX, "step60_out.csv")), M, SM)
    got, nm = oracle.runonestep(veg, soil, clim, np.array([50.0]), np.array([-5.0]), make_cfg(timestep=60.0, edgedist=100.0, spin_steps=0),
                                cvd["jd"][0], cvd["lt"][0], np.array([0.3]), np.array([0.3]), st0, M, SM)
    assert nm[0] > 1, "the 60 s step was meant to take the merged-layer branch"
    assert_state_close(got, want, what="oracle vs R, merged layers")


@needs_fixture
def test_oracle_runmodel_at_height_equals_real_R(oracle):
    from microclimc_b200 import synth as S
    from tests.common import M, SM, make_cfg
    vegp, soilp = _lists()
    veg = S.pack_vegp([vegp], M); soil = S.pack_soilp([soilp], SM)
    f, ts = S.forcing_from_climdata(S.load_weather())
    r = oracle.run_transient(veg, soil, np.ascontiguousarray(f[:72]), np.array([50.0]), np.array([-5.0]),
                             make_cfg(timestep=ts, reqhgt=1.0, spin_steps=50), M, SM, samp_idx=np.arange(1))
    with open(os.path.join(FIX, "runmodel_h1_out.csv")) as fh:
        rows = list(csv.DictReader(fh))
    for k, col in ((0, "tout"), (1, "tleaf")):
        ref = np.array([float(x[col]) for x in rows])
        assert np.max(np.abs(r["series"][:, k, 0] - ref)) <= 1e-6, col


def _steady_case(oracle):
    """runmodelS inputs of dump_reference.R from the fixture directory -> oracle output [T][7]."""
    from microclimc_b200 import synth as S
    from tests.common import M, SM
    vegp, soilp = _lists()
    veg = S.pack_vegp([vegp], M); soil = S.pack_soilp([soilp], SM)
    with open(os.path.join(FIX, "runmodelS_out.csv")) as fh:
        rows = list(csv.DictReader(fh))
    with open(os.path.join(FIX, "nmrout.csv")) as fh:
        nr = list(csv.DictReader(fh))
    sel = np.array([int(float(x["row"])) - 1 for x in rows])
    clim = S.steady_clim_from_climdata(S.load_weather())[sel]
    T = sel.size
    nmr = np.zeros((T, 11))
    cols = list(nr[0].keys())
    for t in range(T):
        nmr[t, :3] = [float(nr[t]["D0cm"]), float(nr[t]["WC0cm"]), float(nr[t]["SNOWDEP"])]
        nmr[t, 3:] = [float(nr[t][c]) for c in cols[3:11]]
    out = oracle.run_steady(veg, soil, clim, nmr, np.array([50.0]), np.array([-5.0]), np.array([1.0, 1, 2, 1, 0.95]), M, SM)["out"][:, :, 0]
    return rows, out, nmr


@needs_fixture
def test_oracle_runmodelS_equals_real_R(oracle):
    from microclimc_b200 import api
    rows, out, nmr = _steady_case(oracle)
    tleaf = api._q16_remap(out[:, 1], out[:, 0], nmr[:, 2], 1.0)
    for k, col, tol, rel in ((0, "Tloc", 1e-6, False), (2, "RHloc", 1e-8, True), (3, "RSWloc", 1e-8, True), (4, "RLWloc", 1e-8, True),
                             (5, "windspeed", 1e-8, True)):
        ref = np.array([float(x[col]) if x[col] not in ("NA", "") else np.nan for x in rows])
        assert np.array_equal(np.isnan(ref), np.isnan(out[:, k])), col
        d = np.abs(np.nan_to_num(out[:, k]) - np.nan_to_num(ref))
        assert np.max(d / np.maximum(np.abs(np.nan_to_num(ref)), 1.0) if rel else d) <= tol, col
    ref = np.array([float(x["tleaf"]) if x["tleaf"] not in ("NA", "") else np.nan for x in rows])
    assert np.array_equal(np.isnan(ref), np.isnan(tleaf)) and np.nanmax(np.abs(ref - tleaf)) <= 1e-6


# ---- microctools call traces (mt_<function>.csv): candidate definitions of SURVEY.md Appendix A against the real package -----------------
def read_trace(path):
    """'<call>,<arg or =field>,<values...>' rows -> list of (args [(name, array)], results {field: array}) per call, in file order."""
    calls = {}
    with open(path) as f:
        for row in csv.reader(f):
            if not row:
                continue
            k = int(row[0]); name = row[1]
            v = np.array([float(x) if x not in ("NA", "NaN", "") else np.nan for x in row[2:]])
            c = calls.setdefault(k, ([], {}))
            if name.startswith("="):
                c[1][name[1:]] = v
            else:
                c[0].append((name, v))
    return [calls[k] for k in sorted(calls)]


def _compat_table(O):
    """microctools function -> (number of leading positional arguments used, defaults for the missing ones, oracle probe)."""
    L = O.lib()
    return {
        "phair": (2, [101.3], lambda a: L.orc_phair(a[0], a[1])),
        "cpair": (1, [], lambda a: L.orc_cpair(a[0])),
        "satvap": (1, [], lambda a: L.orc_satvap(a[0])),
        "dewpoint": (2, [], lambda a: L.orc_dewpoint(a[0], a[1])),
        "zeroplanedis": (2, [], lambda a: L.orc_zeroplanedis(a[0], a[1])),
        "roughlength": (3, [0.003], lambda a: L.orc_roughlength(a[0], a[1], a[2])),
        "mixinglength": (2, [], lambda a: L.orc_mixinglength(a[0], a[1])),
        "attencoef": (5, [1.0], lambda a: L.orc_attencoef(a[0], a[1], a[2], a[3], a[4])),
        "gforcedfree": (6, [101.3, 1.0], lambda a: L.orc_gforcedfree(*a[:6])),
        "gcanopy": (11, [101.3], lambda a: L.orc_gcanopy(*a[:11])),
    }


def check_traces(O, fixdir, rtol=1e-10):
    """Compares every traced call of the functions in _compat_table with the oracle's compat layer; returns {function: calls checked}."""
    done = {}
    for fn, (nargs, defaults, probe) in _compat_table(O).items():
        path = os.path.join(fixdir, "mt_%s.csv" % fn)
        if not os.path.exists(path):
            continue
        n = 0
        for args, res in read_trace(path):
            vals = [v for _, v in args][:nargs]
            need = nargs - len(vals)
            if need > 0:
                vals = vals + [np.array([d]) for d in defaults[len(defaults) - need:]]
            size = max(v.size for v in vals)
            want = res[""]
            for e in range(size):
                a = [float(v[e % v.size]) for v in vals]
                got = probe(a)
                w = float(want[e % want.size])
                assert (np.isnan(got) and np.isnan(w)) or abs(got - w) <= rtol * max(abs(w), 1e-12), (fn, a, got, w)
            n += 1
        done[fn] = n
    return done


@needs_fixture
def test_microctools_traces_equal_the_oracle_compat_layer(oracle):
    done = check_traces(oracle, FIX)
    assert done and all(v > 0 for v in done.values()), done


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
    # the 60 s merged-layer step from a spun-up state
    cfg_spin = make_cfg(timestep=ts, spin_steps=50)
    sp = oracle.run_transient(S.pack_vegp([vegp], M), S.pack_soilp([soilp], SM), f[:1], np.array([50.0]), np.array([-5.0]), cfg_spin, M, SM)["state"]
    i60 = 180 * 24 + 13 - 1
    cv60 = dict(tair=w["temp"][i60], relhum=w["relhum"][i60], pk=w["pres"][i60], u2=w["windspeed"][i60], tsoil=float(np.mean(w["temp"])),
                skyem=w["skyem"][i60], Rsw=w["swrad"][i60], dp=f[i60, 6], jd=f[i60, 8], lt=f[i60, 9])
    wlist(cv60, tmp_path / "step60_climvars.csv")
    clim60 = np.array([[cv60[k]] for k in ("tair", "relhum", "pk", "u2", "tsoil", "skyem", "Rsw", "dp")])
    s60, _ = oracle.runonestep(S.pack_vegp([vegp], M), S.pack_soilp([soilp], SM), clim60, np.array([50.0]), np.array([-5.0]),
                               make_cfg(timestep=60.0, edgedist=100.0, spin_steps=0), cv60["jd"], cv60["lt"], np.array([0.3]), np.array([0.3]), sp, M, SM)
    wlist({k: sp[sl, 0] for k, sl in rows.items()}, tmp_path / "step60_in.csv")
    wlist({k: s60[sl, 0] for k, sl in rows.items()}, tmp_path / "step60_out.csv")
    # runmodel data.frames
    for name, rq in (("runmodel_out.csv", np.nan), ("runmodel_h1_out.csv", 1.0)):
        r = oracle.run_transient(S.pack_vegp([vegp], M), S.pack_soilp([soilp], SM), np.ascontiguousarray(f[:72]), np.array([50.0]), np.array([-5.0]),
                                 make_cfg(timestep=ts, reqhgt=rq, spin_steps=50), M, SM, samp_idx=np.arange(1))["series"][:, :, 0]
        with open(tmp_path / name, "w") as fh:
            fh.write("obs_time,reftemp,tout,tleaf,relhum,SWin,LWin,H,L,G\n")
            for t in range(72):
                fh.write("x,0," + ",".join(repr(float(r[t, k])) for k in (0, 1, 2, 3, 4, 6, 5, 7)) + "\n")
    # runmodelS + nmrout
    sel = 300 + np.arange(240)
    clim = S.steady_clim_from_climdata(w)[sel]
    nmr = np.zeros((240, 11)); nmr[:, 0] = clim[:, 0] + 0.3; nmr[:, 1] = 0.3
    nmr[:, 2] = np.where(clim[:, 0] < 3, 12 + 8 * np.sin(np.arange(1, 241) / 17.0), 0.0)
    nmr[:, 3:] = np.minimum(0, clim[:, :1] - 1 - 0.3 * np.arange(1, 9)[None, :])
    so = oracle.run_steady(S.pack_vegp([vegp], M), S.pack_soilp([soilp], SM), clim, nmr, np.array([50.0]), np.array([-5.0]),
                           np.array([1.0, 1, 2, 1, 0.95]), M, SM)["out"][:, :, 0]
    from microclimc_b200 import api
    tl = api._q16_remap(so[:, 1], so[:, 0], nmr[:, 2], 1.0)
    num = lambda v: "NA" if np.isnan(v) else repr(float(v))
    with open(tmp_path / "nmrout.csv", "w") as fh:
        fh.write("D0cm,WC0cm,SNOWDEP," + ",".join("SN%d" % q for q in range(1, 9)) + "\n")
        for t in range(240):
            fh.write(",".join(num(v) for v in nmr[t]) + "\n")
    with open(tmp_path / "runmodelS_out.csv", "w") as fh:
        fh.write("obs_time,Tref,Tloc,tleaf,RHref,RHloc,RSWloc,RLWloc,windspeed,dp,row\n")
        for t in range(240):
            fh.write("x,0,%s,%s,0,%s,%s,%s,%s,%s,%d\n" % (num(so[t, 0]), num(tl[t]), num(so[t, 2]), num(so[t, 3]), num(so[t, 4]), num(so[t, 5]),
                                                            num(so[t, 6]), sel[t] + 1))
    # microctools traces in dump_traces()'s format
    Lb = oracle.lib()
    with open(tmp_path / "mt_satvap.csv", "w") as fh:
        fh.write("1,tc,-5,0.5,12\n1,=,%r,%r,%r\n" % (Lb.orc_satvap(-5.0), Lb.orc_satvap(0.5), Lb.orc_satvap(12.0)))
    with open(tmp_path / "mt_phair.csv", "w") as fh:
        fh.write("1,tc,10\n1,=,%r\n2,tc,10\n2,pk,99\n2,=,%r\n" % (Lb.orc_phair(10.0, 101.3), Lb.orc_phair(10.0, 99.0)))
    with open(tmp_path / "mt_gcanopy.csv", "w") as fh:
        a = [1.5, 10.0, 8.0, 12.0, 11.0, 15.0, 3.0, 0.2, 0.5, 1.1, 100.0]
        fh.write("".join("1,a%d,%r\n" % (i, v) for i, v in enumerate(a)) + "1,=,%r\n" % Lb.orc_gcanopy(*a))
    monkeypatch.setattr(me, "FIX", str(tmp_path))
    me.test_oracle_runonestep_equals_real_R(oracle)
    me.test_oracle_merged_layer_step_equals_real_R(oracle)
    me.test_oracle_runmodel_equals_real_R(oracle)
    me.test_oracle_runmodel_at_height_equals_real_R(oracle)
    me.test_oracle_runmodelS_equals_real_R(oracle)
    assert me.check_traces(oracle, str(tmp_path)) == {"phair": 2, "satvap": 1, "gcanopy": 1}
    # and a wrong value is caught
    with open(tmp_path / "mt_satvap.csv", "w") as fh:
        fh.write("1,tc,5\n1,=,0.5\n")
    with pytest.raises(AssertionError):
        me.check_traces(oracle, str(tmp_path))
