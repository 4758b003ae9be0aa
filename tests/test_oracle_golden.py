"""Pins the CPU oracle (oracle/*.hpp, test infrastructure) to everything the reference tree itself determines.

The reference has NO tests and no golden vectors, and R cannot run here, so the pins are:
 * known answers of the self-contained in-tree functions (Thomas, paraminit, soilinit);
 * tests/golden/intree_known_answers.json: an independent numpy transliteration of those functions plus leafem,
   .snowtempf, windprofile, .windprofile and tleafS (tools/make_golden.py);
 * decoded contents of the bundled datasets.
Parity of the oracle with real R + microctools remains UNPINNED."""
import json
import os

import numpy as np
import pytest

from microclimc_b200 import layout as L, synth as S

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intree_known_answers.json")))


# ---- SURVEY.md Appendix F -------------------------------------------------------------------------------------------
def test_thomas_appendix_f(oracle):
    tc = [10, 12, 14, 13, 11, 9]; k = [5, 4, 3, 2, 1.5]; cd = [20, 25, 30, 35]
    np.testing.assert_allclose(oracle.thomas(tc, 9.5, 10.5, k, cd, 0.6, 0.0),
                               [10.5, 11.943994456275151, 13.615607995578692, 12.95009265324301, 11.03746930414802, 9.5], rtol=0, atol=1e-14)
    np.testing.assert_allclose(oracle.thomas(tc, 9.5, 10.5, k, cd, 0.6, [0.5, 0, -0.25, 0]),
                               [10.5, 12.374296118813263, 13.627967257440343, 12.715740684503434, 11.032584604350516, 9.5], rtol=0,
                               atol=1e-14)
    # clamp-active case (quirk Q11): nodes 2, 3 and 5 hit the old-stencil limits
    np.testing.assert_allclose(oracle.thomas(tc, 30, -5, [50, 4, 3, 2, 50], 0.1, 1.0, 0.0),
                               [-5, 10, 12, 13.898158373603554, 13, 30], rtol=0, atol=1e-14)


def test_paraminit_appendix_f(oracle):
    st = oracle.paraminit(20, 10, 10, 15, 2, 80, 11, 500)[:, 0]
    r = L.state_rows(20, 10)
    np.testing.assert_allclose(st[r["tc"]][:3], [12.379310344827587, 12.517241379310345, 12.655172413793103], rtol=0, atol=1e-14)
    np.testing.assert_allclose(st[r["tc"]][-2:], [14.862068965517242, 15], rtol=0, atol=1e-14)
    np.testing.assert_allclose(st[r["soiltc"]][[0, 1, 2, -2, -1]],
                               [12.241379310344827, 12.10344827586207, 11.96551724137931, 11.137931034482758, 11], rtol=0, atol=1e-14)
    np.testing.assert_allclose(st[r["Rabs"]][[0, 1, 2, -2, -1]], [120, 141.05263157894737, 162.10526315789474, 498.94736842105266, 520],
                               rtol=1e-15)
    np.testing.assert_allclose(st[r["tleaf"]][[0, -1]], [13.579310344827586, 20.2], rtol=1e-15)
    np.testing.assert_allclose(st[r["gt"]], [20] + [10] * 19 + [2.5], rtol=1e-15)
    np.testing.assert_allclose(st[r["gv"]][[0, 1, -1]], [0.25, 0.253684210526316, 0.32], rtol=1e-14)
    np.testing.assert_allclose(st[r["gha"]][[0, 1, -1]], [0.13, 0.133157894736842, 0.19], rtol=1e-14)
    np.testing.assert_allclose(st[r["z"]][[0, 1, -1]], [0.25, 0.75, 9.75], rtol=1e-15)
    np.testing.assert_allclose(st[r["sz"]][[0, 1, 2, -2, -1]], [0.007603787926411, 0.040693269549982, 0.108558399571067, 1.5498756976766, 2],
                               rtol=1e-12)
    np.testing.assert_allclose(st[r["uz"]][[0, 1, -1]], [0.1, 0.2, 2], rtol=1e-15)
    for k, v in dict(tabove=15, zabove=10, relhum=80, tair=15, tsoil=11, pk=101.3, H=0, G=0, psi_m=0).items():
        assert st[r[k]][0] == v
    assert np.all(st[r["rh"]] == 80) and np.all(st[r["L"]] == 0)
    np.testing.assert_array_equal(st[r["Rswin"]], st[r["Rabs"]])
    np.testing.assert_allclose(st[r["Rlwin"]], 0.2 * st[r["Rabs"]], rtol=1e-15)


def test_soilinit_appendix_f(oracle):
    z = oracle.soilinit_z()
    np.testing.assert_allclose(z, [0.126191468896039, 0.289911865471078, 0.471601851357974, 0.666042565921499, 0.870550563296124,
                                   1.083456541736921, 1.303609881132773, 1.530163999664059, 1.762467052249758, 2], rtol=1e-13)


# ---- independent transliteration (tools/make_golden.py) -------------------------------------------------------------
@pytest.mark.parametrize("i", range(len(GOLD["Thomas"])))
def test_thomas_golden(oracle, i):
    c = GOLD["Thomas"][i]
    cd = c["cd"] if len(c["cd"]) > 1 else c["cd"][0]
    X = c["X"] if len(c["X"]) > 1 else c["X"][0]
    got = oracle.thomas(c["tc"], c["tmsoil"], c["tair"], c["k"], cd, c["f"], X)
    np.testing.assert_allclose(got, c["tn"], rtol=0, atol=1e-13)


def test_leafem_golden(oracle):
    for c in GOLD["leafem"]:
        assert abs(oracle.lib().orc_leafem(c["tc"], c["vegem"]) - c["out"]) <= 1e-12 * c["out"]


@pytest.mark.parametrize("i", range(len(GOLD["paraminit"])))
def test_paraminit_golden(oracle, i):
    c = GOLD["paraminit"][i]
    m, sm = int(c["args"][0]), int(c["args"][1])
    st = oracle.paraminit(m, sm, *c["args"][2:])[:, 0]
    for k, sl in L.state_rows(m, sm).items():
        np.testing.assert_allclose(st[sl], np.atleast_1d(c["out"][k]), rtol=1e-14, atol=1e-14, err_msg=k)


def test_soilinit_golden(oracle):
    for c in GOLD["soilinit_z"]:
        z = oracle.soilinit_z(int(c["m"]), c["sdepth"], float("nan") if c["reqdepth"] is None else c["reqdepth"])
        np.testing.assert_allclose(z, c["z"], rtol=1e-14)


def test_snowtempf_golden(oracle):
    import ctypes as C
    for c in GOLD["snowtempf"]:
        sd = np.array(c["snowdepth"]); SN = np.array(c["SN"]).reshape(sd.size, 8)
        for t in range(sd.size):
            o = C.c_double(0)
            rc = oracle.lib().orc_snowtempf(sd[t], SN[t].ctypes.data_as(C.POINTER(C.c_double)), c["reqhgt"], C.byref(o))
            assert rc == 0
            assert abs(o.value - c["out"][t]) <= 1e-13


def test_snowtempf_layer_lookup_out_of_range(oracle):
    """reqhgt < 0.025 m would index snow[, 11] in R ("undefined columns selected"); the oracle reports it."""
    import ctypes as C
    SN = np.zeros(8); o = C.c_double(0)
    assert oracle.lib().orc_snowtempf(0.5, SN.ctypes.data_as(C.POINTER(C.c_double)), 0.01, C.byref(o)) != 0
    assert oracle.lib().orc_snowtempf(0.5, SN.ctypes.data_as(C.POINTER(C.c_double)), 3.5, C.byref(o)) != 0


def test_windprofile_golden(oracle):
    for c in GOLD["windprofile"]:
        for zo, want in zip(c["zo"], c["out"]):
            got = oracle.lib().orc_windprofile(c["ui"], c["zi"], zo, c["a"], c["PAI"], c["hgt"], c["psi_m"], float("nan"), 0.004)
            assert (np.isnan(got) and np.isnan(want)) or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (c, zo, got, want)
    for c in GOLD["dot_windprofile"]:
        got = oracle.lib().orc_windprofile2(c["ui"], c["zi"], c["zo"], c["a"], c["hgt"], c["PAI"], c["psi_m"])
        assert (np.isnan(got) and np.isnan(c["out"])) or abs(got - c["out"]) <= 1e-12 * max(1.0, abs(c["out"]))


def test_tleafS_golden(oracle):
    import ctypes as C
    g = GOLD["tleafS"]
    order = ["tair", "tground", "relhum", "pk", "theta", "gtt", "gt0", "gha", "gv", "gL", "Rabs", "vegem", "soilb", "Psie", "Smax",
             "surfwet", "leafdens"]
    n = len(g["tleaf"])
    for i in range(n):
        a = np.array([g["args"][k][i] for k in order]); o = np.zeros(3)
        oracle.lib().orc_tleafS(a.ctypes.data_as(C.POINTER(C.c_double)), o.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(o, [g["tleaf"][i], g["tn"][i], g["rh"][i]], rtol=1e-11, atol=1e-11)


# ---- bundled datasets -------------------------------------------------------------------------
def test_bundled_datasets_match_appendix_d():
    w = S.load_weather()
    assert len(w["temp"]) == 8760 and str(w["obs_time"][0]).startswith("1995-01-01 00:00") and str(w["obs_time"][-1]).startswith("1995-12-31 23:00")
    np.testing.assert_allclose(w["temp"][:5], [2.8, 4.5, 5.0, 4.7, 4.8])
    np.testing.assert_allclose(w["relhum"][:5], [82.3555, 73.5754, 60.0152, 48.9327, 50.8530], atol=5e-5)
    np.testing.assert_allclose(w["windspeed"][:3], [3.601108, 6.173328, 8.231104], atol=5e-7)
    assert abs(np.mean(w["temp"]) - 11.3383) < 1e-4 and w["temp"].min() == -3.2 and w["temp"].max() == 27.0
    assert np.count_nonzero(w["swrad"] == 0) == 4388 and abs(w["swrad"].max() - 949.445) < 1e-3
    assert np.count_nonzero(w["difrad"] > w["swrad"]) == 523 and abs(w["difrad"].min() + 36.78) < 1e-2      # quirk Q19 inputs
    dp = S.load_dailyprecip()
    assert dp.size == 365 and abs(dp.sum() - 719.905) < 1e-3
    np.testing.assert_allclose(dp[:3], [0.7775959, 0.02159514, 3.71519309], rtol=1e-6)
    sp = S.load_soilparams()
    assert sp["Soil.type"] == ["Sand", "Loamy sand", "Sandy loam", "Loam", "Silt loam", "Sandy clay loam", "Clay loam", "Silty clay loam",
                               "Sandy clay", "Silty clay", "Clay"]
    i = sp["Soil.type"].index("Loam")
    want = dict(Smax=0.422, Smin=0.074, alpha=0.0322, n=1.55, Ksat=19.69, Vq=0.12, Vm=0.44, Vo=0.018, Mc=0.0844, rho=1.5135064, b=4.5, psi_e=1.1)
    for k, v in want.items():
        assert abs(sp[k][i] - v) < 1e-9, k
    cv = S.load_climvars()
    assert cv["tair"] == 16 and cv["relhum"] == 90 and cv["pk"] == 101.3 and cv["u"] == 2.1 and cv["Rsw"] == 500 and cv["dp"] is None


def test_forcing_keeps_dp_unclamped_q19():
    """runmodel: dp = difrad/swrad, only non-finite -> 0.5 (code.R:1166-1167); values outside [0,1] are kept."""
    f, ts = S.forcing_from_climdata(S.load_weather())
    assert ts == 3600.0 and f.shape == (8760, L.NFORC)
    dp = f[:, L.FORCING.index("dp")]
    assert np.all(np.isfinite(dp)) and dp.max() > 1.0 and dp.min() < 0.0
    assert np.count_nonzero(dp == 0.5) >= 4388
    assert f[0, L.FORCING.index("jd")] == 2449719.0 and f[13, L.FORCING.index("lt")] == 13.0
