"""The R boundary without R: r-package/src/shim.c is compiled against the miniature R runtime of tests/rstub/ and its
.Call entry points are EXECUTED from Python with vectors laid out as R lays them out (column-major), replaying the argument shapes
r-package/R/microclimc_b200.R builds.  CPU part: the shim builds, registers what the R code calls, and turns library errors into R
errors.  GPU part: the replayed calls give the same numbers as the Python mirror of the same wrappers (api.py) and as the oracle."""
import os
import re

import numpy as np
import pytest

from tests import rshim as R
from tests.common import M, SM, make_cfg, transient_inputs, assert_series_close, assert_state_close, loam
from microclimc_b200 import layout as L, synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_builds_and_registers_every_routine_the_r_code_calls():
    R.lib()
    reg = R.routines()
    rsrc = open(os.path.join(ROOT, "r-package", "R", "microclimc_b200.R")).read()
    called = {}
    for m in re.finditer(r"\.Call\((mcb_\w+)((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)", rsrc):
        name, rest = m.group(1), m.group(2)
        depth = 0; nargs = 0
        for ch in rest:                                   # count top-level commas of the argument list
            if ch in "([": depth += 1
            elif ch in ")]": depth -= 1
            elif ch == "," and depth == 0: nargs += 1
        called.setdefault(name, set()).add(nargs)
    assert called, "no .Call found in the R source"
    for name, arities in called.items():
        assert name in reg, "%s is called from R but not registered in shim.c" % name
        assert arities == {reg[name]}, "%s: R passes %s arguments, shim.c registers %d" % (name, sorted(arities), reg[name])
    for name in reg:
        assert hasattr(R.lib(), name)


def test_library_errors_become_r_errors():
    with pytest.raises(R.RError, match="three canopy layers"):                 # the reference's stop() at code.R:699
        R.call("mcb_create", R.r_int([2]), R.r_int([10]), R.r_int([1]))
    from microclimc_b200._lib import lib
    if lib().mc_device_count() == 0:
        with pytest.raises(R.RError, match="no CPU fallback"):
            R.call("mcb_create", R.r_int([20]), R.r_int([10]), R.r_int([1]))


# ---- replay of the R wrappers' calls (needs the GPU) -----------------------------------------------------------------------------------
def _r_forcing(f):                       # .forcing(): rbind(...) = MC_F_N x T matrix
    return R.r_real(f.T)


def _ctx(m, sm, n=1):
    return R.call("mcb_create", R.r_int([m]), R.r_int([sm]), R.r_int([n]))


@pytest.mark.gpu
def test_runmodel_batch_call_sequence_matches_python_mirror_and_oracle(oracle):
    """runmodel_batch(): .pack_vegp -> ncol x NVEG matrix, .pack_soilp, .forcing -> F_N x T, cfg vector, mcb_run_transient with 1-based
    sample indices and want_full; then the same run restarted from `previn` (code.R:1155-1159) and with reqhgt "all" style full tables."""
    ncol, T = 9, 48
    inp = transient_inputs(oracle, ncol=ncol, T=T)
    cfg = make_cfg(reqhgt=2.0, spin_steps=30)
    ctx = _ctx(M, SM)
    R.call("mcb_set_columns", ctx, R.r_real(inp["veg"].T), R.r_real(inp["soil"].T), R.r_real(inp["lat"]), R.r_real(inp["lon"]))
    R.call("mcb_set_forcing", ctx, _r_forcing(inp["f"]), R.r_real(inp["fs"].T), R.r_real(inp["fo"].T), R.r_null())
    R.call("mcb_set_config", ctx, R.r_real(cfg))
    NS = L.nstate(M, SM)
    sample = np.array([1, 4, 9])                                                           # R is 1-based
    r = R.call("mcb_run_transient", ctx, R.r_real([0]), R.r_real([T]), R.r_null(), R.r_real(sample), R.r_real([ncol]), R.r_int([1]), R.r_int([NS]))
    series = R.to_numpy(R.elt(r, 0), (3, 8, T)); stats = R.to_numpy(R.elt(r, 1), (ncol, 6)); flags = R.to_numpy(R.elt(r, 2))
    full = R.to_numpy(R.elt(r, 3), (3, NS, T))
    state = R.to_numpy(R.call("mcb_get_state", ctx, R.r_int([ncol]), R.r_int([NS])), (ncol, NS))
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, fscale=inp["fs"], foffset=inp["fo"],
                                samp_idx=sample - 1, want_full=True)
    assert_series_close(series.transpose(2, 1, 0), want["series"], "shim series")
    assert_state_close(state.T, want["state"], what="shim state")
    np.testing.assert_allclose(stats.T, want["stats"], rtol=0, atol=1e-6)
    assert flags.dtype == np.int32 and flags.shape == (ncol,)
    for t in (0, T - 1):
        assert_state_close(full[:, :, t].T, want["full"][t], what="shim full table")
    # restart: previn given -> no spin-up, state rows passed as ncol x NSTATE (t(sapply(previn, .state_row)))
    cfg2 = cfg.copy(); cfg2[11] = 0
    ctx2 = _ctx(M, SM)
    R.call("mcb_set_columns", ctx2, R.r_real(inp["veg"].T), R.r_real(inp["soil"].T), R.r_real(inp["lat"]), R.r_real(inp["lon"]))
    R.call("mcb_set_forcing", ctx2, _r_forcing(inp["f"]), R.r_real(inp["fs"].T), R.r_real(inp["fo"].T), R.r_null())
    R.call("mcb_set_config", ctx2, R.r_real(cfg2))
    R.call("mcb_set_state", ctx2, R.r_real(state))
    r2 = R.call("mcb_run_transient", ctx2, R.r_real([0]), R.r_real([T]), R.r_null(), R.r_real(sample), R.r_real([ncol]), R.r_int([0]), R.r_int([NS]))
    assert R.lib().rstub_type(R.elt(r2, 3)) == 0                                            # full = NULL when not asked for
    want2 = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg2, M, SM, state=want["state"], fscale=inp["fs"],
                                 foffset=inp["fo"], samp_idx=sample - 1)
    assert_series_close(R.to_numpy(R.elt(r2, 0), (3, 8, T)).transpose(2, 1, 0), want2["series"], "shim restart")
    R.lib().rstub_finalize(ctx); R.lib().rstub_finalize(ctx2)
    with pytest.raises(R.RError, match="already destroyed"):
        R.call("mcb_set_config", ctx, R.r_real(cfg))


@pytest.mark.gpu
def test_time_varying_vegp_and_theta_matrix_through_the_shim(oracle):
    """.vegt(): array c(ncolv, 2m+1, T); theta matrix sm x T passed as the user holds it."""
    from tests.test_host_logic import _vegt_case
    ncol, T = 4, 40
    inp = transient_inputs(oracle, ncol=ncol, T=T, perturb=False)
    vegt = _vegt_case(ncol, T, M, True)                                                    # C order [T][2m+1][ncol]
    theta = np.clip(0.2 + 0.1 * np.linspace(0, 1, SM)[:, None] + 0.03 * np.cos(np.arange(T) / 5.0)[None, :], 0.05, 0.4)   # sm x T as in R
    cfg = make_cfg(spin_steps=20)
    ctx = _ctx(M, SM)
    R.call("mcb_set_columns", ctx, R.r_real(inp["veg"].T), R.r_real(inp["soil"].T), R.r_real(inp["lat"]), R.r_real(inp["lon"]))
    R.call("mcb_set_forcing", ctx, _r_forcing(inp["f"]), R.r_null(), R.r_null(), R.r_null())
    R.call("mcb_set_vegt", ctx, R.r_real(vegt.transpose(2, 1, 0)), R.r_real([T]), R.r_real([ncol]))     # R array dim c(ncol, 2m+1, T)
    R.call("mcb_set_thetaz", ctx, R.r_real(theta), R.r_real([T]))
    R.call("mcb_set_config", ctx, R.r_real(cfg))
    NS = L.nstate(M, SM)
    r = R.call("mcb_run_transient", ctx, R.r_real([0]), R.r_real([T]), R.r_null(), R.r_real(np.arange(1, ncol + 1)), R.r_real([ncol]), R.r_int([0]),
               R.r_int([NS]))
    want = oracle.run_transient(inp["veg"], inp["soil"], inp["f"], inp["lat"], inp["lon"], cfg, M, SM, vegt=vegt, thetaz=theta.T.copy())
    assert_series_close(R.to_numpy(R.elt(r, 0), (ncol, 8, T)).transpose(2, 1, 0), want["series"], "shim vegt + thetaz")


@pytest.mark.gpu
def test_paraminit_onestep_spinup_and_steady_through_the_shim(oracle):
    ncol = 3
    inp = transient_inputs(oracle, ncol=ncol, T=24)
    NS = L.nstate(M, SM)
    # paraminit(): 1 x 6 argument matrix
    ctx = _ctx(M, SM)
    R.call("mcb_set_columns", ctx, R.r_real(np.zeros((1, L.nveg(M)))), R.r_real(np.zeros((1, L.nsoil(SM)))), R.r_real([0.0]), R.r_real([0.0]))
    R.call("mcb_paraminit", ctx, R.r_real(np.array([[12.0, 11.0, 2.0, 80.0, 9.0, 400.0]])))
    st = R.to_numpy(R.call("mcb_get_state", ctx, R.r_int([1]), R.r_int([NS])), (1, NS))
    np.testing.assert_allclose(st[0], oracle.paraminit(M, SM, 12.0, 11.0, 2.0, 80.0, 9.0, 400.0)[:, 0], rtol=1e-14, atol=1e-14)
    # spinup() then runonestep(): state row 1 x NSTATE, climvars 1 x 8
    cfg = make_cfg(spin_steps=25, reqhgt=1.0)
    c1 = _ctx(M, SM)
    R.call("mcb_set_columns", c1, R.r_real(inp["veg"][:, :1].T), R.r_real(inp["soil"][:, :1].T), R.r_real(inp["lat"][:1]), R.r_real(inp["lon"][:1]))
    R.call("mcb_set_forcing", c1, _r_forcing(inp["f"]), R.r_null(), R.r_null(), R.r_null())
    R.call("mcb_set_config", c1, R.r_real(cfg))
    R.call("mcb_spinup", c1, R.r_int([25]))
    s1 = R.to_numpy(R.call("mcb_get_state", c1, R.r_int([1]), R.r_int([NS])), (1, NS))
    cv = np.array([[14.0, 70.0, 101.0, 3.0, 10.0, 0.8, 350.0, 0.4]])
    R.call("mcb_set_state", c1, R.r_real(s1))
    R.call("mcb_run_onestep", c1, R.r_real(cv), R.r_real([2449900.0]), R.r_real([13.0]), R.r_real([0.3]), R.r_real([0.3]))
    s2 = R.to_numpy(R.call("mcb_get_state", c1, R.r_int([1]), R.r_int([NS])), (1, NS))
    cfg0 = cfg.copy(); cfg0[11] = 0
    want, _ = oracle.runonestep(inp["veg"][:, :1], inp["soil"][:, :1], cv.T.copy(), inp["lat"][:1], inp["lon"][:1], cfg0, 2449900.0, 13.0,
                                np.array([0.3]), np.array([0.3]), s1.T.copy(), M, SM)
    assert_state_close(s2.T, want, what="shim runonestep")
    # runmodelS_batch(): clim 9 x T, nmr 11 x T, result array c(ncol, 7, T)
    clim = S.steady_clim_from_climdata(S.load_weather())[500:596]; T = clim.shape[0]
    nmr = np.zeros((T, 11)); nmr[:, 0] = clim[:, 0] + 0.5; nmr[:, 1] = 0.3
    scfg = np.array([1.0, 1, 2, 1, 0.95])
    c2 = _ctx(M, SM)
    R.call("mcb_set_columns", c2, R.r_real(inp["veg"].T), R.r_real(inp["soil"].T), R.r_real(inp["lat"]), R.r_real(inp["lon"]))
    out = R.to_numpy(R.call("mcb_run_steady", c2, R.r_real(clim.T), R.r_real(nmr.T), R.r_real(scfg), R.r_null(), R.r_real([ncol])), (ncol, 7, T))
    ws = oracle.run_steady(inp["veg"], inp["soil"], clim, nmr, inp["lat"], inp["lon"], scfg, M, SM)["out"]
    np.testing.assert_allclose(out.transpose(2, 1, 0), ws, rtol=1e-9, atol=1e-9)
