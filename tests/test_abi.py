"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/microclimc_b200.h
declares; layouts of the header, the Python host mirror and the oracle agree; there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from microclimc_b200 import layout as L, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "microclimc_b200.h")).read()


def _declared_functions():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", code)))


def _enum(prefix):
    """names of the enum whose members start with `prefix`, in order."""
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for body in re.findall(r"enum\s*\{(.*?)\}", code, flags=re.S):
        names = [x.strip().split("=")[0].strip() for x in body.split(",") if x.strip()]
        if names and names[0].startswith(prefix):
            return names
    raise AssertionError(prefix)


def test_header_symbols_are_all_exported():
    names = _declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_version_and_sizes():
    lib = _lib.lib()
    assert lib.mc_version() == 1
    for m, sm in [(3, 2), (20, 10), (64, 32)]:
        assert lib.mc_nveg(m) == L.nveg(m) and lib.mc_nsoil(sm) == L.nsoil(sm) and lib.mc_nstate(m, sm) == L.nstate(m, sm)


def test_header_enums_match_python_layout(oracle):
    assert [n[5:].lower() for n in _enum("MC_V_")[:-1]] == [s.lower() for s in L.VEG_SCALARS]
    assert [n[5:].lower() for n in _enum("MC_S_")[:-1]] == [s.lower().replace("psi_e", "psie") for s in L.SOIL_SCALARS]
    assert [n[5:].lower() for n in _enum("MC_P_")[:-1]] == [s.lower().replace("psi_m", "psim") for s in L.STATE_SCALARS]
    assert [n[5:].lower() for n in _enum("MC_F_")[:-1]] == [s.lower() for s in L.FORCING]
    assert len(_enum("MC_C_")) - 1 == L.NCFG and len(_enum("MC_O_")) - 1 == len(L.SERIES)
    assert len(_enum("MC_SC_")) - 1 == len(L.STEADY_CLIM) and len(_enum("MC_SO_")) - 1 == len(L.STEADY_OUT)
    assert len(_enum("MC_SG_")) - 1 == len(L.STEADY_CFG)
    o = oracle.lib()
    assert o.orc_nstate(20, 10) == L.nstate(20, 10) and o.orc_nveg(20) == L.nveg(20) and o.orc_nsoil(10) == L.nsoil(10)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    lib = _lib.lib()
    if lib.mc_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.MicroclimcError, match="no CPU fallback"):
        _lib.Context(20, 10)
    from microclimc_b200 import api
    with pytest.raises(_lib.MicroclimcError):
        api.paraminit(20, 10, 10, 15, 2, 80, 11, 500)


def test_layer_guards_are_reported_before_device_use():
    with pytest.raises(ValueError, match="three canopy layers"):      # stop() at R/code.R:699
        _lib.Context(2, 10)
    with pytest.raises(_lib.MicroclimcError, match="supported sizes"):
        _lib.Context(65, 10)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "microclimc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "mc_oracle" not in src and "oracle/" not in src, os.path.join(dp, f)


def test_r_shim_compiles_against_stub_headers_and_binds_only_declared_symbols():
    """r-package/src/shim.c is the .Call binding a maintainer adds. R is not installed here, so it is
    syntax-checked against minimal stand-in R headers (tests/rstub) and every mc_* call must be a declared ABI symbol."""
    import subprocess
    shim = os.path.join(ROOT, "r-package", "src", "shim.c")
    subprocess.check_call(["gcc", "-std=c11", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "rstub"), "-I" + os.path.join(ROOT, "include"), shim])
    used = set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", open(shim).read(), flags=re.S)))
    assert used and used <= set(_declared_functions()), used - set(_declared_functions())
    rsrc = open(os.path.join(ROOT, "r-package", "R", "microclimc_b200.R")).read()
    for fn in ("paraminit", "soilinit", "runonestep", "spinup", "runmodel", "runmodelS", "runmodel_batch", "runmodelS_batch"):
        assert re.search(r"^%s <- function\(" % fn, rsrc, flags=re.M), fn
    # reference signatures kept verbatim (R/code.R:687-689, 1125-1129)
    assert "runonestep <- function(climvars, previn, vegp, soilp, timestep, tme, lat, long, edgedist = 1000, sdepth = 2, reqhgt = NA, zu = 2," in rsrc
    assert "runmodel <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, merid = 0," in rsrc
