import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand from oracle/."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def hostdbg():
    """Host (g++) build of the DEVICE column code, for CPU-only logic checks. Test-only; never shipped."""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "hostdbg", "dbg_host.cpp")
    out = os.path.join(ROOT, "tests", "hostdbg", "libdbg_host.so")
    deps = [src] + [os.path.join(ROOT, "microclimc_b200", "csrc", f) for f in
                    ("mc_step.cuh", "mc_driver.cuh", "mc_steady.cuh", "mc_phys.cuh", "mc_fast.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++", "-o", out, src])
    return C.CDLL(out)
