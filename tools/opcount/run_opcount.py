"""Operator census of one column-timestep -> profiles/cost_per_colstep.json (roofline denominators).

Builds tools/opcount/libopcount.so (oracle + shipped column code compiled with a counting scalar), runs the C3
workload of bench.py on a small batch (hourly steps, 20 canopy + 10 soil nodes, synthetic forcing with the per-column
perturbation, state after spin-up) for two run lengths and reports the per-step difference, so one-off per-launch work
(parameter load, hoisted geometry) is excluded.  TEST/MEASUREMENT INFRASTRUCTURE — not part of the product.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
KINDS = ["add", "mul", "div", "sqrt", "exp", "log", "pow", "cbrt", "trig", "cmp"]
# FP64-pipe instruction slots (one DADD/DMUL/DFMA issue each) of an efficient sm_100a implementation of each operator;
# the special-function-unit seeds (MUFU.RCP64H/RSQ64H/EX2) and the integer work are not FP64-pipe slots.
# SURVEY.md §8(d) constants; compares/selects/abs are carried at 0 (standard flop counting ignores them).
WEIGHT = dict(add=1, mul=1, div=8, sqrt=8, exp=20, log=22, pow=45, cbrt=15, trig=25, cmp=0)
PD = C.POINTER(C.c_double)


def build():
    so = os.path.join(HERE, "libopcount.so")
    deps = [os.path.join(HERE, "opcount.cpp")] + [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith(("pp",))] \
        + [os.path.join(ROOT, "microclimc_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "microclimc_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, os.path.join(HERE, "opcount.cpp")])
    return C.CDLL(so)


def counts(lib):
    out = (C.c_ulonglong * len(KINDS))()
    lib.cnt_get(out)
    return np.array(list(out), dtype=np.float64)


def run(lib, fn, w, cfg, T, state):
    from microclimc_b200 import layout as L
    p = lambda a: None if a is None else a.ctypes.data_as(PD)
    ncol = w["veg"].shape[1]
    f = np.ascontiguousarray(w["f"][:T])
    st = state.copy()
    series = np.zeros((T, 8, 1)); stats = np.zeros((6, ncol)); flags = np.zeros(ncol, dtype=np.int32)
    samp = np.zeros(1, dtype=np.int64)
    lib.cnt_reset()
    rc = getattr(lib, fn)(C.c_int64(ncol), 20, 10, C.c_int64(T), p(w["veg"]), p(w["soil"]), p(f), p(w["fs"]), p(w["fo"]), None, p(w["lat"]),
                          p(w["lon"]), p(cfg), None, p(st), p(series), samp.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(1), p(stats), None,
                          flags.ctypes.data_as(C.POINTER(C.c_int32)), 1)
    assert rc == 0
    return counts(lib), st


def main():
    import bench
    lib = build()
    ncol, T1, T2 = 48, 168, 336
    w = bench.workload(ncol, T2, seed=1234)
    cfg_spin = bench.make_cfg(w["ts"], 200)
    cfg = bench.make_cfg(w["ts"], 0)
    out = {"_about": "FP64 operator census per column-timestep; tools/opcount/run_opcount.py", "weights_fp64_slots": WEIGHT,
           "workload": "C3 sample: %d columns, hourly, m=20, sm=10, state after 200-step spin-up, steps %d..%d of the synthetic year" % (ncol, T1, T2),
           "algorithmic_bytes_per_colstep": 152.0}
    for name, fn in (("oracle_literal", "orc_run_transient"), ("general_fused", "dbg_run_transient"), ("shipped_streaming", "dbg_run_transient_fast")):
        _, st0 = run(lib, fn, w, cfg_spin, 1, np.zeros((bench.M * 12 + 10 + 2 * bench.SM, ncol)))
        c1, _ = run(lib, fn, w, cfg, T1, st0)
        c2, _ = run(lib, fn, w, cfg, T2, st0)
        per = (c2 - c1) / (ncol * (T2 - T1))
        ops = dict(zip(KINDS, [float(x) for x in per]))
        slots = float(sum(WEIGHT[k] * ops[k] for k in KINDS))
        out[name] = {"ops_per_colstep": ops, "fp64_slots_per_colstep": slots, "fp64_flop_per_colstep": 2.0 * slots,
                     "one_off_ops_per_column": dict(zip(KINDS, [float(x) for x in (c1 - per * ncol * T1) / ncol]))}
        print(name, json.dumps(out[name]["ops_per_colstep"]), "slots", slots)
    # the roofline uses the SHIPPED algorithm's count (hoisting removes work the literal restatement repeats every step)
    out["fp64_flop_per_colstep"] = out["shipped_streaming"]["fp64_flop_per_colstep"]
    out["fp64_slots_per_colstep"] = out["shipped_streaming"]["fp64_slots_per_colstep"]
    prev = {}
    path = os.path.join(ROOT, "profiles", "cost_per_colstep.json")
    if os.path.exists(path):
        prev = json.load(open(path))
    if "dram_bytes_per_colstep" in prev:
        out["dram_bytes_per_colstep"] = prev["dram_bytes_per_colstep"]; out["dram_bytes_source"] = prev.get("dram_bytes_source")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
