#!/usr/bin/env python
"""Transient run of a habitat-class grid (K parameter sets, columns ordered by class, per-column latitude and forcing perturbation:
the transient counterpart of BASELINE config 4's mixed-habitat grid) with and without mc_set_classes.  With classes, tiles of one class
read their per-layer geometry rows from L2 instead of streaming them from DRAM.  Prints one JSON line; run under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_transient_fast` for the bytes per column-step.
  python tools/bench_classes.py [--columns 1000000] [--classes 4] [--hours 168] [--reps 3] [--mode both|on|off]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1_000_000); ap.add_argument("--classes", type=int, default=4)
    ap.add_argument("--hours", type=int, default=168); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--spin", type=int, default=50)
    ap.add_argument("--mode", default="both")
    a = ap.parse_args()
    from microclimc_b200._lib import Context
    from microclimc_b200 import synth as S
    K, n = a.classes, a.columns
    base = B.workload(K, a.hours * (a.reps + 1), seed=77)
    cls = np.sort(np.random.default_rng(5).integers(0, K, n)).astype(np.int32)            # ordered by class
    veg = np.ascontiguousarray(base["veg"][:, cls]); soil = np.ascontiguousarray(base["soil"][:, cls])
    fs, fo = S.forcing_perturbation(n, seed=78)
    lat = np.random.default_rng(6).uniform(45, 55, n); lon = np.full(n, -5.0)
    out = {"workload": "habitat-class grid, transient", "columns": n, "classes": K, "hours_per_step": a.hours}
    res = {}
    for mode in (("off", "on") if a.mode == "both" else (a.mode,)):
        ctx = Context(B.M, B.SM)
        ctx.set_columns(veg, soil, lat, lon); ctx.set_forcing(base["f"], fs, fo)
        if mode == "on":
            ctx.set_classes(cls)
        ctx.set_config(B.make_cfg(base["ts"], a.spin)); ctx.spinup(a.spin); ctx.set_config(B.make_cfg(base["ts"], 0))
        ms = []
        for k in range(a.reps + 1):
            r = ctx.run_transient(k * a.hours, (k + 1) * a.hours, want_series=False)
            ms.append(ctx.last_kernel_ms())
        res[mode] = r["stats"].copy()
        out[mode] = {"ms_per_step": float(np.mean(ms[1:])), "column_timesteps_per_s": n * a.hours / (np.mean(ms[1:]) * 1e-3),
                     "columns_left_streaming_form": int(np.count_nonzero(r["flags"] & 8))}
        ctx.close()
    if len(res) == 2:
        out["bit_identical"] = bool(np.array_equal(res["on"], res["off"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
