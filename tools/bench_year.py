#!/usr/bin/env python
"""BASELINE config 3 in full: 10^6 columns (1000 x 1000 grid), 20 + 10 layers, transient mode, ONE YEAR of hourly synthetic
forcing (8760 steps) after a 200-step spin-up, on one B200 through the C-ABI — 8.76e9 column-timesteps.

The year is run the way a grid user would run it: one `run_transient` call per week of forcing (state resident on the GPU,
per-column statistics reduced on the device and carried there across the calls: mc_set_stats_accumulate), the hourly series of a sample of columns read back, and that sample compared
with the CPU oracle over the whole year (tolerances of the north_star: 1e-6 K, 1e-8 relative).  Prints one JSON line.

  python tools/bench_year.py [--columns 1000000] [--hours 8760] [--chunk 168] [--check 32] [--gpus 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                              # noqa: E402  (workload generator shared with bench.py)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1_000_000)
    ap.add_argument("--hours", type=int, default=8760)
    ap.add_argument("--chunk", type=int, default=168, help="hours per run_transient call (0: the whole year in ONE call / launch)")
    ap.add_argument("--spin", type=int, default=200)
    ap.add_argument("--check", type=int, default=32, help="columns compared with the CPU oracle over the whole year")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this box used IN PROCESS (static column split, one host thread per GPU)")
    a = ap.parse_args()
    from microclimc_b200._lib import Context, pinned_empty
    M, SM = B.M, B.SM
    if a.chunk <= 0:
        a.chunk = a.hours
    w = B.workload(a.columns, a.hours, seed=1234)
    cfg = B.make_cfg(w["ts"], a.spin)
    nchk = min(a.check, a.columns)
    samp = np.unique(np.linspace(0, a.columns - 1, nchk).astype(np.int64))
    ctx = Context(M, SM, a.gpus, list(range(a.gpus)))
    # caller-owned result buffers in page-locked memory (mc_host_alloc), reused by every call, as a grid driver would keep them
    st = pinned_empty((6, a.columns)); fl = pinned_empty((a.columns,), np.int32)
    t_wall0 = time.perf_counter()
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"])
    ctx.set_forcing(w["f"], w["fs"], w["fo"])
    ctx.set_config(cfg)
    ctx.set_stats_accumulate(True)        # statistics and flags carried on the device across the weekly calls, read back with the last one
    t_upload = time.perf_counter() - t_wall0
    dev_ms, launches, series = 0.0, 0, []
    spin_ms = None
    for t0 in range(0, a.hours, a.chunk):
        t1 = min(a.hours, t0 + a.chunk)
        last = t1 == a.hours
        r = ctx.run_transient(t0, t1, samp_idx=samp, want_series=True, want_stats=last, want_flags=last, out_stats=st, out_flags=fl)
        ms = ctx.last_kernel_ms()
        if t0 == 0:
            spin_ms = ms                                      # first call holds the spin-up launch as well
        dev_ms += ms; launches += ctx.last_kernel_launches()
        series.append(r["series"])
    wall = time.perf_counter() - t_wall0
    series = np.concatenate(series)
    stats, flags = st, fl
    colsteps = float(a.columns) * (a.hours + a.spin)
    out = {"workload": "C3 full year", "n_gpus_in_process": a.gpus, "columns": a.columns, "hours": a.hours, "spin_steps": a.spin, "chunk_hours": a.chunk,
           "device_s": dev_ms * 1e-3, "wall_s_incl_upload_and_readback": wall, "upload_s": t_upload,
           "first_call_ms_incl_spinup": spin_ms, "columns_left_streaming_form": int(np.count_nonzero(flags & 8)),
           "column_timesteps": colsteps, "column_timesteps_per_s_device": colsteps / (dev_ms * 1e-3),
           "column_timesteps_per_s_wall": colsteps / wall, "kernel_launches": launches,
           "nonfinite_columns": int(np.count_nonzero(flags & 1)),
           "tout_annual_mean_range": [float(stats[0].min()), float(stats[0].max())]}
    if nchk:
        from oracle import pyoracle as O
        O.build()
        t0 = time.perf_counter()
        want = O.run_transient(w["veg"][:, samp], w["soil"][:, samp], w["f"], w["lat"][samp], w["lon"][samp], cfg, M, SM,
                               fscale=w["fs"][:, samp], foffset=w["fo"][:, samp])
        cpu_s = time.perf_counter() - t0
        ws = want["series"]
        assert np.array_equal(np.isnan(ws), np.isnan(series))
        dT = float(np.nanmax(np.abs(series[:, :2] - ws[:, :2])))
        rel = float(np.nanmax(np.abs(series[:, 2:] - ws[:, 2:]) / np.maximum(np.abs(ws[:, 2:]), 1.0)))
        dS = float(np.max(np.abs(stats[:, samp] - want["stats"])))
        out["oracle_check"] = {"columns": int(samp.size), "hours": a.hours, "max_abs_dT_K": dT, "max_rel_other": rel,
                               "max_abs_stats_K": dS, "within_tolerance": bool(dT <= 1e-6 and rel <= 1e-8 and dS <= 1e-6),
                               "oracle_column_timesteps_per_s": samp.size * (a.hours + a.spin) / cpu_s,
                               "oracle_threads": B.host_threads()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
