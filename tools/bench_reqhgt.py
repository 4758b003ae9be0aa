"""Throughput of the two instantiations of the streaming kernel on the bench workload (10^6 columns x 168 h from a spun-up state):
reqhgt NA runs the regular-grid instantiation (REG), reqhgt = 1 m moves one node (code.R:719) and runs the general one, which stages
all per-layer rows.  Device time of the run launch (mc_last_kernel_ms), one JSON line per case.
Usage: python tools/bench_reqhgt.py [--columns N] [--hours H]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from microclimc_b200._lib import Context

ap = argparse.ArgumentParser()
ap.add_argument("--columns", type=int, default=1000000); ap.add_argument("--hours", type=int, default=168)
ap.add_argument("--timestep", type=float, default=3600.0, help="seconds per step (below 3600 both cases run the general instantiation; most steps take the wide solve)")
a = ap.parse_args()
w = bench.workload(a.columns, a.hours, 7)
for name, req in (("reqhgt NA (regular grid)", np.nan), ("reqhgt 1 m (one node moved)", 1.0)):
    cfg = bench.make_cfg(a.timestep, 50); cfg[2] = req
    ctx = Context(20, 10)
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"]); ctx.set_forcing(w["f"], w["fs"], w["fo"]); ctx.set_config(cfg)
    ctx.run_transient(want_series=False)                                   # spin-up + first pass
    cfg[11] = 0; ctx.set_config(cfg)
    ms = []
    for rep in range(3):
        r = ctx.run_transient(want_series=False)
        ms.append(ctx.last_kernel_ms())
    ms = min(ms)
    print(json.dumps({"case": name, "timestep": a.timestep, "merged_layer_columns": int(np.count_nonzero(r["flags"] & 2)), "columns": a.columns, "hours": a.hours, "kernel_ms": ms, "column_timesteps_per_s": a.columns * a.hours / ms * 1e3,
                      "left_streaming": int(np.count_nonzero(r["flags"] & 8)), "finite": bool(np.isfinite(r["stats"]).all())}))
    ctx.close()
