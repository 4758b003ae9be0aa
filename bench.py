#!/usr/bin/env python
"""Benchmark of the microclimc hot path (BASELINE.json metric: column-timesteps/s, 20 layers, FP64).

Workload (config.workload = "C3"): a 1000 x 1000 grid (10^6 columns per GPU, weak scaling), transient mode,
m = 20 canopy + sm = 10 soil nodes, hourly synthetic forcing (shared base series + per-column affine
perturbation), parameters from the seeded log-normal ensemble, state taken after a 200-step spin-up.
One bench "step" = one pass of the transient kernel over all columns for `--hours` consecutive hourly
time steps (default one week, 168); consecutive steps walk forward through the year with the state
resident on the GPU, so K steps are K*hours model hours.  The state (2.2 GB per 10^6 columns) is larger
than L2, so no flush is needed between timed iterations.

 * value  : column-timesteps/s with everything already resident in HBM (CUDA events around the kernel launches,
            measured inside the library on the launching stream; max over ranks via all_reduce)
 * e2e    : the same metric through the public C-ABI with HOST buffers: every step uploads that step's forcing
            tile and the per-column perturbation from pinned host memory and reads back the per-column statistics
            and flags (host<->device copies inside the timed region)
 * roofline: this path is FP64-pipe bound; see DESIGN.md "Measurement"
 * cpu_baseline / --impl reference: the repository's CPU oracle (C++ restatement of the R path, OpenMP over
            columns) on the box's host cores, on a bounded sample of the same workload. R itself is not installed
            and ~30 functions of the path live in the absent microctools package, so no interpreted-R timing exists.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, SM = 20, 10
METRIC = "column-timesteps/sec (20 layers, FP64)"
UNIT = "column-timesteps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=1_000_000, help="columns per GPU")
    ap.add_argument("--hours", type=int, default=168, help="hourly model steps per bench step")
    ap.add_argument("--spin", type=int, default=200)
    ap.add_argument("--cpu-sample-cols", type=int, default=0, help="0 = auto (sized for ~10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-steady", action="store_true", help="skip the extra steady-state (BASELINE config 2) measurement")
    return ap.parse_args()


def workload(ncol, T, seed):
    """C3 inputs for one rank: ensemble parameters, synthetic forcing of T hours, affine perturbation."""
    from microclimc_b200 import synth as S, layout as L
    sp = S.load_soilparams(); i = sp["Soil.type"].index("Loam")
    soil = {k: sp[k][i] for k in L.SOIL_SCALARS}
    soil["z"] = (2.0 / 10 ** 1.2) * np.arange(1, 11) ** 1.2           # soilinit("Loam")$z, code.R:622
    veg, so = S.ensemble(S.vegp_deciduous(), soil, ncol, M, SM, seed=seed)
    clim = S.synthetic_forcing(T, seed=7)
    f, ts = S.forcing_from_climdata(clim)
    fs, fo = S.forcing_perturbation(ncol, seed=seed + 1)
    lat = np.full(ncol, 50.0); lon = np.full(ncol, -5.0)
    return dict(veg=veg, soil=so, f=f, ts=ts, fs=fs, fo=fo, lat=lat, lon=lon)


def make_cfg(ts, spin):
    return np.array([ts, 100.0, np.nan, 2.0, 0.0, 0.0, 0.6, 1, 2.0, 1.0, 1.0, spin, 0], dtype=np.float64)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self._halt = gpu_index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set(); self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def oracle_rate(nthreads, ncol_s, hours, steps, warmup, seed=99):
    """CPU oracle throughput (column-timesteps/s) on a bounded sample, all host threads."""
    from oracle import pyoracle as O
    O.build()
    w = workload(ncol_s, hours * (steps + warmup), seed)
    cfg = make_cfg(w["ts"], 0)
    # initial state: oracle spin-up (not timed)
    cfg_spin = make_cfg(w["ts"], 50)
    st = O.run_transient(w["veg"], w["soil"], w["f"][:1], w["lat"], w["lon"], cfg_spin, M, SM, fscale=w["fs"], foffset=w["fo"],
                         samp_idx=np.zeros(0, dtype=np.int64), nthreads=nthreads)["state"]
    times = []
    for k in range(steps + warmup):
        f = np.ascontiguousarray(w["f"][k * hours:(k + 1) * hours])
        t0 = time.perf_counter()
        r = O.run_transient(w["veg"], w["soil"], f, w["lat"], w["lon"], cfg, M, SM, state=st, fscale=w["fs"], foffset=w["fo"],
                            samp_idx=np.zeros(0, dtype=np.int64), nthreads=nthreads)
        dt = time.perf_counter() - t0
        st = r["state"]
        if k >= warmup:
            times.append(dt)
    tot = float(np.sum(times))
    return ncol_s * hours * steps / tot, tot / steps * 1e3


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    ncol_s = args.cpu_sample_cols or max(256, 256 * cores)
    val, ms = oracle_rate(cores, ncol_s, args.hours, args.steps, args.warmup)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "C3", "columns_per_gpu": args.columns, "hours_per_step": args.hours, "m": M, "sm": SM,
                       "mode": "transient", "spin_steps": args.spin,
                       "sample": "bounded: %d columns of the C3 workload per step (the reference path is CPU-only)" % ncol_s},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d columns x %d hourly steps per bench step, OpenMP over columns; C++ restatement of the R "
                                       "path (R and microctools are not installable here)" % (ncol_s, args.hours)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from microclimc_b200._lib import Context
    from microclimc_b200 import layout as L

    ncol, hours, K, W = args.columns, args.hours, args.steps, args.warmup
    nsteps = K + W
    T = hours * nsteps
    w = workload(ncol, T, seed=1234 + 17 * rank)
    cfg = make_cfg(w["ts"], args.spin)

    # pinned host buffers for everything the e2e path uploads per step
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t
    fs_p, fo_p = pinned(w["fs"]), pinned(w["fo"])
    f_p = pinned(w["f"])
    tsoil_p = pinned(w["f"][:, 0].mean() * w["fs"][0] + w["fo"][0])

    ctx = Context(M, SM, n_gpus=1, gpu_ids=[local])
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"])
    ctx.set_forcing(w["f"], w["fs"], w["fo"])
    ctx.set_config(cfg)
    ctx.spinup(args.spin)                       # state after spin-up (untimed)
    cfg0 = cfg.copy(); cfg0[11] = 0             # no spin-up inside the timed steps
    ctx.set_config(cfg0)
    launches = 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only leg: inputs resident, device time from CUDA events on the launching stream ----
    st_after_spin = None
    dev_ms = []
    sampler = ClockSampler(local) if rank == 0 else None
    for k in range(nsteps):
        if k == W:
            barrier()
            if sampler:
                sampler.start()
        r = ctx.run_transient(k * hours, (k + 1) * hours, want_series=False, want_stats=True, want_flags=True)
        if k >= W:
            dev_ms.append(ctx.last_kernel_ms()); launches += ctx.last_kernel_launches()
    barrier()
    clocks = sampler.stop() if sampler else None
    nonfinite = int(np.count_nonzero(r["flags"] & 1))
    dev_total = float(np.sum(dev_ms))

    # ---- e2e leg: same K steps through the C-ABI with host buffers (H2D forcing tile + perturbation, D2H stats/flags) ----
    ctx.spinup(args.spin)
    ctx.set_config(cfg0)
    # result buffers of the e2e leg: pinned host memory owned by the caller and reused every step, as a grid driver would
    stats_p = torch.empty((6, ncol), dtype=torch.float64, pin_memory=True)
    flags_p = torch.empty((ncol,), dtype=torch.int32, pin_memory=True)
    e2e_t = []
    for k in range(nsteps):
        if k == W:
            barrier()
        t0 = time.perf_counter()
        tile = f_p.numpy()[k * hours:(k + 1) * hours]
        ctx.set_forcing(tile, fs_p.numpy(), fo_p.numpy())
        # the tile is a window of the year: the column's mean air temperature (deep-soil boundary) is passed explicitly
        rr = ctx.run_transient(0, hours, tsoil=tsoil_p.numpy(), want_series=False, want_stats=True, want_flags=True,
                               out_stats=stats_p.numpy(), out_flags=flags_p.numpy())
        torch.cuda.synchronize()
        if k >= W:
            e2e_t.append(time.perf_counter() - t0); launches += ctx.last_kernel_launches()
    barrier()
    e2e_total = float(np.sum(e2e_t))
    h2d = hours * L.NFORC * 8 + 2 * L.NPERT * ncol * 8 + ncol * 8
    d2h = 6 * ncol * 8 + ncol * 4

    peak_fp64 = ctx.fp64_peak_tflops() if rank == 0 else 0.0

    # max over ranks
    tt = torch.tensor([dev_total, e2e_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_total, e2e_total = float(tt[0]), float(tt[1])
    units = float(ncol) * hours * K * world
    value = units / (dev_total * 1e-3)
    e2e_val = units / e2e_total

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        cost = {}
        try:
            cost = json.load(open(os.path.join(ROOT, "profiles", "cost_per_colstep.json")))
        except Exception:
            pass
        per_gpu_rate = value / world
        # Roofline of the dominant kernel (k_transient_fast): this path is FP64-pipe bound (DESIGN.md "Measurement").
        #  achieved = algorithmic FP64 work per column-timestep (operator census of the shipped algorithm, 2 flop per FP64 issue slot,
        #             profiles/cost_per_colstep.json) x column-timesteps of the timed launches / their CUDA-event time;
        #  peak     = DFMA probe measured in this run (MEASURED_PEAKS.json carries HBM and bf16 only, no FP64 figure).
        # The HBM view is reported next to it: measured DRAM traffic per column-timestep (ncu) x rate against the measured copy peak.
        flop_per = cost.get("fp64_flop_per_colstep")
        bytes_per = cost.get("algorithmic_bytes_per_colstep", 152.0)
        traffic_per = cost.get("dram_bytes_per_colstep")
        ach = (flop_per * per_gpu_rate / 1e12) if flop_per else None
        roof = {"bound": "fp64", "unit": "TFLOP/s", "achieved": ach, "peak": peak_fp64,
                "peak_source": "of measured: mc_fp64_peak_tflops DFMA probe in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "frac": (ach / peak_fp64) if (ach and peak_fp64 > 0) else None,
                "traffic": (traffic_per * ncol * hours) if traffic_per else None,
                "flop_per_colstep": flop_per, "flop_source": cost.get("fp64_flop_source", "tools/opcount: shipped streaming algorithm, 2 x FP64 issue slots"),
                "hbm": {"achieved": (traffic_per or bytes_per) * per_gpu_rate / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (traffic_per or bytes_per) * per_gpu_rate / 1e9 / hbm_peak,
                        "peak_source": "of measured: MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "of fallback 6650",
                        "bytes_per_colstep_measured": traffic_per, "algorithmic_bytes_per_colstep": bytes_per}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "C3", "columns_per_gpu": ncol, "hours_per_step": hours, "m": M, "sm": SM, "mode": "transient",
                           "spin_steps": args.spin, "forcing": "shared synthetic hourly series + per-column affine perturbation",
                           "outputs": "per-column mean/min/max of tout,tleaf + flags",
                           "l2": "per-launch working set (state + geometry + work rows, 7.4 KB/column = 7.4 GB per 10^6 columns) >> 126 MB L2, no flush",
                           "kernel": "k_geom + k_transient_fast per launch (TMA-staged streaming form)"},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_total / K * 1e3},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "nonfinite_columns": nonfinite}
        if not args.no_cpu_baseline and world == 1:
            cores = host_threads()
            ncs = args.cpu_sample_cols or max(256, 256 * cores)
            cv, cms = oracle_rate(cores, ncs, hours, 2, 1)
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d columns x %d hourly steps x 2, OpenMP over columns; C++ restatement of the R path"
                                              % (ncs, hours)}
        if not args.no_steady and world == 1:
            # BASELINE config 2 beside the headline (not a bench line of its own): steady-state runmodelS, 10^4-column ensemble x the
            # bundled year of hourly forcing, k_steady timed with CUDA events inside the library (tools/bench_steady.py)
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("bench_steady", os.path.join(ROOT, "tools", "bench_steady.py"))
                bs = importlib.util.module_from_spec(spec); spec.loader.exec_module(bs)
                line["other_workloads"] = {"C2_steady": bs.run("C2", 10_000, 8760, 0, False, 1.0, 3, gpu=local, quiet=True)}
            except Exception as e:                                   # the headline must not depend on the side measurement
                line["other_workloads"] = {"C2_steady": {"error": str(e)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
