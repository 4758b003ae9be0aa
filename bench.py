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


WORKLOADS = {"C3": 1_000_000, "C5": 2_000_000}     # columns per GPU: C3 = the 10^6-column grid; C5 = 16 x 10^6 columns / 8 GPUs
NSAMP = 1024                                        # sampled columns whose hourly series travel back every step (BASELINE.md §4, C3 output policy)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS), help="C3: 10^6 columns per GPU; C5: 2 x 10^6 per GPU "
                    "(the per-GPU share of the 16 x 10^6-column tile on 8 GPUs)")
    ap.add_argument("--columns", type=int, default=0, help="columns per GPU (0: the workload's)")
    ap.add_argument("--hours", type=int, default=168, help="hourly model steps per bench step")
    ap.add_argument("--spin", type=int, default=200)
    ap.add_argument("--cpu-sample-cols", type=int, default=0, help="0 = auto (sized for ~10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-steady", action="store_true", help="skip the steady-state side measurements (BASELINE configs 2 and 4)")
    ap.add_argument("--no-extra", action="store_true", help="skip the C5-shard side measurement (N = 1) and the in-process split leg (N > 1)")
    a = ap.parse_args()
    if not a.columns:
        a.columns = WORKLOADS[a.workload]
    return a


def workload_label(args, world):
    n = args.columns
    if args.workload == "C3" and n == WORKLOADS["C3"]:
        return "C3" if world == 1 else "C3-sized shard per GPU: %d x 10^6 columns, weak scaling (BASELINE config 3 is the N = 1 case)" % world
    if args.workload == "C5" and n == WORKLOADS["C5"]:
        return "C5" if world == 8 else "C5 shard: 2 x 10^6 columns per GPU on %d GPU(s) (BASELINE config 5 is the N = 8 case)" % world
    return "%s generator, %d columns per GPU x %d GPU(s)" % (args.workload, n, world)


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
            "config": {"workload": workload_label(args, args.gpus), "columns_per_gpu": args.columns, "hours_per_step": args.hours, "m": M, "sm": SM,
                       "mode": "transient", "spin_steps": args.spin,
                       "sample": "bounded: %d columns of the C3 workload per step (the reference path is CPU-only)" % ncol_s},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d columns x %d hourly steps per bench step, OpenMP over columns; C++ restatement of the R "
                                       "path (R and microctools are not installable here)" % (ncol_s, args.hours)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def e2e_steps(ctx, w, f_p, fs_p, fo_p, tsoil_p, samp, hours, nsteps, W, bufs, sync):
    """The end-to-end leg: every step uploads that step's forcing tile and the per-column perturbation from pinned host memory through the
    C-ABI and reads back per-column statistics, flags and the hourly series of the sampled columns.  Returns (wall seconds of the steps after
    the first W, device seconds of their kernels, kernel launches)."""
    stats_p, flags_p = bufs
    wall, dev, launches = [], 0.0, 0
    for k in range(nsteps):
        if k == W:
            sync()
        t0 = time.perf_counter()
        tile = f_p[k * hours:(k + 1) * hours]
        ctx.set_forcing(tile, fs_p, fo_p)
        # the tile is a window of the year: the column's mean air temperature (deep-soil boundary) is passed explicitly
        ctx.run_transient(0, hours, tsoil=tsoil_p, samp_idx=samp, want_series=True, want_stats=True, want_flags=True,
                          out_stats=stats_p, out_flags=flags_p)
        if k >= W:
            wall.append(time.perf_counter() - t0); dev += ctx.last_kernel_ms() * 1e-3; launches += ctx.last_kernel_launches()
    return float(np.sum(wall)), dev, launches


def e2e_bytes(ncol, hours, L):
    h2d = hours * L.NFORC * 8 + 2 * L.NPERT * ncol * 8 + ncol * 8
    d2h = 6 * ncol * 8 + ncol * 4 + hours * 8 * min(NSAMP, ncol) * 8
    return h2d, d2h


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")        # host-side rendezvous for the in-process leg (no kernel spinning on the GPUs)
    torch.cuda.set_device(local)
    from microclimc_b200._lib import Context, pinned_copy, pinned_empty
    from microclimc_b200 import layout as L

    ncol, hours, K, W = args.columns, args.hours, args.steps, args.warmup
    nsteps = K + W
    T = hours * nsteps
    w = workload(ncol, T, seed=1234 + 17 * rank)
    cfg = make_cfg(w["ts"], args.spin)
    samp = np.unique(np.linspace(0, ncol - 1, min(NSAMP, ncol)).astype(np.int64))

    # pinned host buffers (mc_host_alloc) for everything the e2e path moves per step
    fs_p, fo_p, f_p = pinned_copy(w["fs"]), pinned_copy(w["fo"]), pinned_copy(w["f"])
    tsoil_p = pinned_copy(w["f"][:, 0].mean() * w["fs"][0] + w["fo"][0])

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
    left_streaming = int(np.count_nonzero(r["flags"] & 8))
    dev_total = float(np.sum(dev_ms))

    # ---- e2e leg: same K steps through the C-ABI with host buffers ----
    ctx.spinup(args.spin)
    ctx.set_config(cfg0)
    bufs = (pinned_empty((6, ncol)), pinned_empty((ncol,), np.int32))      # caller-owned result buffers, reused every step
    e2e_total, _, l2 = e2e_steps(ctx, w, f_p, fs_p, fo_p, tsoil_p, samp, hours, nsteps, W, bufs, barrier)
    launches += l2
    barrier()
    h2d, d2h = e2e_bytes(ncol, hours, L)

    peak_fp64 = ctx.fp64_peak_tflops() if rank == 0 else 0.0

    # max over ranks
    tt = torch.tensor([dev_total, e2e_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_total, e2e_total = float(tt[0]), float(tt[1])
    units = float(ncol) * hours * K * world
    value = units / (dev_total * 1e-3)
    e2e_val = units / e2e_total

    # ---- the product's own multi-GPU path: ONE process, Context(n_gpus = N), BASELINE config 3 (10^6 columns in total) split statically over
    # the N GPUs, host buffers in, one gathered host array out.  Rank 0 runs it while the other ranks wait on a host-side barrier. ----
    inproc = None
    if world > 1 and not args.no_extra:
        ctx.close()
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                inproc = inprocess_leg(world, hours, K, W, args.spin)
            except Exception as e:                                   # the headline must not depend on the side measurement
                inproc = {"error": str(e)}
        dist.barrier(group=cpu_group)
        ctx = None

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
        #  peak     = DFMA probe measured in this run (MEASURED_PEAKS.json carries HBM and bf16 only, no FP64 figure).
        # The HBM view is reported next to it: measured DRAM traffic per column-timestep (ncu) x rate against the measured copy peak, with
        # its ratio to the algorithmic bytes (152 B with the state resident on chip, 3.3 KB with the state streamed).
        # `achieved` = ALGORITHMIC FP64 work per column-timestep x column-timesteps/s, the work unit FIXED at 30 332 flop = 2 x the 15 166
        # FP64-pipe instructions the round-1 kernel executed per column-timestep (profiles/r01s; SURVEY.md 8d estimates 3.7e4
        # flop-equivalents, band 3-6e4, so the unit is at the conservative end).  It does not move when instructions are removed from the
        # kernel, so rounds are comparable.  `pipe_utilisation` is the other view: the instructions the SHIPPED kernel executes (ncu,
        # tools/ncu_fp64_count.py) x 2 flop x rate / peak, i.e. how busy the FP64 pipe is — it falls whenever work is removed.
        unit_flop = 2.0 * cost.get("fixed_work_fp64_slots_per_colstep", 15166.0)
        exec_flop = cost.get("fp64_flop_per_colstep")
        bytes_per = cost.get("algorithmic_bytes_per_colstep", 152.0)
        traffic_per = cost.get("dram_bytes_per_colstep")
        ach = unit_flop * per_gpu_rate / 1e12
        roof = {"bound": "fp64", "unit": "TFLOP/s", "achieved": ach, "peak": peak_fp64,
                "peak_source": "of measured: mc_fp64_peak_tflops DFMA probe in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "frac": (ach / peak_fp64) if peak_fp64 > 0 else None,
                "flop_per_colstep": unit_flop,
                "flop_source": "fixed algorithmic unit: 2 flop x 15 166 FP64-pipe instructions per column-timestep (the round-1 kernel's executed count, "
                               "profiles/r01s; SURVEY.md 8d estimate 3.7e4 flop-equivalents); DESIGN.md section 6",
                "pipe_utilisation": {"frac": (exec_flop * per_gpu_rate / 1e12 / peak_fp64) if (exec_flop and peak_fp64 > 0) else None,
                                     "executed_flop_per_colstep": exec_flop, "source": cost.get("fp64_flop_source")},
                "traffic": (traffic_per * ncol * hours) if traffic_per else None,
                "traffic_per_colstep": traffic_per,
                "traffic_over_algorithmic": {"vs_152B_state_on_chip": (traffic_per / 152.0) if traffic_per else None,
                                             "vs_3.3KB_state_streamed": (traffic_per / 3300.0) if traffic_per else None},
                "hbm": {"achieved": (traffic_per or bytes_per) * per_gpu_rate / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (traffic_per or bytes_per) * per_gpu_rate / 1e9 / hbm_peak,
                        "peak_source": "of measured: MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "of fallback 6650",
                        "bytes_per_colstep_measured": traffic_per, "algorithmic_bytes_per_colstep": bytes_per,
                        "frac_on_algorithmic_bytes": bytes_per * per_gpu_rate / 1e9 / hbm_peak}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_label(args, world), "columns_per_gpu": ncol, "columns_total": ncol * world, "hours_per_step": hours,
                           "m": M, "sm": SM, "mode": "transient",
                           "spin_steps": args.spin, "forcing": "shared synthetic hourly series + per-column affine perturbation",
                           "outputs": "per-column mean/min/max of tout,tleaf + flags; e2e also the hourly series of %d sampled columns" % samp.size,
                           "l2": "per-launch working set (state + geometry + work rows, 7.4 KB/column = 7.4 GB per 10^6 columns) >> 126 MB L2, no flush",
                           "kernel": "k_geom + k_transient_fast per launch (TMA-staged streaming form); columns that leave it are redone by k_transient"},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_total / K * 1e3},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "nonfinite_columns": nonfinite,
                "columns_left_streaming_form": left_streaming}
        if inproc is not None:
            line["e2e_inprocess"] = inproc
        if not args.no_cpu_baseline and world == 1:
            cores = host_threads()
            ncs = args.cpu_sample_cols or max(256, 256 * cores)
            cv, cms = oracle_rate(cores, ncs, hours, 2, 1)
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d columns x %d hourly steps x 2, OpenMP over columns; C++ restatement of the R path"
                                              % (ncs, hours)}
        other = {}
        if world == 1 and not args.no_extra and args.workload == "C3":
            # BASELINE config 5's per-GPU share (2 x 10^6 columns; 16 x 10^6 over 8 GPUs) on this one GPU, same kernels, short run
            try:
                ctx.close(); ctx = None
                other["C5_shard"] = c5_shard(local, hours, args.spin)
            except Exception as e:
                other["C5_shard"] = {"error": str(e)}
        if world == 1 and not args.no_extra and args.workload == "C3":
            # a habitat-class grid in transient mode (4 parameter sets, columns ordered by class): with mc_set_classes the per-layer geometry
            # rows are shared per class and come from L2; bit-identical results (tools/bench_classes.py; DRAM bytes per column-step measured
            # with ncu: profiles/r02k_parameter_classes_ncu_dram.csv, 8.44 KB -> 3.69 KB)
            try:
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_classes.py"), "--reps", "2"], capture_output=True, text=True,
                                     timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(local))).stdout.strip().splitlines()
                other["habitat_class_grid_transient"] = json.loads(out[-1])
            except Exception as e:
                other["habitat_class_grid_transient"] = {"error": str(e)}
        if not args.no_steady and world == 1:
            # BASELINE configs 2 and 4 beside the headline (not bench lines of their own): steady-state runmodelS, k_steady timed with CUDA
            # events inside the library, with the same FP64 roofline view (tools/bench_steady.py)
            try:
                if ctx is not None:
                    ctx.close(); ctx = None
                import importlib.util
                spec = importlib.util.spec_from_file_location("bench_steady", os.path.join(ROOT, "tools", "bench_steady.py"))
                bs = importlib.util.module_from_spec(spec); spec.loader.exec_module(bs)
                other["C2_steady"] = bs.run("C2", 10_000, 8760, 0, False, 1.0, 3, gpu=local, quiet=True, peak_fp64=peak_fp64, cost=cost)
                other["C4_steady"] = bs.run("C4", 4_000_000, 8760, 0, True, 1.0, 2, habitats=4, gpu=local, quiet=True, peak_fp64=peak_fp64, cost=cost)
            except Exception as e:
                other["steady_error"] = str(e)
        if other:
            line["other_workloads"] = other
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def c5_shard(gpu, hours, spin):
    """2 x 10^6 columns (BASELINE config 5's share of one GPU) for two bench steps: device-timed rate and the e2e rate."""
    from microclimc_b200._lib import Context, pinned_copy, pinned_empty
    from microclimc_b200 import layout as L
    ncol = WORKLOADS["C5"]
    w = workload(ncol, 3 * hours, seed=4321)
    ctx = Context(M, SM, n_gpus=1, gpu_ids=[gpu])
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"]); ctx.set_forcing(w["f"], w["fs"], w["fo"])
    ctx.set_config(make_cfg(w["ts"], spin)); ctx.spinup(spin)
    ctx.set_config(make_cfg(w["ts"], 0))
    ms = []
    for k in range(3):
        r = ctx.run_transient(k * hours, (k + 1) * hours, want_series=False)
        ms.append(ctx.last_kernel_ms())
    samp = np.unique(np.linspace(0, ncol - 1, NSAMP).astype(np.int64))
    fs_p, fo_p, f_p = pinned_copy(w["fs"]), pinned_copy(w["fo"]), pinned_copy(w["f"])
    tsoil_p = pinned_copy(w["f"][:, 0].mean() * w["fs"][0] + w["fo"][0])
    bufs = (pinned_empty((6, ncol)), pinned_empty((ncol,), np.int32))
    wall, dev, _ = e2e_steps(ctx, w, f_p, fs_p, fo_p, tsoil_p, samp, hours, 3, 1, bufs, lambda: None)
    h2d, d2h = e2e_bytes(ncol, hours, L)
    out = {"workload": "C5 shard: 2 x 10^6 columns on one GPU (16 x 10^6 = 8 such shards)", "columns": ncol, "hours_per_step": hours, "steps": 2,
           "value": ncol * hours * 2 / (sum(ms[1:]) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ms[1:])),
           "e2e": {"value": ncol * hours * 2 / wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "nonfinite_columns": int(np.count_nonzero(r["flags"] & 1)), "device_memory_note": "19 GB of the 180 GB"}
    ctx.close()
    return out


def inprocess_leg(n_gpus, hours, K, W, spin):
    """BASELINE config 3 (10^6 columns in total) through ONE Context(n_gpus): static contiguous split, one worker thread per GPU, no
    collective, results gathered into one host array per output — strong scaling of the product's own multi-GPU path."""
    from microclimc_b200._lib import Context, pinned_copy, pinned_empty
    from microclimc_b200 import layout as L
    ncol = WORKLOADS["C3"]
    nsteps = K + W
    w = workload(ncol, hours * nsteps, seed=1234)
    ctx = Context(M, SM, n_gpus=n_gpus, gpu_ids=list(range(n_gpus)))
    t0 = time.perf_counter()
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"]); ctx.set_forcing(w["f"], w["fs"], w["fo"])
    t_upload = time.perf_counter() - t0
    ctx.set_config(make_cfg(w["ts"], spin)); ctx.spinup(spin)
    spin_ms = ctx.last_kernel_ms()
    ctx.set_config(make_cfg(w["ts"], 0))
    samp = np.unique(np.linspace(0, ncol - 1, NSAMP).astype(np.int64))
    fs_p, fo_p, f_p = pinned_copy(w["fs"]), pinned_copy(w["fo"]), pinned_copy(w["f"])
    tsoil_p = pinned_copy(w["f"][:, 0].mean() * w["fs"][0] + w["fo"][0])
    bufs = (pinned_empty((6, ncol)), pinned_empty((ncol,), np.int32))
    wall, dev, launches = e2e_steps(ctx, w, f_p, fs_p, fo_p, tsoil_p, samp, hours, nsteps, W, bufs, lambda: None)
    h2d, d2h = e2e_bytes(ncol, hours, L)
    units = float(ncol) * hours * K
    out = {"workload": "C3 (10^6 columns in total) split statically over %d GPUs inside one process (mc_create(n_gpus))" % n_gpus,
           "n_gpus": n_gpus, "steps": K, "value": units / wall, "unit": UNIT, "wall_s": wall, "device_s": dev,
           "device_value": units / dev, "ms_per_step_wall": wall / K * 1e3, "ms_per_step_device": dev / K * 1e3,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "gpu_launches": int(launches),
           "upload_columns_and_forcing_s": t_upload, "spinup_device_ms": spin_ms,
           "gathered_nonfinite_columns": int(np.count_nonzero(bufs[1] & 1)), "scaling": "strong"}
    ctx.close()
    return out


if __name__ == "__main__":
    main()
