"""Where the end-to-end step of bench.py spends its time beyond the kernels: python tools/e2e_probe.py (one GPU)."""
import time, sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import bench as B
from microclimc_b200._lib import Context, pinned_copy, pinned_empty
from microclimc_b200 import layout as L
ncol, hours = 1_000_000, 168
w = B.workload(ncol, hours * 4, seed=1234)
ctx = Context(B.M, B.SM, 1, [0])
ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"])
f_p = pinned_copy(w["f"]); fs_p = pinned_copy(w["fs"]); fo_p = pinned_copy(w["fo"])
tsoil = pinned_copy(np.full(ncol, 9.0))
ctx.set_forcing(w["f"], w["fs"], w["fo"]); cfg = B.make_cfg(w["ts"], 200); ctx.set_config(cfg)
ctx.spinup(200)
cfg0 = cfg.copy(); cfg0[11] = 0; ctx.set_config(cfg0)
samp = np.unique(np.linspace(0, ncol - 1, 1024).astype(np.int64))
st = pinned_empty((6, ncol)); fl = pinned_empty((ncol,), np.int32)
import torch
for k in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.set_forcing(f_p[k*hours:(k+1)*hours], fs_p, fo_p)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.run_transient(0, hours, tsoil=tsoil, samp_idx=samp, want_series=True, want_stats=True, want_flags=True, out_stats=st, out_flags=fl)
    t2 = time.perf_counter()
    print("set_forcing %.2f ms  run %.2f ms  kernels %.2f ms  -> overhead %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, ctx.last_kernel_ms(), (t2-t0)*1e3 - ctx.last_kernel_ms()))
