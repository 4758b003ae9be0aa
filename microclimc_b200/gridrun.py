"""Grid runs of `runmodel` with the outputs streamed to disk.

The reference returns every table in memory (R/code.R:1267-1325): fine for one column, impossible for 10^6 columns x 8760 hours
(the default table alone would be 560 GB).  A grid run therefore keeps, per column, the statistics the kernel reduces on the
device (mean / min / max of `tout` and `tleaf`) and the final state, and, for a SAMPLE of columns, the hourly default series
and — for `reqhgt` "all" / "allclim" — the full-state rows from which the reference's tables are rebuilt.  The year is advanced in
chunks of `chunk_hours` with the state resident on the GPU(s); each chunk's sample rows are appended to memory-mapped `.npy`
files, so host memory stays at one chunk.

Directory layout (all little-endian float64 `.npy` unless noted):
  meta.json      m, sm, ncol, T, timestep, chunk_hours, reqhgt, sample indices, row names (layout.SERIES / state rows), obs_time
  stats.npy      [6][ncol]   mean, min, max of tout then of tleaf over the run
  flags.npy      [ncol] int32 (bit 0: non-finite output, bit 1: merged canopy layers occurred)
  state.npy      [NSTATE][ncol] final state (the `previn` of a continuation run)
  series.npy     [T][8][nsample]       default outputs (tout, tleaf, relhum, SWin, LWin, L, H, G) of the sampled columns
  full.npy       [T][NSTATE][nsample]  only for reqhgt "all" / "allclim"
`read_tables(out_dir, k)` rebuilds, for sampled column k, exactly what `api.runmodel` returns for that column.
"""
import json
import os

import numpy as np

from . import api as A
from . import layout as L

__all__ = ["runmodel_grid", "read_tables", "read_meta"]


def runmodel_grid(climdata, vegp, soilp, lat, long, out_dir, chunk_hours=168, sample=None, edgedist=100, reqhgt=None, sdepth=2, zu=2,
                  theta=0.3, merid=0, dst=0, n=0.6, steps=200, tsoil=None, metopen=True, windhgt=2, zlafact=1, surfwet=1, previn=None,
                  snow=None, fscale=None, foffset=None, n_gpus=1):
    """`runmodel` (R/code.R:1125-1328) over stacked columns with outputs written under `out_dir`; arguments as
    `api.runmodel_batch`.  Returns the meta dictionary (also written to meta.json)."""
    if chunk_hours < 1:
        raise ValueError("chunk_hours must be positive")
    ctx, info = A._prepare_transient(climdata, vegp, soilp, lat, long, edgedist, reqhgt, zu, theta, merid, dst, n, steps, metopen, windhgt,
                                     zlafact, surfwet, previn, snow, fscale, foffset, n_gpus)
    m, sm, ncol, T, char = info["m"], info["sm"], info["ncol"], info["T"], info["char"]
    NS = L.nstate(m, sm)
    os.makedirs(out_dir, exist_ok=True)
    samp = np.arange(min(ncol, 16), dtype=np.int64) if sample is None else np.asarray(sample, dtype=np.int64)
    ns = int(samp.size)
    # runmodel's deep-soil boundary is mean(climdata$temp) of the WHOLE run (code.R:1165) unless `tsoil` is given: the library
    # derives it from the forcing it holds, which is the whole series here, so chunking does not change it
    ts_arr = None
    if not A._isna(tsoil):
        ts_arr = np.broadcast_to(np.asarray(tsoil, float), (ncol,)).copy()
    open_mm = np.lib.format.open_memmap
    series = open_mm(os.path.join(out_dir, "series.npy"), mode="w+", dtype=np.float64, shape=(T, len(L.SERIES), ns)) if ns else None
    full = open_mm(os.path.join(out_dir, "full.npy"), mode="w+", dtype=np.float64, shape=(T, NS, ns)) if (ns and char) else None
    stats = None
    flags = np.zeros(ncol, dtype=np.int32)
    kernel_ms = 0.0
    for t0 in range(0, T, chunk_hours):
        t1 = min(T, t0 + chunk_hours)
        r = ctx.run_transient(t0, t1, tsoil=ts_arr, samp_idx=samp if ns else None, want_series=bool(ns), want_full=bool(ns and char))
        kernel_ms += ctx.last_kernel_ms()
        if ns:
            series[t0:t1] = r["series"]
            if full is not None:
                full[t0:t1] = r["full"]
        st = r["stats"]; nh = t1 - t0
        if stats is None:
            stats = st.copy(); stats[[0, 3]] *= nh
        else:
            stats[[0, 3]] += st[[0, 3]] * nh
            stats[[1, 4]] = np.minimum(stats[[1, 4]], st[[1, 4]]); stats[[2, 5]] = np.maximum(stats[[2, 5]], st[[2, 5]])
        flags |= r["flags"]
    stats[[0, 3]] /= T
    for a in (series, full):
        if a is not None:
            a.flush()
    np.save(os.path.join(out_dir, "stats.npy"), stats)
    np.save(os.path.join(out_dir, "flags.npy"), flags)
    np.save(os.path.join(out_dir, "state.npy"), ctx.get_state())
    meta = dict(m=m, sm=sm, ncol=ncol, T=T, timestep=info["timestep"], chunk_hours=int(chunk_hours),
                reqhgt=reqhgt if (char or A._isna(reqhgt)) else float(reqhgt), sample=[int(v) for v in samp], series_rows=L.SERIES,
                stats_rows=["tout_mean", "tout_min", "tout_max", "tleaf_mean", "tleaf_min", "tleaf_max"],
                state_rows={k: [sl.start, sl.stop] for k, sl in L.state_rows(m, sm).items()},
                obs_time=[str(v) for v in np.asarray(climdata["obs_time"])], reftemp=[float(v) for v in np.asarray(climdata["temp"], float)],
                kernel_ms=kernel_ms, n_gpus=int(n_gpus))
    with open(os.path.join(out_dir, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    ctx.close()
    return meta


def read_meta(out_dir):
    with open(os.path.join(out_dir, "meta.json")) as fh:
        return json.load(fh)


def read_tables(out_dir, k=0):
    """What `runmodel` returns (R/code.R:1267-1325) for the k-th SAMPLED column of a grid run: the default data.frame, or the
    dict of data.frames for reqhgt "all" / "allclim"."""
    import pandas as pd
    meta = read_meta(out_dir)
    obs = np.asarray(meta["obs_time"])
    if isinstance(meta["reqhgt"], str):
        full = np.load(os.path.join(out_dir, "full.npy"), mmap_mode="r")[:, :, k]
        return A._all_tables(np.asarray(full), obs, meta["m"], meta["sm"], meta["reqhgt"] == "all")
    s = np.asarray(np.load(os.path.join(out_dir, "series.npy"), mmap_mode="r")[:, :, k])
    return pd.DataFrame(dict(obs_time=obs, reftemp=np.asarray(meta["reftemp"]), tout=s[:, 0], tleaf=s[:, 1], relhum=s[:, 2], SWin=s[:, 3],
                             LWin=s[:, 4], H=s[:, 6], L=s[:, 5], G=s[:, 7]))
