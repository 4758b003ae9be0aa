"""Bit-for-bit repeatability of the streaming kernel at scale: the same transient run (spin-up + hours) is made several times — with one
library, or with several builds of it (e.g. the two proxy-fence variants, -DMC_FENCE=0/1) — in separate processes, and the final states,
per-column statistics and flags are compared as raw bytes (SHA-256).  Every mutable row of the kernel is lane-private and only travels
through memory (generic-proxy store, TMA bulk load a sweep later), so a copy that overtook its store would show up as a difference.
Usage: python tools/determinism_check.py [--columns N] [--hours H] [--spin S] [--reps R] lib1.so [lib2.so ...]"""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(a):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from microclimc_b200._lib import Context
    w = bench.workload(a.columns, a.hours, 7)
    ctx = Context(20, 10)
    ctx.set_columns(w["veg"], w["soil"], w["lat"], w["lon"]); ctx.set_forcing(w["f"], w["fs"], w["fo"]); ctx.set_config(bench.make_cfg(3600.0, a.spin))
    r = ctx.run_transient(want_series=False)
    st = ctx.get_state()
    h = hashlib.sha256(); h.update(np.ascontiguousarray(st).tobytes()); h.update(np.ascontiguousarray(r["stats"]).tobytes()); h.update(np.ascontiguousarray(r["flags"]).tobytes())
    print(json.dumps({"sha256": h.hexdigest(), "finite": bool(np.isfinite(st).all()), "left_streaming": int(np.count_nonzero(r["flags"] & 8))}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1000000); ap.add_argument("--hours", type=int, default=48); ap.add_argument("--spin", type=int, default=20)
    ap.add_argument("--reps", type=int, default=2); ap.add_argument("--child", action="store_true"); ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        child(a); sys.exit(0)
    res = []
    for rep in range(a.reps):
        for lib in (a.libs or [os.path.join(ROOT, "microclimc_b200", "libmicroclimc_b200.so")]):
            env = dict(os.environ, MC_B200_LIB=os.path.realpath(lib))
            out = subprocess.run([sys.executable, __file__, "--child", "--columns", str(a.columns), "--hours", str(a.hours), "--spin", str(a.spin)],
                                 env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
            print(os.path.basename(lib), rep, line); res.append(line)
    shas = {json.loads(l)["sha256"] for l in res}
    print("IDENTICAL" if len(shas) == 1 else "DIFFERENT: %d distinct results" % len(shas))
    sys.exit(0 if len(shas) == 1 else 1)
