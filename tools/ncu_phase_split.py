"""Split an .ncu-rep's SASS listing of k_transient_fast into the phases delimited by the clock reads of MC_TICK
(CS2R ... SR_CLOCKLO) and report, per segment: static instructions, executed warp instructions, FP64 instructions,
local/global/shared/TMEM accesses, samples and the top stall reasons.
Usage: python tools/ncu_phase_split.py rep.ncu-rep [warp_steps]"""
import csv, subprocess, sys, re
from collections import Counter
rep = sys.argv[1]
wsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 151552 * 24 / 32
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
segs = []
cur = None
def new(name):
    return dict(name=name, n=0, ex=0, fp64=0, lds=0, sts=0, ldl=0, stl=0, ldg=0, stg=0, tm=0, bar=0, samples=0, stalls=Counter(), mufu=0, ctl=0)
infn = False
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        infn = "k_transient_fast" in r[1]; continue
    if r[0] == "Address":
        hdr = r; cur = new("seg0"); segs.append(cur); continue
    if not infn or hdr is None or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r))
    sass = d["Source"].strip()
    ex = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
    if "SR_CLOCK" in sass or "PMTRIG" in sass:
        cur = new("after %s @%s" % (sass.replace(";", "").strip()[:18], r[0][-6:])); segs.append(cur)
    cur["n"] += 1; cur["ex"] += ex; cur["samples"] += s
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")): cur["fp64"] += ex
    elif op.startswith("MUFU"): cur["mufu"] += ex
    elif op.startswith("LDS"): cur["lds"] += ex
    elif op.startswith("STS"): cur["sts"] += ex
    elif op.startswith("LDL"): cur["ldl"] += ex
    elif op.startswith("STL"): cur["stl"] += ex
    elif op.startswith("LDG") or op.startswith("LD."): cur["ldg"] += ex
    elif op.startswith("STG") or op.startswith("ST."): cur["stg"] += ex
    elif op.startswith(("LDTM", "STTM")): cur["tm"] += ex
    elif op.startswith(("BAR", "SYNCS", "WARPSYNC")): cur["bar"] += ex
    elif op.startswith(("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "JMP")): cur["ctl"] += ex
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = d.get(h)
            if v and v != "0": cur["stalls"][h[6:]] += int(v)
tot = sum(s["samples"] for s in segs) or 1
print("%-28s %6s %8s %7s %6s %5s %5s %5s %5s %5s %5s %5s %5s %6s  stalls" % ("segment", "static", "ex/wstep", "fp64", "mufu", "lds", "sts", "ldl", "stl", "ldg", "stg", "tmem", "bar", "smp%"))
for s in segs:
    if s["ex"] == 0 and s["samples"] == 0: continue
    f = lambda k: s[k] / wsteps
    print("%-28s %6d %8.0f %7.0f %6.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %6.2f  %s" % (
        s["name"], s["n"], f("ex"), f("fp64"), f("mufu"), f("lds"), f("sts"), f("ldl"), f("stl"), f("ldg"), f("stg"), f("tm"), f("bar"),
        100.0 * s["samples"] / tot, ", ".join("%s %.0f%%" % (k, 100.0 * v / max(s["samples"], 1)) for k, v in s["stalls"].most_common(4))))
