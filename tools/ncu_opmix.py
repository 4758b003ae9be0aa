"""Executed-instruction mix per phase of k_transient_fast from an .ncu-rep (SASS page): opcode histogram, weighted by the executed count,
for the segments delimited by the PMTRIG markers.  Usage: python tools/ncu_opmix.py rep.ncu-rep [warp_steps] [top]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
wsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 151552 * 24 / 32
top = int(sys.argv[3]) if len(sys.argv) > 3 else 22
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None; segs = []; cur = None; infn = False
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == "Kernel Name": infn = "k_transient_fast" in r[1]; continue
    if r[0] == "Address": hdr = r; cur = ["seg0", Counter()]; segs.append(cur); continue
    if not infn or hdr is None or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r)); sass = d["Source"].strip(); ex = int(d["Instructions Executed"] or 0)
    if "PMTRIG" in sass: cur = [sass.replace(";", "")[:16], Counter()]; segs.append(cur)
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.rstrip(";")
    base = op.split(".")[0]
    if base == "IMAD" and "MOV" in op: base = "IMAD.MOV"
    cur[1][base] += ex
for name, c in segs:
    tot = sum(c.values())
    if tot / wsteps < 100: continue
    print("%s: %.0f per warp-step" % (name, tot / wsteps))
    print("   " + "  ".join("%s %.0f" % (k, v / wsteps) for k, v in c.most_common(top)))
