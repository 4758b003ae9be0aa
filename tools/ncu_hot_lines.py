"""Per-source-line hot spots of an .ncu-rep compiled with -lineinfo: samples, dominant stall reasons, executed instructions.
Usage: python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
files = []
cur = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; hdr = None; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and cur and r[0].strip().isdigit():
        d = {}
        for h, v in zip(hdr, r):
            d.setdefault(h, v)          # "Source" appears twice (CUDA line, SASS); keep the first
        files.append((cur, d, r))
stall_cols = [h for h in (hdr or []) if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
items = []
for f, d, r in files:
    try:
        s = int(d.get("# Samples", "0") or 0)
        ins = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    tot += s
    st = {}
    for h in stall_cols:
        try:
            st[h[6:]] = int(r[hdr.index(h)] or 0)
        except ValueError:
            pass
    items.append((s, ins, f.split("/")[-1], d["Line No"], d["Source"].strip()[:110], st))
items.sort(reverse=True)
print("total samples", tot)
for s, ins, f, ln, src, st in items[:top]:
    dom = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %9d %s:%s  %s   [%s]" % (100.0 * s / max(tot, 1), ins, f, ln, src, ", ".join("%s %d" % kv for kv in dom if kv[1])))
