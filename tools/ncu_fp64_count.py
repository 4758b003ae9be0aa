"""FP64-pipe warp instructions (DFMA + DMUL + DADD + DSETP + DMNMX) and all warp instructions one kernel launch executed, from the SASS page of
an .ncu-rep (ncu --set full --import-source on); `units` = work items per thread of that launch (e.g. hours per thread) to get a per-item
count.  Usage: python tools/ncu_fp64_count.py rep.ncu-rep kernel-substring threads items_per_launch"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
items = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0          # work items (column-steps, column-hours) of the launch
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; inf = False; fp = tot = ldl = stl = 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": inf = kern in r[1]; continue
    if r[0] == "Address": hdr = r; continue
    if not inf or hdr is None or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r)); ex = int(d["Instructions Executed"] or 0)
    src = d["Source"].strip(); op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    tot += ex
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): fp += ex
    if op == "LDL": ldl += ex
    if op == "STL": stl += ex
print({"kernel": kern, "warp_instructions": tot, "fp64_pipe_warp_instructions": fp, "per_item": {"instructions": tot * 32 / items, "fp64_pipe": fp * 32 / items,
       "local_loads": ldl * 32 / items, "local_stores": stl * 32 / items}})
