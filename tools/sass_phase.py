"""Static per-phase census of k_transient_fast's SASS (phases delimited by the PMTRIG markers of MC_TICK): instructions, FP64-pipe
instructions, local / shared / global / tensor-memory accesses per phase.  CPU-side check before spending GPU time.
Usage: python tools/sass_phase.py [lib.so] [kernel-substring]"""
import re, subprocess, sys
from collections import Counter
lib = sys.argv[1] if len(sys.argv) > 1 else "microclimc_b200/libmicroclimc_b200.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "k_transient_fastILi20ELi10ELb1"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = [b for b in blocks if b.startswith("_Z") and kern in b.split("\n")[0]][0]
names = {0x80: "phase0", 0x1: "sweepA+cap", 0x2: "Bprep", 0x4: "sweepB (per layer)", 0x8: "solve", 0x10: "sweepC", 0x20: "end", 0x40: "between steps + epilogue"}   # the code AFTER each marker
segs = []; cur = ["prologue", Counter()]; segs.append(cur)
for line in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if not m: continue
    ins = m.group(2).strip()
    ops = ins.split()
    op = ops[1] if ops[0].startswith("@") else ops[0]
    if op == "PMTRIG":
        v = int(ops[-1], 16); cur = [names.get(v, hex(v)), Counter()]; segs.append(cur); continue
    c = cur[1]; c["n"] += 1
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): c["fp64"] += 1
    if base == "MUFU": c["mufu"] += 1
    if base in ("LDL",): c["ldl"] += 1
    if base in ("STL",): c["stl"] += 1
    if base in ("LDS",): c["lds"] += 1
    if base in ("STS",): c["sts"] += 1
    if base in ("LDG", "LD"): c["ldg"] += 1
    if base in ("STG", "ST"): c["stg"] += 1
    if base in ("LDTM", "STTM"): c["tmem"] += 1
    if base in ("BAR", "SYNCS", "WARPSYNC"): c["sync"] += 1
    if base in ("MOV", "IMAD") and ("MOV" in op): c["mov"] += 1
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET"): c["ctl"] += 1
print("%-34s %6s %6s %5s %4s %4s %4s %4s %4s %4s %5s %4s %4s" % ("segment", "instr", "fp64", "mufu", "ldl", "stl", "lds", "sts", "ldg", "stg", "tmem", "mov", "ctl"))
for name, c in segs:
    print("%-34s %6d %6d %5d %4d %4d %4d %4d %4d %4d %5d %4d %4d" % (name, c["n"], c["fp64"], c["mufu"], c["ldl"], c["stl"], c["lds"], c["sts"], c["ldg"], c["stg"], c["tmem"], c["mov"], c["ctl"]))
