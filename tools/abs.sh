#!/bin/bash
# A/B builds of the library on the steady path: tools/abs.sh NAME1 NAME2 ...  (build/ab/NAME.so), C2 and a 48-hour C4, two rounds
cd "$(dirname "$0")/.."
for rep in 1 2; do for N in "$@"; do
  MC_B200_LIB=$(realpath build/ab/$N.so) python tools/bench_steady.py --c4-hours 48 2>/dev/null | python -c '
import sys, json
r = [json.loads(l) for l in sys.stdin if l.startswith("{")]
print("'$N'  " + "   ".join("%s %.2f ms = %.2f G/s" % (x["config"], x["kernel_ms"], x["column_hours_per_s"] / 1e9) for x in r))'
done; done
