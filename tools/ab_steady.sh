#!/bin/bash
# A/B two library builds on the steady-state bench: tools/ab_steady.sh libA.so libB.so
for rep in 1 2; do for L in $1 $2; do echo "$L $(MC_B200_LIB=$(realpath $L) python tools/bench_steady.py --c4-cols 1000000 2>/dev/null | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["config"], round(d["column_hours_per_s"]/1e9,3), end="  ")')"; done; done
