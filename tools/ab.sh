#!/bin/bash
# A/B two builds of the CUDA library on the same box: tools/ab.sh libA.so libB.so [bench args...]; prints column-steps/s of each, alternating.
A=$1; B=$2; shift 2
ARGS=${@:---columns 303104 --hours 48 --steps 3 --warmup 2 --no-cpu-baseline --no-steady --no-extra --spin 20}
for rep in 1 2; do
  for L in $A $B; do
    v=$(MC_B200_LIB=$(realpath $L) python bench.py $ARGS 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1fM  fp64 %.3f"%(d["value"]/1e6, d["roofline"]["frac"]))')
    echo "$L  $v"
  done
done
