#!/bin/bash
# quick A/B of library builds on the same box (bench only): tools/abq.sh "libs..." [reps]
LIBS=$1; REPS=${2:-2}
ARGS="${ABQ_ARGS:---columns 303104 --hours 48 --steps 3 --warmup 2 --no-cpu-baseline --no-steady --spin 20}"
for rep in $(seq $REPS); do
  for L in $LIBS; do
    v=$(MC_B200_LIB=$(realpath $L) python bench.py $ARGS 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1fM  fp64 %.3f nonfinite %d"%(d["value"]/1e6, d["roofline"]["frac"], d["nonfinite_columns"]))')
    echo "$L  $v"
  done
done
