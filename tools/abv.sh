#!/bin/bash
# Build variants of the CUDA library with extra -D flags (CPU side) or A/B them on the GPU box.
#   tools/abv.sh build NAME "-DFLAG=1 ..."   -> build/ab/NAME.so (+ ptxas summary line)
#   tools/abv.sh run NAME1 NAME2 ...          -> alternating bench runs of build/ab/NAME*.so (on the GPU box), 2 rounds
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
if [ "$1" = build ]; then
  N=$2; shift 2
  (cd microclimc_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $@ -shared -o ../../build/ab/$N.so mc_lib.cu -lcudart 2> ../../build/ab/$N.ptxas)
  grep -A3 "k_transient_fastILi20ELi10ELb1" build/ab/$N.ptxas | sed -n 3,4p | tr '\n' ' '; echo
else
  shift
  ARGS=${ABARGS:---columns 303104 --hours 48 --steps 3 --warmup 2 --no-cpu-baseline --no-steady --no-extra --spin 20}
  for rep in 1 2; do
    for N in "$@"; do
      v=$(MC_B200_LIB=$(realpath build/ab/$N.so) python bench.py $ARGS 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f M  clk %s"%(d["value"]/1e6, d["clocks"]["sm_mhz"]))')
      echo "$N  $v"
    done
  done
fi
