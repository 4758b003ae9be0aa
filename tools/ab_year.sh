#!/bin/bash
# Same-box A/B of the full-year run (BASELINE config 3) between builds of the library: tools/ab_year.sh NAME1 NAME2 ...  (build/ab/NAME.so)
for rep in 1 2; do
  for n in "$@"; do
    MC_B200_LIB=$(realpath build/ab/$n.so) python tools/bench_year.py --gpus 1 --check 1 2> gpurun_out/ab_year_$n.err | tail -1 > gpurun_out/ab_year_$n.json
    python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_year_%s.json" % n))
    print(n, "device %.2f s  wall %.2f s  launches %d" % (d["device_s"], d["wall_s_incl_upload_and_readback"], d["kernel_launches"]))
except Exception as e:
    print(n, "failed", e, open("gpurun_out/ab_year_%s.err" % n).read()[-300:])
P
  done
done
