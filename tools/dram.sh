#!/bin/bash
# DRAM bytes of one 24-step launch of the streaming kernel for a given build: tools/dram.sh lib.so
MC_B200_LIB=$(realpath $1) ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_transient_fast -s 1 -c 1 python bench.py --columns 151552 --hours 24 --steps 2 --warmup 1 --no-cpu-baseline --no-steady --spin 20 2>&1 | grep -E "dram__|gpu__time|lts__" | awk -v L=$1 '{print L, $0}'
