#!/bin/bash
# One profiling pass of the bench kernel on the GPU box: tools/profile_round.sh TAG
#   plain run (must exit 0) -> launch list (gpu__time_duration per launch) -> one `--set full` capture of k_transient_fast.
# Outputs land in gpurun_out/; summarise them here with tools/ncu_summary.py and copy the summaries to profiles/.
TAG=$1
ARGS="--columns 151552 --hours 24 --steps 2 --warmup 1 --no-cpu-baseline --no-steady --no-extra --spin 20"
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_transient_fast -s 1 -c 1 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
