#!/bin/bash
# 8-GPU records of a round: tools/gpu_round8.sh TAG   (run under `gpurun --gpus 8`)
TAG=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_$TAG.log 2> gpurun_out/bench_n8_$TAG.err; echo "bench n8 rc=$?"; tail -c 1500 gpurun_out/bench_n8_$TAG.log
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload C5 --no-extra --no-steady > gpurun_out/bench_c5_n8_$TAG.log 2> gpurun_out/bench_c5_n8_$TAG.err; echo "bench c5 n8 rc=$?"; tail -c 600 gpurun_out/bench_c5_n8_$TAG.log
timeout 900 python tools/bench_year.py --gpus 8 > gpurun_out/year_n8_$TAG.log 2> gpurun_out/year_n8_$TAG.err; echo "year n8 rc=$?"; tail -c 1500 gpurun_out/year_n8_$TAG.log
timeout 900 python tools/bench_year.py --gpus 1 > gpurun_out/year_n1_$TAG.log 2> gpurun_out/year_n1_$TAG.err; echo "year n1 rc=$?"; tail -c 1500 gpurun_out/year_n1_$TAG.log
