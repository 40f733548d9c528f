#!/bin/bash
# BASELINE config 4 shape (512x512x400) on 2 GPUs: throughput mode (one volume per GPU, replicas only) vs latency mode
set -u
TAG=${1:-r02cfg4}
OUT=gpurun_out
mkdir -p $OUT
for mode in throughput latency; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 2 --quick --mode $mode --shape 400 512 512 > $OUT/${TAG}_$mode.json 2> $OUT/${TAG}_$mode.err
echo "$mode exit $?"; tail -c 1200 $OUT/${TAG}_$mode.json; tail -3 $OUT/${TAG}_$mode.err
done
