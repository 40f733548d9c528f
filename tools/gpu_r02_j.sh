#!/bin/bash
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_network.py -x -q > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -3 $OUT/${TAG}_pytest_network.log
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused.txt
BOA_B200_TAPS_NOTRIM=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_notrim.txt 2>&1; tail -3 $OUT/${TAG}_layers_notrim.txt
paste <(awk '{print $1, $2, $5}' $OUT/${TAG}_layers_fused.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_notrim.txt | head -27)
