#!/bin/bash
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -s > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -25 $OUT/${TAG}_pytest_network.log
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused.txt
awk '{print $1, $2, $5}' $OUT/${TAG}_layers_fused.txt | head -27
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
