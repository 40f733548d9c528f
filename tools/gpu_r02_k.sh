#!/bin/bash
set -u
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 $OUT/${TAG}_pytest_gpu.log | cut -c1-400
