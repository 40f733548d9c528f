#!/bin/bash
# round-2 closing check on one GPU: default bench first (the number the driver takes), then smoke and the GPU suite
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
BOA_BENCH_WATCHDOG=420 timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cut -c1-1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 $OUT/${TAG}_pytest_gpu.log | cut -c1-400
