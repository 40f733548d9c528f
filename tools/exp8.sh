#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== unrolled issue loop"
timeout 300 python tools/perf_probe.py 8 2>&1 | tail -31
