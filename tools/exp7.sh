#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== resident weights on"
timeout 300 python tools/perf_probe.py 8 2>&1 | tail -31
echo "== resident weights off"
BOA_B200_BRES=0 timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "stages.0.0.convs|stages.4.convs|conv kernels|forward_accumulate"
