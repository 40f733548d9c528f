#!/bin/bash
set -u
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== mma head"
timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "conv kernels|forward_accumulate"
echo "== ncu head"
timeout 300 ncu --set full --clock-control none --csv --page raw -k regex:head_ --launch-count 2 python tools/perf_probe.py 8 > gpurun_out/head_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/head_raw.csv
