#!/bin/bash
set -u
echo "== pytest -m gpu (merged tensor maps, thin passes)"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== overlap probe"
BOA_B200_THIN=1 timeout 300 python tools/overlap_probe2.py 2>&1 | tail -2
BOA_B200_THIN=0 timeout 300 python tools/overlap_probe2.py 2>&1 | tail -2
echo "== layers, merged tmap, thin=0"
BOA_B200_THIN=0 timeout 300 python tools/perf_probe.py 8 2>&1 | tail -32
echo "== layers, 5-D tmap, thin=0"
BOA_B200_THIN=0 BOA_B200_TMAP5D=1 timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "conv kernels|forward_accumulate"
