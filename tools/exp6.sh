#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== overlap probe (convs <= 192 regs)"
BOA_B200_THIN=1 timeout 300 python tools/overlap_probe2.py 2>&1 | tail -1
BOA_B200_THIN=0 timeout 300 python tools/overlap_probe2.py 2>&1 | tail -1
for L in 1 2 3 4; do
echo "== thin=1 lanes=$L"
BOA_B200_THIN=1 BOA_B200_LANES=$L timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "conv kernels|forward_accumulate graph=False"
done
echo "== thin=0 lanes=2"
BOA_B200_THIN=0 BOA_B200_LANES=2 timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "conv kernels|forward_accumulate graph=False"
