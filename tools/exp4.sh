#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== probe"
timeout 300 python tools/perf_probe.py 8 2>&1 | grep -E "conv kernels|forward_accumulate"
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench_exp4.json 2> gpurun_out/bench_exp4.err; tail -c 2500 gpurun_out/bench_exp4.json; tail -3 gpurun_out/bench_exp4.err
