#!/bin/bash
# Short sanity pass after the last code change of a round: GPU tests, smoke, one bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['achieved'], d['roofline']['frac'])
PY
tail -2 gpurun_out/final_bench.err
