#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_exp13.json 2> gpurun_out/bench_exp13.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_exp13.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','stage_seconds')})
PY
tail -3 gpurun_out/bench_exp13.err
