#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/coresidency_probe
echo "== bench"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_exp5.json 2> gpurun_out/bench_exp5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_exp5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','stage_seconds')})
PY
tail -3 gpurun_out/bench_exp5.err
