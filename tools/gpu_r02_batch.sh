#!/bin/bash
# patches per launch group: 8 (round-1 default) against 16 / 25 (the deep layers fill more of the 148 SMs)
set -u
TAG=${1:-r02batch}
OUT=gpurun_out
mkdir -p $OUT
for B in $2; do
  BOA_BENCH_WATCHDOG=200 timeout 240 python bench.py --steps 3 --warmup 2 --quick --batch $B > $OUT/${TAG}_b$B.json 2> $OUT/${TAG}_b$B.err
  echo "batch $B exit $?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_b$B.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['label_checksum'], d['stage_seconds'])
except Exception as e:
    print('no line', e)
PY
  tail -3 $OUT/${TAG}_b$B.err | cut -c1-300
done
nvidia-smi --query-gpu=memory.used --format=csv
