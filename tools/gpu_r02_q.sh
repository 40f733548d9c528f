#!/bin/bash
set -u
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_cli.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log | cut -c1-400
BOA_BENCH_WATCHDOG=200 timeout 240 python bench.py --steps 3 --warmup 3 --quick > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
b=json.loads(open("$OUT/${TAG}_bench.json").read())
print(b["value"], b["ms_per_step"], b["e2e"], b["stage_seconds"], b["label_checksum"])
PY
tail -5 $OUT/${TAG}_bench.err | cut -c1-300
