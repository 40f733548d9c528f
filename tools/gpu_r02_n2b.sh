#!/bin/bash
# 2 GPUs: the distributed GPU tests, then the driver's own N=2 invocation (no --quick)
set -u
TAG=${1:-r02n2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s > $OUT/${TAG}_pytest_dist.log 2>&1
echo "pytest exit $?"; tail -8 $OUT/${TAG}_pytest_dist.log | cut -c1-300
BOA_BENCH_WATCHDOG=260 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/${TAG}_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'], d['vs_single_gpu'], d['stage_seconds'])
except Exception as e:
    print('no line', e)
PY
tail -4 $OUT/${TAG}_bench.err | cut -c1-300
