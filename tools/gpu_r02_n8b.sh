#!/bin/bash
# 8 GPUs, closing measurement of the round (pair post-processing, sharded measurements, maps delivered to rank 0)
set -u
TAG=${1:-r02n8b}
OUT=gpurun_out
mkdir -p $OUT
BOA_BENCH_WATCHDOG=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --quick > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/${TAG}_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'], d['vs_single_gpu'], d['stage_seconds'])
except Exception as e:
    print('no line', e)
PY
tail -4 $OUT/${TAG}_bench.err | cut -c1-300
