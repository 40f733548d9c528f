#!/bin/bash
# 4 GPUs, the driver's own invocation (no --quick): pair post-processing + sharded measurements are active from 4 ranks;
# then throughput mode (replicas)
set -u
TAG=${1:-r02n4}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name nproc port extra-args
  BOA_BENCH_WATCHDOG=260 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 $4 > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  echo "$1 exit $?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_$1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'], d['vs_single_gpu'], d['stage_seconds'])
except Exception as e:
    print('no line', e)
PY
  tail -4 $OUT/${TAG}_$1.err | cut -c1-300
}
run n4_full 4 29531 "--steps 3 --warmup 3"
run n4_throughput 4 29532 "--steps 2 --warmup 2 --quick --mode throughput"
