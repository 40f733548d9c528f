#!/bin/bash
# 8 GPUs: bench (peer-memory path), then the NCCL path, then N=4
set -u
TAG=${1:-r02n8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
run() {  # name nproc port extra-env
  env $4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 5 --warmup 3 --quick > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  echo "$1 exit $?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_$1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['vs_single_gpu'], d['stage_seconds'])
except Exception as e:
    print('no line', e)
PY
  tail -4 $OUT/${TAG}_$1.err
}
run n8_peer 8 29521 BOA_X=1
run n8_nccl 8 29522 BOA_B200_EXCHANGE=nccl
run n4_peer 4 29523 BOA_X=1
