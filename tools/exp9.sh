#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== pytest dist"; timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','stage_seconds','n_gpus')})
PY
tail -3 gpurun_out/bench_n2.err
