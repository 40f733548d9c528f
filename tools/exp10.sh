#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench 1 GPU"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_n1.json','gpurun_out/bench_n2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print({k:d[k] for k in ('n_gpus','value','ms_per_step','e2e','stage_seconds')})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_n1.err gpurun_out/bench_n2.err | grep -v OMP
