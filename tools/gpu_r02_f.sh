#!/bin/bash
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_network.py -x -q > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -3 $OUT/${TAG}_pytest_network.log
BOA_B200_UNFUSED=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_unfused.txt 2>&1; tail -3 $OUT/${TAG}_layers_unfused.txt
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused.txt
BOA_B200_XF_DEBUG=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_relay.txt 2>&1; tail -3 $OUT/${TAG}_layers_relay.txt
paste <(awk '{print $1, $5}' $OUT/${TAG}_layers_unfused.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_fused.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_relay.txt | head -27)
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 2 --quick > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['stage_seconds'])
except Exception as e:
    print('no bench line', e)
PY
tail -3 $OUT/${TAG}_bench.err
