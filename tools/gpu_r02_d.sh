#!/bin/bash
# fused (transform groups) vs unfused on one box + all GPU tests + bench
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_network.py -x -q > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -3 $OUT/${TAG}_pytest_network.log
BOA_B200_UNFUSED=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_unfused.txt 2>&1; tail -3 $OUT/${TAG}_layers_unfused.txt
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused.txt
paste <(awk '{print $1, $5}' $OUT/${TAG}_layers_unfused.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_fused.txt | head -27)
echo "== pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['achieved'], {k:(v.get('tflops') or v.get('gbs')) for k,v in d['roofline']['kernels'].items()})
    print(d['stage_seconds']); print({k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['roofline_hbm']['kernels'].items()})
except Exception as e:
    print('no bench line', e)
PY
tail -3 $OUT/${TAG}_bench.err
