#!/bin/bash
# Round-2 second GPU call: fused-normalisation schedule - parity tests, per-layer timings fused vs unfused, head RED vs
# RMW, bench line.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest network tests first (fail fast)"
timeout 900 python -m pytest tests/test_gpu_network.py -x -q -s > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -15 $OUT/${TAG}_pytest_network.log
echo "== per-layer timings (fused, default)"
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -32 $OUT/${TAG}_layers_fused.txt
echo "== per-layer timings (BOA_B200_UNFUSED=1)"
BOA_B200_UNFUSED=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_unfused.txt 2>&1; tail -4 $OUT/${TAG}_layers_unfused.txt
echo "== head: RMW in the SM instead of L2 reductions (BOA_B200_HEAD_RMW=1)"
BOA_B200_HEAD_RMW=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_head_rmw.txt 2>&1; tail -3 $OUT/${TAG}_layers_head_rmw.txt
echo "== pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 $OUT/${TAG}_pytest_gpu.log; grep -A12 "bench configuration" $OUT/${TAG}_pytest_gpu.log | head -14
echo "== bench (ours)"
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
    print(json.dumps(d['roofline']['kernels'], indent=0)); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['isolated_burst'])
    print(d['roofline_hbm']['kernels'].get('head_mma_kernel(in step)')); print(d['stage_seconds'])
except Exception as e:
    print('no bench line', e)
PY
tail -5 $OUT/${TAG}_bench.err
du -sh $OUT
