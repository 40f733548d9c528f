#!/bin/bash
# Round-2 third GPU call: A/B of the conv kernels on ONE box - r01 kernels (copy under _r01/), new kernels unfused, new
# kernels fused; bit-identity test; bench.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== bit-identity + network tests"
timeout 900 python -m pytest tests/test_gpu_network.py -x -q > $OUT/${TAG}_pytest_network.log 2>&1
echo "network tests exit $?"; tail -3 $OUT/${TAG}_pytest_network.log
for rep in 1 2; do
echo "== r01 kernels (rep $rep)"
(cd _r01 && timeout 300 python tools/perf_probe.py 8) > $OUT/${TAG}_layers_r01_$rep.txt 2>&1; tail -3 $OUT/${TAG}_layers_r01_$rep.txt
echo "== new kernels, unfused (rep $rep)"
BOA_B200_UNFUSED=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_unfused_$rep.txt 2>&1; tail -3 $OUT/${TAG}_layers_unfused_$rep.txt
echo "== new kernels, fused (rep $rep)"
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused_$rep.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused_$rep.txt
done
paste <(awk '{print $1, $5}' $OUT/${TAG}_layers_r01_2.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_unfused_2.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_fused_2.txt | head -27)
echo "== bench fused / unfused"
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_fused.json 2> $OUT/${TAG}_bench_fused.err
BOA_B200_UNFUSED=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_unfused.json 2> $OUT/${TAG}_bench_unfused.err
python - <<'PY'
import json
for n in ('fused','unfused'):
    try:
        d=json.loads(open(f'gpurun_out/r02c_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['achieved'], {k:(v.get('tflops') or v.get('gbs')) for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(n, 'no bench line', e)
PY
tail -3 $OUT/${TAG}_bench_fused.err
