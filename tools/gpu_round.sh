#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line, the reference arm, per-layer timings, the ncu launch list of
# the bench command (bounded) and `ncu --set full` captures of the hot kernels.  Everything lands in gpurun_out/
# (kept under 64 MiB so that it comes back).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r01'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log

echo "== smoke"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log

echo "== bench (ours)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 3500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

echo "== per-layer timings"
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers.txt 2>&1
tail -32 $OUT/${TAG}_layers.txt

echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference exit $?"; tail -c 1500 $OUT/${TAG}_bench_reference.json

echo "== ncu launch list of the bench command (first ${NLAUNCH:-4000} launches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-4000} --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
gzip -f $OUT/${TAG}_launches.csv
cat $OUT/${TAG}_launches_summary.txt

bash tools/ncu_capture.sh $TAG
ls -la $OUT; du -sh $OUT
