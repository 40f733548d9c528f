#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the reference arm, per-layer timings, the ncu launch list of the
# bench command and `ncu --set full` captures of the hot kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
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
echo "bench exit $?"; tail -c 3000 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

echo "== per-layer timings"
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers.txt 2>&1
tail -60 $OUT/${TAG}_layers.txt

echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference exit $?"; tail -c 1500 $OUT/${TAG}_bench_reference.json

echo "== ncu launch list of the bench command (one whole volume)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
gzip -f $OUT/${TAG}_launches.csv
cat $OUT/${TAG}_launches_summary.txt

echo "== ncu --set full: fold conv (all layers of one forward)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_fold -c 17 -f \
  -o $OUT/${TAG}_fold_full python tools/perf_probe.py 8 > $OUT/${TAG}_ncu_fold.log 2>&1
echo "ncu fold exit $?"
echo "== ncu --set full: the other kernels of one forward"
timeout 900 ncu --set full --clock-control none -k regex:'norm_lrelu|head_kernel|conv_taps|extract|stats_finalize|finalize_argmax' \
  -c 70 -f -o $OUT/${TAG}_rest_full python tools/perf_probe.py 8 > $OUT/${TAG}_ncu_rest.log 2>&1
echo "ncu rest exit $?"
for f in fold_full rest_full; do
  ncu -i $OUT/${TAG}_$f.ncu-rep --page raw --csv --metrics \
gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active \
    > $OUT/${TAG}_$f.csv 2>&1
done
ls -la $OUT
