#!/bin/bash
# Round-2 first GPU call: parity tests (incl. the benchmarked configuration), smoke, bench line with the new records,
# ncu --set full rows of the HBM passes, compute-sanitizer on the small network test.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log; grep -A12 "bench configuration" $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench (ours)"
timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 6000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== ncu --set full of the HBM passes"
timeout 900 ncu --set full --clock-control none --csv --page raw \
  -k regex:'finalize_argmax|tissue_kernel|slice_stats|label_hist|erode_axis|ct_normalize|head_mma' --launch-skip 14 --launch-count 24 \
  python tools/hbm_probe.py > $OUT/${TAG}_hbm_raw.csv 2> $OUT/${TAG}_ncu_hbm.log
echo "ncu hbm exit $?"
python tools/ncu_pick.py $OUT/${TAG}_hbm_raw.csv > $OUT/${TAG}_hbm_passes.txt 2>&1
gzip -f $OUT/${TAG}_hbm_raw.csv
cat $OUT/${TAG}_hbm_passes.txt
echo "== compute-sanitizer (memcheck, racecheck) on the small all-kernels network test"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_network.py -x -q -k small_net > $OUT/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 $OUT/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_network.py -x -q -k small_net > $OUT/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?"; tail -4 $OUT/${TAG}_racecheck.log
du -sh $OUT
