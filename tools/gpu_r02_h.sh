#!/bin/bash
# Evidence round: ncu launch list of the bench command, ncu --set full rows of the HBM passes and of the conv kernels of
# one forward, BASELINE configs 2 and 5 (incl. the triple z-split on hardware), the full bench line.
set -u
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== bench (full line)"
timeout 1500 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 1500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference exit $?"; tail -c 600 $OUT/${TAG}_bench_reference.json
echo "== config 2: 512x512x300, --models total"
timeout 900 python bench.py --shape 300 512 512 --models total --steps 3 --warmup 2 --quick > $OUT/${TAG}_cfg2.json 2> $OUT/${TAG}_cfg2.err
echo "cfg2 exit $?"; tail -c 900 $OUT/${TAG}_cfg2.json; tail -2 $OUT/${TAG}_cfg2.err
echo "== config 5 shape: 1024x1024x800, patch 192^3, --models total (triple z-split path)"
timeout 1500 python bench.py --shape 800 1024 1024 --patch 192 --batch 4 --models total --steps 1 --warmup 1 --quick > $OUT/${TAG}_cfg5.json 2> $OUT/${TAG}_cfg5.err
echo "cfg5 exit $?"; tail -c 900 $OUT/${TAG}_cfg5.json; tail -3 $OUT/${TAG}_cfg5.err
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-6000} --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --quick > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
gzip -f $OUT/${TAG}_launches.csv
cat $OUT/${TAG}_launches_summary.txt
echo "== ncu --set full: HBM passes"
timeout 900 ncu --set full --clock-control none --csv --page raw \
  -k regex:'finalize_argmax|tissue_kernel|slice_stats|label_hist|erode_axis|ct_normalize|head_mma' --launch-skip 16 --launch-count 16 \
  python tools/hbm_probe.py > $OUT/${TAG}_hbm_raw.csv 2> $OUT/${TAG}_ncu_hbm.log
python tools/ncu_pick.py $OUT/${TAG}_hbm_raw.csv > $OUT/${TAG}_hbm_passes.txt 2>&1
gzip -f $OUT/${TAG}_hbm_raw.csv
cat $OUT/${TAG}_hbm_passes.txt
echo "== ncu --set full: conv kernels of one forward"
timeout 900 ncu --set full --clock-control none --csv --page raw -k regex:'conv_taps|conv3_fold|extract_patches|stats_finalize' --launch-count 60 \
  python tools/perf_probe.py 8 > $OUT/${TAG}_forward_raw.csv 2> $OUT/${TAG}_ncu_forward.log
python tools/ncu_pick.py $OUT/${TAG}_forward_raw.csv > $OUT/${TAG}_forward_summary.txt 2>&1
gzip -f $OUT/${TAG}_forward_raw.csv
cat $OUT/${TAG}_forward_summary.txt
du -sh $OUT
