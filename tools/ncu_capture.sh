#!/bin/bash
# `ncu --set full` captures of the hot kernels of one forward (batch of 8 patches), kept small enough to come back
# through gpurun_out/ (64 MiB): reports with source only for three fold-conv launches, raw CSV pages for the rest.
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r01'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none"
# fold conv: enc0.conv1 (32->32 @128^3, N=96), enc1.conv1 (64->64 @64^3, N=192), enc2.conv1
timeout 600 $NCU --import-source on -k regex:conv3_fold --launch-skip 1 --launch-count 3 -f -o $OUT/${TAG}_fold \
  python tools/perf_probe.py 8 > $OUT/${TAG}_ncu_fold.log 2>&1
echo "fold exit $?"
ncu -i $OUT/${TAG}_fold.ncu-rep --page raw --csv > $OUT/${TAG}_fold_raw.csv 2>&1
ncu -i $OUT/${TAG}_fold.ncu-rep --page details > $OUT/${TAG}_fold_details.txt 2>&1
ncu -i $OUT/${TAG}_fold.ncu-rep --page source --csv --kernel-name regex:conv3_fold --launch-count 1 > $OUT/${TAG}_fold_source.csv 2>&1
# everything else of the first forward: raw CSV only (no report kept)
timeout 600 $NCU --csv --page raw -k regex:'norm_lrelu|head_kernel|conv_taps|extract_patches|conv3_fold' --launch-count 60 \
  python tools/perf_probe.py 8 > $OUT/${TAG}_forward_raw.csv 2> $OUT/${TAG}_ncu_forward.log
echo "forward exit $?"
python tools/ncu_pick.py $OUT/${TAG}_forward_raw.csv > $OUT/${TAG}_forward_summary.txt 2>&1
python tools/ncu_pick.py $OUT/${TAG}_fold_raw.csv > $OUT/${TAG}_fold_summary.txt 2>&1
gzip -f $OUT/${TAG}_forward_raw.csv $OUT/${TAG}_fold_raw.csv $OUT/${TAG}_fold_source.csv
cat $OUT/${TAG}_forward_summary.txt
du -sh $OUT
