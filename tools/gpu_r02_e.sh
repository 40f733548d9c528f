#!/bin/bash
# Why are the tap-list kernels slow with the fused transform?  relay-only knob + ncu source view.
set -u
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
BOA_B200_XF_DEBUG=1 timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_relay.txt 2>&1; tail -3 $OUT/${TAG}_layers_relay.txt
timeout 300 python tools/perf_probe.py 8 > $OUT/${TAG}_layers_fused.txt 2>&1; tail -3 $OUT/${TAG}_layers_fused.txt
paste <(awk '{print $1, $5}' $OUT/${TAG}_layers_relay.txt | head -27) <(awk '{print $5}' $OUT/${TAG}_layers_fused.txt | head -27)
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:conv_taps --launch-skip 0 --launch-count 2 -f -o $OUT/${TAG}_taps python tools/perf_probe.py 8 > $OUT/${TAG}_ncu_taps.log 2>&1
echo "ncu taps exit $?"
ncu -i $OUT/${TAG}_taps.ncu-rep --page details > $OUT/${TAG}_taps_details.txt 2>&1
ncu -i $OUT/${TAG}_taps.ncu-rep --page source --csv --launch-count 1 > $OUT/${TAG}_taps_source.csv 2>&1
ncu -i $OUT/${TAG}_taps.ncu-rep --page raw --csv > $OUT/${TAG}_taps_raw.csv 2>&1
python tools/ncu_pick.py $OUT/${TAG}_taps_raw.csv
grep -E "Stall|stall|Warp Cycles Per|Issued Warp|No Eligible|Eligible Warps|Active Warps" $OUT/${TAG}_taps_details.txt | head -40
rm -f $OUT/${TAG}_taps.ncu-rep
gzip -f $OUT/${TAG}_taps_source.csv $OUT/${TAG}_taps_raw.csv
du -sh $OUT
