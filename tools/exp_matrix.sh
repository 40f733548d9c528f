#!/bin/bash
# A/B matrix of the forward schedule knobs (one process per configuration; the knobs are read once per process).
#   gpurun --timeout 900 -- 'bash tools/exp_matrix.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() {  # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python tools/perf_probe.py ${BATCH:-8} 2>&1 | grep -E "conv kernels|forward_accumulate|taps|tconv|Error|error" 
}
run "thin=1 lanes=2" BOA_B200_THIN=1
run "thin=0 lanes=2" BOA_B200_THIN=0
run "thin=1 lanes=1" BOA_B200_THIN=1 BOA_B200_LANES=1
run "thin=1 lanes=2 taps_stages=4" BOA_B200_THIN=1 BOA_B200_TAPS_STAGES=4
BATCH=16 run "thin=1 lanes=2 batch=16" BOA_B200_THIN=1
