#!/bin/bash
# compute-sanitizer over four small parity cases: memcheck, racecheck (shared-memory hazards of the staged row / scan kernels),
# synccheck, initcheck.  Usage (under gpurun):  bash tools/gpu_sanitize.sh [tag]
TAG=${1:-r03}
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for TOOL in ${QNB_SAN_TOOLS:-memcheck racecheck synccheck initcheck}; do
  echo "=== $TOOL" 
  (timeout ${QNB_SAN_TIMEOUT:-120} $CS --tool $TOOL --print-limit 30 --error-exitcode 9 python tools/sanitize_case.py "${@:2}") > $OUT/sanitize_${TAG}_$TOOL.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^[a-z_0-9]+ \[" $OUT/sanitize_${TAG}_$TOOL.log | tail -8
done
