#!/bin/bash
# Quick GPU iteration: parity tests, per-kernel times of C2/C3/C5, optional ncu capture.
# Usage (under gpurun, from the repo root): bash tools/gpu_quick.sh TAG [ncu-kernel-regex] [workload] [pytest-args]
# Extra environment for A/B comparisons: QNB_AB="VAR=val VAR2=val" re-runs the kernel table with those set.
TAG=${1:-q}; KREG=$2; W=${3:-C5}; PYT=${4:-tests}
OUT=gpurun_out
mkdir -p $OUT
(timeout 1200 python -m pytest $PYT -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
tail -4 $OUT/pytest_$TAG.log
for w in C2 C3 C5; do timeout 300 python tools/exp_kernels.py $w flush; done > $OUT/kernels_$TAG.txt 2>&1
cat $OUT/kernels_$TAG.txt
if [ -n "$QNB_AB" ]; then
  for w in C2 C5; do env $QNB_AB timeout 300 python tools/exp_kernels.py $w flush; done > $OUT/kernels_${TAG}_ab.txt 2>&1
  echo "--- with $QNB_AB"; cat $OUT/kernels_${TAG}_ab.txt
fi
if [ -n "$KREG" ]; then
  (timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREG" -c 12 -f -o $OUT/prof_${TAG} \
      python tools/exp_kernels.py $W) > $OUT/ncu_full_${TAG}.log 2>&1
  tail -3 $OUT/ncu_full_${TAG}.log
fi
