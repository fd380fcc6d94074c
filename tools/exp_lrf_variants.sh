#!/bin/bash
# A/B of builds of libqnb.so that differ in the LRF shell kernel (q6_b200/csrc/variants/*.so): one ncu pass with a short metric
# list per variant on the C5 list build.  Usage (under gpurun): bash tools/exp_lrf_variants.sh TAG
TAG=${1:-lrf}; OUT=gpurun_out; mkdir -p $OUT
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for s in long_scoreboard short_scoreboard wait no_instruction branch_resolving not_selected math_pipe_throttle barrier dispatch_stall lg_throttle mio_throttle; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
cp q6_b200/csrc/libqnb.so /tmp/libqnb_keep.so
for v in q6_b200/csrc/variants/*.so; do
  n=$(basename $v .so)
  cp $v q6_b200/csrc/libqnb.so
  (timeout 150 ncu --metrics $M --clock-control none -k regex:k_lrf_accumulate -c 1 --csv --log-file $OUT/lrfvar_${TAG}_$n.csv python tools/exp_kernels.py C5) > $OUT/lrfvar_${TAG}_$n.log 2>&1
  echo "== $n"; python - $OUT/lrfvar_${TAG}_$n.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; mi, vi = h.index("Metric Name"), h.index("Metric Value")
for r in rows[1:]:
    print("   %-78s %s" % (r[mi].replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), r[vi]))
PY
done
cp /tmp/libqnb_keep.so q6_b200/csrc/libqnb.so
