#!/bin/bash
# Quick GPU iteration: parity tests, step/build timing of C2 and C5, per-kernel time + instruction counts.
# Usage (under gpurun, from the repo root): bash tools/gpu_quick.sh [tag]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
tail -4 $OUT/pytest_$TAG.log
for W in C2 C3 C5; do timeout 300 python tools/exp_step.py $W; done 2>&1 | tee $OUT/step_$TAG.txt
timeout 300 python tools/exp_e2e.py 2>&1 | tee -a $OUT/step_$TAG.txt
(timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
    -k 'regex:k_water_force|k_solute_force|k_q_atom|k_q_partner|k_lrf_taylor|k_qq_static|k_pack_coords' -s 14 -c 14 --csv \
    --log-file $OUT/inst_$TAG.csv python tools/exp_step.py C2) > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/inst_$TAG.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
seen={}
for r in rows[1:]:
    k=r[ki].split('(')[0]; seen.setdefault(k,{}).setdefault(r[mi],r[vi])
for k,v in seen.items(): print(k[:40], v)
PY
