#!/bin/bash
# ncu captures for profiles/: launch list of one bench run, full sets of the hot kernels on C2 and C5.  Usage: bash tools/gpu_prof.sh TAG
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
KREG='regex:k_water_rows|k_solute_rows|k_pair_energy|k_lrf_allpairs|k_lrf_accumulate|k_rows_scan|k_q_atom|k_q_partner|k_lrf_taylor|k_qq_static|k_pack_step'
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_${TAG}_c2.csv python tools/exp_kernels.py C2) > $OUT/ncu_launch_${TAG}_c2.log 2>&1
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_${TAG}_c5.csv python tools/exp_kernels.py C5) > $OUT/ncu_launch_${TAG}_c5.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -c 14 -f -o $OUT/prof_${TAG}_c2 python tools/exp_kernels.py C2) > $OUT/ncu_full_${TAG}_c2.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -c 8 -f -o $OUT/prof_${TAG}_c5 python tools/exp_kernels.py C5) > $OUT/ncu_full_${TAG}_c5.log 2>&1
tail -2 $OUT/ncu_full_${TAG}_c5.log
ls -la $OUT | tail -8
