#!/bin/bash
# One GPU visit: parity tests, per-kernel times, bench line, ncu launch list and a full capture of the hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [skip-tests]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
(which gfortran mpif90 flang nvfortran ifort ifx 2>&1 || true) > $OUT/fortran_probe_$TAG.txt
if [ -z "$2" ]; then
  (timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
  tail -6 $OUT/pytest_$TAG.log
fi
for W in C2 C3 C5; do timeout 300 python tools/exp_kernels.py $W flush; done > $OUT/kernels_$TAG.txt 2>&1
cat $OUT/kernels_$TAG.txt
(timeout 900 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
(timeout 600 python bench.py --workload C5 --steps 50 --warmup 25 --no-cpu-baseline) > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --windows-per-gpu 0) > $OUT/ncu_launch_$TAG.log 2>&1
KREG='regex:k_water_rows|k_solute_rows|k_pair_energy|k_lrf_allpairs|k_lrf_accumulate|k_build_rows|k_q_atom|k_q_partner'
(timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -s 16 -c 16 -f -o $OUT/prof_${TAG}_c2 \
    python tools/exp_kernels.py C2) > $OUT/ncu_full_${TAG}_c2.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" -s 16 -c 16 -f -o $OUT/prof_${TAG}_c5 \
    python tools/exp_kernels.py C5) > $OUT/ncu_full_${TAG}_c5.log 2>&1
tail -3 $OUT/ncu_full_${TAG}_c5.log
ls -la $OUT
