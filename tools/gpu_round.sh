#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list and a full capture of the hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
(timeout 1200 python -m pytest tests -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
tail -15 $OUT/pytest_$TAG.log
(timeout 600 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 1500 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
(timeout 600 python bench.py --workload C5 --steps 50 --warmup 25 --no-cpu-baseline) > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
(timeout 600 python bench.py --workload C3 --steps 200 --warmup 25 --no-cpu-baseline) > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --windows-per-gpu 0) > $OUT/ncu_launch_$TAG.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_water_force|k_solute_force|k_lrf_accumulate|k_lrf_allpairs|k_build_rows|k_q_atom|k_q_partner' -s 14 -c 14 -f -o $OUT/prof_$TAG \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --windows-per-gpu 0) > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT
