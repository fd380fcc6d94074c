#!/bin/bash
# One GPU visit: parity tests, per-kernel times, bench lines (C2 default, C5, C3), ncu launch lists and full captures of the hot
# kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
(which gfortran mpif90 flang nvfortran ifort ifx 2>&1 || true) > $OUT/fortran_probe_$TAG.txt
(timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
tail -4 $OUT/pytest_$TAG.log
for W in C2 C3 C5; do timeout 300 python tools/exp_kernels.py $W flush; done > $OUT/kernels_$TAG.txt 2>&1
for W in C2 C3; do timeout 300 python tools/exp_batch.py $W 7; done >> $OUT/kernels_$TAG.txt 2>&1
cat $OUT/kernels_$TAG.txt
(timeout 900 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 300 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
(timeout 600 python bench.py --workload C5 --steps 50 --warmup 25 --no-cpu-baseline --batch-windows 0 --farm-steps 0 --repeats 9) > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
(timeout 600 python bench.py --workload C3 --steps 200 --warmup 25 --no-cpu-baseline --farm-steps 0 --repeats 9) > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err
bash tools/gpu_prof.sh $TAG > $OUT/prof_$TAG.log 2>&1
ls -la $OUT | tail -12
