#!/bin/bash
# Round-4 visit after the LRF shell kernel rework: parity tests, smoke, kernel table, C5 bench line, C5 launch list and one full
# capture of the shell kernel.  Usage (under gpurun): bash tools/gpu_r04.sh TAG
TAG=${1:-r04j}; OUT=gpurun_out; mkdir -p $OUT
(timeout 600 python -m pytest tests -m gpu -x -q; timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > $OUT/pytest_$TAG.log 2>&1
tail -3 $OUT/pytest_$TAG.log
for W in C5 C2; do timeout 200 python tools/exp_kernels.py $W flush; done > $OUT/kernels_$TAG.txt 2>&1
grep "^C" $OUT/kernels_$TAG.txt
(timeout 400 python bench.py --workload C5 --steps 50 --warmup 25 --no-cpu-baseline --batch-windows 0 --farm-steps 0 --repeats 9) > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
tail -c 400 $OUT/bench_${TAG}_c5.json; tail -2 $OUT/bench_${TAG}_c5.err
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_${TAG}_c5.csv python tools/exp_kernels.py C5) > $OUT/ncu_launch_${TAG}_c5.log 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_lrf_accumulate|k_pack_lrf_planes" -c 2 -f -o $OUT/prof_${TAG}_c5 python tools/exp_kernels.py C5) > $OUT/ncu_full_${TAG}_c5.log 2>&1
tail -1 $OUT/ncu_full_${TAG}_c5.log
