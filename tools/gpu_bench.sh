#!/bin/bash
# GPU visit: parity tests, kernel table (+ A/B), the default bench line.  Usage: bash tools/gpu_bench.sh TAG
TAG=${1:-b}
OUT=gpurun_out
mkdir -p $OUT
(timeout 1200 python -m pytest tests -m gpu -x -q) > $OUT/pytest_$TAG.log 2>&1
tail -6 $OUT/pytest_$TAG.log
for w in C2 C5; do timeout 300 python tools/exp_kernels.py $w flush; done > $OUT/kernels_$TAG.txt 2>&1
cat $OUT/kernels_$TAG.txt
if [ -n "$QNB_AB" ]; then
  for w in C5; do env $QNB_AB timeout 300 python tools/exp_kernels.py $w flush; done > $OUT/kernels_${TAG}_ab.txt 2>&1
  echo "--- with $QNB_AB"; cat $OUT/kernels_${TAG}_ab.txt
fi
(timeout 900 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -5 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'value',d['value'],'repeats',d['repeats']['ms_per_window'])
print('roof',r['kernel'],r['frac'],'step',r['step'],'lrf',{k:r['lrf'][k] for k in ('ms','frac_fp32','frac_fp64')},'rows',r['rows']['ms'], 'build', r['list_build']['ms'])
print('batched',d.get('batched_windows'))
print('farm',d.get('fep_farm'))
print('compiled', d.get('e2e_compiled_host'), 'cpu', d.get('cpu_baseline',{}).get('ms_per_step'))
PY
