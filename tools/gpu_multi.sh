#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded parity tests on >= 2 GPUs, sharded C5 bench in its variants.
# Usage: bash tools/gpu_multi.sh TAG N
TAG=${1:-r02m}; N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt; nvidia-smi topo -m >> $OUT/gpus_$TAG.txt 2>&1
(timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x) > $OUT/pytest_multi_$TAG.log 2>&1
tail -5 $OUT/pytest_multi_$TAG.log
run() {  # name, env...
  name=$1; shift
  (env "$@" NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload C5 --steps 50 --warmup 25 --no-cpu-baseline) > $OUT/bench_${TAG}_c5_$name.json 2> $OUT/bench_${TAG}_c5_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_${TAG}_c5_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'build', d['roofline']['list_build']['ms'], d['roofline']['kernels_ms_l2_flushed'])
except Exception as e:
    print('$name', 'failed', e); print(open('$OUT/bench_${TAG}_c5_$name.err').read()[-1500:])
PY
}
run pairs QNB_X=0
run rows QNB_SHARD_ROWS=1
run rows_graph QNB_SHARD_ROWS=1 QNB_SHARD_GRAPH=1
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}_c5.csv \
    python tools/exp_kernels.py C5) > $OUT/ncu_launch_${TAG}_c5.log 2>&1
ls $OUT | head -50
