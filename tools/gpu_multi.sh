#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded parity tests on >= 2 GPUs (NCCL and peer-memory all-reduce, both partitions),
# the driver's own bench command at N (its line carries sharded_c5 and fep_farm), sharded C5 with NCCL for comparison.
# Usage: bash tools/gpu_multi.sh TAG N
TAG=${1:-r02m}; N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt; nvidia-smi topo -m >> $OUT/gpus_$TAG.txt 2>&1
(timeout 420 python -m pytest tests/test_multi_gpu.py -m gpu -q -x) > $OUT/pytest_multi_$TAG.log 2>&1
tail -5 $OUT/pytest_multi_$TAG.log
run() {  # name, bench args, env...
  name=$1; bargs=$2; shift; shift
  (env "$@" NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N $bargs) > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'])
    for k in ('sharded_c5','fep_farm','batched_windows'):
        if k in d: print('   ', k, {a:b for a,b in d[k].items() if a not in ('note','kernels_ms_rank0_l2_flushed','workload')})
except Exception as e:
    print('$name', 'failed', e); print(open('$OUT/bench_${TAG}_$name.err').read()[-1500:])
PY
}
run default "--steps 20 --warmup 3" QNB_X=0
run c5_nccl "--workload C5 --steps 50 --warmup 25 --no-cpu-baseline --repeats 7" QNB_COMM=nccl
run c5_p2p "--workload C5 --steps 50 --warmup 25 --no-cpu-baseline --repeats 7" QNB_COMM=p2p
ls $OUT | head -50
