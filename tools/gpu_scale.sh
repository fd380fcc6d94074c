#!/bin/bash
# The driver's bench command at N GPUs (+ the all-reduce alone): bash tools/gpu_scale.sh TAG N
TAG=${1:-s}; N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
(NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3) > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/exp_allreduce.py 2>&1 | grep "N=\|rror\|phases") > $OUT/allreduce_${TAG}_n$N.txt
cat $OUT/allreduce_${TAG}_n$N.txt
python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_${TAG}_n$N.json').read().strip().splitlines()[-1])
    print('N=$N ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))
    for k in ('sharded_c5','fep_farm','batched_windows'):
        if k in d: print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ('note','kernels_ms_rank0_l2_flushed','workload','spread_ms_per_window','spread_s','host_breakdown_last_call_us')})
except Exception as e:
    print('failed', e); print(open('$OUT/bench_${TAG}_n$N.err').read()[-1500:])
PY
