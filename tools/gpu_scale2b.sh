#!/bin/bash
# chunked all-reduce overlap of the sharded KL-metric step: chunk counts x NCCL stream priority (N = 2)
mkdir -p gpurun_out
: > gpurun_out/scale2b.log
port=29620
for pr in 0 1; do for ch in 1 2 4; do
  port=$((port+1))
  echo "=== TORCH_NCCL_HIGH_PRIORITY=$pr NB200_REDUCE_CHUNKS=$ch" >> gpurun_out/scale2b.log
  TORCH_NCCL_HIGH_PRIORITY=$pr NB200_REDUCE_CHUNKS=$ch NB200_BENCH_EXTRAS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 50 --warmup 3 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['collective']; print(d['value'], d['ms_per_step'], 'prod', c['product_ms_alone'], 'ar', c['allreduce_ms_alone'], 'exposed', c['exposed_ms'])" >> gpurun_out/scale2b.log
done; done
cat gpurun_out/scale2b.log
