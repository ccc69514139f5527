#!/bin/bash
# driver-style launch of the default bench at N = 8 (sample-sharded KL-metric step + config-3 and 1024^3 slab extras)
mkdir -p gpurun_out
{
NB200_BENCH_EXTRAS_TIMEOUT=${NB200_BENCH_EXTRAS_TIMEOUT:-300} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/scale8_full.log 2>&1
echo "rc=$?"
grep "^{" gpurun_out/scale8_full.log | cut -c1-9000
grep -B5 "Error" gpurun_out/scale8_full.log | grep -v "^\[W\|^W" | head -40
} > gpurun_out/scale8.log 2>&1
cut -c1-9000 gpurun_out/scale8.log | tail -5
