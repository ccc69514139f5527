#!/bin/bash
# driver-style launch of both bench arms at N=2 (sample-sharded KL-metric step with its all-reduce)
mkdir -p gpurun_out
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep "^{" | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/scale2_full.log 2>&1
grep "^{" gpurun_out/scale2_full.log | cut -c1-6000
grep -B30 "AssertionError\|Error" gpurun_out/scale2_full.log | grep -v "^\[W\|^W" | head -80
} > gpurun_out/scale2.log 2>&1
cut -c1-6000 gpurun_out/scale2.log
