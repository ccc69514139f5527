#!/bin/bash
# BASELINE.json configs[4]: 1024^3 float64 slab-decomposed over 8 GPUs (one product = 2 pipelined NCCL exchanges + 1 all-reduce)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --workload cf3d_1024_f64_slab --steps 10 --warmup 3 2>&1 | grep "^{\|Error\|error" | cut -c1-3000 > gpurun_out/slab8_r2.log
cut -c1-400 gpurun_out/slab8_r2.log
