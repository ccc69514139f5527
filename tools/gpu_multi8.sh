#!/bin/bash
mkdir -p gpurun_out
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --workload cf3d_1024_f64_slab --steps 10 --warmup 3 2>&1 | grep "^{\|Error\|error" | cut -c1-3000
NB200_SLAB_CHUNKS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --workload cf3d_1024_f64_slab --steps 10 --warmup 3 2>&1 | grep "^{\|Error\|error" | cut -c1-600
} > gpurun_out/multi8.log 2>&1
