#!/bin/bash
# N = 2: slab-decomposed workload incl. the MGVI sample draw (uneven local sizes), then the default sharded bench line
mkdir -p gpurun_out
{
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --workload cf3d_256_f64_slab --steps 10 --warmup 3 2>&1 | grep "^{\|Error" | cut -c1-3000
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 2 --steps 30 --warmup 3 2>&1 | grep "^{\|Error" | cut -c1-6000
} > gpurun_out/scale2c.log 2>&1
cut -c1-6000 gpurun_out/scale2c.log
