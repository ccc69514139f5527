#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
{
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 --no-cpu-baseline
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --impl reference
} > gpurun_out/multi_$N.log 2>&1
