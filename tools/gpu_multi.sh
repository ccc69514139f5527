#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{"
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_vi.py -x -q -m gpu -s -k "config4" 2>&1 | tail -6
} > gpurun_out/multi_$N.log 2>&1
