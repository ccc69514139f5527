#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; SHAPE=${2:-256,256,256}; EXTRA=${3:-}
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dist_slab_check.py --shape $SHAPE $EXTRA 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$"
} > gpurun_out/slab_$N.log 2>&1
