#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; SHAPE=${2:-512,512,512}
run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 tools/dist_slab_check.py --shape $SHAPE --steps 10 $3 2>&1 | grep "^{\|slab parity\|Error\|error" | cut -c60-330; }
{
run "NB200_SLAB_CHUNKS=2" 29530 ""
run "NB200_SLAB_CHUNKS=1" 29531 --no-parity
run "NB200_SLAB_CHUNKS=4 TORCH_NCCL_HIGH_PRIORITY=1" 29532 --no-parity
run "NB200_SLAB_CHUNKS=4 TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MAX_NCHANNELS=8" 29533 --no-parity
run "NB200_SLAB_CHUNKS=2 TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MAX_NCHANNELS=16" 29534 --no-parity
} > gpurun_out/slab_tune.log 2>&1
