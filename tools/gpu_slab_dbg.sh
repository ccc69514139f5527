#!/bin/bash
mkdir -p gpurun_out
NB200_SLAB_CHUNKS=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_slab_check.py --shape 64,64,64 --steps 3 > gpurun_out/slab_dbg.log 2>&1
