#!/bin/bash
mkdir -p gpurun_out
{
python tools/gpu_probe.py --shape 4096,4096
NB200_NO_PREFETCH=1 python tools/gpu_probe.py --shape 4096,4096 | head -6
python tools/gpu_probe.py --shape 256,256,256
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/run4.log 2>&1
