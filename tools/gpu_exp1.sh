#!/bin/bash
mkdir -p gpurun_out
{
python tools/gpu_probe.py --shape 4096,4096 | head -12
python tools/gpu_probe.py --shape 256,256,256 | head -8
python tools/gpu_probe.py --shape 128,128 | head -9
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/exp1.log 2>&1
