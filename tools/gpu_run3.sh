#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/gpu_probe.py --shape 4096,4096
python tools/gpu_probe.py --shape 256,256,256
python tools/gpu_probe.py --shape 128,128
} > gpurun_out/run3.log 2>&1
