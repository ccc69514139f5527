#!/bin/bash
mkdir -p gpurun_out
{
for s in 4096,4096 2048,2048 256,256,256 1024,1024 128,128,128 512,512; do python tools/gpu_probe.py --shape $s 2>/dev/null | sed -n 1,5p; done
python tools/gpu_probe.py --shape 4096,4096 --dtype f32 2>/dev/null | sed -n 1,6p
python tools/gpu_probe.py --shape 256,256,256 --dtype f32 2>/dev/null | sed -n 1,6p
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/exp2.log 2>&1
