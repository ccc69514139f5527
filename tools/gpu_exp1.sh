#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_probe.py --shape 4096,4096 2>&1 | tail -8 > gpurun_out/exp1.log
python tools/gpu_probe.py --shape 256,256,256 2>&1 | tail -6 >> gpurun_out/exp1.log
