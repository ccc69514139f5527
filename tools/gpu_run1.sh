#!/bin/bash
# first GPU pass: smoke, gpu tests, probe timings -> gpurun_out/
mkdir -p gpurun_out
{
which python; python --version
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python __graft_entry__.py smoke
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python tools/gpu_probe.py --shape 4096,4096
python tools/gpu_probe.py --shape 256,256,256
python tools/gpu_probe.py --shape 128,128
python tools/gpu_probe.py --shape 2048,2048
} > gpurun_out/run1.log 2>&1
tail -5 gpurun_out/run1.log
