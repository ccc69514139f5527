#!/bin/bash
mkdir -p gpurun_out
{
python bench.py --workload cf2d_128_f64 --steps 200 --no-cpu-baseline
python tools/gpu_probe.py --shape 128,128 --steps 200
} > gpurun_out/small.log 2>&1
cut -c1-1500 gpurun_out/small.log | tail -30
