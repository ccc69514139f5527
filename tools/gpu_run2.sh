#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 20 --warmup 3
python bench.py --impl reference --steps 2 --warmup 1
python bench.py --workload cf3d_256_f64 --steps 20 --no-cpu-baseline
python bench.py --workload cf2d_128_f64 --steps 50 --no-cpu-baseline
} > gpurun_out/run2.log 2>&1
