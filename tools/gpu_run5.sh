#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 20 --warmup 3
python bench.py --workload cf3d_256_f64 --steps 20 --no-cpu-baseline
python bench.py --workload cf2d_2048_f64 --steps 20 --no-cpu-baseline
python bench.py --workload cf2d_128_f64 --steps 50 --no-cpu-baseline
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits
} > gpurun_out/run5.log 2>&1
