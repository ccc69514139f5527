#!/bin/bash
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py
python bench.py --impl reference --steps 3 --warmup 1
} > gpurun_out/final.log 2>&1
