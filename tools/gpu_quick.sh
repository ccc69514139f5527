#!/bin/bash
# smoke + GPU tests + default bench (no profile)
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py
} > gpurun_out/quick.log 2>&1
tail -c 2500 gpurun_out/quick.log | cut -c1-700
