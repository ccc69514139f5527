#!/bin/bash
# round 2 final state: smoke + GPU tests + both bench arms, then ncu (launch list + full capture of every kernel of a product)
mkdir -p gpurun_out
SHAPE=${1:-4096,4096}
TAG=${2:-r3_4096}
{
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py
python bench.py --impl reference --steps 3 --warmup 1
} > gpurun_out/final_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 30 -c 40 --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/gpu_probe.py --shape $SHAPE --steps 2 --quick > gpurun_out/probe_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"Body" -s 15 -c 5 \
    -f -o gpurun_out/prof_$TAG python tools/gpu_probe.py --shape $SHAPE --steps 2 --quick >> gpurun_out/probe_ncu_$TAG.log 2>&1
tail -c 3000 gpurun_out/final_$TAG.log | cut -c1-1500
