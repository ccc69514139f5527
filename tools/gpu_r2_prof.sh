#!/bin/bash
# round-2: full ncu capture of the staged passes (one launch each) on a given grid
mkdir -p gpurun_out
SHAPE=${1:-256,256,256}
TAG=${2:-r2s_256}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"P[135C]FBody" -s 8 -c 5 \
    -f -o gpurun_out/prof_$TAG python tools/gpu_probe.py --shape $SHAPE --steps 1 --quick > gpurun_out/probe_ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
