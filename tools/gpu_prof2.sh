#!/bin/bash
# ncu: launch list of one probe run + full capture of the main passes and the amplitude-chain kernels
mkdir -p gpurun_out
SHAPE=${1:-4096,4096}
TAG=${2:-r2a}
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 40 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python tools/gpu_probe.py --shape $SHAPE --steps 2 --quick > gpurun_out/probe_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"P[135C]M?Body|SegSum|ScanA" -s 20 -c 10 \
    -f -o gpurun_out/prof_$TAG python tools/gpu_probe.py --shape $SHAPE --steps 2 --quick >> gpurun_out/probe_ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -5
