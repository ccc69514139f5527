#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ScanApply|SegSum|ScanAgg" -s 10 -c 5 \
    -f -o gpurun_out/prof_kchain python tools/gpu_probe.py --shape 4096,4096 --steps 2 > gpurun_out/probe_ncu_kchain.log 2>&1
