#!/bin/bash
mkdir -p gpurun_out
{
for v in "NB200_P3_SJL=1" "NB200_P3_SJL=0" "NB200_P5F=1" "NB200_L2PF=0"; do
  for shp in 4096,4096 2048,2048; do
  echo "=== $v $shp"; env $v timeout 300 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|Body|checksum"
  done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r3_p3.log 2>&1
tail -c 5000 gpurun_out/r3_p3.log | cut -c1-200
