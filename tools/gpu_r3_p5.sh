#!/bin/bash
mkdir -p gpurun_out
{
for v in "NB200_P5_SIDX=1" "NB200_P5_SIDX=0" "NB200_P5F=1" "NB200_P5F=1 NB200_P5F_SE=0"; do
  for shp in 4096,4096 2048,2048 256,256,256; do
  echo "=== $v $shp"; env $v timeout 300 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|P5|checksum"
  done
done
NB200_P5F=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/r3_p5.log 2>&1
tail -c 5000 gpurun_out/r3_p5.log | cut -c1-200
