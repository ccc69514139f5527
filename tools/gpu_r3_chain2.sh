#!/bin/bash
mkdir -p gpurun_out
{
for shp in 4096,4096 256,256,256 2048,2048 128,128; do
  echo "=== $shp"; NB200_COOP_DBG=1 timeout 300 python tools/gpu_probe.py --shape $shp --quick --steps 3 2>&1 | grep -E "cot-chain" | tail -1
  timeout 300 python tools/gpu_probe.py --shape $shp --quick 2>&1 | grep -E "MVP|Chain|checksum"
done
} > gpurun_out/r3_chain2.log 2>&1
tail -c 4000 gpurun_out/r3_chain2.log | cut -c1-500
