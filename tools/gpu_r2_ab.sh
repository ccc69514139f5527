#!/bin/bash
# round-2 A/B: staged chain (forced on: NB200_CHAIN=1) vs generic bodies (NB200_FAST=0) on the BASELINE grids + GPU parity tests
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()"
for shp in 4096,4096 2048,2048 256,256,256; do
  echo "=== staged $shp"; NB200_CHAIN=1 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|Body"
  echo "=== generic $shp"; NB200_FAST=0 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|P[135C]"
done
NB200_CHAIN=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6
} > gpurun_out/r2_ab.log 2>&1
tail -c 6000 gpurun_out/r2_ab.log | cut -c1-400
