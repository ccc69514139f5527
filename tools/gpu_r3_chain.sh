#!/bin/bash
# round-2 (session 2): fused cooperative chain kernels (nb_chain.cuh) vs the five-launch chain, + GPU tests
mkdir -p gpurun_out
{
python -c "import __graft_entry__ as g; g.smoke()"
for shp in 4096,4096 2048,2048 256,256,256 128,128; do
  echo "=== coop $shp"; timeout 300 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|Body|checksum"
  echo "=== legacy $shp"; NB200_COOP=0 timeout 300 python tools/gpu_probe.py --shape $shp --quick | grep -E "MVP|Seg|Scan|checksum"
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
} > gpurun_out/r3_chain.log 2>&1
tail -c 6000 gpurun_out/r3_chain.log | cut -c1-300
