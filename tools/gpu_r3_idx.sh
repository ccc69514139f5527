#!/bin/bash
# (historical: the variant this script measured was removed again, see profiles/r3_notes.md)
# bin-index rows in shared memory via cp.async (first pass / register-resident last pass): A/B + staged-chain parity tests
mkdir -p gpurun_out
{
for v in "NB200_X=1" "NB200_P5F_IDX=0" "NB200_P1F_IDX=0" "NB200_P5F_IDX=0 NB200_P1F_IDX=0"; do
  for shp in 4096,4096 2048,2048; do
  echo "=== $v $shp"; env $v timeout 100 python tools/gpu_probe.py --shape $shp --quick --steps 15 | grep -E "MVP|P5|P1|checksum"
  done
done
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "staged or full_size or baseline" 2>&1 | tail -2
} > gpurun_out/r3_idx.log 2>&1
cut -c1-150 gpurun_out/r3_idx.log
