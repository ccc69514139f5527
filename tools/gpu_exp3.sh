#!/bin/bash
mkdir -p gpurun_out
B="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared nifty_b200/csrc/nb_api.cu -o nifty_b200/lib/libniftyb200.so"
P="python tools/gpu_probe.py --shape 4096,4096"
{
echo "== base"; $P | sed -n 2,5p
echo "== P5_U=4"; $B -DNB_P5_U=4; $P | sed -n 2,5p
echo "== P3_U=4"; $B -DNB_P3_U=4; $P | sed -n 2,5p
echo "== P1_HALF"; $B -DNB_P1_HALF; $P | sed -n 2,5p
echo "== P1_HALF + MINB3"; $B -DNB_P1_HALF -DNB_P1_MINB=3; $P | sed -n 2,5p
echo "== P5_U=1 P3_U=1"; $B -DNB_P5_U=1 -DNB_P3_U=1; $P | sed -n 2,5p
} > gpurun_out/exp3.log 2>&1
