#!/bin/bash
# Developer tool: build a tuning variant of the library: tools/build_variant.sh <name> [-DFLAG ...]
# -> nifty_b200/lib/variants/lib_<name>.so (git-ignored; travels with gpurun).  tools/gpu_probe.py loads it
# when NB200_PROBE_LIB points at it.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=${NB200_VARIANT_SRC:-nifty_b200/csrc}
mkdir -p nifty_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -Iinclude $src/nb_api.cu -o nifty_b200/lib/variants/lib_$name.so
echo built lib_$name.so
