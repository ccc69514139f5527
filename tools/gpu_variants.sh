#!/bin/bash
# Developer tool: A/B the tuning variants of tools/build_variant.sh inside ONE gpurun job (boxes differ by 10-20 %).
# usage: gpurun -- bash tools/gpu_variants.sh "name[:ENV=V,...] ..." "shape ..."
mkdir -p gpurun_out
out=gpurun_out/variants.log
: > $out
for shape in $2; do
  for v in $1; do
    name=${v%%:*}; envs=""
    if [[ "$v" == *:* ]]; then envs=$(echo "${v#*:}" | tr ',' ' '); fi
    echo "== $name $envs shape=$shape" >> $out
    env $envs NB200_PROBE_LIB=nifty_b200/lib/variants/lib_$name.so timeout 300 python tools/gpu_probe.py --shape $shape --quick --steps 30 >> $out 2>&1
  done
done
grep -E "^==|^MVP|checksum" $out
