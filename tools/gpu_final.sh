#!/bin/bash
# Round-end validation on one B200: smoke, GPU parity tests, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out
TAG=${1:-r2b}
{
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py
python bench.py --impl reference --steps 3 --warmup 1
python bench.py --workload cf3d_256_f64 --steps 50
python bench.py --workload cf2d_2048_f64 --steps 50
} > gpurun_out/final.log 2>&1
bash tools/gpu_prof2.sh 4096,4096 $TAG > /dev/null 2>&1
tail -c 3000 gpurun_out/final.log | cut -c1-400
