#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tables_hartley" 2>&1 | tail -40
cat > /tmp/small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, nifty_b200 as nb, parity_checks as pc
rt = nb.default_runtime()
pc.check_hartley(rt, (2,))
pc.check_bilinear(rt, (2,), 1.0)
pc.check_hartley(rt, (4, 4))
print("ok")
PY
compute-sanitizer --tool memcheck python /tmp/small.py 2>&1 | tail -40
} > gpurun_out/dbg.log 2>&1
