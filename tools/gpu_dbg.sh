#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/dbg.log 2>&1
export PYTHONFAULTHANDLER=1
which nvcc; echo "PATH=$PATH"
python -u - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
print("a", flush=True)
import torch
import nifty_b200 as nb
print("imported", flush=True)
rt = nb.default_runtime()
print("runtime", rt.device, flush=True)
plan = nb.Plan((64, 64), 1/64)
print("plan K", plan.K, flush=True)
x = torch.randn(64, 64, dtype=torch.float64, device="cuda")
y = plan.hartley(x)
torch.cuda.synchronize()
f = torch.fft.fftn(x); ref = f.real + f.imag
print("hartley err", float((y-ref).abs().max()), flush=True)
PY
echo "rc=$?"
python -u __graft_entry__.py smoke
echo "rc=$?"
