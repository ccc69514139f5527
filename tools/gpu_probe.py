"""Developer tool: per-kernel timing of one metric-vector product (CUDA events) on a named grid."""
import argparse
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nifty_b200 as nb  # noqa: E402
import nifty_b200._capi as _capi  # noqa: E402

if os.environ.get("NB200_PROBE_LIB"):      # tuning variants built by tools/build_variant.sh
    _capi.DEFAULT_LIB = os.path.abspath(os.environ["NB200_PROBE_LIB"])


def build(shape, dist, dtype, lh="gauss"):
    cfm = nb.CorrelatedFieldMaker("cf", dtype=dtype)
    cfm.set_amplitude_total_offset(0.0, (1e-3, 1e-4))
    cfm.add_fluctuations(shape, dist, fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5),
                         asperity=(0.5, 0.05), prefix="ax1", non_parametric_kind="power")
    cf = cfm.finalize()
    sig = nb.SignalModel(cf, "exp")
    gen = torch.Generator(cf.rt.device).manual_seed(0)
    data = torch.randn(shape, dtype=dtype, device=cf.rt.device, generator=gen)
    lhm = nb.Gaussian(data, noise_cov_inv=100.0).amend(sig)
    return lhm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4096,4096")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="product timing only")
    a = ap.parse_args()
    shape = tuple(int(s) for s in a.shape.split(","))
    dtype = torch.float64 if a.dtype == "f64" else torch.float32
    t0 = time.time()
    lh = build(shape, 1.0 / shape[0], dtype)
    rt = lh.rt
    gen = torch.Generator(rt.device).manual_seed(1)
    L = lh.layout.size
    pos = 0.1 * torch.randn(L, dtype=dtype, device=rt.device, generator=gen)
    t = torch.randn(L, dtype=dtype, device=rt.device, generator=gen)
    lin, _ = lh.lin_at(pos)
    out = torch.empty_like(t)
    torch.cuda.synchronize()
    print(f"setup {time.time()-t0:.1f}s  N={np.prod(shape)} K={lh.signal.cf.plan.K} L={L}")
    t_w = time.time()
    while time.time() - t_w < 0.5:          # clocks ramp up on a fresh box
        lin.metric(t, add_identity=True, out=out)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(a.steps):
            lin.metric(t, add_identity=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    ms = best
    w = 8 if dtype == torch.float64 else 4
    d = len(shape) if len(shape) > 1 else 2
    bytes_mvp = w * np.prod(shape) * (2 * (2 * d - 1) + 4)
    print(f"MVP {ms:.4f} ms  -> {1e3/ms:.1f} MVP/s ; algorithmic {bytes_mvp/1e9:.3f} GB -> {bytes_mvp/ms/1e6:.1f} GB/s")
    rt.timing_begin()
    for _ in range(a.steps):
        lin.metric(t, add_identity=True, out=out)
    tm = rt.timing_end()
    tot = sum(v[1] for v in tm.values())
    for k, (c, msk) in sorted(tm.items(), key=lambda kv: -kv[1][1]):
        print(f"  {msk/c*1e3:9.1f} us x{c//a.steps}  {100*msk/tot:5.1f}%  {k}")
    print(f"  checksum: sum {out.double().sum().item():.12e}  |out| {out.double().norm().item():.12e}")
    if a.quick:
        return
    # one MGVI-style CG solve with per-kernel timing
    j = torch.randn(L, dtype=dtype, device=rt.device, generator=gen)
    x, res = lin.cg_solve(j, j.clone(), absdelta=1e-4 * L / 10, maxiter=100, raise_nonposdef=False)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    x, res = lin.cg_solve(j, j.clone(), absdelta=1e-4 * L / 10, maxiter=100, raise_nonposdef=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"CG solve: {res.nit} iterations, {1e3*(t2-t1):.2f} ms wall -> {1e3*(t2-t1)/max(res.nit,1):.3f} ms / iteration")
    rt.timing_begin()
    x, res = lin.cg_solve(j, j.clone(), absdelta=1e-4 * L / 10, maxiter=100, raise_nonposdef=False)
    tm = rt.timing_end()
    for k, (c, msk) in sorted(tm.items(), key=lambda kv: -kv[1][1]):
        if "Cg" in k or "Dot" in k:
            print(f"  {msk/c*1e3:9.1f} us x{c}  {k}")
    # update (linearise + gradient)
    e0.record()
    for _ in range(5):
        lin.update(pos, want_grad=True, add_prior=True)
    e1.record()
    torch.cuda.synchronize()
    try:
        print(f"linearise+grad {e0.elapsed_time(e1)/5:.4f} ms")
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
