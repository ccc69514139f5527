"""GPU check + timing of the non-power-of-two path (run through gpurun; writes gpurun_out/r4_nonpow2.{log,json}).
Correctness first (small shapes against the oracle / nifty.cl fixture), then the paper's 3618^2 point
(nifty.re 10.8 ms per product on an A100)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "r4_nonpow2.log"), "w")
T0 = time.time()


def log(*a):
    msg = f"[{time.time() - T0:6.1f}s] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()


import numpy as np  # noqa: E402
import torch  # noqa: E402
import nifty_b200 as nb  # noqa: E402
import parity_checks as pc  # noqa: E402
import vi_checks as vc  # noqa: E402

GPU = torch.cuda.is_available()
if GPU:
    rt = nb.default_runtime()
    log("runtime", rt.device, torch.cuda.get_device_name(0))
else:       # dry run of this script on the host emulator (tests/emu), wall-clock timing
    from emu.build_emu import build
    from nifty_b200._capi import CApi
    rt = nb.Runtime(CApi(build()), "cpu")
    log("runtime: host emulator (dry run)")
out = {"checks": {}}


def run(name, fn, *a, **k):
    try:
        fn(*a, **k)
        out["checks"][name] = "ok"
        log("ok  ", name)
    except Exception as e:  # noqa: BLE001
        out["checks"][name] = f"FAILED: {type(e).__name__}: {e}"[:400]
        log("FAIL", name, repr(e)[:400])


for shp in [(3, 3), (6, 7), (3, 5, 7), (100, 37), (1, 5)]:
    run(f"hartley{shp}", pc.check_nonpow2_hartley, rt, shp)
run("hartley_f32_errors", lambda: (pc.check_nonpow2_hartley(rt, (12, 7), dtype=torch.float32), pc.check_nonpow2_errors(rt)))
run("golden_3x3", pc.check_nonpow2_golden, rt)
run("model_6x10_gauss", pc.check_nonpow2_model, rt)
run("vi_nonpow2_gauss", vc.check_host_composed_vi, rt, "nonpow2", "gauss")
run("vi_outer_poisson", vc.check_host_composed_vi, rt, "outer", "poisson")
run("outer_golden", pc.check_outer_golden, rt, "o_8x16_x_4")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r4_nonpow2.json"), "w"), indent=1)


def timed(fn, warm=2, reps=5):
    if not GPU:
        t = time.time()
        fn()
        return 1e3 * (time.time() - t)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]


n = int(os.environ.get("NB200_NONPOW2_N", "3618"))
shape = (n, n)
log("building", shape)
H = nb.BluesteinHartley(shape, runtime=rt)
log("padded plan", H.pad)
x = torch.randn(shape, dtype=torch.float64, device=rt.device)
ms_h = timed(lambda: H(x))
ref = torch.fft.fft2(x)
err = float(((H(x) - (ref.real + ref.imag)).abs().max() / ref.abs().max()))
ms_fft = timed(lambda: torch.fft.fft2(x))
log(f"hartley {n}^2: {ms_h:.3f} ms (cuFFT complex fft2 of the same grid, for scale: {ms_fft:.3f} ms), rel err vs cuFFT {err:.2e}")
out["hartley_ms"], out["cufft_fft2_ms"], out["hartley_rel_err_vs_cufft"], out["n"], out["pad"] = ms_h, ms_fft, err, n, list(H.pad)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r4_nonpow2.json"), "w"), indent=1)
del ref
cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
cfm.set_amplitude_total_offset(0.0, (1e-3, 1e-4))
cfm.add_fluctuations(shape, 1.0 / n, fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05),
                     prefix="ax1", non_parametric_kind="power")
cf = cfm.finalize()
log("model built, K =", cf._tabs["ell"].numel())
data = torch.randn(shape, dtype=torch.float64, device=rt.device)
lh = nb.Gaussian(data, noise_cov_inv=100.0).amend(nb.SignalModel(cf, "exp"))
pos = 0.1 * lh.layout.random(1, torch.float64, rt.device)
tan = lh.layout.random(2, torch.float64, rt.device)
lin = lh.new_lin()
t0 = time.time()
lin.update(pos)
if GPU:
    torch.cuda.synchronize()
log(f"linearise: {1e3 * (time.time() - t0):.1f} ms")
ms_m = timed(lambda: lin.metric(tan, add_identity=True), warm=2, reps=5)
log(f"metric product {n}^2 (J^T M J t + t, host-composed around 2 chirp-z transforms): {ms_m:.3f} ms  [nifty.re on an A100: 10.8 ms at 3618^2]")
out["metric_ms"] = ms_m
out["launches"] = rt.launch_count()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r4_nonpow2.json"), "w"), indent=1)
log("done")
