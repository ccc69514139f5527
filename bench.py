#!/usr/bin/env python3
"""bench.py -- metric-vector products per second of the MGVI inner loop (BASELINE.json metric).

A *step* is one application of M_p = F_p + 1 (``lh.metric(pos, t) + t``: Fisher metric of
Gaussian(data, N^-1) o exp o CorrelatedField plus identity) on the configured grid -- the quantity
the reference itself benchmarks (misc/re/paper/minimal_benchmark.py:93-114, paper.md:283-307) and
the operation every CG iteration of an MGVI sample draw runs (nifty/re/evi.py:83-85,139-144).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

* N=1 workload: BASELINE.json configs[1], 2-D correlated field 4096x4096 float64.
* N>1 (torchrun): the independent MGVI samples are sharded over the ranks (one CG solve per GPU, no
  data-path collective) -> weak scaling; value = products of all ranks / max time.
* ``--workload cf3d_1024_f64_slab`` (BASELINE.json configs[4], needs --gpus >= 2): ONE field slab-decomposed
  over the ranks, two pipelined NCCL exchanges + one all-reduce per product -> strong scaling.
* ``value``: device-resident inputs, CUDA events.  ``e2e``: the public Python API with HOST (pinned)
  buffers, H2D of the tangent and D2H of the result inside the timed region.
* ``roofline``: the dominant kernel's algorithmic bytes / its CUDA-event duration against the
  measured HBM peak of MEASURED_PEAKS.json; ``cpu_baseline``: the NumPy/scipy.fft oracle port on
  the host cores (bounded sample).  ``--impl reference`` times that CPU path as its own arm.
"""
import argparse
import json
import os
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (rank 0 only) is meant to use all host threads
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape, description)
    "cf2d_4096_f64": ((4096, 4096), "2D correlated field 4096x4096 float64 (BASELINE.json configs[1])"),
    "cf2d_2048_f64": ((2048, 2048), "2D correlated field 2048x2048 float64"),
    "cf3d_256_f64": ((256, 256, 256), "3D correlated field 256^3 float64 (BASELINE.json configs[2])"),
    "cf2d_128_f64": ((128, 128), "2D correlated field 128x128 float64 (BASELINE.json configs[0])"),
    # slab-decomposed over the ranks (needs --gpus >= 2 under torchrun): ONE field spread over all GPUs
    "cf3d_1024_f64_slab": ((1024, 1024, 1024), "3D correlated field 1024^3 float64, slab-decomposed (BASELINE.json configs[4])"),
    "cf3d_512_f64_slab": ((512, 512, 512), "3D correlated field 512^3 float64, slab-decomposed"),
    "cf3d_256_f64_slab": ((256, 256, 256), "3D correlated field 256^3 float64, slab-decomposed"),
}
CF_KW = dict(fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
OFFSET = (0.0, (1e-3, 1e-4))
NOISE_STD = 0.1


def algorithmic_bytes_mvp(shape, w=8):
    """w*N*[2(2d-1)+4]"""
    d = max(len(shape), 2)
    return w * int(np.prod(shape)) * (2 * (2 * d - 1) + 4)


def kernel_bytes(name, shape, w=8):
    """Algorithmic bytes of one launch of the named pass."""
    N = int(np.prod(shape))
    if "P1Body" in name or "P1MBody" in name or "P1FBody" in name:
        return 3 * w * N      # read t, xi; write half spectrum
    if "P3Body" in name or "P3FBody" in name:
        return 3 * w * N      # read half spectrum, Jacobian weight; write half spectrum
    if "P5Body" in name or "P5FBody" in name:
        return 4 * w * N      # read half spectrum, t, xi; write out
    if "PCBody" in name or "PCFBody" in name:
        return 2 * w * N
    return None


def short_kernel_name(name):
    """Body name of a mangled kernel name as recorded by the library's launch timer."""
    for k in ("P1FBody", "P3FBody", "PCFBody", "P5FBody", "P1MBody", "P1Body", "P3Body", "P5Body", "PCBody", "CotChain", "TanChain", "SegSum", "ScanApply", "ScanAgg",
              "CgStep", "CgDir"):
        if k in name:
            return k
    return name[:40]


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the GPU is under load (the recipe's
    clocks line of B200_PROFILING.md; NVML instead of a piped `nvidia-smi -lms`, whose stdout is block
    buffered).  Runs from the first warm-up step to the end of the end-to-end region."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0, period=0.01):
        self.index, self.period, self.rows, self.stop_flag, self.thread, self.err = index, period, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = torch.cuda.get_device_properties(self.index).uuid
            except Exception:
                pass
            h = None
            if uuid is not None:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(mhz), int(mask)))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for _, m in self.rows for bit, name in self.REASONS.items() if m & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "window": "warm-up + timed + instrumented + end-to-end regions"}


# -------------------------------------------------------------------------------------------------
# CPU arm: the NumPy/scipy.fft oracle port of the reference's path (nifty.re itself needs JAX, which
# is not installed here or on the GPU box; /root/reference does not exist on the GPU box either).
# -------------------------------------------------------------------------------------------------
def cpu_setup(shape, seed=42):
    import oracle
    cores = os.cpu_count() or 1
    oracle.correlated_field.set_nthreads(cores)
    cf = oracle.CorrelatedFieldOracle("cf")
    cf.set_amplitude_total_offset(*OFFSET)
    cf.add_fluctuations(shape, 1.0 / shape[0], prefix="ax1", non_parametric_kind="power", **CF_KW)
    cf.finalize()
    sig = oracle.SignalOracle(cf, "exp")
    lay = oracle.Layout(sig.domain)
    rng = np.random.default_rng(seed)
    truth = lay.random(rng)
    data = sig(truth) + NOISE_STD * rng.standard_normal(shape)
    lh = oracle.GaussianOracle(data, NOISE_STD**-2, sig)
    pos = {k: 0.1 * v for k, v in lay.random(np.random.default_rng(seed + 2)).items()}
    tan = lay.random(np.random.default_rng(seed + 3))
    lh.bench_data = data
    return lh, lay, pos, tan, cores


def cpu_mvp_once(lh, lay, pos, tan):
    out = lh.metric(pos, tan)          # J^T N^-1 J t, re-linearising like nifty.re (likelihood.py:613-621)
    return {k: out[k] + tan[k] for k in out}


def _nifty_re_leg():
    """baseline/run_nifty_re_cpu.py when `nifty.re` on JAX-CPU is importable on this box (it is not in this image), else None."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("run_nifty_re_cpu", os.path.join(ROOT, "baseline", "run_nifty_re_cpu.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod if mod.available() else None
    except Exception:
        return None


def run_reference(args, shape, wname, rank, world):
    if rank != 0:
        return
    re_leg = _nifty_re_leg()
    if re_leg is not None:      # the reference itself (kind = "reference"); protocol of misc/re/paper/minimal_benchmark.py
        cores = os.cpu_count() or 1
        dt = re_leg.time_metric_products(shape, cores)
        val = 1.0 / dt
        line = {"impl": "reference", "metric": "metric_vector_products_per_sec", "value": val, "unit": "MVP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wname, "shape": list(shape), "what": WORKLOADS[wname][1], "cpu_path": "nifty.re on JAX-CPU"},
                "cpu_baseline": {"value": val, "unit": "MVP/s", "cores": cores, "kind": "reference",
                                 "sample": f"median of 7 timeit repeats of lh.metric(pos, t) + t at {wname}"},
                "e2e": {"value": val, "unit": "MVP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    t0 = time.time()
    lh, lay, pos, tan, cores = cpu_setup(shape)
    setup_s = time.time() - t0
    for _ in range(max(args.warmup, 1) if np.prod(shape) <= 2**22 else 1):
        cpu_mvp_once(lh, lay, pos, tan)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_mvp_once(lh, lay, pos, tan)
    dt = (time.perf_counter() - t0) / args.steps
    val = 1.0 / dt
    sample = f"{args.steps} full {wname} products on {cores} host threads (scipy.fft workers), setup {setup_s:.1f}s excluded"
    line = {"impl": "reference", "metric": "metric_vector_products_per_sec", "value": val, "unit": "MVP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "shape": list(shape), "what": WORKLOADS[wname][1],
                       "cpu_path": "NumPy/scipy.fft oracle port of nifty.re (JAX not installable offline)"},
            "cpu_baseline": {"value": val, "unit": "MVP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "MVP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def build_b200_lh(nb, torch, shape, dev, dtype, data=None, seed=42):
    cfm = nb.CorrelatedFieldMaker("cf", dtype=dtype)
    cfm.set_amplitude_total_offset(*OFFSET)
    cfm.add_fluctuations(shape, 1.0 / shape[0], prefix="ax1", non_parametric_kind="power", **CF_KW)
    cf = cfm.finalize()
    sig = nb.SignalModel(cf, "exp")
    if data is None:     # synthetic data: signal(truth) + noise
        truth = sig.layout.random(seed, dtype, dev)
        tmp_lh = nb.Gaussian(torch.zeros(shape, dtype=dtype, device=dev), noise_cov_inv=NOISE_STD**-2).amend(sig)
        s_truth = tmp_lh.signal_response(truth)
        gen = torch.Generator(dev).manual_seed(seed + 1)
        data = s_truth + NOISE_STD * torch.randn(shape, dtype=dtype, device=dev, generator=gen)
        del tmp_lh
    return nb.Gaussian(data, noise_cov_inv=NOISE_STD**-2).amend(sig), sig


def run_b200(args, shape, wname, rank, world, local_rank):
    import torch
    import nifty_b200 as nb
    from nifty_b200._runtime import metric_multi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if world == 8:
            for k, v in SLAB_ENV.items():           # (NCCL settings of the slab-decomposed extra measurement, read at init)
                os.environ.setdefault(k, v)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    dtype = torch.float64
    lh, sig = build_b200_lh(nb, torch, shape, dev, dtype)
    rt = sig.cf.rt
    L = sig.layout.size
    # every rank = one sample point of the KL (own residual around a common position)
    pos = 0.1 * sig.layout.random(44, dtype, dev) + (0.01 * sig.layout.random(500 + rank, dtype, dev) if world > 1 else 0.0)
    t = sig.layout.random(45, dtype, dev)               # the common tangent (CG direction): replicated
    lin, _ = lh.lin_at(pos)
    out = torch.empty_like(t)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce(buf):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def allreduce_async(buf):
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)

    reduce_chunks = int(os.environ.get("NB200_REDUCE_CHUNKS", "1"))
    if world > 1:
        sig.cf.plan.set_reduce_chunks(reduce_chunks)

    if world == 1:
        def step(tin, o):      # M_p t = lh.metric(pos, t) + t
            return lin.metric(tin, add_identity=True, out=o)
    else:
        # one application of the sample-averaged KL metric (`_kl_met`, optimize_kl.py:117-144) with one sample point per rank:
        # local fused product (scaled by 1/world, the identity added on rank 0) + NCCL all-reduce of the L-sized result,
        # enqueued in-stream from the library's reduction hook -- what every KL-CG iteration of a sharded run executes
        def step(tin, o):
            return metric_multi([lin], tin, scale=1.0 / world, identity_here=(rank == 0), out=o, reduce_fn=allreduce_async)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(t, out)
    barrier()
    n0 = rt.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(t, out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = rt.launch_count() - n0
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax) / args.steps
    value = world * 1e3 / ms_step

    def timed(fn, n):
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    collective = None
    if world > 1:
        scratch = torch.empty_like(out)
        prod_ms = timed(lambda: metric_multi([lin], t, scale=1.0 / world, identity_here=(rank == 0), out=scratch), max(10, args.steps // 4))
        ar_ms = timed(lambda: allreduce(scratch), max(10, args.steps // 4))
        nbytes_ar = out.numel() * out.element_size()
        collective = {"what": f"NCCL all-reduce (SUM) of the metric output, once per step: the last pass runs in {reduce_chunks} launches and "
                              "every finished range of the result is all-reduced on the communication stream beside the remaining launches "
                              "and the cotangent chain (include/nifty_b200.h nb200_reduce_hook)",
                      "bytes": nbytes_ar, "allreduce_ms_alone": ar_ms, "product_ms_alone": prod_ms, "step_ms": ms_step,
                      "bus_GBps_alone": 2.0 * (world - 1) / world * nbytes_ar / (ar_ms * 1e-3) / 1e9,
                      "exposed_ms": ms_step - prod_ms,
                      "overlap": "within the product only: the next CG product needs the all-reduced vector (sequential dependency of the "
                                 "recurrence)", "reduce_chunks": reduce_chunks}

    # per-kernel event timing over the same steps (instrumented pass; the library records events
    # around each of its launches on the stream it launches on)
    rt.timing_begin()
    for _ in range(args.steps):
        step(t, out)
    tm = rt.timing_end()
    tot = sum(v[1] for v in tm.values())
    dom = max(tm.items(), key=lambda kv: kv[1][1])
    dom_name, (dom_cnt, dom_ms) = dom
    kb = kernel_bytes(dom_name, shape)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    achieved = kb / (dom_ms / dom_cnt * 1e-3) / 1e9 if kb else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wname, {}).get(short_kernel_name(dom_name))
    except Exception:
        pass
    short = short_kernel_name(dom_name)
    ab = algorithmic_bytes_mvp(shape)
    prod_only_ms = tot / args.steps          # sum of the product's kernels (events), without collective
    roofline = {"bound": "hbm", "kernel": short, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "kernel_share_of_step": dom_ms / tot,
                "whole_step": {"algorithmic_bytes": ab, "achieved": ab / (ms_step * 1e-3) / 1e9, "frac": ab / (ms_step * 1e-3) / 1e9 / peak},
                "kernels_us": {k.split("nb")[-1][:40]: round(v[1] / v[0] * 1e3, 1) for k, v in sorted(tm.items(), key=lambda kv: -kv[1][1])},
                "product_kernels_ms": prod_only_ms}

    # linearisation (lin.update): paid once per CG solve / Newton step, not per product -- the CPU arm re-linearises on
    # every call like nifty.re (likelihood.py:613-621), this arm reuses the cached linearisation (the optimisation)
    lin_ms = timed(lambda: lin.update(pos, want_grad=True, add_prior=True), 5)

    # end to end through the public API with HOST buffers: every step copies its tangent from pinned
    # host memory to the device, applies the product and copies the result back to pinned host memory.
    # Copies run on their own streams (double-buffered) so that step i+1's H2D and step i-1's D2H overlap
    # step i's kernels -- all K transfers in both directions are inside the timed region.
    t_host = [t.cpu().pin_memory() for _ in range(2)]
    out_host = [torch.empty_like(t_host[0]).pin_memory() for _ in range(2)]
    t_dev = [torch.empty_like(t) for _ in range(2)]
    out_dev = [torch.empty_like(t) for _ in range(2)]
    s_main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n):
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[b])            # the product that last read t_dev[b] is done
                t_dev[b].copy_(t_host[b], non_blocking=True)
                ev_in[b].record(s_in)
            s_main.wait_event(ev_in[b])
            s_main.wait_event(ev_out[b])              # out_dev[b] has been drained
            step(t_dev[b], out_dev[b])
            ev_cmp[b].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[b])
                out_host[b].copy_(out_dev[b], non_blocking=True)
                ev_out[b].record(s_out)
        s_main.wait_stream(s_in)
        s_main.wait_stream(s_out)

    e2e_steps(4)
    barrier()
    e0.record()
    e2e_steps(args.steps)
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * 1e3 / (float(te) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    nbytes = t.numel() * t.element_size()
    assert torch.equal(out_host[(args.steps - 1) & 1], out.cpu())   # the copied-back result is the product

    # sample-draw seconds (second half of the BASELINE metric): one draw_linear_residual-style CG
    # solve with the demo's settings (absdelta = 1e-4 * L / 10, maxiter = 100; demos/re/0_intro.py:105-108)
    j = sig.layout.random(1000 + rank, dtype, dev)
    lin.cg_solve(j, j.clone(), absdelta=1e-4 * L / 10, maxiter=2, raise_nonposdef=False)   # untimed: allocates the CG work vectors
    torch.cuda.synchronize()
    ts = time.perf_counter()
    x, cgres = lin.cg_solve(j, j.clone(), absdelta=1e-4 * L / 10, maxiter=100, raise_nonposdef=False)
    torch.cuda.synchronize()
    draw_s = time.perf_counter() - ts

    if rank == 0:
        cpu, parity = None, None
        if args.cpu_baseline and world == 1:
            t0 = time.time()
            olh, lay, opos, otan, cores = cpu_setup(shape)
            ref = cpu_mvp_once(olh, lay, opos, otan)
            nrep = 2
            t1 = time.perf_counter()
            for _ in range(nrep):
                cpu_mvp_once(olh, lay, opos, otan)
            dt = (time.perf_counter() - t1) / nrep
            cpu = {"value": 1.0 / dt, "unit": "MVP/s", "cores": cores, "kind": "port",
                   "sample": f"{nrep} full {wname} products with the NumPy/scipy.fft oracle port ({cores} threads), {time.time()-t0:.0f}s incl. setup"}
            # parity of the timed product path: the SAME inputs through the B200 path, compared with the oracle's product
            lh2, sig2 = build_b200_lh(nb, torch, shape, dev, dtype, data=torch.as_tensor(olh.bench_data, device=dev))
            lin2, _ = lh2.lin_at(torch.as_tensor(lay.pack(opos), device=dev))
            got = lin2.metric(torch.as_tensor(lay.pack(otan), device=dev), add_identity=True).cpu().numpy()
            want = lay.pack(ref)
            parity = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
            del lh2, lin2
        line = {"metric": "metric_vector_products_per_sec", "value": value, "unit": "MVP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wname, "shape": list(shape), "what": WORKLOADS[wname][1], "latent_size": L,
                           "parallelism": (f"sample-sharded KL metric x{world}: one fused product per rank + NCCL all-reduce of the "
                                           f"{L * 8 / 1e6:.0f} MB result per step (optimize_kl.py:117-144)") if world > 1 else "single GPU",
                           "l2": f"working set per product {ab/1e6:.0f} MB > 126 MB L2 (no explicit flush)",
                           "linearisation": "cached (lin.update once per solve, timed separately as linearize_ms); the CPU arm "
                                            "re-linearises on every call like nifty.re"},
                "clocks": clocks, "e2e": {"value": e2e_val, "unit": "MVP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity_rel_err": parity,
                "linearize_ms": lin_ms, "collective": collective,
                "sample_draw": {"seconds": draw_s, "cg_iterations": int(cgres.nit), "info": int(cgres.info), "nfev": int(cgres.nfev)}}
    else:
        line = None

    # extra measurements of the multi-GPU runs (keys of the same JSON line).  They involve further collectives; a watchdog
    # guarantees the line: if they do not finish in time every rank leaves, rank 0 after printing the line with what it has.
    extras = {}
    if world > 1 and os.environ.get("NB200_BENCH_EXTRAS", "1") != "0":
        limit = float(os.environ.get("NB200_BENCH_EXTRAS_TIMEOUT", "300"))

        def give_up():
            if rank == 0:
                extras.setdefault("extras_error", f"extra measurements did not finish within {limit:.0f} s")
                line.update(extras)
                print(json.dumps(line), flush=True)
            else:
                time.sleep(3.0)
            os._exit(0)

        dog = threading.Timer(limit, give_up)
        dog.daemon = True
        dog.start()
        del t_host, out_host, t_dev, out_dev
        try:
            extras["kl_config3"] = kl_config3(nb, torch, dist, metric_multi, rank, world, dev, peak)
        except Exception as e:  # pragma: no cover
            extras["kl_config3"] = {"error": repr(e)[:300]}
        if world == 8 and os.environ.get("NB200_BENCH_SLAB", "1") != "0":
            # BASELINE.json configs[4] / the north-star target: 1024^3 slab-decomposed product + one MGVI sample draw
            try:
                del lin, lh, sig
                torch.cuda.empty_cache()
                extras["slab_1024"] = slab_measure(nb, torch, dist, (1024, 1024, 1024), "cf3d_1024_f64_slab", rank, world, dev, 10, 3)
            except Exception as e:  # pragma: no cover
                extras["slab_1024"] = {"error": repr(e)[:300]}
        dog.cancel()
    if rank == 0:
        line.update(extras)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def kl_config3(nb, torch, dist, metric_multi, rank, world, dev, peak):
    """BASELINE.json configs[2]: 3-D correlated field 256^3, 16 antithetic samples (8 keys) sharded over the ranks.  One step =
    one application of the sample-averaged KL metric: 16 / world local products accumulated in the last-pass epilogues +
    one NCCL all-reduce of the 135 MB result.  Reported as per-sample products per second of the whole job."""
    shape = (256, 256, 256)
    n_total = 16
    if n_total % world:
        return {"skipped": f"16 samples do not divide over {world} ranks"}
    dtype = torch.float64
    lh, sig = build_b200_lh(nb, torch, shape, dev, dtype)
    pos = 0.1 * sig.layout.random(44, dtype, dev)
    n_loc = n_total // world
    lins = []
    for i in range(n_loc // 2):          # mirrored pairs around pos
        r = 0.05 * sig.layout.random(700 + rank * 8 + i, dtype, dev)
        for sgn in (1.0, -1.0):
            l = lh.new_lin()
            l.update(pos + sgn * r, want_grad=False)
            lins.append(l)
    t = sig.layout.random(45, dtype, dev)
    out = torch.empty_like(t)

    sig.cf.plan.set_reduce_chunks(int(os.environ.get("NB200_REDUCE_CHUNKS", "1")))

    def allreduce(buf):
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)

    def step(hook):
        metric_multi(lins, t, scale=1.0 / n_total, identity_here=(rank == 0), out=out, reduce_fn=hook)

    def timed(fn, n):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    for _ in range(3):
        step(allreduce)
    ms = timed(lambda: step(allreduce), 20)
    ms_local = timed(lambda: step(None), 20)
    ms_ar = timed(lambda: allreduce(out).wait(), 20)
    ab = algorithmic_bytes_mvp(shape)
    return {"workload": "cf3d_256_f64, 16 antithetic samples sharded", "samples_per_rank": n_loc, "ms_per_kl_metric": ms,
            "products_per_sec": n_total * 1e3 / ms, "local_products_ms": ms_local, "allreduce_ms_alone": ms_ar,
            "allreduce_bytes": out.numel() * 8, "exposed_collective_ms": ms - ms_local,
            "hbm_frac_per_gpu": n_loc * ab / (ms * 1e-3) / 1e9 / peak}


SLAB_ENV = {"NCCL_P2P_NVL_CHUNKSIZE": "4194304",    # measured on 8 x B200: exchanges 4.9 -> 3.9 ms at 1024^3
            "NCCL_BUFFSIZE": "16777216",
            "TORCH_NCCL_HIGH_PRIORITY": "1",           # lets the exchange kernels run beside the passes (3.98 vs 4.54 ms at 512^3 x 2)
            "NB200_SLAB_CHUNKS": "4"}


def slab_measure(nb, torch, dist, shape, wname, rank, world, dev, steps, warmup, want_draw=True):
    """One field slab-decomposed over all ranks (strong scaling): step = one metric-vector product with its two chunked
    NCCL exchanges and one all-reduce; plus one MGVI sample draw (CG on metric + 1, demo settings) on that field."""
    dtype = torch.float64
    cfm = nb.CorrelatedFieldMaker("cf", dtype=dtype, comm=True)
    cfm.set_amplitude_total_offset(*OFFSET)
    cfm.add_fluctuations(shape, 1.0 / shape[0], prefix="ax1", non_parametric_kind="power", **CF_KW)
    cf = cfm.finalize()
    plan, rt = cf.plan, cf.rt
    sig = nb.SignalModel(cf, "exp")
    gen = torch.Generator(dev).manual_seed(100 + rank)
    data = torch.randn(plan.local_pos_shape, dtype=dtype, device=dev, generator=gen)
    lh = nb.Gaussian(data, noise_cov_inv=NOISE_STD**-2).amend(sig)
    L = sig.layout.size
    o, n_xi = sig.layout.offsets["cfxi"], sig.layout.numel("cfxi")
    rows_ok = torch.as_tensor(plan.row_map >= 0, device=dev)
    hyper = np.random.default_rng(7)

    def latent(scale_hyper, scale_xi):
        """A local latent vector: the replicated hyper-parameter leaves from ONE host generator, leaf by leaf (bit-identical on
        every rank whatever its local size: the reductions of a slab CG rely on it), the local excitation rows from the
        rank's own device generator."""
        v = torch.zeros(L, dtype=dtype, device=dev)
        for k in sig.layout.keys:
            if k == "cfxi":
                continue
            ko, kn = sig.layout.offsets[k], sig.layout.numel(k)
            v[ko:ko + kn] = torch.as_tensor(scale_hyper * hyper.standard_normal(kn), dtype=dtype, device=dev)
        blk = v[o:o + n_xi].view(plan.local_shape)
        blk.copy_(scale_xi * torch.randn(plan.local_shape, dtype=dtype, device=dev, generator=gen))
        blk[~rows_ok] = 0
        return v

    pos, t = latent(0.1, 0.1), latent(1.0, 0.1)
    lin, _ = lh.lin_at(pos)
    out = torch.empty_like(t)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        lin.metric(t, add_identity=True, out=out)
    n0 = rt.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        lin.metric(t, add_identity=True, out=out)
    e1.record()
    barrier()
    launches = rt.launch_count() - n0
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax) / steps
    rt.timing_begin()
    for _ in range(steps):
        lin.metric(t, add_identity=True, out=out)
    tm = rt.timing_end()
    kern_ms = sum(v[1] for v in tm.values()) / steps
    # end to end: host tangent (local block) -> device -> product -> host
    t_host, out_host = t.cpu().pin_memory(), torch.empty_like(t, device="cpu").pin_memory()
    t_dev = torch.empty_like(t)
    barrier()
    e0.record()
    for _ in range(steps):
        t_dev.copy_(t_host, non_blocking=True)
        out_host.copy_(lin.metric(t_dev, add_identity=True, out=out), non_blocking=True)
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    draw = None
    if want_draw:
        # MGVI sample draw on the slab-decomposed field: CG on (metric + 1) with all-reduced reductions, demo settings
        from nifty_b200.conjugate_gradient import HamiltonianMetric, _cg
        Lg = lh.global_size()
        j = latent(1.0, 1.0)
        op = HamiltonianMetric(lin, likelihood=lh)
        barrier()
        ts = time.perf_counter()
        res = _cg(op, j, x0=j.clone(), absdelta=1e-4 * Lg / 10, maxiter=100, _raise_nonposdef=False)
        barrier()
        draw = {"seconds": time.perf_counter() - ts, "cg_iterations": int(res.nit), "info": int(res.info), "nfev": int(res.nfev),
                "global_latent_size": int(Lg)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    N = int(np.prod(shape))
    ab = algorithmic_bytes_mvp(shape)
    nv = 2 * 8 * (N / world) * (world - 1) / world
    nbytes = t.numel() * t.element_size()
    return {"ms_per_step": ms_step, "value": 1e3 / ms_step, "latent_size_per_rank": L, "launches": int(launches),
            "e2e": {"value": 1e3 / (float(te) / steps), "unit": "MVP/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world},
            "roofline": {"bound": "hbm", "kernel": "whole product (aggregate over ranks)", "achieved": ab / (ms_step * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": ab / (ms_step * 1e-3) / 1e9 / (peak * world), "traffic": None,
                         "peak_source": "measured (MEASURED_PEAKS.json) x n_gpus", "kernel_ms_rank0": kern_ms,
                         "exchange_and_host_ms": ms_step - kern_ms,
                         "nvlink": {"bytes_per_gpu_per_direction": nv, "achieved_GBps_over_whole_step": nv / (ms_step * 1e-3) / 1e9,
                                    "peak_GBps": 770.0, "chunks": plan.nchunks,
                                    "note": "exchanges are pipelined in chunks beside the passes (kernel_ms is measured "
                                            "with per-kernel events, i.e. under contention with the exchange kernels)"}},
            "sample_draw": draw}


def run_b200_slab(args, shape, wname, rank, world, local_rank):
    """One field slab-decomposed over all ranks (strong scaling): value = products per second of the whole job."""
    import torch
    import torch.distributed as dist
    import nifty_b200 as nb
    if world < 2:
        raise SystemExit("slab workloads need --gpus >= 2 under torchrun (one field spread over the ranks)")
    torch.cuda.set_device(local_rank)
    for k, v in SLAB_ENV.items():
        os.environ.setdefault(k, v)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r = slab_measure(nb, torch, dist, shape, wname, rank, world, dev, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = {"metric": "metric_vector_products_per_sec", "value": r["value"], "unit": "MVP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wname, "shape": list(shape), "what": WORKLOADS[wname][1], "latent_size_per_rank": r["latent_size_per_rank"],
                           "parallelism": f"slab-decomposed x{world} (2 chunked NCCL exchanges + 1 all-reduce per product)",
                           "l2": "working set per product per GPU >> 126 MB L2 (no explicit flush)"},
                "clocks": clocks, "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": r["roofline"], "cpu_baseline": None,
                "sample_draw": r["sample_draw"]}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cf2d_4096_f64", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    shape = WORKLOADS[args.workload][0]
    if args.impl == "reference":
        if args.steps > 10 and np.prod(shape) >= 2**24:
            args.steps = 10     # bounded sample: ~1-3 s per product on the host cores
        run_reference(args, shape, args.workload, rank, world)
    elif args.workload.endswith("_slab"):
        run_b200_slab(args, shape, args.workload, rank, world, local_rank)
    else:
        run_b200(args, shape, args.workload, rank, world, local_rank)


if __name__ == "__main__":
    main()
