"""CPU baseline with the reference itself: nifty.re on JAX-CPU, protocol of misc/re/paper/minimal_benchmark.py:56-74,156-165
(warm-up, timeit autorange, 7 repeats, median, block_until_ready; threads through XLA_FLAGS :25-30).

`nifty.re` needs JAX, which is neither installed nor installable in this image (no network, not in /opt/wheelhouse), so
`bench.py --impl reference` times the NumPy/scipy.fft oracle port instead (`cpu_baseline.kind = "port"`).  This script is the
drop-in replacement for that leg the moment `import jax` and `import nifty.re` work: `bench.py` calls `available()` and,
if true, `time_metric_products(...)`, and reports `cpu_baseline.kind = "reference"`.

Usage:  python baseline/run_nifty_re_cpu.py [--shape 4096,4096] [--threads N]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import timeit


def available() -> bool:
    try:
        import jax  # noqa: F401
        import nifty.re  # noqa: F401
        return True
    except Exception:
        return False


def build(shape, seed=42):
    """The synthetic workload of SURVEY.md section 8(d) (demos/re/0_intro.py:23-80): power-kind correlated field, exp, Gaussian
    noise 0.1."""
    import jax
    import jax.numpy as jnp
    import nifty.re as jft
    jax.config.update("jax_enable_x64", True)
    dims = tuple(shape)
    cfm = jft.CorrelatedFieldMaker("cf")
    cfm.set_amplitude_total_offset(offset_mean=0.0, offset_std=(1e-3, 1e-4))
    cfm.add_fluctuations(dims, distances=1.0 / dims[0], fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2),
                         flexibility=(1.0, 0.5), asperity=(0.5, 0.05), prefix="ax1", non_parametric_kind="power")
    cf = cfm.finalize()

    class Signal(jft.Model):
        def __init__(self):
            self.cf = cf
            super().__init__(init=cf.init)

        def __call__(self, x):
            return jnp.exp(self.cf(x))

    sig = Signal()
    key = jax.random.PRNGKey(seed)
    key, k1, k2, k3, k4 = jax.random.split(key, 5)
    truth = jft.random_like(k1, sig.domain)
    data = sig(truth) + 0.1 * jax.random.normal(k2, dims)
    lh = jft.Gaussian(data, noise_cov_inv=lambda x: x / 0.01).amend(sig)
    pos = 0.1 * jft.Vector(jft.random_like(k3, sig.domain))
    tan = jft.Vector(jft.random_like(k4, sig.domain))
    return lh, pos, tan


def time_metric_products(shape, threads=None):
    """Median seconds per `lh.metric(pos, tan) + tan` (re-linearising on every call, as nifty.re does)."""
    if threads:
        os.environ["XLA_FLAGS"] = (f"--xla_cpu_multi_thread_eigen={'true' if threads > 1 else 'false'} "
                                   f"intra_op_parallelism_threads={threads}")
    import jax
    lh, pos, tan = build(shape)
    met = jax.jit(lambda p, t: lh.metric(p, t) + t)
    jax.block_until_ready(met(pos, tan))
    tm = timeit.Timer(lambda: jax.block_until_ready(met(pos, tan)))
    n, _ = tm.autorange()
    return sorted(t / n for t in tm.repeat(7, n))[3]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4096,4096")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    if not available():
        print(json.dumps({"impl": "reference", "unavailable": "nifty.re needs jax / jaxlib, which are not installed in this image"}))
        return 0
    shape = tuple(int(s) for s in a.shape.split(","))
    s = time_metric_products(shape, a.threads)
    print(json.dumps({"impl": "reference", "kind": "reference", "metric": "metric_vector_products_per_sec", "value": 1.0 / s,
                      "unit": "MVP/s", "cores": a.threads, "shape": list(shape)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
