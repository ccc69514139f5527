"""Golden vectors from nifty.re ITSELF (JAX-CPU), for the day a JAX wheel is present next to /root/reference.

Today the committed fixtures `tests/golden/*.npz` come from the reference's second implementation, `nifty.cl`
(`tests/golden/make_golden.py`), because `nifty.re` cannot be imported here (no JAX).  This script writes the same
quantities -- field, JVP, VJP, energy, gradient, metric-vector product for the cases of `tests/golden/golden_cases.py`
-- with `nifty.re`, into `tests/golden/re_<case>.npz`; `tests/test_oracle_golden.py` picks those files up when they exist
(tolerance 1e-10, the north-star parity bar).  It exits with a message and status 0 when JAX is missing.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    try:
        import jax
        import jax.numpy as jnp
        sys.path.insert(0, "/root/reference")
        import nifty.re as jft
    except Exception as e:  # pragma: no cover
        print(f"nifty.re is not importable here ({e.__class__.__name__}: {e}); nothing written")
        return 0
    jax.config.update("jax_enable_x64", True)
    from golden_util import CASES, load           # the cases and inputs of the nifty.cl fixtures
    for name, c in CASES.items():
        if c.get("matern"):
            continue
        g = load(name)
        cfm = jft.CorrelatedFieldMaker("cf")
        cfm.set_amplitude_total_offset(offset_mean=c["offset_mean"], offset_std=tuple(c["offset_std"]))
        cfm.add_fluctuations(tuple(c["shape"]), distances=c["distances"], prefix="ax1", non_parametric_kind="power", **c["fluct"])
        cf = cfm.finalize()
        pos = {k: jnp.asarray(v) for k, v in g["pos"].items()}
        tan = {k: jnp.asarray(v) for k, v in g["tan"].items()}
        sig = lambda x: jnp.exp(cf(x))
        field = cf(pos)
        _, jvp = jax.jvp(cf, (pos,), (tan,))
        if c["lh"] == "gauss":
            lh = jft.Gaussian(jnp.asarray(g["data"]), noise_cov_inv=lambda x: x * c["noise_cov_inv"]).amend(jft.Model(sig, init=cf.init))
        else:
            lh = jft.Poissonian(jnp.asarray(g["data"]).astype(int)).amend(jft.Model(sig, init=cf.init))
        energy, grad = jax.value_and_grad(lh)(pos)
        metric = lh.metric(pos, tan)
        out = {"field": np.asarray(field), "jvp": np.asarray(jvp), "energy": float(energy)}
        out.update({f"grad/{k}": np.asarray(v) for k, v in grad.items()})
        out.update({f"metric/{k}": np.asarray(v) for k, v in metric.items()})
        np.savez(os.path.join(ROOT, "tests", "golden", f"re_{name}.npz"), **out)
        print("wrote", name)
    return 0


if __name__ == "__main__":
    sys.exit(main())
