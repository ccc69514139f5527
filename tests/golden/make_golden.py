#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  ``nifty.re`` needs JAX,
which is not installed; the reference's second implementation of the same maths,
``nifty.cl``, is importable through two throw-away shims created in a temp dir
(a ``nifty-9.2.0.dist-info`` because ``nifty/__init__.py`` asks importlib.metadata
for the version, and a ``ducc0`` stub without an ``fft`` sub-module so that
``nifty/cl/ducc_dispatch.py:152-156`` falls back to ``scipy.fft``).  Nothing is
copied from or written into the reference tree.

The ``nifty.re`` <-> ``nifty.cl`` parameter mapping is the one of the reference's own
parity test, test/test_re/test_correlated_field.py:116-192: identical keys, the
``spectrum`` leaf transposed ((2, K-2) in cl, (K-2, 2) in re) and
``non_parametric_kind="power"`` on the re side.

Usage:  python tests/golden/make_golden.py        (rewrites tests/golden/*.npz)
"""

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_nifty_cl():
    shim = tempfile.mkdtemp(prefix="nifty_shim_")
    os.makedirs(os.path.join(shim, "nifty-9.2.0.dist-info"))
    with open(os.path.join(shim, "nifty-9.2.0.dist-info", "METADATA"), "w") as f:
        f.write("Metadata-Version: 2.1\nName: nifty\nVersion: 9.2.0\n")
    os.makedirs(os.path.join(shim, "ducc0"))
    with open(os.path.join(shim, "ducc0", "__init__.py"), "w") as f:
        f.write("__version__ = '0.0'\nfrom . import misc\n")
    with open(os.path.join(shim, "ducc0", "misc.py"), "w") as f:
        f.write("import os\ndef resize_thread_pool(n): pass\n"
                "def available_hardware_threads(): return os.cpu_count()\n")
    sys.path.insert(0, REF)
    sys.path.insert(0, shim)
    import nifty.cl as ift
    return ift


CASES = {
    # name: (shape, distances, offset_mean, offset_std, fluct-kwargs, likelihood, seed, scaling)
    "g2d_16x16": dict(shape=(16, 16), distances=1.0 / 16, offset_mean=0.0, offset_std=(1e-3, 1e-4),
                      fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5),
                      asperity=(0.5, 0.05), lh="gauss", seed=42),
    "g2d_8x32": dict(shape=(8, 32), distances=(0.3, 0.11), offset_mean=0.5, offset_std=(0.1, 0.1),
                     fluctuations=(1.0, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.1),
                     asperity=None, lh="gauss", seed=7),
    "p2d_32x32": dict(shape=(32, 32), distances=1.0 / 32, offset_mean=2.0, offset_std=(0.1, 0.03),
                      fluctuations=(1.0, 0.5), loglogavgslope=(-3.0, 0.2), flexibility=(1.0, 0.2),
                      asperity=(0.5, 0.05), lh="poisson", seed=3),
    "g3d_8x8x8": dict(shape=(8, 8, 8), distances=1.0 / 8, offset_mean=0.0, offset_std=(1e-3, 1e-4),
                      fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5),
                      asperity=(0.5, 0.05), lh="gauss", seed=11),
    "g3d_4x8x16": dict(shape=(4, 8, 16), distances=(0.2, 0.1, 0.05), offset_mean=-0.3,
                       offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1), loglogavgslope=(-2.5, 0.5),
                       flexibility=(2.0, 1.0), asperity=(0.2, 0.02), lh="gauss", seed=5),
    "g1d_64": dict(shape=(64,), distances=0.1, offset_mean=0.0, offset_std=(0.1, 0.1),
                   fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1),
                   asperity=(0.2, 2e-2), lh="gauss", seed=0),
    "g2d_3x3": dict(shape=(3, 3), distances=0.1, offset_mean=0.0, offset_std=(0.1, 0.1),
                    fluctuations=(3.0, 2.0), loglogavgslope=(4.0, 1.0), flexibility=(3.0, 2.0),
                    asperity=(0.2, 2e-2), lh="gauss", seed=42),
    # Matern amplitude (test/test_re/test_correlated_field.py:195-228: re kind="amplitude",
    # renormalize_amplitude=False == cl add_fluctuations_matern defaults)
    "m2d_16x8": dict(shape=(16, 8), distances=(0.1, 0.3), offset_mean=0.2, offset_std=(0.1, 0.1),
                     matern=dict(scale=(1.0, 1.0), cutoff=(1.0, 0.5), loglogslope=(-3.0, 0.5)), lh="gauss", seed=9),
    "m3d_8x4x8": dict(shape=(8, 4, 8), distances=0.5, offset_mean=0.0, offset_std=(0.1, 0.1),
                      matern=dict(scale=(3.0, 2.0), cutoff=(0.3, 0.05), loglogslope=(-4.0, 0.5)), lh="poisson", seed=10),
    # Matern amplitude on the reference's (3, 3) grid with distances 5.0 (test_correlated_field.py:195-229)
    "m2d_3x3": dict(shape=(3, 3), distances=5.0, offset_mean=0.0, offset_std=(0.1, 0.1),
                    matern=dict(scale=(3.0, 2.0), cutoff=(0.1, 0.01), loglogslope=(5.0, 0.5)), lh="gauss", seed=42),
}


def build_cl(ift, c):
    sp = ift.RGSpace(c["shape"], c["distances"])
    cfm = ift.CorrelatedFieldMaker("cf")
    cfm.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    if "matern" in c:
        cfm.add_fluctuations_matern(sp, prefix="ax1", **c["matern"])
    else:
        cfm.add_fluctuations(sp, c["fluctuations"], c["flexibility"], c["asperity"],
                             c["loglogavgslope"], prefix="ax1")
    cf = cfm.finalize(prior_info=0)
    return cf


def to_cl(ift, dom, tree):
    d = {}
    for k, v in tree.items():
        v = np.asarray(v)
        if k.endswith("spectrum"):
            v = v.T
        d[k] = ift.makeField(dom[k], np.array(v, order='C'))
    return ift.MultiField.from_dict(d, dom)


def from_cl(mf):
    out = {}
    for k, v in mf.to_dict().items():
        a = np.asarray(v.asnumpy() if hasattr(v, "asnumpy") else v.val)
        if k.endswith("spectrum"):
            a = a.T
        out[k] = np.array(a, order='C')
    return out


def main(only=None):
    ift = _import_nifty_cl()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import CorrelatedFieldOracle, Layout  # only to draw inputs in oracle key order

    for name, c in CASES.items():
        if only and name not in only:
            continue
        cf = build_cl(ift, c)
        orc = CorrelatedFieldOracle("cf")
        orc.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
        if "matern" in c:
            orc.add_fluctuations_matern(c["shape"], c["distances"], renormalize_amplitude=False, prefix="ax1",
                                        non_parametric_kind="amplitude", **c["matern"])
        else:
            orc.add_fluctuations(c["shape"], c["distances"], c["fluctuations"], c["loglogavgslope"],
                                 c["flexibility"], c["asperity"], prefix="ax1", non_parametric_kind="power")
        orc.finalize()
        lay = Layout(orc.domain)
        rng = np.random.default_rng(c["seed"])
        pos = lay.random(rng)
        tan = lay.random(rng)
        cot = rng.standard_normal(c["shape"])
        out = {}
        for k in lay.keys:
            out["pos/" + k] = pos[k]
            out["tan/" + k] = tan[k]
        out["cot"] = cot

        npos = to_cl(ift, cf.domain, pos)
        ntan = to_cl(ift, cf.domain, tan)
        ncot = ift.makeField(cf.target, cot)
        lin = cf(ift.Linearization.make_var(npos))
        out["field"] = lin.val.asnumpy()
        out["field_jvp"] = lin.jac(ntan).asnumpy()
        for k, v in from_cl(lin.jac.adjoint(ncot)).items():
            out["field_vjp/" + k] = v

        signal = cf.exp()
        slin = signal(ift.Linearization.make_var(npos, want_metric=True))
        sig = slin.val.asnumpy()
        out["signal"] = sig
        if c["lh"] == "gauss":
            noise_std = 0.1
            data = sig + noise_std * rng.standard_normal(c["shape"])
            out["data"] = data
            out["noise_cov_inv"] = np.array(noise_std**-2)
            N_inv = ift.ScalingOperator(cf.target, noise_std**-2, float)
            lh = ift.GaussianEnergy(data=ift.makeField(cf.target, data), inverse_covariance=N_inv) @ signal
        else:
            data = rng.poisson(sig).astype(np.int64)
            out["data"] = data
            lh = ift.PoissonianEnergy(ift.makeField(cf.target, data)) @ signal
        elin = lh(ift.Linearization.make_var(npos, want_metric=True))
        out["energy"] = np.array(elin.val.asnumpy())
        for k, v in from_cl(elin.gradient).items():
            out["grad/" + k] = v
        for k, v in from_cl(elin.metric(ntan)).items():
            out["metric/" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, "L =", lay.size)


# Outer products of two sub-grids (test/test_re/test_correlated_field.py:245-283: re == cl field values, both sub-grids with
# non_parametric_kind="power" on the re side)
OUTER_CASES = {
    "o_8x16_x_4": dict(shapes=((8, 16), (4,)), distances=((0.2, 0.1), (0.5,)), offset_mean=0.3, offset_std=(0.2, 0.1), seed=21,
                       fluct=(dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05)),
                              dict(fluctuations=(0.3, 0.2), loglogavgslope=(-1.5, 0.2), flexibility=(0.8, 0.3), asperity=(0.2, 0.02)))),
    "o_16_x_8x8": dict(shapes=((16,), (8, 8)), distances=((1.0,), (0.1, 0.1)), offset_mean=0.0, offset_std=(0.1, 0.1), seed=42,
                       fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                              dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)))),
    # the shapes of the reference's own product test (test_correlated_field.py:232-283: CFG_OFFSET, CFG_FLUCT, FLUCTUATIONS_CHOICES):
    # sub-grids whose extents are not powers of two
    "o_3x3_x_6": dict(shapes=((3, 3), (6,)), distances=((0.1, 0.1), (1.0,)), offset_mean=0.0, offset_std=(0.1, 0.1), seed=0,
                      fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                             dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)))),
    "o_6_x_3x3": dict(shapes=((6,), (3, 3)), distances=((1.0,), (0.1, 0.1)), offset_mean=0.0, offset_std=(0.1, 0.1), seed=42,
                      fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                             dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)))),
    "o_6_x_6": dict(shapes=((6,), (6,)), distances=((1.0,), (1.0,)), offset_mean=0.0, offset_std=(0.1, 0.1), seed=0,
                    fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                           dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)))),
    # three sub-grids (three 1-D grids, one of them not a power of two)
    "o_4_x_6_x_8": dict(shapes=((4,), (6,), (8,)), distances=((0.5,), (1.0,), (0.25,)), offset_mean=0.2, offset_std=(0.1, 0.1), seed=5,
                        fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                               dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05)),
                               dict(fluctuations=(0.3, 0.2), loglogavgslope=(-1.5, 0.2), flexibility=(0.8, 0.3), asperity=(0.2, 0.02)))),
    # four axes in total: the (3, 3) x (3, 3) case of the reference's product test (every sub-grid transformed slice by slice)
    "o_3x3_x_3x3": dict(shapes=((3, 3), (3, 3)), distances=((0.1, 0.1), (0.1, 0.1)), offset_mean=0.0, offset_std=(0.1, 0.1), seed=42,
                        fluct=(dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)),
                               dict(fluctuations=(1.0, 0.1), loglogavgslope=(-1.0, 0.1), flexibility=(1.0, 0.1), asperity=(0.2, 2e-2)))),
    # a Matern sub-grid inside an outer product (Matern x non-parametric; cl's add_fluctuations_matern = re's kind "amplitude", no renormalisation)
    "o_m8_x_3x4": dict(shapes=((8,), (3, 4)), distances=((0.5,), (0.25, 0.25)), offset_mean=0.1, offset_std=(0.2, 0.1), seed=13,
                       fluct=(dict(matern=dict(scale=(1.0, 0.5), cutoff=(1.0, 0.5), loglogslope=(-3.0, 0.5))),
                              dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05)))),
}


def main_outer(only=None):
    ift = _import_nifty_cl()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import CorrelatedFieldOracle, Layout
    for name, c in OUTER_CASES.items():
        if only and name not in only:
            continue
        cfm = ift.CorrelatedFieldMaker("cf")
        cfm.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
        orc = CorrelatedFieldOracle("cf")
        orc.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
        for i, (shp, dist, f) in enumerate(zip(c["shapes"], c["distances"], c["fluct"])):
            if "matern" in f:
                cfm.add_fluctuations_matern(ift.RGSpace(shp, dist), prefix=f"space{i}", **f["matern"])
                orc.add_fluctuations_matern(shp, dist, renormalize_amplitude=False, prefix=f"space{i}", non_parametric_kind="amplitude",
                                            **f["matern"])
                continue
            cfm.add_fluctuations(ift.RGSpace(shp, dist), f["fluctuations"], f["flexibility"], f["asperity"], f["loglogavgslope"],
                                 prefix=f"space{i}")
            orc.add_fluctuations(shp, dist, prefix=f"space{i}", non_parametric_kind="power", **f)
        cf = cfm.finalize(prior_info=0)
        orc.finalize()
        lay = Layout(orc.domain)
        rng = np.random.default_rng(c["seed"])
        pos, tan = lay.random(rng), lay.random(rng)
        full = sum((tuple(shp) for shp in c["shapes"]), ())
        cot = rng.standard_normal(full)
        out = {"cot": cot}
        for k in lay.keys:
            out["pos/" + k] = pos[k]
            out["tan/" + k] = tan[k]
        npos, ntan = to_cl(ift, cf.domain, pos), to_cl(ift, cf.domain, tan)
        lin = cf(ift.Linearization.make_var(npos))
        out["field"] = lin.val.asnumpy()
        out["field_jvp"] = lin.jac(ntan).asnumpy()
        for k, v in from_cl(lin.jac.adjoint(ift.makeField(cf.target, cot))).items():
            out["field_vjp/" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, "L =", lay.size)


if __name__ == "__main__":
    if len(sys.argv) > 1:            # `make_golden.py o_3x3_x_6 m2d_3x3 ...`: (re)write only the named fixtures
        main(only=set(sys.argv[1:]))
        main_outer(only=set(sys.argv[1:]))
    else:
        main()
        main_outer()
