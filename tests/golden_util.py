"""Helpers shared by the CPU (oracle) and GPU parity tests: load the nifty.cl fixtures."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must stay in sync with tests/golden/make_golden.py:CASES (the generator is the source of truth;
# the hyper-parameters are repeated here because /root/reference is absent on the GPU box)
CASES = {
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
    "m2d_16x8": dict(shape=(16, 8), distances=(0.1, 0.3), offset_mean=0.2, offset_std=(0.1, 0.1),
                     matern=dict(scale=(1.0, 1.0), cutoff=(1.0, 0.5), loglogslope=(-3.0, 0.5)), lh="gauss", seed=9),
    "m3d_8x4x8": dict(shape=(8, 4, 8), distances=0.5, offset_mean=0.0, offset_std=(0.1, 0.1),
                      matern=dict(scale=(3.0, 2.0), cutoff=(0.3, 0.05), loglogslope=(-4.0, 0.5)), lh="poisson", seed=10),
    # Matern amplitude on the reference's (3, 3) grid with distances 5.0 (test_correlated_field.py:195-229)
    "m2d_3x3": dict(shape=(3, 3), distances=5.0, offset_mean=0.0, offset_std=(0.1, 0.1),
                    matern=dict(scale=(3.0, 2.0), cutoff=(0.1, 0.01), loglogslope=(5.0, 0.5)), lh="gauss", seed=42),
}


# outer products of two sub-grids (make_golden.py:OUTER_CASES; reference check test/test_re/test_correlated_field.py:245-283)
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


def build_outer(c, maker):
    """Two `add_fluctuations` on a `CorrelatedFieldOracle("cf")` / `nb.CorrelatedFieldMaker("cf", ...)` (same call protocol)."""
    maker.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    for i, (shp, dist, f) in enumerate(zip(c["shapes"], c["distances"], c["fluct"])):
        if "matern" in f:
            maker.add_fluctuations_matern(shp, dist, renormalize_amplitude=False, prefix=f"space{i}", non_parametric_kind="amplitude", **f["matern"])
        else:
            maker.add_fluctuations(shp, dist, prefix=f"space{i}", non_parametric_kind="power", **f)
    return maker.finalize()


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"pos": {}, "tan": {}, "field_vjp": {}, "grad": {}, "metric": {}}
    for k in z.files:
        if "/" in k:
            grp, key = k.split("/", 1)
            out[grp][key] = z[k]
        else:
            out[k] = z[k]
    return out


def build_oracle(c):
    from oracle import CorrelatedFieldOracle, GaussianOracle, PoissonianOracle, SignalOracle
    cf = CorrelatedFieldOracle("cf")
    cf.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    if "matern" in c:
        cf.add_fluctuations_matern(c["shape"], c["distances"], renormalize_amplitude=False, prefix="ax1",
                                   non_parametric_kind="amplitude", **c["matern"])
    else:
        cf.add_fluctuations(c["shape"], c["distances"], c["fluctuations"], c["loglogavgslope"],
                            c["flexibility"], c["asperity"], prefix="ax1", non_parametric_kind="power")
    cf.finalize()
    return cf


def build_oracle_lh(c, g):
    from oracle import GaussianOracle, PoissonianOracle, SignalOracle
    cf = build_oracle(c)
    sig = SignalOracle(cf, "exp")
    if c["lh"] == "gauss":
        return GaussianOracle(g["data"], float(g["noise_cov_inv"]), sig)
    return PoissonianOracle(g["data"], sig)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
