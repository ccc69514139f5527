"""CPU tier: the kernel sources executed sequentially on the host (tests/emu) against the oracle and
the nifty.cl fixtures.  Checks host logic, index arithmetic and call sequences without a GPU; the
GPU tier (test_gpu_parity.py) runs the same checks through libniftyb200.so."""
import numpy as np
import pytest
import torch

import nifty_b200 as nb
import parity_checks as pc
from nifty_b200._capi import CApi


@pytest.fixture(scope="module")
def rt():
    from emu.build_emu import build
    return nb.Runtime(CApi(build()), "cpu")


@pytest.mark.parametrize("shape,dist", [((16,), 0.1), ((2,), 1.0), ((4, 4), 1.0), ((8, 32), (0.3, 0.11)),
                                        ((2, 2), 1.0), ((4, 8, 16), (0.2, 0.1, 0.05)), ((2, 2, 2), 1.0),
                                        ((32, 4), 1.0), ((4, 2, 32), 1.0), ((128, 128), 1.0 / 128)])
def test_tables_hartley_bilinear(rt, shape, dist):
    pc.check_mode_tables(rt, shape, dist)
    pc.check_hartley(rt, shape)
    pc.check_hartley(rt, shape, convention="canonical_hartley")
    pc.check_bilinear(rt, shape, dist)


def test_hartley_float32(rt):
    pc.check_hartley(rt, (16, 32), dtype=torch.float32)
    pc.check_bilinear(rt, (8, 4, 16), 0.5, dtype=torch.float32)


@pytest.mark.parametrize("name", pc.POW2_CASES)
def test_golden(rt, name):
    pc.check_golden(rt, name)


def test_golden_float32(rt):
    pc.check_golden(rt, "g2d_16x16", dtype=torch.float32)


@pytest.mark.parametrize("kind", ["amplitude", "power"])
def test_kind_scaling_convention(rt, kind):
    pc.check_kind_and_scaling(rt, kind=kind)


def test_metric_properties(rt):
    pc.check_metric_properties(rt, (32, 64), (0.1, 0.05))
    pc.check_metric_properties(rt, (8, 16, 8), 0.3, lh_kind="poisson")


def test_model_surface(rt):
    pc.check_model_surface(rt)
    pc.check_model_surface(rt, "g3d_8x8x8")


def test_cg(rt):
    pc.check_cg(rt)


def test_matern_variants(rt):
    pc.check_matern_variants(rt)


def test_multichunk_amplitude_chain(rt):
    """K > 2048 mode bins: the amplitude scans span several chunks (carry-in from chunk aggregates)."""
    pc.check_against_oracle(rt, (128, 256), (0.01, 0.02))
    pc.check_against_oracle(rt, (256, 256), 1.0 / 256, lh_kind="poisson", asperity=None)
    pc.check_against_oracle(rt, (64, 64), 0.1, flexibility=None, asperity=None)


@pytest.mark.parametrize("maxgrid", ["1", "2", "5"])
def test_fused_chain_rounds(rt, monkeypatch, maxgrid):
    """Fused chain kernels (nb_chain.cuh) with a capped grid: several scan rounds and segment-sum windows per CTA,
    carry-in across CTAs, thread prefixes of every round kept across the grid barrier."""
    monkeypatch.setenv("NB200_COOP_MAXGRID", maxgrid)
    pc.check_against_oracle(rt, (256, 256), 1.0 / 256)
    pc.check_against_oracle(rt, (32, 32, 32), 0.1, lh_kind="poisson")


def test_legacy_chain_kernels(rt, monkeypatch):
    """The five-launch chain (segment sum, aggregate / apply scans) that slab-decomposed plans still use."""
    monkeypatch.setenv("NB200_COOP", "0")
    pc.check_against_oracle(rt, (128, 256), (0.01, 0.02))
    pc.check_against_oracle(rt, (16, 16, 16), 0.1, lh_kind="poisson")


@pytest.mark.parametrize("lg_r", ["1", "2", "3"])
def test_p1_mirror_quads(rt, monkeypatch, lg_r):
    """P1MBody (one bin lookup per mirror quad) needs >= 2 lines per CTA, which the launch heuristic only picks
    for grids with many lines: force it (developer knob read at plan creation) on small grids."""
    monkeypatch.setenv("NB200_LGR1", lg_r)
    for shape, dist in [((8, 32), (0.3, 0.11)), ((16, 16), 1.0), ((8, 8, 16), 0.2), ((64, 2), 1.0)]:
        pc.check_bilinear(rt, shape, dist)
    pc.check_against_oracle(rt, (32, 64), (0.1, 0.05))
    pc.check_against_oracle(rt, (8, 16, 8), 0.3, lh_kind="poisson")
    monkeypatch.setenv("NB200_SCAN_E", "4")
    pc.check_against_oracle(rt, (64, 128), (0.01, 0.02))


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations(rt, monkeypatch, seed):
    """Seeded random draws of the configuration space the path supports -- 1 to 3 power-of-two axes of unequal extent and
    spacing, both likelihoods, optional flexibility / asperity, forced lines-per-CTA choices and scan chunk sizes -- each
    checked against the oracle (energy, gradient, metric, both sqrt-metrics)."""
    rng = np.random.default_rng(1000 + seed)
    ndim = int(rng.integers(1, 4))
    shape = tuple(int(2 ** rng.integers(1, 6 if ndim < 3 else 5)) for _ in range(ndim))
    dist = tuple(float(d) for d in rng.uniform(0.05, 0.7, ndim))
    kw = {}
    r = rng.integers(0, 3)
    if r == 1:
        kw["asperity"] = None
    elif r == 2:
        kw["flexibility"], kw["asperity"] = None, None
    for knob in ("NB200_LGR1", "NB200_LGR3", "NB200_LGR5", "NB200_LGRC"):
        if rng.integers(0, 2):
            monkeypatch.setenv(knob, str(int(rng.integers(0, 4))))
    if rng.integers(0, 2):
        monkeypatch.setenv("NB200_SCAN_E", "4")
    pc.check_against_oracle(rt, shape if ndim > 1 else shape, dist if ndim > 1 else dist[0], lh_kind="gauss" if rng.integers(0, 2) else "poisson",
                            seed=int(rng.integers(0, 1 << 30)), **kw)


def test_unsupported_shapes_fail_loudly(rt):
    with pytest.raises(nb.NB200Error, match="power of two"):
        nb.Plan((3, 3), 0.1, runtime=rt)
    with pytest.raises(nb.NB200Error):
        nb.Plan((4, 4, 4, 4), 0.1, runtime=rt)


@pytest.mark.parametrize("name", ["small_gauss", "small_poisson"])
def test_baseline_config_check_small(rt, name):
    """The full-size check of the GPU tier (energy / gradient / metric / metric + 1 vs the oracle) at oracle-friendly sizes."""
    pc.check_config_vs_oracle(rt, name)


@pytest.mark.parametrize("shape,dist,kind", [((64, 128), (0.01, 0.02), "gauss"), ((512, 64), 0.1, "poisson"),
                                             ((32, 128, 64), (0.1, 0.2, 0.3), "poisson")])
def test_staged_chain_vs_oracle(rt, monkeypatch, shape, dist, kind):
    """The staged metric chain (nb_passes2.cuh) forced on (the launch heuristic reserves it for long 2-D lines): P1F / PCF /
    P3F, the tensor-map gather of the generic last pass and the register-resident last pass (NB200_P5F), 256 virtual
    threads per block on the host."""
    monkeypatch.setenv("NB200_CHAIN", "1")
    pc.check_against_oracle(rt, shape, dist, lh_kind=kind)
    monkeypatch.setenv("NB200_P5F", "1")
    pc.check_against_oracle(rt, shape, dist, lh_kind=kind)


@pytest.mark.parametrize("which,lh_kind,shape", [("softplus", "gauss", (16, 32)), ("sigmoid_scaled", "poisson", (32, 16)),
                                                 ("square_plus", "gauss", (8, 8, 16))])
def test_custom_pointwise_nonlinearity(rt, which, lh_kind, shape):
    pc.check_custom_nonlinearity(rt, shape, 0.1, lh_kind=lh_kind, which=which)


@pytest.mark.parametrize("shapes,lh_kind,conv", [(((8, 16), (4,)), "gauss", "non_canonical_hartley"),
                                                 (((16,), (8, 8)), "poisson", "canonical_hartley"),
                                                 (((6, 5), (3,)), "poisson", "non_canonical_hartley"),
                                                 (((4, 2), (3, 2)), "gauss", "non_canonical_hartley")])
def test_outer_product_of_two_subgrids(rt, shapes, lh_kind, conv):
    pc.check_outer_product(rt, shapes=shapes, lh_kind=lh_kind, conv=conv)


@pytest.mark.parametrize("name", ["o_8x16_x_4", "o_16_x_8x8", "o_3x3_x_6", "o_6_x_3x3", "o_6_x_6", "o_4_x_6_x_8", "o_3x3_x_3x3", "o_m8_x_3x4"])
def test_outer_product_golden(rt, name):
    pc.check_outer_golden(rt, name)


@pytest.mark.parametrize("shape", [(3,), (6,), (3, 3), (6, 7), (5, 12), (3, 5, 7), (100, 37), (1, 5)])
def test_nonpow2_hartley(rt, shape):
    pc.check_nonpow2_hartley(rt, shape)


def test_nonpow2_hartley_float32_and_errors(rt):
    pc.check_nonpow2_hartley(rt, (12, 7), dtype=torch.float32)
    pc.check_nonpow2_errors(rt)


@pytest.mark.parametrize("name", ["g2d_3x3", "m2d_3x3"])
def test_nonpow2_golden_3x3(rt, name):
    pc.check_nonpow2_golden(rt, name)


def test_reference_cf_cases(rt):
    pc.check_reference_cf_cases(rt)


@pytest.mark.parametrize("shape,dist,lh_kind,conv", [((6, 10), (0.2, 0.3), "gauss", "non_canonical_hartley"),
                                                     ((5, 3, 6), 0.4, "poisson", "canonical_hartley"), ((12,), 0.4, "gauss", "non_canonical_hartley")])
def test_nonpow2_model(rt, shape, dist, lh_kind, conv):
    pc.check_nonpow2_model(rt, shape, dist, lh_kind, conv)


def test_host_composed_matern(rt):
    pc.check_host_composed_matern(rt)


def test_host_composed_scaling_leaf(rt):
    pc.check_host_composed_scaling(rt)


@pytest.mark.parametrize("shape,nl", [((8, 8), "exp"), ((4, 8, 4), "identity"), ((6, 5), "exp")])
def test_gaussian_with_non_diagonal_covariance(rt, shape, nl):
    pc.check_operator_gaussian(rt, shape, nl)
