"""GPU tier (-m gpu), host-composed model families: outer products of sub-grids (nifty_b200/outer.py) and grids whose extents are
not powers of two (nifty_b200/bluestein.py, the chirp-z transform `nb200_hartley_chirpz`), against the nifty.cl fixtures and the
oracle.  Collected after the tests of the fused path (file name order)."""
import pytest
import torch

import nifty_b200 as nb
import parity_checks as pc
import vi_checks as vc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    return nb.default_runtime()


@pytest.mark.parametrize("shapes,lh_kind,conv", [(((8, 16), (4,)), "gauss", "non_canonical_hartley"),
                                                 (((16,), (8, 8)), "poisson", "canonical_hartley"),
                                                 (((6, 5), (3,)), "poisson", "non_canonical_hartley"),
                                                 (((4, 2), (3, 2)), "gauss", "non_canonical_hartley")])
def test_outer_product_of_two_subgrids(rt, shapes, lh_kind, conv):
    pc.check_outer_product(rt, shapes=shapes, lh_kind=lh_kind, conv=conv)


@pytest.mark.parametrize("name", ["o_8x16_x_4", "o_16_x_8x8", "o_3x3_x_6", "o_6_x_3x3", "o_6_x_6", "o_4_x_6_x_8", "o_3x3_x_3x3", "o_m8_x_3x4"])
def test_outer_product_golden(rt, name):
    pc.check_outer_golden(rt, name)


@pytest.mark.parametrize("shape", [(3,), (6,), (3, 3), (6, 7), (5, 12), (3, 5, 7), (100, 37), (1000, 3), (33, 65, 9), (1, 5)])
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


@pytest.mark.parametrize("which,lh_kind", [("nonpow2", "gauss"), ("outer", "gauss"), ("nonpow2", "poisson"), ("outer", "poisson")])
def test_host_composed_fields_through_the_vi_drivers(rt, which, lh_kind):
    vc.check_host_composed_vi(rt, which, lh_kind)


def test_sample_consistency_and_constants_host_composed(rt):
    lh = vc._host_composed_pair(rt, "nonpow2", "gauss")[0]
    vc.check_sample_consistency(rt, "nonlinear_resample", ("cfax1loglogavgslope",), lh=lh)
    vc.check_constants_do_not_move(rt, ("cfax1fluctuations",), lh=lh)


def test_host_composed_wiener_filter_and_slq(rt):
    vc.check_host_composed_wiener_and_elbo(rt, "nonpow2")


@pytest.mark.parametrize("shape,nl", [((8, 8), "exp"), ((4, 8, 4), "identity"), ((6, 5), "exp")])
def test_gaussian_with_non_diagonal_covariance(rt, shape, nl):
    pc.check_operator_gaussian(rt, shape, nl)


def test_evidence_lower_bound_hybrid_slq_with_radau_bounds(rt):
    vc.check_elbo_hybrid(rt)


def test_likelihood_sum(rt):
    vc.check_likelihood_sum(rt)


def test_freeze_likelihood_partial(rt):
    vc.check_freeze(rt)
    vc.check_freeze(rt, "p2d_32x32", ("cfzeromode",))
