"""CPU tier: MGVI / geoVI host drivers on the host emulation of the kernels, against the oracle."""
import pytest

import nifty_b200 as nb
import vi_checks as vc
from nifty_b200._capi import CApi


@pytest.fixture(scope="module")
def rt():
    from emu.build_emu import build
    return nb.Runtime(CApi(build()), "cpu")


def test_draw_linear_residual(rt):
    vc.check_draw_linear_residual(rt)


def test_wiener_filter(rt):
    vc.check_wiener_filter(rt)
    vc.check_wiener_filter(rt, "g3d_8x8x8")
    vc.check_wiener_filter(rt, "m2d_16x8")


def test_minisanity(rt):
    vc.check_minisanity(rt)
    vc.check_minisanity(rt, "g3d_8x8x8")


def test_point_estimates_and_constants(rt):
    vc.check_point_estimates(rt)
    vc.check_point_estimates(rt, frozen=("cfxi",))
    vc.check_point_estimates(rt, "g3d_8x8x8", frozen=("cfzeromode", "cfax1loglogavgslope", "cfax1flexibility"))


def test_map_and_schedules(rt):
    vc.check_map_and_schedules(rt)


def test_evidence_lower_bound(rt):
    vc.check_elbo(rt, "g2d_8x32")


def test_nonlinear_update(rt):
    vc.check_nonlinear_update(rt)


def test_kl_value_grad_metric(rt):
    vc.check_kl(rt)
    vc.check_kl(rt, "p2d_32x32")


def test_optimize_kl_and_resume(rt, tmp_path):
    vc.check_optimize_kl(rt, tmp_path)


def test_final_kl_matches_oracle(rt):
    """configs[0] in miniature: final KL of one optimize_kl iteration within 1e-10 of the oracle (north star)."""
    assert vc.check_final_kl(rt, (32, 32), 4) <= 1e-10
    assert vc.check_final_kl(rt, (16, 64), 2, scaling=(3.0, 1.0), noise_cov_inv=0.01) <= 1e-10


def test_schedule_indices(rt):
    vc.check_schedule_indices(rt)


def test_kl_cg_on_device(rt):
    vc.check_kl_cg_on_device(rt)


def test_reduce_pieces(rt):
    vc.check_reduce_pieces(rt)
    vc.check_reduce_pieces(rt, "g3d_8x8x8")


@pytest.mark.parametrize("which,lh_kind", [("nonpow2", "gauss"), ("outer", "gauss"), ("nonpow2", "poisson"), ("outer", "poisson")])
def test_host_composed_fields_through_the_vi_drivers(rt, which, lh_kind):
    vc.check_host_composed_vi(rt, which, lh_kind)


@pytest.mark.parametrize("sample_mode,point_estimates", [("nonlinear_resample", ()), ("linear_resample", ()),
                                                         ("nonlinear_resample", ("cfax1fluctuations", "cfzeromode")),
                                                         ("linear_resample", ("cfax1spectrum",))])
def test_sample_consistency(rt, sample_mode, point_estimates):
    vc.check_sample_consistency(rt, sample_mode, point_estimates)


def test_sample_consistency_and_constants_host_composed(rt):
    lh = vc._host_composed_pair(rt, "nonpow2", "gauss")[0]
    vc.check_sample_consistency(rt, "nonlinear_resample", ("cfax1loglogavgslope",), lh=lh)
    vc.check_constants_do_not_move(rt, ("cfax1fluctuations",), lh=lh)


def test_constants_do_not_move(rt):
    vc.check_constants_do_not_move(rt)


def test_host_composed_wiener_filter_and_slq(rt):
    vc.check_host_composed_wiener_and_elbo(rt, "nonpow2")


def test_evidence_lower_bound_hybrid_slq_with_radau_bounds(rt):
    vc.check_elbo_hybrid(rt)


def test_likelihood_sum(rt):
    vc.check_likelihood_sum(rt)


def test_freeze_likelihood_partial(rt):
    vc.check_freeze(rt)
    vc.check_freeze(rt, "p2d_32x32", ("cfzeromode",))
