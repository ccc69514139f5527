"""CPU tier: pieces of the host-composed model families (nifty_b200/outer.py, bluestein.py) that need no device."""
import itertools

import numpy as np
import pytest
import torch

import nifty_b200 as nb
import oracle
from nifty_b200._capi import CApi


@pytest.fixture(scope="module")
def rt():
    from emu.build_emu import build
    return nb.Runtime(CApi(build()), "cpu")


@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_reflect_coefficients_identity(m):
    """prod_i cas(t_i) = sum_s c_s cas(sum_i s_i t_i): the identity behind the composite transform of an outer product."""
    coef = nb.OuterCorrelatedField._reflect_coefficients(m)
    cas = lambda x: np.cos(x) + np.sin(x)      # noqa: E731
    rng = np.random.default_rng(m)
    for _ in range(20):
        t = rng.uniform(-10, 10, size=m)
        lhs = np.prod(cas(t))
        rhs = sum(c * cas(float(np.dot(s, t))) for s, c in coef.items())
        assert abs(lhs - rhs) < 1e-12
    assert abs(sum(coef.values()) - 1.0) < 1e-15          # t = 0


@pytest.mark.parametrize("shape,dist", [((8,), 0.3), ((4, 16), (0.2, 0.7)), ((2, 4, 8), 0.5)])
def test_mode_tables_numpy_restatement_matches_the_plan(rt, shape, dist):
    """`fourier_mode_tables` (host NumPy, any extents) and the C++ tables of a power-of-two plan are the same tables."""
    a = nb.fourier_mode_tables(shape, dist)
    b = nb.bluestein.grid_tables(shape, dist, runtime=rt)
    assert np.array_equal(a["power_distributor"], b["power_distributor"])
    assert np.array_equal(a["mode_multiplicity"], b["mode_multiplicity"])
    np.testing.assert_allclose(a["mode_lengths"], b["mode_lengths"], rtol=0, atol=0)
    np.testing.assert_allclose(a["relative_log_mode_lengths"], b["relative_log_mode_lengths"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(a["log_volume"], b["log_volume"], rtol=0, atol=1e-14)
    assert a["total_volume"] == b["total_volume"]


@pytest.mark.parametrize("shape,dist", [((3,), 0.1), ((3, 3), 0.1), ((6, 10), (0.2, 0.3)), ((5, 3, 4), 0.4), ((1, 7), 1.0)])
def test_mode_tables_numpy_restatement_matches_the_oracle(shape, dist):
    a = nb.fourier_mode_tables(shape, dist)
    idx, um, cnt = oracle.fourier_mode_distributor(shape, dist)
    assert np.array_equal(a["power_distributor"], idx) and np.array_equal(a["mode_lengths"], um) and np.array_equal(a["mode_multiplicity"], cnt)
    g = oracle.make_fourier_grid(shape, dist)
    np.testing.assert_allclose(a["relative_log_mode_lengths"], g.relative_log_mode_lengths, rtol=0, atol=1e-15)
    np.testing.assert_allclose(a["log_volume"], g.log_volume, rtol=0, atol=1e-15)


def test_differentiable_model_call(rt):
    """`cf(p)` of a host-composed field is differentiable with torch autograd (the composite transform is a self-adjoint autograd
    function): the gradient of a scalar function of the field against the explicit cotangent formula."""
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    cfm.set_amplitude_total_offset(0.3, (0.2, 0.1))
    cfm.add_fluctuations((6,), 1.0, fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), prefix="a")
    cfm.add_fluctuations((3, 4), 0.5, fluctuations=(0.3, 0.2), loglogavgslope=(-1.5, 0.2), flexibility=(0.8, 0.3), asperity=None, prefix="b")
    cf = cfm.finalize()
    p = {k: v.clone().requires_grad_(True) for k, v in cf.init(4).items()}
    c = torch.as_tensor(np.random.default_rng(0).standard_normal(cf.target))
    (cf(p) * c).sum().backward()
    lh = nb.Gaussian(np.zeros(cf.target), noise_cov_inv=1.0).amend(cf)           # identity signal: left sqrt-metric = J^T
    want = lh.left_sqrt_metric({k: v.detach() for k, v in p.items()}, c)
    for k in p:
        np.testing.assert_allclose(p[k].grad.numpy(), want[k].numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("subs", [[((6, 10), 0.3)], [((8,), 0.5), ((3, 4), 0.25)]])
def test_float32_host_composed_metric(rt, subs):
    """float32 models of both host-composed families: the metric against the float64 oracle at single-precision accuracy."""
    kw = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    ocf = oracle.CorrelatedFieldOracle("cf")
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt, dtype=torch.float32)
    for m in (ocf, cfm):
        m.set_amplitude_total_offset(0.1, (0.2, 0.1))
        for i, (shp, d) in enumerate(subs):
            m.add_fluctuations(shp, d, prefix=f"s{i}", **kw)
    ocf.finalize()
    cf = cfm.finalize()
    shape = sum((tuple(s[0]) for s in subs), ())
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(0)
    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.5 * v for k, v in pos.items()}
    data = osig(pos) + 0.3 * rng.standard_normal(shape)
    olh = oracle.GaussianOracle(data, 1 / 0.09, osig)
    lh = nb.Gaussian(data, noise_cov_inv=1 / 0.09).amend(nb.SignalModel(cf, "exp"))
    got = lh.metric({k: torch.as_tensor(v) for k, v in pos.items()}, {k: torch.as_tensor(v) for k, v in tan.items()})
    want = olh.metric(pos, tan)
    scale = max(np.abs(v).max() for v in want.values())
    assert all(got[k].dtype == torch.float32 for k in got)
    assert max(np.abs(got[k].numpy() - want[k]).max() for k in want) / scale < 2e-5


@pytest.mark.parametrize("subs", [[((16, 8), 0.3)], [((6, 10), 0.3)], [((8,), 0.5), ((3, 4), 0.25)]])
def test_maker_accessors(rt, subs):
    """`CorrelatedFieldMaker.azm` / `amplitude_total_offset` / `fluctuations` (correlated_field.py:785-805) on the fused and on both
    host-composed model families against the oracle's amplitude models."""
    kw = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    ocf = oracle.CorrelatedFieldOracle("cf")
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    with pytest.raises(NotImplementedError):
        cfm.azm
    for m in (ocf, cfm):
        m.set_amplitude_total_offset(0.1, (0.2, 0.1))
        for i, (shp, d) in enumerate(subs):
            m.add_fluctuations(shp, d, prefix=f"s{i}", **kw)
    ocf.finalize()
    cfm.finalize()
    lay = oracle.Layout(ocf.domain)
    pos = lay.random(np.random.default_rng(1))
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    assert len(cfm.fluctuations) == len(subs)
    for f, oa in zip(cfm.fluctuations, ocf.amps):
        np.testing.assert_allclose(f(tp).numpy(), oa(pos), rtol=1e-12)
    a, b = nb.lognormal_moments(0.2, 0.1) if hasattr(nb, "lognormal_moments") else (None, None)
    z = float(cfm.amplitude_total_offset(tp))
    assert z > 0 and abs(z - float(cfm.azm(tp))) == 0.0


def test_vmodel_over_host_composed_field(rt):
    """`VModel(cf, n, in_axes="cfxi")` (model.py:370-417; test_empirical_power_spectrum.py:39) on a non-power-of-two field: the mapped
    model equals the per-sample evaluations."""
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    cfm.set_amplitude_total_offset(0.3, (0.2, 0.1))
    cfm.add_fluctuations((6, 5), 0.5, fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), prefix="a")
    cf = cfm.finalize()
    vm = nb.VModel(cf, 3, in_axes="cfxi")
    x = vm.init(11)
    assert tuple(x["cfxi"].shape) == (3, 6, 5) and tuple(vm.target) == (3, 6, 5)
    out = vm(x)
    for i in range(3):
        xi = dict(x)
        xi["cfxi"] = x["cfxi"][i]
        assert torch.equal(out[i], cf(xi))


def test_vmodel_axes(rt):
    """Sample axes other than 0 (model.py:370-417 `in_axes` / `out_axes`): mapped leaf with its sample axis in the middle, output axis last."""
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    cfm.set_amplitude_total_offset(0.3, (0.2, 0.1))
    cfm.add_fluctuations((8, 4), 0.5, fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), prefix="a")
    cf = cfm.finalize()
    vm = nb.VModel(cf, 3, in_axes={"cfxi": 1}, out_axes=-1)
    assert vm.domain["cfxi"] == (8, 3, 4) and vm.domain["cfzeromode"] == () and tuple(vm.target) == (8, 4, 3)
    x = vm.init(2)
    out = vm(x)
    assert tuple(out.shape) == (8, 4, 3)
    for i in range(3):
        xi = dict(x)
        xi["cfxi"] = x["cfxi"][:, i]
        assert torch.equal(out[..., i], cf(xi))
    with pytest.raises(ValueError):
        nb.VModel(cf, 3, in_axes=1)                # scalar leaves have no axis 1


def test_likelihood_surface(rt):
    """`init`, `lsm_tangents_shape`, `rsm_tangents_shape` of amended likelihoods (likelihood.py:352-376, 546-598), also on a sum."""
    import vi_checks as vc
    c, g, lh, olh, lay = vc._setup(rt, "g2d_16x16")
    p = lh.init(3)
    assert set(p) == set(lh.domain) and lh.lsm_tangents_shape == (16, 16) == lh.left_sqrt_metric_tangents_shape
    assert lh.rsm_tangents_shape == lh.domain == lh.right_sqrt_metric_tangents_shape
    s = lh + lh
    assert s.lsm_tangents_shape == (512,) and abs(s.energy(s.init(3)) - 2.0 * lh.energy(p)) <= 1e-12 * abs(lh.energy(p))


@pytest.mark.parametrize("shape,axes", [((8, 16), None), ((6, 5), None), ((4, 8, 3), (0, 1)), ((3, 4, 5, 2), (1, 2, 3)), ((7, 4), (-1,)), ((2, 3, 4, 4), (2, 3))])
def test_module_level_hartley(rt, shape, axes):
    """`hartley(p, axes)` (correlated_field.py:24-30) on the device for any extents and axis subsets, both sign conventions."""
    x = np.random.default_rng(1).standard_normal(shape)
    for conv in ("non_canonical_hartley", "canonical_hartley"):
        got = nb.hartley(torch.as_tensor(x), axes, hartley_convention=conv, runtime=rt)
        want = oracle.hartley(x, axes=axes, convention=conv)
        assert got.shape == x.shape and np.max(np.abs(got.numpy() - want)) < 1e-12 * np.max(np.abs(want))
    with pytest.raises(NotImplementedError):
        nb.hartley(torch.zeros((2, 2, 2, 2)), runtime=rt)


@pytest.mark.parametrize("shape,dist", [((16,), 0.1), ((3, 3), 0.1), ((8, 32), (0.3, 0.11)), ((5, 3, 4), 0.4)])
def test_grid_functions_with_the_reference_names(shape, dist):
    """`get_fourier_mode_distributor` / `make_grid` (correlated_field.py:134-176, 238-265) against the oracle."""
    idx, um, cnt = nb.get_fourier_mode_distributor(shape, dist)
    oidx, oum, ocnt = oracle.fourier_mode_distributor(shape, dist)
    assert np.array_equal(idx, oidx) and np.array_equal(um, oum) and np.array_equal(cnt, ocnt)
    g, og = nb.make_grid(shape, dist), oracle.make_fourier_grid(shape, dist)
    assert tuple(g.shape) == tuple(shape) and g.total_volume == og.total_volume
    np.testing.assert_allclose(g.harmonic_grid.relative_log_mode_lengths, og.relative_log_mode_lengths, rtol=0, atol=1e-15)
    np.testing.assert_allclose(g.harmonic_grid.log_volume, og.log_volume, rtol=0, atol=1e-15)
    with pytest.raises(NotImplementedError):
        nb.make_grid((4,), None, "spherical")
    with pytest.raises(ValueError):
        nb.make_grid((4,), 1.0, "cubic")


def test_sum_of_a_fused_and_a_host_composed_likelihood(rt):
    """`lh_a + lh_b` with a summand on the fused device path and one on a non-power-of-two (host-composed) field: energies and metrics add
    on the union of the domains, and the standard driver runs on the sum."""
    import vi_checks as vc
    c, g, lh_a, olh, lay = vc._setup(rt, "g2d_16x16")
    cfm = nb.CorrelatedFieldMaker("other", runtime=rt)
    cfm.set_amplitude_total_offset(0.5, (0.2, 0.1))
    cfm.add_fluctuations((6, 5), 0.3, (0.3, 0.1), (-1.5, 0.3), (1.0, 0.5), None, prefix="ax")
    lh_h = nb.Poissonian(np.random.default_rng(0).poisson(2.0, size=(6, 5))).amend(nb.SignalModel(cfm.finalize(), "exp"))
    both = lh_a + lh_h
    pos = {k: 0.3 * v for k, v in both.init(3).items()}
    t = both.init(4)
    sub = lambda lh, tree: {k: tree[k] for k in lh.domain}          # noqa: E731
    assert abs(both.energy(pos) - (lh_a.energy(sub(lh_a, pos)) + lh_h.energy(sub(lh_h, pos)))) <= 1e-12 * abs(both.energy(pos))
    m, want = both.metric(pos, t), {**lh_a.metric(sub(lh_a, pos), sub(lh_a, t)), **lh_h.metric(sub(lh_h, pos), sub(lh_h, t))}
    assert all(torch.allclose(m[k], want[k], rtol=1e-12, atol=1e-14) for k in m)
    s, st = nb.optimize_kl(both, both.signal.as_flat(pos), key=1, n_total_iterations=1, n_samples=1, sample_mode="linear_resample",
                           draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=100)),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))))
    assert st.nit == 1 and len(s) == 2
