"""Parity checks shared by the CPU tier (host emulation of the kernel sources, tests/emu) and the
GPU tier (libniftyb200.so through the C ABI).  Every check compares the product path with the
committed nifty.cl fixtures and / or the NumPy oracle on identical inputs.

Tolerances: float64 1e-10 relative (north star), float32 1e-5, relative to the largest entry of
the compared quantity (leaves that are structurally ~0 only carry cancellation noise).
"""
import numpy as np
import pytest
import torch

import nifty_b200 as nb
import oracle
from golden_util import CASES, build_oracle, build_oracle_lh, load, rel_err

TOL = {torch.float64: 1e-10, torch.float32: 2e-5}
POW2_CASES = [n for n in sorted(CASES) if all((s & (s - 1)) == 0 for s in CASES[n]["shape"])]


def build_product(c, rt, dtype=torch.float64, kind="power", convention="non_canonical_hartley"):
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt, dtype=dtype, hartley_convention=convention)
    cfm.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    if "matern" in c:
        cfm.add_fluctuations_matern(c["shape"], c["distances"], renormalize_amplitude=c.get("renorm", False), prefix="ax1",
                                    non_parametric_kind=c.get("kind", "amplitude"), **c["matern"])
    else:
        cfm.add_fluctuations(c["shape"], c["distances"], c["fluctuations"], c["loglogavgslope"], c["flexibility"],
                             c["asperity"], prefix="ax1", non_parametric_kind=kind)
    return cfm.finalize()


def build_product_lh(c, g, rt, dtype=torch.float64, scaling=None):
    cf = build_product(c, rt, dtype)
    sig = nb.SignalModel(cf, "exp", scaling=scaling)
    if c["lh"] == "gauss":
        return nb.Gaussian(g["data"], noise_cov_inv=float(g["noise_cov_inv"])).amend(sig)
    return nb.Poissonian(g["data"]).amend(sig)


def tree_err(got, want):
    scale = max(float(np.max(np.abs(v))) for v in want.values())
    assert set(got) == set(want)
    return max(float(np.max(np.abs(got[k].detach().cpu().numpy().astype(np.float64) - want[k]))) / scale for k in want)


def t2n(x):
    return x.detach().cpu().numpy().astype(np.float64)


def check_mode_tables(rt, shape, distances):
    plan = nb.Plan(shape, distances, runtime=rt)
    idx, um, cnt = oracle.fourier_mode_distributor(shape, distances)
    assert plan.K == um.size
    assert np.array_equal(plan.power_distributor, idx)
    assert np.array_equal(plan.mode_lengths, um)
    assert np.array_equal(plan.mode_multiplicity, cnt)
    grid = oracle.make_fourier_grid(shape, distances)
    # libm vs numpy log differ in the last ulp
    np.testing.assert_allclose(plan.relative_log_mode_lengths, grid.relative_log_mode_lengths, rtol=0, atol=1e-14)
    np.testing.assert_allclose(plan.log_volume, grid.log_volume, rtol=0, atol=1e-14)
    assert plan.total_volume == grid.total_volume


def check_hartley(rt, shape, dtype=torch.float64, convention="non_canonical_hartley", seed=0):
    plan = nb.Plan(shape, 1.0, dtype=dtype, hartley_convention=convention, runtime=rt)
    x = np.random.default_rng(seed).standard_normal(shape)
    out = t2n(plan.hartley(torch.as_tensor(x)))
    ref = oracle.hartley(x, convention=convention)
    assert rel_err(out, ref) < TOL[dtype], (shape, rel_err(out, ref))
    # H H = N (self-inverse up to N) -- size independent property
    if dtype == torch.float64:
        back = t2n(plan.hartley(plan.hartley(torch.as_tensor(x)))) / x.size
        assert rel_err(back, x) < 1e-12


def check_bilinear(rt, shape, distances, dtype=torch.float64, seed=1):
    plan = nb.Plan(shape, distances, dtype=dtype, runtime=rt)
    idx, um, _ = oracle.fourier_mode_distributor(shape, distances)
    rng = np.random.default_rng(seed)
    amp, xi, cot = rng.standard_normal(um.size) ** 2, rng.standard_normal(shape), rng.standard_normal(shape)
    V = plan.total_volume
    out = t2n(plan.cf_apply(torch.as_tensor(amp), torch.as_tensor(xi), 0.3))
    ref = 0.3 + oracle.hartley(amp[idx] * xi) / V
    assert rel_err(out, ref) < TOL[dtype]
    xb, ab = plan.cf_apply_adjoint(torch.as_tensor(amp), torch.as_tensor(xi), torch.as_tensor(cot))
    g = oracle.hartley(cot) / V
    assert rel_err(t2n(xb), amp[idx] * g) < TOL[dtype]
    assert rel_err(t2n(ab), np.bincount(idx.ravel(), weights=(xi * g).ravel(), minlength=um.size)) < 10 * TOL[dtype]
    xb2, none = plan.cf_apply_adjoint(torch.as_tensor(amp), None, torch.as_tensor(cot))
    assert none is None and rel_err(t2n(xb2), amp[idx] * g) < TOL[dtype]
    # batched forms (the vmap rules of a jax.ffi binding): identical to looped single calls, with the amplitude table shared
    # by the batch (`VModel(cf, in_axes="xi")`) or batched with it (the sample axis of optimize_kl.py:106,135)
    B = 3
    amps = rng.standard_normal((B, um.size)) ** 2
    xis, cots = rng.standard_normal((B,) + tuple(shape)), rng.standard_normal((B,) + tuple(shape))
    for a_in in (amps, amps[1]):
        ob = plan.cf_apply_batch(torch.as_tensor(a_in), torch.as_tensor(xis), 0.3)
        xbb, abb = plan.cf_apply_adjoint_batch(torch.as_tensor(a_in), torch.as_tensor(xis), torch.as_tensor(cots))
        xbb2, nb2 = plan.cf_apply_adjoint_batch(torch.as_tensor(a_in), None, torch.as_tensor(cots))
        assert nb2 is None
        for i in range(B):
            a_i = torch.as_tensor(a_in[i] if a_in.ndim == 2 else a_in)
            assert torch.equal(ob[i], plan.cf_apply(a_i, torch.as_tensor(xis[i]), 0.3))
            xs, as_ = plan.cf_apply_adjoint(a_i, torch.as_tensor(xis[i]), torch.as_tensor(cots[i]))
            assert torch.equal(xbb[i], xs) and torch.equal(abb[i], as_) and torch.equal(xbb2[i], xs)


def check_golden(rt, name, dtype=torch.float64):
    """field / signal / energy / gradient / metric against the nifty.cl fixtures; sqrt-metrics,
    transformation and residual against the oracle."""
    c, g = CASES[name], load(name)
    tol = TOL[dtype]
    lh = build_product_lh(c, g, rt, dtype)
    pos = {k: torch.as_tensor(v) for k, v in g["pos"].items()}
    tan = {k: torch.as_tensor(v) for k, v in g["tan"].items()}
    assert rel_err(t2n(lh.signal.cf(pos)), g["field"]) < tol
    assert rel_err(t2n(lh.signal_response(pos)), g["signal"]) < tol
    e, grad = lh.energy_and_gradient(pos)
    assert abs(e - float(g["energy"])) <= tol * abs(float(g["energy"]))
    assert tree_err(grad, g["grad"]) < tol
    assert tree_err(lh.metric(pos, tan), g["metric"]) < tol
    olh = build_oracle_lh(c, g)
    assert rel_err(t2n(lh.right_sqrt_metric(pos, tan)), olh.right_sqrt_metric(g["pos"], g["tan"])) < tol
    assert tree_err(lh.left_sqrt_metric(pos, g["cot"]), olh.left_sqrt_metric(g["pos"], g["cot"])) < tol
    assert rel_err(t2n(lh.transformation(pos)), olh.transformation(g["pos"])) < tol
    assert rel_err(t2n(lh.normalized_residual(pos)), olh.normalized_residual(g["pos"])) < 100 * tol
    # field JVP / VJP of the bare correlated field (fixtures field_jvp / field_vjp)
    lin, _ = lh.lin_at(pos)
    assert rel_err(t2n(lin.rsm(lh.signal.as_flat(tan), scaled=False)), g["field_jvp"]) < tol
    assert tree_err(lh.layout.unpack(lin.lsm(torch.as_tensor(g["cot"]), scaled=False)), g["field_vjp"]) < tol


def check_model_surface(rt, name="g2d_16x16"):
    """The `jft.Model` surface of the finalised field (correlated_field.py:807-845, 913-920): amplitude / power spectrum /
    normalized amplitudes against the oracle, `target_grids` against the oracle's grid tables, `domain` / `target` / `init`."""
    from golden_util import build_oracle
    c, g = CASES[name], load(name)
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    cfm.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    with pytest.raises(ValueError, match="finalize"):
        cfm.amplitude
    cfm.add_fluctuations(c["shape"], c["distances"], c["fluctuations"], c["loglogavgslope"], c["flexibility"], c["asperity"],
                         prefix="ax1", non_parametric_kind="power")
    cf = cfm.finalize()
    ocf = build_oracle(c)
    pos = {k: torch.as_tensor(v) for k, v in g["pos"].items()}
    assert rel_err(t2n(cfm.amplitude(pos)), ocf.amplitude(g["pos"])) < 1e-12
    assert rel_err(t2n(cfm.power_spectrum(pos)), ocf.amplitude(g["pos"]) ** 2) < 1e-12
    (na,), (ona,) = cfm.get_normalized_amplitudes(), ocf.normalized_amplitudes(g["pos"])
    assert cf.normalized_amplitudes[0] is not None and rel_err(t2n(na(pos)), ona) < 1e-12
    (tg,), og = cf.target_grids, ocf.grids[0]
    assert tuple(tg.shape) == tuple(c["shape"]) and abs(tg.total_volume - og.total_volume) <= 1e-14 * og.total_volume
    hg, ohg = tg.harmonic_grid, og          # the oracle keeps both grids in one record
    assert np.array_equal(hg.power_distributor, ohg.power_distributor) and np.array_equal(hg.mode_multiplicity, ohg.mode_multiplicity)
    np.testing.assert_allclose(hg.mode_lengths, ohg.mode_lengths, rtol=0, atol=0)
    np.testing.assert_allclose(hg.relative_log_mode_lengths, ohg.relative_log_mode_lengths, rtol=0, atol=1e-14)
    np.testing.assert_allclose(hg.log_volume, ohg.log_volume, rtol=0, atol=1e-14)
    assert cf.domain == {k: tuple(np.shape(v)) for k, v in g["pos"].items()} and tuple(cf.target) == tuple(c["shape"])
    p0 = cf.init(3)
    assert sorted(p0) == sorted(cf.domain) and all(tuple(p0[k].shape) == cf.domain[k] for k in p0)
    # the container types of the reference's model API (model.py:32-340, tree_math/vector.py:79-188)
    assert isinstance(cf, nb.LazyModel) and isinstance(nb.SignalModel(cf, "exp"), nb.LazyModel)
    lh = nb.Gaussian(g["data"], noise_cov_inv=float(g["noise_cov_inv"])).amend(nb.Model.pointwise(cf, "exp"))
    tan = {k: torch.as_tensor(v) for k, v in g["tan"].items()}
    mv = lh.metric(nb.Vector(pos), nb.Vector(tan))            # Vector in -> Vector out, same numbers as with plain trees
    assert isinstance(mv, nb.Vector) and tree_err(mv.tree, g["metric"]) < 1e-10
    assert rel_err(t2n(cf(nb.Vector(pos))), g["field"]) < 1e-10
    # VModel (model.py:370-417): only the excitations mapped -> one batched device call with a shared amplitude table
    # (test/test_re/test_empirical_power_spectrum.py:39); every leaf mapped -> the in-order loop
    xi_key = "cfxi"
    rng = np.random.default_rng(3)
    xis = torch.as_tensor(rng.standard_normal((3,) + tuple(c["shape"])))
    vm = nb.VModel(cf, 3, in_axes=xi_key)
    assert vm.domain[xi_key] == (3,) + tuple(c["shape"]) and vm.domain["cfzeromode"] == cf.domain["cfzeromode"]
    out = vm({**pos, xi_key: xis})
    for i in range(3):
        assert rel_err(t2n(out[i]), t2n(cf({**pos, xi_key: xis[i]}))) < 1e-12
    allmapped = nb.VModel(cf, 2, in_axes=0)
    batch = {k: torch.stack([v, 0.5 * v]) for k, v in pos.items()}
    out2 = allmapped(batch)
    assert rel_err(t2n(out2[1]), t2n(cf({k: 0.5 * v for k, v in pos.items()}))) < 1e-12
    p_init = vm.init(5)
    assert tuple(p_init[xi_key].shape) == (3,) + tuple(c["shape"]) and tuple(p_init["cfzeromode"].shape) == tuple(cf.domain["cfzeromode"])


def check_kind_and_scaling(rt, name="g2d_16x16", kind="amplitude", scaling=(3.0, 1.0)):
    """'amplitude' kind, the multiplicative log-normal scaling leaf and the canonical convention
    (not covered by the nifty.cl fixtures) against the oracle."""
    c, g = CASES[name], load(name)
    for conv in ("non_canonical_hartley", "canonical_hartley"):
        ocf = oracle.CorrelatedFieldOracle("cf", hartley_convention=conv)
        ocf.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
        ocf.add_fluctuations(c["shape"], c["distances"], c["fluctuations"], c["loglogavgslope"], c["flexibility"],
                             c["asperity"], prefix="ax1", non_parametric_kind=kind)
        ocf.finalize()
        osig = oracle.SignalOracle(ocf, "exp", scaling=scaling)
        olh = oracle.GaussianOracle(g["data"], float(g["noise_cov_inv"]), osig)
        cf = build_product(c, rt, kind=kind, convention=conv)
        lh = nb.Gaussian(g["data"], noise_cov_inv=float(g["noise_cov_inv"])).amend(nb.SignalModel(cf, "exp", scaling=scaling))
        lay = oracle.Layout(olh.domain)
        rng = np.random.default_rng(3)
        pos, tan = lay.random(rng), lay.random(rng)
        tp = {k: torch.as_tensor(v) for k, v in pos.items()}
        tt = {k: torch.as_tensor(v) for k, v in tan.items()}
        assert rel_err(t2n(lh.signal_response(tp)), osig(pos)) < 1e-10
        e, grad = lh.energy_and_gradient(tp)
        oe, ograd = olh.energy_and_gradient(pos)
        assert abs(e - oe) <= 1e-10 * abs(oe)
        assert tree_err(grad, ograd) < 1e-10
        assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < 1e-10
        assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)), olh.right_sqrt_metric(pos, tan)) < 1e-10
        u = rng.standard_normal(c["shape"])
        assert tree_err(lh.left_sqrt_metric(tp, u), olh.left_sqrt_metric(pos, u)) < 1e-10


def check_metric_properties(rt, shape, distances, seed=0, dtype=torch.float64, lh_kind="gauss"):
    """Size-independent properties at sizes the oracle cannot reach: symmetry <u, M t> = <t, M u>,
    positivity, metric == LSM(RSM(.)) (likelihood.py:263-282) and linearity."""
    c = dict(shape=shape, distances=distances, offset_mean=0.0, offset_std=(1e-3, 1e-4), fluctuations=(1e-1, 5e-3),
             loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh=lh_kind)
    cf = build_product(c, rt, dtype)
    sig = nb.SignalModel(cf, "exp")
    rng = np.random.default_rng(seed)
    pos = sig.layout.random(rng, dtype, rt.device)
    if lh_kind == "gauss":
        data = torch.as_tensor(rng.standard_normal(shape))
        lh = nb.Gaussian(data, noise_cov_inv=100.0).amend(sig)
    else:
        lh = nb.Poissonian(rng.poisson(1.0, size=shape)).amend(sig)
    t, u = sig.layout.random(rng, dtype, rt.device), sig.layout.random(rng, dtype, rt.device)
    lin, _ = lh.lin_at(pos)
    Mt, Mu = lin.metric(t), lin.metric(u)
    a, b = float(torch.dot(u, Mt)), float(torch.dot(t, Mu))
    tol = 1e-10 if dtype == torch.float64 else 1e-3
    assert abs(a - b) <= tol * max(abs(a), abs(b)), (a, b)
    assert float(torch.dot(t, Mt)) > 0
    M2 = lin.lsm(lin.rsm(t, scaled=True), scaled=True)
    assert float((M2 - Mt).abs().max()) <= tol * float(Mt.abs().max())
    lin_comb = lin.metric(2.0 * t - 3.0 * u)
    assert float((lin_comb - (2.0 * Mt - 3.0 * Mu)).abs().max()) <= tol * float(lin_comb.abs().max())
    if dtype == torch.float64 and int(np.prod(shape)) <= (1 << 20):
        # the Jacobian of `transformation` is the right sqrt-metric (test/test_re/test_likelihood_impl.py:288-322, 435-470), by a
        # central difference along t
        rs = lin.rsm(t, scaled=True).clone()
        h = 1e-5 / max(1.0, float(t.abs().max()))
        la, lb = lh.new_lin(), lh.new_lin()
        la.update(pos + h * t)
        lb.update(pos - h * t)
        fd = (la.transformation() - lb.transformation()) / (2.0 * h)
        assert float((fd - rs).abs().max()) <= 1e-6 * float(rs.abs().max())
    return lh, lin, pos


def check_cg(rt):
    """Device CG against the oracle restatement of `_cg`: identical iteration counts, info and nfev,
    solutions to 1e-10 on a well-conditioned case; the N_RESET branch (iteration 20) on a case that
    is not yet converged there (rounding differences grow with the iteration count, hence 1e-6)."""
    runs = [("g2d_16x16", dict(absdelta=1e-6, maxiter=100), 1e-10), ("g2d_16x16", dict(), 1e-10),
            ("g2d_16x16", dict(resnorm=1e-3, norm_ord=1, maxiter=50), 1e-10),
            ("g2d_16x16", dict(absdelta=1e-8, maxiter=45), 1e-10),
            ("g2d_16x16", dict(absdelta=1e-30, maxiter=7, miniter=7), 1e-10),
            ("g3d_4x8x16", dict(absdelta=1e-30, maxiter=23, miniter=24), 1e-6),
            ("g3d_4x8x16", dict(resnorm=10.0, norm_ord=np.inf, maxiter=60), 1e-7)]
    cache = {}
    for name, kw, tol in runs:
        if name not in cache:
            c, g = CASES[name], load(name)
            lh = build_product_lh(c, g, rt)
            olh = build_oracle_lh(c, g)
            lay = oracle.Layout(olh.domain)
            rng = np.random.default_rng(5)
            j, x0 = rng.standard_normal(lay.size), rng.standard_normal(lay.size)
            lin, _ = lh.lin_at(torch.as_tensor(lay.pack(g["pos"])))
            cache[name] = (g, lh, olh, lay, j, x0, lin)
        g, lh, olh, lay, j, x0, lin = cache[name]

        def mat(v):
            return lay.pack(olh.metric(g["pos"], lay.unpack(v))) + v

        tj, tx0 = rt.asarray(j, torch.float64), rt.asarray(x0, torch.float64)
        for use_x0 in (True, False):
            ores = oracle.cg(mat, j, x0=x0 if use_x0 else None, **kw)
            x, res = lin.cg_solve(tj, tx0 if use_x0 else None, check_every=3, **kw)
            assert (res.nit, res.info, res.nfev) == (ores.nit, ores.info, ores.nfev), (name, kw, use_x0)
            assert rel_err(t2n(x), ores.x) < tol, (name, kw, use_x0)


def check_against_oracle(rt, shape, distances, lh_kind="gauss", seed=11, tol=1e-10, **cf_kw):
    """Energy / gradient / metric / sqrt-metrics against the oracle on a grid without fixture
    (used for grids with several scan chunks, K > 2048, and other shapes)."""
    c = dict(shape=shape, distances=distances, offset_mean=0.3, offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1),
             loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh=lh_kind)
    c.update(cf_kw)
    ocf = build_oracle(c)
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.5 * v for k, v in pos.items()}
    if lh_kind == "gauss":
        data = osig(pos) + 0.3 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
        g = dict(data=data, noise_cov_inv=1.0 / 0.09)
    else:
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        g = dict(data=data)
    lh = build_product_lh(c, g, rt)
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    tt = {k: torch.as_tensor(v) for k, v in tan.items()}
    e, grad = lh.energy_and_gradient(tp)
    oe, ograd = olh.energy_and_gradient(pos)
    assert abs(e - oe) <= tol * abs(oe)
    assert tree_err(grad, ograd) < tol
    assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol
    u = rng.standard_normal(shape)
    assert tree_err(lh.left_sqrt_metric(tp, u), olh.left_sqrt_metric(pos, u)) < tol
    assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)), olh.right_sqrt_metric(pos, tan)) < tol


NONLINEARITIES = {
    # name: (torch map, NumPy map, NumPy derivative)
    "softplus": (torch.nn.functional.softplus, lambda f: np.logaddexp(0.0, f), lambda f: 1.0 / (1.0 + np.exp(-f))),
    "sigmoid_scaled": (lambda f: 3.0 * torch.sigmoid(f) + 0.1, lambda f: 3.0 / (1.0 + np.exp(-f)) + 0.1,
                       lambda f: 3.0 * np.exp(-f) / (1.0 + np.exp(-f)) ** 2),
    "square_plus": (lambda f: f * f + 0.5, lambda f: f * f + 0.5, lambda f: 2.0 * f),
}


def check_custom_nonlinearity(rt, shape, distances, lh_kind="gauss", which="softplus", seed=5, tol=1e-10):
    """`SignalModel(cf, nonlinearity=<torch callable>)` -- an arbitrary pointwise map of the field, f and f' evaluated on the
    host side at every linearisation and handed over as tables (nb200_lin_set_pointwise) -- against the oracle with the
    same map and its analytic derivative: signal, energy, gradient, metric, both sqrt-metrics, a CG solve on the metric."""
    tfn, nfn, ndfn = NONLINEARITIES[which]
    c = dict(shape=shape, distances=distances, offset_mean=0.3, offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1),
             loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh=lh_kind)
    ocf = build_oracle(c)
    osig = oracle.SignalOracle(ocf, (nfn, ndfn))
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.5 * v for k, v in pos.items()}
    sig = nb.SignalModel(build_product(c, rt), tfn)
    if lh_kind == "gauss":
        data = osig(pos) + 0.3 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
        lh = nb.Gaussian(data, noise_cov_inv=1.0 / 0.09).amend(sig)
    else:
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        lh = nb.Poissonian(data).amend(sig)
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    tt = {k: torch.as_tensor(v) for k, v in tan.items()}
    assert rel_err(t2n(lh.signal_response(tp)), osig(pos)) < tol
    e, grad = lh.energy_and_gradient(tp)
    oe, ograd = olh.energy_and_gradient(pos)
    assert abs(e - oe) <= tol * abs(oe)
    assert tree_err(grad, ograd) < tol
    assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol
    u = rng.standard_normal(shape)
    assert tree_err(lh.left_sqrt_metric(tp, u), olh.left_sqrt_metric(pos, u)) < tol
    assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)), olh.right_sqrt_metric(pos, tan)) < tol
    lin, _ = lh.lin_at(tp)
    j = lh.signal.layout.random(3, torch.float64, rt.device)
    # (fixed iteration count, away from any stopping threshold: two CG implementations drift apart with the iteration count)
    x, res = lin.cg_solve(j, None, absdelta=1e-30, miniter=6, maxiter=6)
    ores = oracle.cg(lambda v: lay.pack(olh.metric(pos, lay.unpack(v))) + v, t2n(j), absdelta=1e-30, miniter=6, maxiter=6)
    assert res.nit == ores.nit == 6 and rel_err(t2n(x), ores.x) < 1e-6
    import pytest
    with pytest.raises(ValueError):
        nb.SignalModel(build_product(c, rt), tfn, scaling=(3.0, 1.0))
    bad = nb.Gaussian(data.astype(np.float64), noise_cov_inv=1.0).amend(nb.SignalModel(build_product(c, rt), lambda f: f.sum()))
    with pytest.raises(ValueError, match="pointwise"):
        bad.energy(tp)


def check_matern_variants(rt):
    """Matern amplitude with renormalisation and both kinds (no nifty.cl fixture: cl has no such switch) vs the oracle."""
    for kind in ("amplitude", "power"):
        for renorm in (True, False):
            c = dict(shape=(16, 16), distances=0.2, offset_mean=0.1, offset_std=(0.1, 0.1), lh="gauss", kind=kind, renorm=renorm,
                     matern=dict(scale=(1.0, 0.5), cutoff=(0.8, 0.3), loglogslope=(-3.0, 0.5)))
            ocf = oracle.CorrelatedFieldOracle("cf")
            ocf.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
            ocf.add_fluctuations_matern(c["shape"], c["distances"], renormalize_amplitude=renorm, prefix="ax1",
                                        non_parametric_kind=kind, **c["matern"])
            ocf.finalize()
            osig = oracle.SignalOracle(ocf, "exp")
            lay = oracle.Layout(osig.domain)
            rng = np.random.default_rng(17)
            pos, tan = lay.random(rng), lay.random(rng)
            pos = {k: 0.5 * v for k, v in pos.items()}
            data = osig(pos) + 0.3 * rng.standard_normal(c["shape"])
            olh = oracle.GaussianOracle(data, 4.0, osig)
            lh = build_product_lh(c, dict(data=data, noise_cov_inv=4.0), rt)
            tp = {k: torch.as_tensor(v) for k, v in pos.items()}
            tt = {k: torch.as_tensor(v) for k, v in tan.items()}
            e, grad = lh.energy_and_gradient(tp)
            oe, ograd = olh.energy_and_gradient(pos)
            assert abs(e - oe) <= 1e-10 * abs(oe), (kind, renorm)
            assert tree_err(grad, ograd) < 1e-10, (kind, renorm)
            assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < 1e-10, (kind, renorm)


FULL_SIZE = {
    # BASELINE.json configs at FULL size, hyper-parameters of SURVEY.md 8(d)
    "cfg2_4096_gauss": dict(shape=(4096, 4096), distances=1.0 / 4096, offset_mean=0.0, offset_std=(1e-3, 1e-4), fluctuations=(1e-1, 5e-3),
                            loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh="gauss"),
    "cfg3_256c_gauss": dict(shape=(256, 256, 256), distances=1.0 / 256, offset_mean=0.0, offset_std=(1e-3, 1e-4), fluctuations=(1e-1, 5e-3),
                            loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh="gauss"),
    # misc/re/paper/minimal_benchmark.py:94-105 (the geoVI / Poisson configuration)
    "cfg4_2048_poisson": dict(shape=(2048, 2048), distances=1.0 / 2048, offset_mean=2.0, offset_std=(0.1, 0.03), fluctuations=(1.0, 0.5),
                              loglogavgslope=(-3.0, 0.2), flexibility=(1.0, 0.2), asperity=(0.5, 0.05), lh="poisson"),
    # CPU-tier stand-ins of the same three configurations (the same code path of this check at oracle-friendly sizes)
    "small_gauss": dict(shape=(64, 128), distances=1.0 / 64, offset_mean=0.0, offset_std=(1e-3, 1e-4), fluctuations=(1e-1, 5e-3),
                        loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh="gauss"),
    "small_poisson": dict(shape=(64, 64), distances=1.0 / 64, offset_mean=2.0, offset_std=(0.1, 0.03), fluctuations=(1.0, 0.5),
                          loglogavgslope=(-3.0, 0.2), flexibility=(1.0, 0.2), asperity=(0.5, 0.05), lh="poisson"),
}


def check_config_vs_oracle(rt, name, tol=1e-10):
    """BASELINE.json's configurations at their FULL size against the oracle on the same seeded inputs: energy, gradient,
    metric-vector product and metric + 1 (the operator CG runs on) at 1e-10 relative -- the north-star parity bar."""
    import time
    c = FULL_SIZE[name]
    shape = c["shape"]
    t0 = time.time()
    ocf = build_oracle(c)
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(42)
    truth = lay.random(rng)
    if c["lh"] == "gauss":
        data = osig(truth) + 0.1 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 100.0, osig)
        g = dict(data=data, noise_cov_inv=100.0)
    else:
        data = rng.poisson(osig(truth)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        g = dict(data=data)
    pos = {k: 0.1 * v for k, v in lay.random(np.random.default_rng(44)).items()}
    tan = lay.random(np.random.default_rng(45))
    lh = build_product_lh(c, g, rt)
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    tt = {k: torch.as_tensor(v) for k, v in tan.items()}
    e, grad = lh.energy_and_gradient(tp)
    oe, ograd = olh.energy_and_gradient(pos)
    errs = {"energy": abs(e - oe) / abs(oe), "gradient": tree_err(grad, ograd)}
    om = olh.metric(pos, tan)
    errs["metric"] = tree_err(lh.metric(tp, tt), om)
    lin, _ = lh.lin_at(lh.signal.as_flat(tp))
    m1 = lh.layout.unpack(lin.metric(lh.signal.as_flat(tt), add_identity=True))
    errs["metric_plus_1"] = tree_err(m1, {k: om[k] + tan[k] for k in om})
    for k, v in errs.items():
        assert v < tol, (name, k, v)
    return errs, time.time() - t0


def check_outer_product(rt, shapes=((8, 16), (4,)), distances=(0.2, 0.5), lh_kind="gauss", conv="non_canonical_hartley", seed=21,
                        tol=1e-10):
    """Correlated field on the outer product of two sub-grids (correlated_field.py:856-912; reference check
    test/test_re/test_correlated_field.py:245-283) against the oracle: field, normalized amplitudes, and the operator-level
    likelihood interface (energy, gradient, metric, sqrt-metrics, a CG solve on metric + 1)."""
    kw1 = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    kw2 = dict(fluctuations=(0.3, 0.2), loglogavgslope=(-1.5, 0.2), flexibility=(0.8, 0.3), asperity=None)
    ocf = oracle.CorrelatedFieldOracle("cf", hartley_convention=conv)
    ocf.set_amplitude_total_offset(0.3, (0.2, 0.1))
    ocf.add_fluctuations(shapes[0], distances[0], prefix="space", non_parametric_kind="power", **kw1)
    ocf.add_fluctuations(shapes[1], distances[1], prefix="freq", non_parametric_kind="amplitude", **kw2)
    ocf.finalize()
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt, hartley_convention=conv)
    cfm.set_amplitude_total_offset(0.3, (0.2, 0.1))
    cfm.add_fluctuations(shapes[0], distances[0], prefix="space", non_parametric_kind="power", **kw1)
    cfm.add_fluctuations(shapes[1], distances[1], prefix="freq", non_parametric_kind="amplitude", **kw2)
    cf = cfm.finalize()
    assert isinstance(cf, nb.OuterCorrelatedField) and isinstance(cf, nb.LazyModel)
    assert cf.domain == {k: tuple(v) for k, v in ocf.domain.items()} and tuple(cf.target) == tuple(shapes[0]) + tuple(shapes[1])
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.5 * v for k, v in pos.items()}
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    tt = {k: torch.as_tensor(v) for k, v in tan.items()}
    assert rel_err(t2n(cf(tp)), ocf(pos)) < tol
    for na, ona in zip(cf.normalized_amplitudes, ocf.normalized_amplitudes(pos)):
        assert rel_err(t2n(na(tp)), ona) < 1e-12
    assert len(cf.target_grids) == 2 and tuple(cf.target_grids[1].shape) == tuple(shapes[1])
    full = tuple(shapes[0]) + tuple(shapes[1])
    if lh_kind == "gauss":
        data = osig(pos) + 0.3 * rng.standard_normal(full)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
        lh = nb.Gaussian(data, noise_cov_inv=1.0 / 0.09).amend(nb.SignalModel(cf, "exp"))
    else:
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        lh = nb.Poissonian(data).amend(nb.SignalModel(cf, "exp"))
    assert isinstance(lh, nb.OuterLikelihood)
    e, grad = lh.energy_and_gradient(tp)
    oe, ograd = olh.energy_and_gradient(pos)
    assert abs(e - oe) <= tol * abs(oe) and abs(lh.energy(tp) - oe) <= tol * abs(oe)
    assert tree_err(grad, ograd) < tol
    assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol
    u = rng.standard_normal(full)
    assert tree_err(lh.left_sqrt_metric(tp, u), olh.left_sqrt_metric(pos, u)) < tol
    assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)), olh.right_sqrt_metric(pos, tan)) < tol
    j = lay.pack(lay.random(rng))
    res = lh.cg_on_metric(tp, torch.as_tensor(j, device=rt.device), absdelta=1e-30, miniter=6, maxiter=6)
    ores = oracle.cg(lambda v: lay.pack(olh.metric(pos, lay.unpack(v))) + v, j, absdelta=1e-30, miniter=6, maxiter=6)
    assert res.nit == ores.nit == 6 and rel_err(t2n(res.x), ores.x) < 1e-6
    # an MGVI sample draw on the outer-product model against the oracle, same white noise (evi.py:88-150)
    wd, wp = rng.standard_normal(full), lay.random(rng)
    cgkw = dict(absdelta=1e-30, miniter=8, maxiter=8)
    ores_s, oinfo, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=cgkw)
    smp, info = lh.draw_linear_residual(tp, 0, cg_kwargs=cgkw, _white=(wd, {k: torch.as_tensor(v) for k, v in wp.items()}))
    assert info == oinfo and tree_err(smp, ores_s) < 1e-6
    ms, _ = lh.draw_linear_residual(tp, 0, from_inverse=False, _white=(wd, {k: torch.as_tensor(v) for k, v in wp.items()}))
    oms, _, _ = oracle.draw_linear_residual(olh, pos, wd, wp, from_inverse=False)
    assert tree_err(ms, oms) < 1e-10
    # sample-averaged KL (value, gradient, metric) and one MGVI iteration against the oracle on identical white noise
    residuals = [ores_s, {k: -v for k, v in ores_s.items()}]
    tres = [{k: torch.as_tensor(v) for k, v in r.items()} for r in residuals]
    ov, og = oracle.kl_value_and_grad(olh, pos, residuals)
    v, gr = lh.kl_value_and_grad(tp, tres)
    assert abs(v - ov) <= 1e-10 * abs(ov) and rel_err(t2n(gr), lay.pack(og)) < 1e-9
    om = oracle.kl_metric(olh, pos, tan, residuals)
    assert rel_err(t2n(lh.kl_metric(tp, tt, tres)), lay.pack(om)) < 1e-9
    mk = dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=6))
    whites = [(wd, {k: torch.as_tensor(v) for k, v in wp.items()})]
    newpos, res_rows, states = lh.mgvi(tp, key=3, n_total_iterations=1, n_samples=1, draw_linear_kwargs=dict(cg_kwargs=cgkw),
                                       kl_kwargs=dict(minimize_kwargs=mk), _whites=whites)
    opos, oopt = oracle.kl_minimize(olh, pos, residuals, minimize_kwargs=mk)
    assert res_rows.shape[0] == 2 and states[0].nit == oopt.nit
    assert tree_err(newpos, opos) < 1e-5
    v1, _ = lh.kl_value_and_grad(newpos, res_rows)
    assert v1 < v
    with pytest.raises(NotImplementedError):
        cfm3 = nb.CorrelatedFieldMaker("cf", runtime=rt)
        cfm3.set_amplitude_total_offset(0.0, (0.1, 0.1))
        cfm3.add_fluctuations((2, 2, 2, 2), 1.0, prefix="a0", **kw1)        # a sub-grid with four axes
        cfm3.add_fluctuations((4,), 1.0, prefix="a1", **kw1)
        cfm3.finalize()


def check_outer_golden(rt, name):
    """The outer-product model against the nifty.cl fixtures: field, and J t / J^T c through the likelihood-free autograd path."""
    from golden_util import OUTER_CASES, build_outer
    c, g = OUTER_CASES[name], load(name)
    cf = build_outer(c, nb.CorrelatedFieldMaker("cf", runtime=rt))
    pos = {k: torch.as_tensor(v) for k, v in g["pos"].items()}
    assert rel_err(t2n(cf(pos)), g["field"]) < 1e-10
    lh = nb.Gaussian(np.zeros(g["field"].shape), noise_cov_inv=1.0).amend(cf)          # identity signal: its sqrt-metrics are J, J^T
    tan = {k: torch.as_tensor(v) for k, v in g["tan"].items()}
    assert rel_err(t2n(lh.right_sqrt_metric(pos, tan)), g["field_jvp"]) < 1e-10
    assert tree_err(lh.left_sqrt_metric(pos, g["cot"]), g["field_vjp"]) < 1e-10


def check_nonpow2_hartley(rt, shape, dtype=torch.float64):
    """Hartley transform on extents that are not powers of two (chirp convolution around the power-of-two device transform,
    nifty_b200/bluestein.py) against ``oracle.hartley`` (correlated_field.py:24-30), both sign conventions, and H H = N."""
    rng = np.random.default_rng(int(np.prod(shape)))
    x = rng.standard_normal(shape)
    for conv in ("non_canonical_hartley", "canonical_hartley"):
        H = nb.BluesteinHartley(shape, dtype=dtype, convention=conv, runtime=rt)
        got = H(torch.as_tensor(x, dtype=dtype, device=rt.device))
        assert got.dtype == dtype and tuple(got.shape) == tuple(shape)
        assert rel_err(t2n(got), oracle.hartley(x, convention=conv)) < 100 * TOL[dtype]
        assert rel_err(t2n(H(got)) / np.prod(shape), x) < 100 * TOL[dtype]


def check_nonpow2_golden(rt, name="g2d_3x3"):
    """The reference's own (3, 3) parity case (test/test_re/test_correlated_field.py:123-124) on the device path: every quantity of
    the nifty.cl fixture, sqrt-metrics against the oracle."""
    c, g = CASES[name], load(name)
    tol = 1e-10
    lh = build_product_lh(c, g, rt)
    assert isinstance(lh.cf, nb.BluesteinCorrelatedField)
    pos = {k: torch.as_tensor(v) for k, v in g["pos"].items()}
    tan = {k: torch.as_tensor(v) for k, v in g["tan"].items()}
    assert lh.cf.domain == {k: tuple(np.shape(v)) for k, v in g["pos"].items()}
    assert rel_err(t2n(lh.cf(pos)), g["field"]) < tol
    assert rel_err(t2n(lh.signal_response(pos)), g["signal"]) < tol
    e, grad = lh.energy_and_gradient(pos)
    assert abs(e - float(g["energy"])) <= tol * abs(float(g["energy"]))
    assert tree_err(grad, g["grad"]) < tol
    assert tree_err(lh.metric(pos, tan), g["metric"]) < tol
    olh = build_oracle_lh(c, g)
    assert rel_err(t2n(lh.right_sqrt_metric(pos, tan)), olh.right_sqrt_metric(g["pos"], g["tan"])) < tol
    assert tree_err(lh.left_sqrt_metric(pos, g["cot"]), olh.left_sqrt_metric(g["pos"], g["cot"])) < tol
    bare = nb.Gaussian(np.zeros(c["shape"]), noise_cov_inv=1.0).amend(lh.cf)            # identity signal: its sqrt-metrics are J, J^T
    assert rel_err(t2n(bare.right_sqrt_metric(pos, tan)), g["field_jvp"]) < tol
    assert tree_err(bare.left_sqrt_metric(pos, g["cot"]), g["field_vjp"]) < tol


def check_nonpow2_model(rt, shape=(6, 10), distances=(0.2, 0.3), lh_kind="gauss", conv="non_canonical_hartley", seed=5, tol=1e-10):
    """A correlated field on a grid with non-power-of-two extents against the oracle: mode tables, field, amplitudes, the
    operator-level likelihood interface, one MGVI sample draw and one MGVI iteration on identical white noise."""
    kw = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    ocf = oracle.CorrelatedFieldOracle("cf", hartley_convention=conv)
    ocf.set_amplitude_total_offset(0.3, (0.2, 0.1))
    ocf.add_fluctuations(shape, distances, prefix="ax1", non_parametric_kind="power", **kw)
    ocf.finalize()
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt, hartley_convention=conv)
    cfm.set_amplitude_total_offset(0.3, (0.2, 0.1))
    cfm.add_fluctuations(shape, distances, prefix="ax1", non_parametric_kind="power", **kw)
    cf = cfm.finalize()
    assert isinstance(cf, nb.BluesteinCorrelatedField) and isinstance(cf, nb.LazyModel)
    assert cf.domain == {k: tuple(v) for k, v in ocf.domain.items()} and tuple(cf.target) == tuple(shape)
    idx, um, cnt = oracle.fourier_mode_distributor(shape, distances)
    hg = cf.target_grids[0].harmonic_grid
    assert np.array_equal(hg.power_distributor, idx) and np.array_equal(hg.mode_lengths, um) and np.array_equal(hg.mode_multiplicity, cnt)
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.5 * v for k, v in pos.items()}
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    tt = {k: torch.as_tensor(v) for k, v in tan.items()}
    assert rel_err(t2n(cf(tp)), ocf(pos)) < tol
    assert rel_err(t2n(cf.normalized_amplitudes[0](tp)), ocf.normalized_amplitudes(pos)[0]) < 1e-12
    assert rel_err(t2n(cfm.amplitude(tp)), ocf.amplitude(pos)) < 1e-12 and rel_err(t2n(cfm.power_spectrum(tp)), ocf.amplitude(pos) ** 2) < 1e-12
    if lh_kind == "gauss":
        data = osig(pos) + 0.3 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
        lh = nb.Gaussian(data, noise_cov_inv=1.0 / 0.09).amend(nb.SignalModel(cf, "exp"))
    else:
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        lh = nb.Poissonian(data).amend(nb.SignalModel(cf, "exp"))
    e, grad = lh.energy_and_gradient(tp)
    oe, ograd = olh.energy_and_gradient(pos)
    assert abs(e - oe) <= tol * abs(oe)
    assert tree_err(grad, ograd) < tol
    assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol
    wd, wp = rng.standard_normal(shape), lay.random(rng)
    cgkw = dict(absdelta=1e-30, miniter=8, maxiter=8)
    white = (wd, {k: torch.as_tensor(v) for k, v in wp.items()})
    ores_s, oinfo, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=cgkw)
    smp, info = lh.draw_linear_residual(tp, 0, cg_kwargs=cgkw, _white=white)
    assert info == oinfo and tree_err(smp, ores_s) < 1e-6
    residuals = [ores_s, {k: -v for k, v in ores_s.items()}]
    mk = dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=6))
    newpos, res_rows, states = lh.mgvi(tp, key=3, n_total_iterations=1, n_samples=1, draw_linear_kwargs=dict(cg_kwargs=cgkw),
                                       kl_kwargs=dict(minimize_kwargs=mk), _whites=[white])
    opos, oopt = oracle.kl_minimize(olh, pos, residuals, minimize_kwargs=mk)
    assert states[0].nit == oopt.nit and tree_err(newpos, opos) < 1e-5


def check_nonpow2_errors(rt):
    """`nb200_hartley_chirpz` refuses a padded plan that is too short for the cyclic convolution; shapes are checked."""
    import ctypes
    with pytest.raises(NotImplementedError):
        nb.BluesteinHartley((5078,), runtime=rt)                 # padded line of 16384 points: above the float64 limit of the passes
    H = nb.BluesteinHartley((5, 6), runtime=rt)
    with pytest.raises(ValueError):
        H(torch.zeros((6, 5), dtype=torch.float64))
    small = nb.Plan((8, 8), 1.0, runtime=rt)                     # needs >= 9 and >= 11
    x = torch.zeros((5, 6), dtype=torch.float64, device=rt.device)
    with pytest.raises(nb.NB200Error):
        rt.api.call("nb200_hartley_chirpz", small._h, rt.stream(), (ctypes.c_int64 * 2)(5, 6), rt.ptr(H._tab), rt.ptr(x),
                    rt.ptr(torch.zeros(128, dtype=torch.float64, device=rt.device)), rt.ptr(torch.empty_like(x)))


def check_host_composed_matern(rt, tol=1e-10):
    """Matern amplitudes (correlated_field.py:302-395) on host-composed fields -- a grid with non-power-of-two extents and an
    outer product Matern x non-parametric -- against the oracle: field, normalized amplitudes, energy / gradient / metric."""
    mk = dict(scale=(1.0, 0.5), cutoff=(1.0, 0.5), loglogslope=(-3.0, 0.5))
    kw = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    cases = [[("m", (6, 10), (0.2, 0.3), dict(renormalize_amplitude=True, non_parametric_kind="power"))],
             [("m", (5, 3, 4), 0.4, dict(renormalize_amplitude=False, non_parametric_kind="amplitude"))],
             [("m", (8,), 0.5, dict(renormalize_amplitude=True, non_parametric_kind="amplitude")), ("n", (4, 4), 0.25, None)],
             [("n", (6,), 1.0, None), ("m", (3, 3), 0.1, dict(renormalize_amplitude=False, non_parametric_kind="power"))]]
    for ci, subs in enumerate(cases):
        ocf = oracle.CorrelatedFieldOracle("cf")
        cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
        for m in (ocf, cfm):
            m.set_amplitude_total_offset(0.1, (0.2, 0.1))
            for i, (typ, shp, dist, opt) in enumerate(subs):
                if typ == "m":
                    m.add_fluctuations_matern(shp, dist, prefix=f"s{i}", **mk, **opt)
                else:
                    m.add_fluctuations(shp, dist, prefix=f"s{i}", non_parametric_kind="power", **kw)
        ocf.finalize()
        cf = cfm.finalize()
        assert isinstance(cf, (nb.OuterCorrelatedField, nb.BluesteinCorrelatedField))
        assert cf.domain == {k: tuple(v) for k, v in ocf.domain.items()}
        shape = sum((tuple(s[1]) for s in subs), ())
        osig = oracle.SignalOracle(ocf, "exp")
        lay = oracle.Layout(osig.domain)
        rng = np.random.default_rng(100 + ci)
        pos, tan = lay.random(rng), lay.random(rng)
        pos = {k: 0.5 * v for k, v in pos.items()}
        tp = {k: torch.as_tensor(v) for k, v in pos.items()}
        tt = {k: torch.as_tensor(v) for k, v in tan.items()}
        assert rel_err(t2n(cf(tp)), ocf(pos)) < tol, ci
        for na, ona in zip(cf.normalized_amplitudes, ocf.normalized_amplitudes(pos)):
            assert rel_err(t2n(na(tp)), ona) < 1e-12
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
        lh = nb.Poissonian(data).amend(nb.SignalModel(cf, "exp"))
        e, grad = lh.energy_and_gradient(tp)
        oe, ograd = olh.energy_and_gradient(pos)
        assert abs(e - oe) <= tol * abs(oe) and tree_err(grad, ograd) < tol, ci
        assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol, ci


def check_reference_cf_cases(rt):
    """The reference's remaining correlated-field tests restated on this path (test/test_re/test_correlated_field.py):
    :17-47 / :50-78 construction on (2,) and (3, 3) with optional asperity / flexibility and with Matern amplitudes,
    :81-113 TypeError for malformed prior tuples, :286-322 `renormalize_amplitude=True` gives fields of standard deviation
    `scale` (200 prior draws on a 12-point grid)."""
    import itertools
    for shape, asp, flx in itertools.product([(2,), (3, 3)], [None, (1.0, 1.0)], [None, (1.0, 1.0)]):
        cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
        cfm.set_amplitude_total_offset(offset_mean=0, offset_std=(0.1, 0.1))
        cfm.add_fluctuations(shape, distances=0.1, fluctuations=(1.0, 1.0), loglogavgslope=(1.0, 1.0), asperity=asp, flexibility=flx)
        cf = cfm.finalize()
        assert cf is not None and cf.domain                       # what the reference asserts
        if shape == (3, 3):
            f = cf(cf.init(3))
            assert tuple(f.shape) == shape and bool(torch.isfinite(f).all())
    for shape in [(2,), (3, 3)]:
        cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
        cfm.set_amplitude_total_offset(offset_mean=0, offset_std=(0.1, 0.1))
        cfm.add_fluctuations_matern(shape, distances=0.1, scale=(1.0, 1.0), loglogslope=(1.0, 1.0), cutoff=(1.0, 1.0), renormalize_amplitude=False)
        cf = cfm.finalize()
        assert cf is not None and cf.domain and (shape != (3, 3) or bool(torch.isfinite(cf(cf.init(1))).all()))
    choices = ([1e-1], [1e-1, 5e-3], [1e-1, 5e-3, 5e-3], 1e-1)
    for flu, slp, flx, asp in itertools.product(choices, repeat=4):
        ok = all(isinstance(el, (tuple, list)) and len(el) == 2 and all(isinstance(v, float) for v in el) for el in (flu, slp, flx, asp))
        if ok:
            continue
        with pytest.raises(TypeError):
            cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
            cfm.set_amplitude_total_offset(offset_mean=0.0, offset_std=(1e-3, 1e-4))
            cfm.add_fluctuations(shape=(16,), distances=(1.0 / 16,), fluctuations=flu, loglogavgslope=slp, flexibility=flx, asperity=asp,
                                 prefix="ax1", non_parametric_kind="power")
            cfm.finalize()
    for scale, slope, cutoff, kind in [((1.0, 1e-10), (-1.0, 1.0), (1.0, 1.0), "amplitude"), ((3.0, 1e-10), (-5.0, 0.5), (3.3, 0.01), "power"),
                                       ((3.0, 1e-10), (-1.0, 1.0), (3.3, 0.01), "amplitude"), ((1.0, 1e-10), (-5.0, 0.5), (1.0, 1.0), "power")]:
        cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
        cfm.set_amplitude_total_offset(offset_mean=0.0, offset_std=(0.1, 1e-10))
        cfm.add_fluctuations_matern((12,), distances=0.1, scale=scale, cutoff=cutoff, loglogslope=slope, non_parametric_kind=kind,
                                    renormalize_amplitude=True)
        cf = cfm.finalize()
        fields = np.stack([t2n(cf(cf.init(k))) for k in range(200)])
        avg_scale = float(np.mean(np.std(fields, axis=0)))
        assert abs(avg_scale - scale[0]) < 2e-1, (avg_scale, scale)


def check_host_composed_scaling(rt, tol=1e-10):
    """The demo's log-normal `scaling` leaf (signal = scaling * exp(cf), demos/re/0_intro.py:39-57) on host-composed fields against
    the oracle: energy, gradient, metric, both sqrt-metrics, on a non-power-of-two grid and on an outer product."""
    kw = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    for ci, subs in enumerate([[((6, 10), (0.2, 0.3))], [((8,), 0.5), ((3, 4), 0.25)]]):
        ocf = oracle.CorrelatedFieldOracle("cf")
        cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
        for m in (ocf, cfm):
            m.set_amplitude_total_offset(0.1, (0.2, 0.1))
            for i, (shp, dist) in enumerate(subs):
                m.add_fluctuations(shp, dist, prefix=f"s{i}", non_parametric_kind="power", **kw)
        ocf.finalize()
        cf = cfm.finalize()
        shape = sum((tuple(s[0]) for s in subs), ())
        osig = oracle.SignalOracle(ocf, "exp", scaling=(1.0, 0.5))
        sig = nb.SignalModel(cf, "exp", scaling=(1.0, 0.5))
        assert sig.domain == {k: tuple(v) for k, v in osig.domain.items()}
        lay = oracle.Layout(osig.domain)
        rng = np.random.default_rng(200 + ci)
        pos, tan = lay.random(rng), lay.random(rng)
        pos = {k: 0.5 * v for k, v in pos.items()}
        data = osig(pos) + 0.3 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
        lh = nb.Gaussian(data, noise_cov_inv=1.0 / 0.09).amend(sig)
        tp = {k: torch.as_tensor(v) for k, v in pos.items()}
        tt = {k: torch.as_tensor(v) for k, v in tan.items()}
        assert rel_err(t2n(lh.signal_response(tp)), osig(pos)) < tol
        e, grad = lh.energy_and_gradient(tp)
        oe, ograd = olh.energy_and_gradient(pos)
        assert abs(e - oe) <= tol * abs(oe) and tree_err(grad, ograd) < tol
        assert tree_err(lh.metric(tp, tt), olh.metric(pos, tan)) < tol
        u = rng.standard_normal(shape)
        assert tree_err(lh.left_sqrt_metric(tp, u), olh.left_sqrt_metric(pos, u)) < tol
        assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)), olh.right_sqrt_metric(pos, tan)) < tol
        flat = rt.asarray(lay.pack(pos), torch.float64)                      # flat positions in the order of the SIGNAL layout
        assert abs(lh.energy(flat) - oe) <= tol * abs(oe)


def check_operator_gaussian(rt, shape=(8, 8), lh_nl="exp", seed=17):
    """`Gaussian(data, noise_cov_inv=<non-diagonal operator>, noise_std_inv=<its square root>)` (likelihood_impl.py:35-138 accepts
    arbitrary callables) on the fused model: energy, gradient, metric, sqrt-metrics, transformation, residual and an MGVI draw
    against DENSE linear algebra built from the oracle's signal Jacobian."""
    c = dict(shape=shape, distances=1.0 / shape[0], offset_mean=0.1, offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3),
             flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh="gauss")
    from golden_util import build_oracle
    ocf = build_oracle(c)
    osig = oracle.SignalOracle(ocf, lh_nl)
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    nd, L = int(np.prod(shape)), lay.size
    B = rng.standard_normal((nd, nd))
    Ninv = B @ B.T / nd + 2.0 * np.eye(nd)                       # symmetric positive definite, couples every pair of pixels
    ev, U = np.linalg.eigh(Ninv)
    Nsq = (U * np.sqrt(ev)) @ U.T                                # symmetric square root
    tN, tS = torch.as_tensor(Ninv), torch.as_tensor(Nsq)

    def op(m):            # device-agnostic callables: they are probed on the host and applied to device tensors later
        return lambda x: (m.to(device=x.device, dtype=x.dtype) @ x.reshape(-1)).reshape(x.shape)

    pos, tan = lay.random(rng), lay.random(rng)
    pos = {k: 0.3 * v for k, v in pos.items()}
    data = osig(pos) + 0.3 * rng.standard_normal(shape)
    lh = nb.Gaussian(data, noise_cov_inv=op(tN), noise_std_inv=op(tS)).amend(nb.SignalModel(build_product(c, rt), lh_nl))
    pow2 = all(n & (n - 1) == 0 for n in shape)
    assert isinstance(lh, nb.OperatorLikelihood if pow2 else nb.OuterLikelihood) and isinstance(lh, nb.LikelihoodWithModel)
    pv, tv = lay.pack(pos), lay.pack(tan)
    J = np.stack([osig.jvp(pos, lay.unpack(e)).reshape(-1) for e in np.eye(L)], axis=1)          # nd x L
    s = osig(pos).reshape(-1)
    r = s - data.reshape(-1)
    tp, tt = rt.asarray(pv, torch.float64), rt.asarray(tv, torch.float64)
    tol = 1e-10
    e, g = lh.energy_and_gradient(tp)
    assert abs(e - 0.5 * r @ Ninv @ r) <= tol * abs(e)
    assert rel_err(t2n(g), J.T @ Ninv @ r) < tol
    M = J.T @ Ninv @ J
    assert rel_err(t2n(lh.metric(tp, tt)), M @ tv) < tol
    u = rng.standard_normal(shape)
    assert rel_err(t2n(lh.left_sqrt_metric(tp, u)), J.T @ Nsq @ u.reshape(-1)) < tol
    assert rel_err(t2n(lh.right_sqrt_metric(tp, tt)).reshape(-1), Nsq @ J @ tv) < tol
    assert rel_err(t2n(lh.transformation(tp)).reshape(-1), Nsq @ s) < tol
    assert rel_err(t2n(lh.normalized_residual(tp)).reshape(-1), Nsq @ (data.reshape(-1) - s)) < tol
    # MGVI draw: (M + 1)^-1 (J^T N^-1/2 w_d + w_p) by CG (host loop) against the dense solve
    wd, wp = rng.standard_normal(shape), rng.standard_normal(L)
    res, info = nb.draw_linear_residual(lh, tp, 0, cg_kwargs=dict(resnorm=1e-10, maxiter=4 * L),
                                        _white=(rt.asarray(wd, torch.float64), rt.asarray(wp, torch.float64)))
    dense = np.linalg.solve(M + np.eye(L), J.T @ Nsq @ wd.reshape(-1) + wp)
    assert info == 0 and rel_err(t2n(res), dense) < 1e-7
    # Wiener-filter posterior mean of the linearised problem (evi.py:399-476) against the dense solve
    wf, (winfo, _) = nb.wiener_filter_posterior(lh, tp, key=5, n_samples=0, model_is_linear=False,
                                                draw_linear_kwargs=dict(cg_kwargs=dict(resnorm=1e-10, maxiter=4 * L)))
    wdense = np.linalg.solve(M + np.eye(L), J.T @ Ninv @ (data.reshape(-1) - s + J @ pv))
    assert winfo == 0 and rel_err(t2n(wf.pos), wdense) < 1e-7
    # one MGVI iteration of the standard driver on this likelihood: the KL decreases
    kw = dict(n_samples=1, key=5, sample_mode="linear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-8, maxiter=2 * L)),
              kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=30))))
    smp, st = nb.optimize_kl(lh, tp, n_total_iterations=1, **kw)
    vi = nb.OptimizeVI(lh, 1)
    e1, _ = vi.kl_value_and_grad(smp.pos, smp.residuals)
    e0, _ = vi.kl_value_and_grad(tp, smp.residuals)
    assert st.nit == 1 and np.isfinite(e1) and e1 < e0
    # without the square root: energy / metric work, sqrt-metrics say what is missing
    lh2 = nb.Gaussian(data, noise_cov_inv=op(tN)).amend(nb.SignalModel(build_product(c, rt), lh_nl))
    assert rel_err(t2n(lh2.metric(tp, tt)), M @ tv) < tol
    with pytest.raises(NotImplementedError, match="noise_std_inv"):
        lh2.left_sqrt_metric(tp, u)
    with pytest.raises(NotImplementedError):
        nb.Gaussian(data, noise_std_inv=op(tS))                  # a non-diagonal square root alone
