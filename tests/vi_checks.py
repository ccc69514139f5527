"""Sampling / KL checks shared by the CPU (emulator) and GPU tiers: the product's host drivers
(nifty_b200.evi / optimize_kl, mirrors of nifty.re) against the oracle restatement on identical
white-noise inputs."""
import os

import numpy as np
import torch

import nifty_b200 as nb
import oracle
from golden_util import CASES, build_oracle_lh, load, rel_err
from parity_checks import build_product_lh, t2n

CG_KW = dict(absdelta=1e-8, maxiter=60)


def _setup(rt, name):
    c, g = CASES[name], load(name)
    lh = build_product_lh(c, g, rt)
    olh = build_oracle_lh(c, g)
    lay = oracle.Layout(olh.domain)
    return c, g, lh, olh, lay


def check_draw_linear_residual(rt, name="g2d_16x16"):
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(21)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    wd, wp = rng.standard_normal(c["shape"]), lay.random(rng)
    ores, oinfo, ocg = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=CG_KW)
    tpos = rt.asarray(lay.pack(pos), torch.float64)
    white = (rt.asarray(wd, torch.float64), rt.asarray(lay.pack(wp), torch.float64))
    res, info = nb.draw_linear_residual(lh, tpos, 0, cg_kwargs=CG_KW, _white=white)
    assert info == oinfo
    assert rel_err(t2n(res), lay.pack(ores)) < 1e-9
    # from_inverse=False is the metric sample itself
    ms, _ = nb.draw_linear_residual(lh, tpos, 0, from_inverse=False, _white=white)
    oms, _, _ = oracle.draw_linear_residual(olh, pos, wd, wp, from_inverse=False)
    assert rel_err(t2n(ms), lay.pack(oms)) < 1e-11
    return lh, olh, lay, pos, tpos, white, (wd, wp), res, ores


def check_wiener_filter(rt, name="g2d_16x16"):
    """wiener_filter_posterior (evi.py:399-517, signal-space branch) against the oracle restatement and, through the
    oracle's operators, against the DENSE solve of the linearised problem (the reference's own check,
    test/test_re/test_evi.py:206-232, compares with a dense Wiener filter)."""
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(33)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    kw = dict(resnorm=1e-9, maxiter=400)      # (tighter: the N_RESET residual refresh at iteration 20 trips "energy increased")
    omean, oinfo, ores = oracle.wiener_filter_posterior_mean(olh, pos, cg_kwargs=kw)
    # dense cross-check of the oracle: (J^T M J + 1) mean = J^T M (d - f + J p)
    L = lay.size
    pv = lay.pack(pos)
    H = np.stack([lay.pack(olh.metric(pos, lay.unpack(e))) + e for e in np.eye(L)], axis=1)
    rhs = lay.pack(olh.left_sqrt_metric(pos, olh.normalized_residual(pos) + olh.right_sqrt_metric(pos, pos)))
    dense = np.linalg.solve(H, rhs)
    assert rel_err(lay.pack(omean), dense) < 1e-7
    tpos = rt.asarray(pv, torch.float64)
    smp, (info, sinfo) = nb.wiener_filter_posterior(lh, tpos, key=5, n_samples=2, model_is_linear=False,
                                                    draw_linear_kwargs=dict(cg_kwargs=kw))
    assert info == oinfo and sinfo == [0, 0]
    assert rel_err(t2n(smp.pos), dense) < 1e-7
    assert rel_err(t2n(smp.pos), lay.pack(omean)) < 1e-9
    s, r = smp.samples, smp.residuals
    assert s.shape == (4, L) and torch.equal(r[0], -r[1]) and torch.equal(r[2], -r[3])     # mirrored pairs around the mean
    assert torch.equal(s[2], smp.pos + r[2])
    # the draws are MGVI samples at the posterior mean: same as draw_linear_residual with the same key
    k0 = nb.random_split(5, 2)[0]
    r0, _ = nb.draw_linear_residual(lh, smp.pos, k0, cg_kwargs=kw)
    assert torch.equal(r[0], r0)
    # error behaviour of the reference
    import pytest
    with pytest.raises(ValueError, match="position to linearize"):
        nb.wiener_filter_posterior(lh, key=1, model_is_linear=False)
    with pytest.raises(ValueError, match="noise_covariance"):
        nb.wiener_filter_posterior(lh, tpos, key=1, signal_space=False)
    # data-space branch (evi.py:477-497): same mean, as the reference checks (test_evi.py:218-231, atol 7e-7)
    ncov = 1.0 / float(g["noise_cov_inv"])
    dsp, (dinfo, _) = nb.wiener_filter_posterior(lh, tpos, key=1, n_samples=0, model_is_linear=False, signal_space=False,
                                                 noise_covariance=lambda x: ncov * x,
                                                 draw_linear_kwargs=dict(cg_kwargs=dict(resnorm=1e-9, maxiter=600)))
    assert dinfo == 0 and rel_err(t2n(dsp.pos), dense) < 1e-6
    with pytest.raises(TypeError):
        nb.wiener_filter_posterior(object(), tpos, key=1)


def check_minisanity(rt, name="g2d_16x16"):
    """reduced_residual_stats / minisanity (minisanity.py:17-129) against the formulas evaluated with NumPy on the
    oracle's residuals: mean = sum/size, reduced chi^2 = <x,x>/size per leaf, [mean, std] over the samples."""
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(8)
    pos = lay.pack({k: 0.3 * v for k, v in lay.random(rng).items()})
    res = 0.1 * np.stack([lay.pack(lay.random(rng)) for _ in range(3)])
    smp = nb.Samples(pos=rt.asarray(pos, torch.float64), samples=rt.asarray(res, torch.float64))
    plan, layout = lh.signal.cf.plan, lh.layout
    # prior residuals: per latent leaf
    st, msg = nb.minisanity(smp, plan=plan, layout=layout)
    pts = pos[None] + res
    for k in lay.keys:
        leaf = np.stack([lay.unpack(p)[k].reshape(-1) for p in pts])
        m, rx = leaf.sum(1) / leaf.shape[1], (leaf * leaf).sum(1) / leaf.shape[1]
        assert st[k].ndof == leaf.shape[1]
        np.testing.assert_allclose(st[k].mean, [m.mean(), m.std()], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(st[k].reduced_chisq, [rx.mean(), rx.std()], rtol=1e-10, atol=1e-14)
    assert msg.count("reduced Chi²:") == len(lay.keys) and all(f"{k:24s}::" in msg for k in lay.keys)
    # likelihood residuals through func
    st_r, msg_r = nb.minisanity(smp, lh.normalized_residual, plan=plan)
    r = np.stack([olh.normalized_residual(lay.unpack(p)).reshape(-1) for p in pts])
    np.testing.assert_allclose(st_r.mean[0], (r.sum(1) / r.shape[1]).mean(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st_r.reduced_chisq, [((r * r).sum(1) / r.shape[1]).mean(), ((r * r).sum(1) / r.shape[1]).std()], rtol=1e-9)
    assert st_r.ndof == r.shape[1] and msg_r.startswith("reduced Chi²:")
    # a bare position: the sample std is zero (minisanity.py:52-53)
    st0 = nb.reduced_residual_stats(rt.asarray(pos, torch.float64), plan=plan, layout=layout)
    assert all(v.mean[1] == 0 and v.reduced_chisq[1] == 0 for v in st0.values())


def check_nonlinear_update(rt, name="g2d_16x16"):
    lh, olh, lay, pos, tpos, white, (wd, wp), res, ores = check_draw_linear_residual(rt, name)
    mk = dict(xtol=1e-6, maxiter=4, cg_kwargs=dict(maxiter=40))
    for sign in (1.0, -1.0):
        onew, oopt = oracle.nonlinearly_update_residual(olh, pos, {k: sign * v for k, v in ores.items()}, wd, wp, sign, minimize_kwargs=mk)
        new, opt = nb.nonlinearly_update_residual(lh, tpos, sign * res, 0, sign, minimize_kwargs=mk, _white=white)
        assert opt.nit == oopt.nit and opt.status == oopt.status
        assert abs(opt.fun - oopt.fun) <= 1e-6 * max(abs(oopt.fun), 1e-12) + 1e-12
        assert rel_err(t2n(new), lay.pack(onew)) < 1e-6


def check_kl(rt, name="g2d_16x16"):
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(5)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    residuals = [{k: 0.1 * v for k, v in lay.random(rng).items()} for _ in range(3)]
    residuals = [r for pair in ((r, {k: -v for k, v in r.items()}) for r in residuals) for r in pair]
    tan = lay.random(rng)
    ov, og = oracle.kl_value_and_grad(olh, pos, residuals)
    om = oracle.kl_metric(olh, pos, tan, residuals)
    vi = nb.OptimizeVI(lh, 1)
    tpos = rt.asarray(lay.pack(pos), torch.float64)
    tres = rt.asarray(np.stack([lay.pack(r) for r in residuals]), torch.float64)
    v, gr = vi.kl_value_and_grad(tpos, tres)
    assert abs(v - ov) <= 1e-11 * abs(ov)
    assert rel_err(t2n(gr), lay.pack(og)) < 1e-10
    m = vi.kl_metric(rt.asarray(lay.pack(tan), torch.float64))
    assert rel_err(t2n(m), lay.pack(om)) < 1e-10
    # no samples: the Hamiltonian at pos itself (optimize_kl.py:99-103)
    ov0, og0 = oracle.kl_value_and_grad(olh, pos, [])
    v0, g0 = vi.kl_value_and_grad(tpos, None)
    assert abs(v0 - ov0) <= 1e-11 * abs(ov0) and rel_err(t2n(g0), lay.pack(og0)) < 1e-10


def check_point_estimates(rt, name="g2d_16x16", frozen=("cfax1fluctuations", "cfax1spectrum")):
    """point_estimates / constants (likelihood.py:399-499 LikelihoodPartial + partial_insert_and_remove :119-177;
    evi.py:62-85, 109-111, 149, 224-254; optimize_kl.py:553-590).  The oracle follows the reference: every solve runs
    on vectors with the frozen leaves REMOVED; the product keeps full-length vectors and clears the frozen entries
    (device CG: nb200_cg_opts.frozen) -- two formulations of the same restricted operators."""
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(77)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    wd, wp = rng.standard_normal(c["shape"]), lay.random(rng)
    tpos = rt.asarray(lay.pack(pos), torch.float64)
    white = (rt.asarray(wd, torch.float64), rt.asarray(lay.pack(wp), torch.float64))
    fr = lh.frozen_ranges(frozen)
    assert sum(hi - lo for lo, hi in fr) == sum(int(np.prod(lay.shapes[k])) for k in frozen)
    # MGVI draw
    ores, oinfo, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=CG_KW, point_estimates=frozen)
    res, info = nb.draw_linear_residual(lh, tpos, 0, cg_kwargs=CG_KW, point_estimates=frozen, _white=white)
    assert info == oinfo
    assert rel_err(t2n(res), lay.pack(ores)) < 1e-9
    assert all(float(res[lo:hi].abs().max()) == 0.0 for lo, hi in fr)
    full, _ = nb.draw_linear_residual(lh, tpos, 0, cg_kwargs=CG_KW, _white=white)
    assert rel_err(t2n(res), t2n(full)) > 1e-3          # freezing changes the draw
    # geoVI update
    mk = dict(xtol=1e-6, maxiter=3, cg_kwargs=dict(maxiter=40))
    for sign in (1.0, -1.0):
        onew, oopt = oracle.nonlinearly_update_residual(olh, pos, {k: sign * v for k, v in ores.items()}, wd, wp, sign,
                                                        minimize_kwargs=mk, point_estimates=frozen)
        new, opt = nb.nonlinearly_update_residual(lh, tpos, sign * res, 0, sign, minimize_kwargs=mk, point_estimates=frozen,
                                                  _white=white)
        assert opt.nit == oopt.nit and opt.status == oopt.status
        assert rel_err(t2n(new), lay.pack(onew)) < 1e-6
        assert all(float(new[lo:hi].abs().max()) == 0.0 for lo, hi in fr)
    # KL minimisation with constants: the constant leaves do not move, the others follow the oracle
    residuals = [{k: 0.1 * v for k, v in lay.random(rng).items()} for _ in range(2)]
    residuals = [r for pair in ((r, {k: -v for k, v in r.items()}) for r in residuals) for r in pair]
    kmk = dict(xtol=1e-8, maxiter=3, cg_kwargs=dict(maxiter=30))
    opos, oopt = oracle.kl_minimize(olh, pos, residuals, constants=frozen, minimize_kwargs=kmk)
    vi = nb.OptimizeVI(lh, 1)
    smp = nb.Samples(pos=tpos, samples=rt.asarray(np.stack([lay.pack(r) for r in residuals]), torch.float64))
    kopt = vi.kl_minimize(smp, minimize_kwargs=kmk, constants=frozen)
    assert kopt.nit == oopt.nit
    assert rel_err(t2n(kopt.x), lay.pack(opos)) < 1e-7
    assert all(torch.equal(kopt.x[lo:hi], tpos[lo:hi]) for lo, hi in fr)
    # unknown keys are an error, as in the reference's _parse_point_estimates
    import pytest
    with pytest.raises(ValueError, match="not leaves"):
        nb.draw_linear_residual(lh, tpos, 0, point_estimates=("nope",))
    # end to end: one VI iteration; point-estimated leaves carry no sample spread, constant leaves do not move
    pe, const = (frozen[0],), (frozen[-1],)
    smp, st = nb.optimize_kl(lh, tpos.clone(), key=3, n_total_iterations=1, n_samples=2, point_estimates=pe, constants=const,
                             draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=40)),
                             nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))),
                             kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=20))),
                             sample_mode="nonlinear_resample")
    assert st.nit == 1 and len(smp) == 4
    for lo, hi in lh.frozen_ranges(pe):
        assert float(smp.residuals[:, lo:hi].abs().max()) == 0.0
    for lo, hi in lh.frozen_ranges(const):
        assert torch.equal(smp.pos[lo:hi], tpos[lo:hi])
    moved = torch.ones(lay.size, dtype=torch.bool)
    for lo, hi in lh.frozen_ranges(const):
        moved[lo:hi] = False
    assert float((smp.pos - tpos)[moved].abs().max()) > 0


def check_optimize_kl(rt, tmpdir, name="g2d_16x16", comm=None):
    """Two VI iterations: the KL decreases, samples are interleaved antithetic pairs, resume continues."""
    c, g, lh, olh, lay = _setup(rt, name)
    pos0 = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    kw = dict(n_samples=2, key=42, draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
              nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=30))),
              kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=4, cg_kwargs=dict(maxiter=30))), sample_mode="nonlinear_resample")
    vi = nb.OptimizeVI(lh, 2, comm=comm)
    samples, state = nb.optimize_kl(lh, pos0, n_total_iterations=1, odir=str(tmpdir), comm=comm, **kw)
    assert state.nit == 1
    nloc = len(samples)
    assert nloc % 2 == 0
    r = samples.residuals
    e1, _ = vi.kl_value_and_grad(samples.pos, r)
    e0, _ = vi.kl_value_and_grad(pos0, r)       # same residuals around the start position
    assert np.isfinite(e1) and e1 < e0
    # resume: picks up at iteration 1 and runs the second one
    samples2, state2 = nb.optimize_kl(lh, pos0, n_total_iterations=2, odir=str(tmpdir), resume=True, comm=comm, **kw)
    assert state2.nit == 2 and len(samples2) == nloc
    e2, _ = vi.kl_value_and_grad(samples2.pos, samples2.residuals)
    assert np.isfinite(e2)
    # minisanity.txt (optimize_kl.py:803, 866-870): one report per iteration, appended on resume
    import os
    if vi.comm.rank == 0:
        txt = open(os.path.join(str(tmpdir), "minisanity.txt")).read()
        assert txt.count("OPTIMIZE_KL: Iteration") == 2 and "Iteration 0001" in txt and "Iteration 0002" in txt
        assert "#(Nonlinear sampling steps)" in txt and "Likelihood residual(s):" in txt and "Prior residual(s):" in txt
        assert txt.count("reduced Chi²:") == 2 * (1 + len(lh.layout.keys))
    return samples2, state2


def check_map_and_schedules(rt, name="g2d_16x16"):
    """n_samples = 0 is a MAP run (optimize_kl.py:488-489, 530-531); n_samples / sample_mode may be callables of the
    iteration index (:166-170); the MAP energy agrees with the oracle's Newton-CG on the Hamiltonian."""
    c, g, lh, olh, lay = _setup(rt, name)
    pos0 = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    kmk = dict(xtol=1e-6, maxiter=6, cg_kwargs=dict(maxiter=30))
    s, st = nb.optimize_kl(lh, pos0, key=1, n_total_iterations=1, n_samples=0, kl_kwargs=dict(minimize_kwargs=kmk))
    assert st.nit == 1 and len(s) == 0 and st.sample_state == 0
    opos, oopt = oracle.kl_minimize(olh, lay.unpack(t2n(pos0)), [], minimize_kwargs=kmk)
    assert oopt.nit == st.minimization_state.nit
    assert abs(st.minimization_state.fun - oopt.fun) <= 1e-9 * abs(oopt.fun)
    assert rel_err(t2n(s.pos), lay.pack(opos)) < 1e-7
    s, st = nb.optimize_kl(lh, pos0, key=1, n_total_iterations=2, n_samples=lambda i: 1 if i < 1 else 2,
                           sample_mode=lambda i: "linear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=30)),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=20))))
    assert st.nit == 2 and len(s) == 4


def check_elbo(rt, name="g2d_16x16"):
    """estimate_evidence_lower_bound (evidence_lower_bound.py:414-1042, eigenvalue path) against dense linear algebra on the
    oracle's metric: ELBO_i = -1/2 sum log eig(M + 1) + L/2 - H(s_i)  (:826-832, 956-960), the unresolved tail in
    `lower_error`; signal- and data-space operators give the same eigenvalues (:154-173)."""
    import pytest
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(12)
    pos = lay.pack({k: 0.3 * v for k, v in lay.random(rng).items()})
    res = 0.05 * np.stack([lay.pack(lay.random(rng)) for _ in range(3)])
    smp = nb.Samples(pos=rt.asarray(pos, torch.float64), samples=rt.asarray(res, torch.float64))
    L, nd = lay.size, int(np.prod(c["shape"]))
    pd = lay.unpack(pos)
    H = np.stack([lay.pack(olh.metric(pd, lay.unpack(e))) + e for e in np.eye(L)], axis=1)
    eig = np.sort(np.linalg.eigvalsh(0.5 * (H + H.T)))[::-1]
    nrel = min(nd, L)
    ham = np.array([olh.energy(lay.unpack(pos + r)) + 0.5 * (pos + r) @ (pos + r) for r in res])
    # all relevant eigenvalues: exact trace-log, no tail
    el, st = nb.estimate_evidence_lower_bound(lh, smp, 0, compute_all=True, verbose=False)
    want = -0.5 * np.sum(np.log(eig[:nrel])) + 0.5 * L - ham
    np.testing.assert_allclose(el, want, rtol=1e-9)
    assert st["lower_error"] == 0.0 and abs(st["elbo_mean"] - want.mean()) <= 1e-9 * abs(want.mean())
    assert abs(st["elbo_std"] - want.std(ddof=1)) <= 1e-6 * want.std(ddof=1) + 1e-12
    assert abs(st["elbo_up"] - (want.mean() + want.std(ddof=1))) < 1e-6 * abs(want.mean())
    # the largest n by ARPACK in batches with deflation; the tail bound of :828-832
    n = 24
    for space in ("signal", "data"):
        el2, st2 = nb.estimate_evidence_lower_bound(lh, smp, n, n_batches=5, min_lh_eval=1e-12, verbose=False, trace_log_space=space)
        top = eig[:n]
        want2 = -0.5 * np.sum(np.log(top)) + 0.5 * L - ham
        np.testing.assert_allclose(el2, want2, rtol=1e-8)
        assert abs(st2["lower_error"] - 0.5 * (nrel - n) * np.log(top.min())) <= 1e-7 * abs(st2["lower_error"]) + 1e-10
        assert abs(st2["elbo_lw"] - (want2.mean() - want2.std(ddof=1) - st2["lower_error"])) <= 1e-7 * abs(want2.mean())
    # early stop: once the smallest eigenvalue found is within min_lh_eval of 1 the remaining batches are skipped
    el3, st3 = nb.estimate_evidence_lower_bound(lh, smp, 40, n_batches=40, min_lh_eval=float(eig[5] - 1.0 + 1e-9), verbose=False)
    k = 6          # one eigenvalue per batch: the sixth is the first one within min_lh_eval of 1
    want3 = -0.5 * np.sum(np.log(eig[:k])) + 0.5 * L - ham
    np.testing.assert_allclose(el3, want3, rtol=1e-8)
    # stochastic Lanczos quadrature of the same trace-log (lanczos.py): full Krylov order makes every probe exact up to the
    # probe variance; signal- and data-space operators estimate the same log-determinant
    exact = -0.5 * np.sum(np.log(eig)) + 0.5 * L - ham
    for space in ("signal", "data"):
        el4, st4 = nb.estimate_evidence_lower_bound(lh, smp, 0, trace_log_method="slq", trace_log_space=space, slq_order=64,
                                                    slq_num_samples=24, slq_key=3, verbose=False)
        assert st4["lower_error"] == 0.0 and st4["slq_stochastic_se"] > 0
        assert abs(st4["elbo_mean"] - exact.mean()) < 5.0 * st4["slq_stochastic_se"] + 1e-3 * abs(exact.mean())
    tri, vecs = nb.lanczos_tridiag(lambda v: lh.lin_at(smp.pos)[0].metric(v, add_identity=True), rt.asarray(np.ones(L), torch.float64), order=8)
    q = t2n(vecs)
    np.testing.assert_allclose(q @ q.T, np.eye(8), atol=1e-10)                     # orthonormal Krylov basis of the device product
    np.testing.assert_allclose(q @ (0.5 * (H + H.T)) @ q.T, t2n(tri), atol=1e-8 * np.abs(H).max())
    # refused / invalid options
    with pytest.raises(ValueError, match="analytic_prior_term requires"):
        nb.estimate_evidence_lower_bound(lh, smp, 4, analytic_prior_term=True, verbose=False)
    with pytest.raises(NotImplementedError):
        nb.estimate_evidence_lower_bound(lh, smp, 4, trace_log_method="slq", slq_jit=True, verbose=False)
    with pytest.raises(ValueError, match="at least one eigenvalue"):
        nb.estimate_evidence_lower_bound(lh, smp, 0, verbose=False)
    with pytest.raises(ValueError, match="exceeds"):
        nb.estimate_evidence_lower_bound(lh, smp, nrel + 1, verbose=False)
    with pytest.raises(TypeError):
        nb.estimate_evidence_lower_bound(lh, t2n(smp.pos), 4, verbose=False)


def check_final_kl(rt, shape=(32, 32), n_samples=4, scaling=None, noise_cov_inv=100.0):
    """BASELINE.json configs[0] in miniature (demos/re/0_intro.py: exp(correlated field) [x log-normal scaling], Gaussian
    noise, 4 MGVI samples, one optimize_kl iteration with the demo's CG / Newton settings): the FINAL KL of the product path
    against the oracle's restatement of the same iteration on identical white noise -- north_star: "matching nifty.re's KL
    to within 1e-10 relative".  Conditioning caveat: through ~30 CG iterations on an ill-conditioned
    metric ANY two implementations drift apart by many orders of magnitude above their 1e-16 operator difference (measured
    here for the demo's scaling leaf at noise std 0.1: operators agree to 2e-16, 27-iteration solves to 4e-4), so trajectory
    parity is asserted where the solves are well conditioned: the plain exp(cf) model at the demo's noise level, and the
    scaling model at a noise level where its metric is close to the identity."""
    kw = dict(fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    c = dict(shape=shape, distances=1.0 / shape[0], offset_mean=0.0, offset_std=(1e-3, 1e-4), lh="gauss", **kw)
    from golden_util import build_oracle
    from parity_checks import build_product
    ocf = build_oracle(c)
    osig = oracle.SignalOracle(ocf, "exp", scaling=scaling)
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(42)
    truth = lay.random(rng)
    data = osig(truth) + noise_cov_inv ** -0.5 * rng.standard_normal(shape)
    olh = oracle.GaussianOracle(data, noise_cov_inv, osig)
    lh = nb.Gaussian(data, noise_cov_inv=noise_cov_inv).amend(nb.SignalModel(build_product(c, rt), "exp", scaling=scaling))
    L = lay.size
    pos0 = {k: 0.1 * v for k, v in lay.random(np.random.default_rng(44)).items()}
    tpos0 = rt.asarray(lay.pack(pos0), torch.float64)
    whites = [(rng.standard_normal(shape), lay.random(rng)) for _ in range(n_samples)]
    dkw = dict(absdelta=1e-4 * L / 10, maxiter=100)                       # 0_intro.py:105-108
    kmk = dict(xtol=1e-4, maxiter=35, cg_kwargs=dict(name=None))          # 0_intro.py:120-124
    calls = []

    def dlr(lh_, pos, key, **kwargs):          # the reference's injection seam (optimize_kl.py:228-246)
        wd, wp = whites[len(calls)]
        calls.append(key)
        return nb.draw_linear_residual(lh_, pos, key, _white=(rt.asarray(wd, torch.float64), rt.asarray(lay.pack(wp), torch.float64)), **kwargs)

    vi = nb.OptimizeVI(lh, 1, _draw_linear_residual=dlr)
    samples, state = nb.optimize_kl(lh, tpos0, key=42, n_total_iterations=1, n_samples=n_samples, sample_mode="linear_resample",
                                    draw_linear_kwargs=dict(cg_kwargs=dkw), kl_kwargs=dict(minimize_kwargs=kmk), _optimize_vi=vi)
    assert len(calls) == n_samples and len(samples) == 2 * n_samples
    residuals = []
    for wd, wp in whites:
        r, info, _ = oracle.draw_linear_residual(olh, pos0, wd, wp, cg_kwargs=dkw)
        residuals += [r, {k: -v for k, v in r.items()}]
    for i, r in enumerate(residuals):
        assert rel_err(t2n(samples.residuals[i]), lay.pack(r)) < 1e-9, i
    okmk = dict(kmk); okmk["cg_kwargs"] = {}
    opos, oopt = oracle.kl_minimize(olh, pos0, residuals, minimize_kwargs=okmk)
    ms = state.minimization_state
    assert (ms.nit, ms.status) == (oopt.nit, oopt.status), (ms.nit, oopt.nit)
    rel = abs(ms.fun - oopt.fun) / abs(oopt.fun)
    assert rel <= 1e-10, rel
    assert rel_err(t2n(samples.pos), lay.pack(opos)) < 1e-8
    return rel


def check_schedule_indices(rt, name="g2d_16x16"):
    """Schedules are evaluated at the PRE-increment iteration index (optimize_kl.py:691-701): the first iteration sees 0, so the
    idiom `lambda i: "nonlinear_resample" if i >= 1 else "linear_resample"` runs one linear iteration first; only callables
    of exactly one argument are evaluated (:166-170)."""
    c, g, lh, olh, lay = _setup(rt, name)
    pos0 = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    seen, modes = [], []

    def n_samples(i):
        seen.append(i)
        return 1

    def sample_mode(i):
        modes.append(i)
        return "nonlinear_resample" if i >= 1 else "linear_resample"

    s, st = nb.optimize_kl(lh, pos0, key=1, n_total_iterations=2, n_samples=n_samples, sample_mode=sample_mode,
                           draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=30)),
                           nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=1, cg_kwargs=dict(maxiter=10))),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))))
    assert seen == [0, 1] and modes == [0, 1] and st.nit == 2
    # iteration 0 was linear: its sample state holds CG infos, iteration 1 non-linear: OptimizeResults per sample
    assert isinstance(st.sample_state, (list, tuple)) and all(hasattr(el, "nit") for el in st.sample_state)


def check_kl_cg_on_device(rt, name="g2d_16x16", x_tol=1e-9, kws=None):
    """`_cg` on the sample-averaged operator: the device solve (nb200_cg_solve_multi: products accumulated in the last-pass
    epilogues, fused curvature reduction) against the generic host loop on the same operator -- identical iteration
    counts / info, solutions to 1e-10; with frozen ranges too."""
    from nifty_b200.conjugate_gradient import _cg
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(5)
    pos = rt.asarray(lay.pack({k: 0.3 * v for k, v in lay.random(rng).items()}), torch.float64)
    res = rt.asarray(0.1 * np.stack([lay.pack(lay.random(rng)) for _ in range(3)]), torch.float64)
    vi = nb.OptimizeVI(lh, 1)
    vi.kl_value_and_grad(pos, torch.cat([res, -res]))
    j = rt.asarray(lay.pack(lay.random(rng)), torch.float64)
    for frozen in (None, lh.frozen_ranges(("cfax1spectrum",))):
        op = vi._kl_operator(frozen=frozen)
        jj = j.clone()
        if frozen:
            for lo, hi in frozen:
                jj[lo:hi] = 0
        for kw in kws or (dict(absdelta=1e-7, maxiter=50), dict(resnorm=1e-4, norm_ord=1, maxiter=25), dict(absdelta=1e-30, miniter=12, maxiter=12)):
            dev = _cg(op, jj, **kw)
            host = _cg(lambda t: op(t), jj, **kw)
            assert (dev.nit, dev.info, dev.nfev) == (host.nit, host.info, host.nfev), kw
            assert rel_err(t2n(dev.x), t2n(host.x)) < x_tol, kw


def check_reduce_pieces(rt, name="g2d_16x16"):
    """Sample-averaged product with the last pass cut into pieces (nb200_plan_set_reduce_chunks): the ranges handed to the
    reduction hook partition the latent vector exactly once, arrive before the final (flush) call, and the result equals
    the one-piece product bit for bit."""
    from nifty_b200._runtime import metric_multi
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(11)
    pos = rt.asarray(lay.pack({k: 0.3 * v for k, v in lay.random(rng).items()}), torch.float64)
    t = rt.asarray(lay.pack(lay.random(rng)), torch.float64)
    lins = []
    for sgn in (1.0, -1.0):
        l = lh.new_lin()
        l.update(pos + sgn * 0.1, want_grad=False)
        lins.append(l)
    plan = lh.signal.cf.plan
    plan.set_reduce_chunks(1)
    ref = metric_multi(lins, t, scale=0.5, identity_here=True).clone()
    L = t.numel()
    for nch in (1, 2, 3, 4, 64):
        plan.set_reduce_chunks(nch)
        seen = []
        out = torch.full_like(t, float("nan"))

        def reduce_fn(piece, out=out, seen=seen):
            off = (piece.data_ptr() - out.data_ptr()) // piece.element_size()
            seen.append((int(off), int(piece.numel())))
            return None

        got = metric_multi(lins, t, scale=0.5, identity_here=True, out=out, reduce_fn=reduce_fn)
        cover = np.zeros(L, dtype=np.int64)
        for off, n in seen:
            assert 0 <= off and off + n <= L and n > 0
            cover[off:off + n] += 1
        assert np.all(cover == 1), (nch, seen)
        if nch > 1:
            assert len(seen) > 2
        assert torch.equal(got, ref), nch
    plan.set_reduce_chunks(1)


def _host_composed_pair(rt, which, lh_kind="gauss", seed=8):
    """(likelihood of this package, oracle likelihood, oracle layout, shape) for a host-composed field: 'outer' = outer product
    of two sub-grids, 'nonpow2' = one grid whose extents are not powers of two."""
    kw1 = dict(fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    kw2 = dict(fluctuations=(0.3, 0.2), loglogavgslope=(-1.5, 0.2), flexibility=(0.8, 0.3), asperity=None)
    subs = [((8, 4), (0.2, 0.3), "space", kw1), ((4,), 0.5, "freq", kw2)] if which == "outer" else [((6, 10), (0.2, 0.3), "ax1", kw1)]
    ocf = oracle.CorrelatedFieldOracle("cf")
    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt)
    for m in (ocf, cfm):
        m.set_amplitude_total_offset(0.3, (0.2, 0.1))
        for shp, dist, pf, kw in subs:
            m.add_fluctuations(shp, dist, prefix=pf, non_parametric_kind="power", **kw)
    ocf.finalize()
    cf = cfm.finalize()
    shape = sum((tuple(s[0]) for s in subs), ())
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    truth = {k: 0.5 * v for k, v in lay.random(rng).items()}
    if lh_kind == "gauss":
        data = osig(truth) + 0.3 * rng.standard_normal(shape)
        return nb.Gaussian(data, noise_cov_inv=1.0 / 0.09).amend(nb.SignalModel(cf, "exp")), oracle.GaussianOracle(data, 1.0 / 0.09, osig), lay, shape
    data = rng.poisson(osig(truth)).astype(np.int64)
    return nb.Poissonian(data).amend(nb.SignalModel(cf, "exp")), oracle.PoissonianOracle(data, osig), lay, shape


def check_host_composed_vi(rt, which, lh_kind="gauss"):
    """Outer-product and non-power-of-two fields through the SAME drivers as the fused single-grid path: `nb.draw_linear_residual`,
    `nb.nonlinearly_update_residual` (geoVI) and one `nb.optimize_kl` iteration against the oracle on identical white noise
    (optimize_kl.py:672-729, evi.py:88-255); transformation / residual / minisanity message on the way."""
    lh, olh, lay, shape = _host_composed_pair(rt, which, lh_kind)
    assert isinstance(lh, nb.LikelihoodWithModel) and isinstance(lh, nb.OuterLikelihood)
    rng = np.random.default_rng(31)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    tpos = rt.asarray(lay.pack(pos), torch.float64)
    assert rel_err(t2n(lh.transformation(tpos)), olh.transformation(pos)) < 1e-11
    assert rel_err(t2n(lh.normalized_residual(tpos)), olh.normalized_residual(pos)) < 1e-9
    e, g = lh.energy_and_gradient(tpos, add_prior=True)
    oe, og = olh.energy_and_gradient(pos)
    assert abs(e - (oe + 0.5 * float(np.dot(lay.pack(pos), lay.pack(pos))))) <= 1e-10 * abs(e)
    assert rel_err(t2n(g), lay.pack(og) + lay.pack(pos)) < 1e-10
    # MGVI draw and geoVI update of both signs
    wd, wp = rng.standard_normal(shape), lay.random(rng)
    white = (rt.asarray(wd, torch.float64), rt.asarray(lay.pack(wp), torch.float64))
    cgkw = dict(absdelta=1e-30, miniter=8, maxiter=8)       # fixed iteration count: trajectories, not stopping rules, are compared
    ores, oinfo, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=cgkw)
    res, info = nb.draw_linear_residual(lh, tpos, 0, cg_kwargs=cgkw, _white=white)
    assert info == oinfo and rel_err(t2n(res), lay.pack(ores)) < 1e-6
    mk = dict(xtol=1e-6, maxiter=3, cg_kwargs=dict(maxiter=30))
    for sign in (1.0, -1.0):
        onew, oopt = oracle.nonlinearly_update_residual(olh, pos, {k: sign * v for k, v in ores.items()}, wd, wp, sign, minimize_kwargs=mk)
        new, opt = nb.nonlinearly_update_residual(lh, tpos, sign * res, 0, sign, minimize_kwargs=mk, _white=white)
        assert opt.nit == oopt.nit and opt.status == oopt.status
        assert rel_err(t2n(new), lay.pack(onew)) < 1e-5
    # one optimize_kl iteration (MGVI, 2 samples) with injected white noise: residuals, Newton iterations, final KL, position
    n_samples = 2
    whites = [(rng.standard_normal(shape), lay.random(rng)) for _ in range(n_samples)]
    dkw = cgkw
    kmk = dict(xtol=1e-4, maxiter=5, cg_kwargs=dict(maxiter=30))
    calls = []

    def dlr(lh_, p, key, **kwargs):
        wd_, wp_ = whites[len(calls)]
        calls.append(key)
        return nb.draw_linear_residual(lh_, p, key, _white=(rt.asarray(wd_, torch.float64), rt.asarray(lay.pack(wp_), torch.float64)), **kwargs)

    vi = nb.OptimizeVI(lh, 1, _draw_linear_residual=dlr)
    samples, state = nb.optimize_kl(lh, tpos, key=42, n_total_iterations=1, n_samples=n_samples, sample_mode="linear_resample",
                                    draw_linear_kwargs=dict(cg_kwargs=dkw), kl_kwargs=dict(minimize_kwargs=kmk), _optimize_vi=vi)
    assert len(calls) == n_samples and len(samples) == 2 * n_samples
    residuals = []
    for wd_, wp_ in whites:
        r, _, _ = oracle.draw_linear_residual(olh, pos, wd_, wp_, cg_kwargs=dkw)
        residuals += [r, {k: -v for k, v in r.items()}]
    for i, r in enumerate(residuals):
        assert rel_err(t2n(samples.residuals[i]), lay.pack(r)) < 1e-6, i
    opos, oopt = oracle.kl_minimize(olh, pos, residuals, minimize_kwargs=kmk)
    ms = state.minimization_state
    assert (ms.nit, ms.status) == (oopt.nit, oopt.status)
    assert abs(ms.fun - oopt.fun) <= 1e-7 * abs(oopt.fun)
    assert rel_err(t2n(samples.pos), lay.pack(opos)) < 1e-5
    msg = vi.get_status_message(samples, state, name="OPTIMIZE_KL")
    assert "Likelihood residual(s):" in msg and "Prior residual(s):" in msg
    # a geoVI iteration of the state machine runs (nonlinear_resample), and a point estimate is honoured
    s2, st2 = nb.optimize_kl(lh, tpos, key=1, n_total_iterations=1, n_samples=1, sample_mode="nonlinear_resample",
                             point_estimates=("cfzeromode",), draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-8, maxiter=40)),
                             nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))),
                             kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))))
    assert st2.nit == 1 and len(s2) == 2
    lo = lh.layout.offsets["cfzeromode"]
    assert float(s2.residuals[:, lo].abs().max()) == 0.0


def check_sample_consistency(rt, sample_mode="nonlinear_resample", point_estimates=(), lh=None):
    """test/test_re/test_optimize_kl.py:120-250 restated: `draw_residual` == `draw_linear_residual` followed by the two
    `nonlinearly_update_residual` calls (sign +1 / -1, same key) == the samples `optimize_kl` holds after one iteration with a
    KL minimisation of zero steps; residuals vanish on point-estimated leaves.  Linear modes run the update with maxiter = 0."""
    if lh is None:
        lh = _setup(rt, "g2d_16x16")[2]
    pos = 0.1 * lh.layout.random(3, torch.float64, rt.device)
    delta = 1e-3
    dkw = dict(cg_name="SL", cg_kwargs=dict(miniter=2, absdelta=delta * lh.layout.size / 10.0, maxiter=100))
    mk = dict(name="SN", xtol=delta, cg_kwargs=dict(name=None, miniter=2), maxiter=5 if sample_mode == "nonlinear_resample" else 0)
    frozen = lh.frozen_ranges(point_estimates)

    def close(x, y):
        # the reference compares element-wise at rtol 1e-7 on a deterministic back end; the segment sums of the host-composed path
        # use atomics on the GPU (summation order varies from run to run), so the bound is taken in the maximum norm
        x, y = t2n(x), t2n(y)
        assert np.max(np.abs(x - y)) <= 1e-7 * np.max(np.abs(y)) + 1e-12

    def zero_on_frozen(r):
        for lo, hi in frozen:
            assert float(r[..., lo:hi].abs().max()) == 0.0

    key = 7
    draw, _ = nb.draw_residual(lh, pos, key, point_estimates=point_estimates, minimize_kwargs=mk, **dkw)
    zero_on_frozen(draw)
    l1, _ = nb.draw_linear_residual(lh, pos, key, point_estimates=point_estimates, **dkw)
    zero_on_frozen(l1)
    n1, _ = nb.nonlinearly_update_residual(lh, pos, l1, key, +1.0, point_estimates=point_estimates, minimize_kwargs=mk)
    n2, _ = nb.nonlinearly_update_residual(lh, pos, -l1, key, -1.0, point_estimates=point_estimates, minimize_kwargs=mk)
    diy = torch.stack((n1, n2))
    zero_on_frozen(diy)
    close(draw, diy)
    if sample_mode != "nonlinear_resample":
        close(draw, torch.stack((l1, -l1)))
    s, _ = nb.optimize_kl(lh, pos, key=11, n_total_iterations=1, n_samples=1, point_estimates=point_estimates, draw_linear_kwargs=dkw,
                          nonlinearly_update_kwargs=dict(minimize_kwargs=mk), kl_kwargs=dict(minimize_kwargs=dict(name="M", maxiter=0)),
                          sample_mode=sample_mode)
    zero_on_frozen(s.residuals)
    again, _ = nb.draw_residual(lh, pos, s.keys[0], point_estimates=point_estimates, minimize_kwargs=mk, **dkw)
    close(s.residuals, again)
    np.testing.assert_array_equal(t2n(s.pos), t2n(pos))          # zero KL steps: the expansion point stays


def check_constants_do_not_move(rt, constants=("cfax1fluctuations", "cfax1spectrum"), lh=None):
    """test/test_re/test_optimize_kl.py:253-323 restated: after one `optimize_kl` iteration the constant leaves have not moved at
    all and every other entry has."""
    if lh is None:
        lh = _setup(rt, "g2d_16x16")[2]
    pos = 0.1 * lh.layout.random(5, torch.float64, rt.device)
    delta = 1e-3
    dkw = dict(cg_name="SL", cg_kwargs=dict(miniter=2, absdelta=delta * lh.layout.size / 10.0, maxiter=100))
    s, _ = nb.optimize_kl(lh, pos, key=13, n_total_iterations=1, n_samples=1, constants=constants, draw_linear_kwargs=dkw,
                          kl_kwargs=dict(minimize_kwargs=dict(name="M", maxiter=5)), sample_mode="linear_resample")
    move = t2n(s.pos - pos)
    mask = np.zeros(move.shape, dtype=bool)
    for lo, hi in lh.frozen_ranges(constants):
        mask[lo:hi] = True
    assert np.all(move[mask] == 0.0) and np.all(move[~mask] != 0.0)


def check_host_composed_wiener_and_elbo(rt, which="nonpow2"):
    """`wiener_filter_posterior` (both branches), stochastic Lanczos quadrature and `estimate_evidence_lower_bound` accept host-composed likelihoods: the posterior
    mean against the dense solve of the oracle's linearised problem (as check_wiener_filter does for the fused path)."""
    lh, olh, lay, shape = _host_composed_pair(rt, which, "gauss")
    rng = np.random.default_rng(33)
    pos = {k: 0.3 * v for k, v in lay.random(rng).items()}
    L, pv = lay.size, lay.pack(pos)
    H = np.stack([lay.pack(olh.metric(pos, lay.unpack(e))) + e for e in np.eye(L)], axis=1)
    rhs = lay.pack(olh.left_sqrt_metric(pos, olh.normalized_residual(pos) + olh.right_sqrt_metric(pos, pos)))
    dense = np.linalg.solve(H, rhs)
    tpos = rt.asarray(pv, torch.float64)
    kw = dict(resnorm=1e-9, maxiter=400)
    smp, (info, sinfo) = nb.wiener_filter_posterior(lh, tpos, key=5, n_samples=1, model_is_linear=False, draw_linear_kwargs=dict(cg_kwargs=kw))
    assert info == 0 and sinfo == [0] and rel_err(t2n(smp.pos), dense) < 1e-7
    assert smp.samples.shape == (2, L) and torch.equal(smp.residuals[0], -smp.residuals[1])
    dsp, (dinfo, _) = nb.wiener_filter_posterior(lh, tpos, key=1, n_samples=0, model_is_linear=False, signal_space=False,
                                                 noise_covariance=lambda x: 0.09 * x, draw_linear_kwargs=dict(cg_kwargs=dict(resnorm=1e-9, maxiter=600)))
    assert dinfo == 0 and rel_err(t2n(dsp.pos), dense) < 1e-6
    # trace-log of metric + 1 by stochastic Lanczos quadrature on the host-applied operator against the dense log-determinant
    lin, _ = lh.lin_at(tpos)
    est = nb.stochastic_lq_logdet(lambda v: lin.metric(v, add_identity=True), 20, 16, 3, shape0=L, dtype=torch.float64, device=rt.device)
    want = float(np.linalg.slogdet(H)[1])
    assert abs(float(est) - want) < 0.2 * abs(want) + 1.0
    # estimate_evidence_lower_bound, all relevant eigenvalues (eigsh path) against dense eigenvalues of the oracle metric
    res = 0.05 * np.stack([lay.pack(lay.random(rng)) for _ in range(2)])
    smp2 = nb.Samples(pos=tpos, samples=rt.asarray(res, torch.float64))
    el, st = nb.estimate_evidence_lower_bound(lh, smp2, 0, compute_all=True, verbose=False)
    eig = np.sort(np.linalg.eigvalsh(0.5 * (H + H.T)))[::-1]
    nrel = min(int(np.prod(shape)), L)
    ham = np.array([olh.energy(lay.unpack(pv + r)) + 0.5 * (pv + r) @ (pv + r) for r in res])
    np.testing.assert_allclose(el, -0.5 * np.sum(np.log(eig[:nrel])) + 0.5 * L - ham, rtol=1e-8)


def check_elbo_hybrid(rt, name="g2d_16x16"):
    """The hybrid trace-log estimator of estimate_evidence_lower_bound (evidence_lower_bound.py:833-937): the largest eigenvalues
    exactly, the remainder by stochastic Lanczos quadrature on probes deflated by their eigenvectors, optionally bracketed by
    Gauss-Radau quadratures -- against dense eigenvalues of the oracle's metric."""
    import pytest
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(12)
    pos = lay.pack({k: 0.3 * v for k, v in lay.random(rng).items()})
    res = 0.05 * np.stack([lay.pack(lay.random(rng)) for _ in range(3)])
    smp = nb.Samples(pos=rt.asarray(pos, torch.float64), samples=rt.asarray(res, torch.float64))
    L = lay.size
    pd = lay.unpack(pos)
    H = np.stack([lay.pack(olh.metric(pd, lay.unpack(e))) + e for e in np.eye(L)], axis=1)
    eig = np.sort(np.linalg.eigvalsh(0.5 * (H + H.T)))[::-1]
    ham = np.array([olh.energy(lay.unpack(pos + r)) + 0.5 * (pos + r) @ (pos + r) for r in res])
    exact = -0.5 * np.sum(np.log(eig)) + 0.5 * L - ham
    n = 24
    for space in ("signal", "data"):
        for radau in (False, True):
            # (a Radau node that a Ritz value has converged onto makes the modified tridiagonal singular: the reference raises then,
            # :930-937, and so does this path -- hence the lower order with the bound)
            el, st = nb.estimate_evidence_lower_bound(lh, smp, n, trace_log_method="slq", trace_log_space=space, slq_order=10 if radau else 40,
                                                      slq_num_samples=12, slq_key=3, use_radau_as_bound=radau, n_batches=4, verbose=False)
            assert abs(st["exact_log"] - np.sum(np.log(eig[:n]))) <= 1e-8 * abs(st["exact_log"])
            rem = float(np.sum(np.log(eig[n:])))
            assert abs(st["slq_remainder"] - rem) <= 5.0 * 2.0 * st["slq_stochastic_se"] + 1e-3 * abs(rem) + 1e-6
            assert st["lower_error"] >= 0.0 and np.isfinite(st["lower_error"])
            assert abs(st["elbo_mean"] - exact.mean()) <= 5.0 * st["slq_stochastic_se"] + 1e-3 * abs(exact.mean())
            assert st["elbo_lw"] <= st["elbo_mean"] <= st["elbo_up"]
    # resuming a stored eigensystem (:199-268): 10 eigenpairs written by a first call, the remaining 14 computed by the second one
    import tempfile
    d = tempfile.mkdtemp()
    nb.estimate_evidence_lower_bound(lh, smp, 10, n_batches=2, min_lh_eval=1e-12, output_directory=d, verbose=False)
    vals, vecs = np.load(os.path.join(d, "metric_signal_eigenvalues.npy")), np.load(os.path.join(d, "metric_signal_eigenvectors.npy"))
    assert vals.shape == (10,) and vecs.shape == (L, 10)
    el_r, st_r = nb.estimate_evidence_lower_bound(lh, smp, n, n_batches=2, min_lh_eval=1e-12, resume_eigenvectors=vecs, resume_eigenvalues=vals,
                                                  verbose=False)
    np.testing.assert_allclose(el_r, -0.5 * np.sum(np.log(eig[:n])) + 0.5 * L - ham, rtol=1e-8)
    with pytest.raises(ValueError, match="does not match the selected operator"):
        nb.estimate_evidence_lower_bound(lh, smp, n, resume_eigenvectors=vecs, resume_eigenvalues=2.0 * vals, verbose=False)
    with pytest.raises(ValueError, match="mismatched sizes"):
        nb.estimate_evidence_lower_bound(lh, smp, n, resume_eigenvectors=vecs, resume_eigenvalues=vals[:5], verbose=False)
    with pytest.raises(ValueError, match="requires resume_eigenvectors"):
        nb.estimate_evidence_lower_bound(lh, smp, n, resume_eigenvalues=vals, verbose=False)
    with pytest.raises(ValueError, match="resume_eigenvalues is required"):
        nb.estimate_evidence_lower_bound(lh, smp, n, resume_eigenvectors=vecs, verbose=False)
    # analytic_prior_term (:809-824, 953-979): the prior energy in closed form, 1/2 (tr (M + 1)^-1 + <mean, mean>)
    lh_e = np.array([olh.energy(lay.unpack(pos + r)) for r in res])
    prior = 0.5 * (np.sum(1.0 / eig) + pos @ pos)
    want = -0.5 * np.sum(np.log(eig)) + 0.5 * L - lh_e - prior
    el, st = nb.estimate_evidence_lower_bound(lh, smp, 0, compute_all=True, analytic_prior_term=True, verbose=False)
    np.testing.assert_allclose(el, want, rtol=1e-8)
    assert abs(st["trace_inv_total"] - np.sum(1.0 / eig)) <= 1e-8 * st["trace_inv_total"] and abs(st["prior_term"] - prior) <= 1e-8 * prior
    for space in ("signal", "data"):
        el, st = nb.estimate_evidence_lower_bound(lh, smp, n, trace_log_method="slq", trace_log_space=space, slq_order=40, slq_num_samples=12,
                                                  slq_key=3, analytic_prior_term=True, n_batches=4, verbose=False)
        assert abs(st["trace_inv_total"] - np.sum(1.0 / eig)) <= 5.0 * st["trace_inv_se"] + 1e-3 * np.sum(1.0 / eig)
        assert abs(st["elbo_mean"] - want.mean()) <= 5.0 * (st["slq_stochastic_se"] + 0.5 * st["trace_inv_se"]) + 1e-3 * abs(want.mean())
        assert st["trace_inv_const"] == L - min(L, int(np.prod(c["shape"])))
    with pytest.raises(ValueError, match="upper spectral endpoint"):
        nb.estimate_evidence_lower_bound(lh, smp, 0, trace_log_method="slq", use_radau_as_bound=True, verbose=False)
    with pytest.raises(ValueError, match="too close to the Lanczos spectrum"):
        nb.estimate_evidence_lower_bound(lh, smp, n, trace_log_method="slq", slq_order=60, slq_num_samples=4, use_radau_as_bound=True, verbose=False)
    with pytest.raises(ValueError, match="at least two probes"):
        nb.estimate_evidence_lower_bound(lh, smp, 4, trace_log_method="slq", slq_num_samples=1, verbose=False)


def check_likelihood_sum(rt):
    """`lh_a + lh_b` (likelihood.py:661-757; reference check test/test_re/test_likelihood.py:103-168): (1) two data sets seen through
    ONE correlated field against the single Gaussian they are equivalent to, (2) two different fields with different likelihoods on
    the union of their latent domains against the summands evaluated one by one, (3) the drivers on the sum."""
    import pytest
    from parity_checks import build_product, tree_err
    c = dict(shape=(16, 8), distances=(0.1, 0.2), offset_mean=0.1, offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1), loglogavgslope=(-2.0, 0.3),
             flexibility=(1.0, 0.5), asperity=(0.5, 0.05), lh="gauss")
    rng = np.random.default_rng(9)
    sig = nb.SignalModel(build_product(c, rt), "exp")
    L = sig.layout.size
    da, db = rng.standard_normal(c["shape"]) + 1.0, rng.standard_normal(c["shape"]) + 1.0
    wa, wb = 4.0, 9.0
    lh_a, lh_b = nb.Gaussian(da, noise_cov_inv=wa).amend(sig), nb.Gaussian(db, noise_cov_inv=wb).amend(sig)
    lh_ab = lh_a + lh_b
    assert isinstance(lh_ab, nb.LikelihoodSum) and lh_ab.layout.size == L and lh_ab.domain == sig.domain
    one = nb.Gaussian((wa * da + wb * db) / (wa + wb), noise_cov_inv=wa + wb).amend(sig)
    p1, p2, t = (0.3 * sig.layout.random(s, torch.float64, rt.device) for s in (1, 2, 3))
    (e1, g1), (e2, _) = lh_ab.energy_and_gradient(p1), lh_ab.energy_and_gradient(p2)
    (f1, h1), (f2, _) = one.energy_and_gradient(p1), one.energy_and_gradient(p2)
    assert abs((e1 - f1) - (e2 - f2)) <= 1e-10 * abs(e1) and rel_err(t2n(g1), t2n(h1)) < 1e-10        # energies differ by a constant
    assert abs(e1 - (lh_a.energy(p1) + lh_b.energy(p1))) <= 1e-12 * abs(e1)
    assert rel_err(t2n(lh_ab.metric(p1, t)), t2n(one.metric(p1, t))) < 1e-10
    rs = lh_ab.right_sqrt_metric(p1, t)
    assert sorted(rs) == ["lh_0", "lh_1"] and rel_err(t2n(rs["lh_1"]), t2n(lh_b.right_sqrt_metric(p1, t))) < 1e-12
    assert rel_err(t2n(lh_ab.left_sqrt_metric(p1, rs)), t2n(lh_ab.metric(p1, t))) < 1e-10            # metric == LSM o RSM
    nr = lh_ab.normalized_residual(p1)
    assert rel_err(t2n(nr["lh_0"]), t2n(lh_a.normalized_residual(p1))) < 1e-12
    assert rel_err(t2n(lh_ab.transformation(p1)["lh_1"]), t2n(lh_b.transformation(p1))) < 1e-12
    # (2) different fields and likelihoods: block structure on the union of the domains
    c2 = dict(c, shape=(8, 8), distances=0.25)
    cfm2 = nb.CorrelatedFieldMaker("other", runtime=rt)
    cfm2.set_amplitude_total_offset(0.5, (0.2, 0.1))
    cfm2.add_fluctuations(c2["shape"], c2["distances"], (0.3, 0.1), (-1.5, 0.3), (1.0, 0.5), None, prefix="ax", non_parametric_kind="amplitude")
    sig2 = nb.SignalModel(cfm2.finalize(), "exp")
    lh_c = nb.Poissonian(rng.poisson(2.0, size=c2["shape"])).amend(sig2)
    both = lh_a + lh_c
    assert set(both.domain) == set(sig.domain) | set(sig2.domain) and both.layout.size == L + sig2.layout.size
    tree = {**{k: 0.3 * v for k, v in sig.layout.unpack(sig.layout.random(4, torch.float64, rt.device)).items()},
            **{k: 0.3 * v for k, v in sig2.layout.unpack(sig2.layout.random(5, torch.float64, rt.device)).items()}}
    tan = {**sig.layout.unpack(sig.layout.random(6, torch.float64, rt.device)), **sig2.layout.unpack(sig2.layout.random(7, torch.float64, rt.device))}
    pa, pc = {k: tree[k] for k in sig.domain}, {k: tree[k] for k in sig2.domain}
    e, g = both.energy_and_gradient(tree)
    (ea, ga), (ec, gc) = lh_a.energy_and_gradient(pa), lh_c.energy_and_gradient(pc)
    assert abs(e - (ea + ec)) <= 1e-12 * abs(e) and tree_err(g, {**{k: t2n(v) for k, v in ga.items()}, **{k: t2n(v) for k, v in gc.items()}}) < 1e-12
    m = both.metric(tree, tan)
    ma, mc = lh_a.metric(pa, {k: tan[k] for k in sig.domain}), lh_c.metric(pc, {k: tan[k] for k in sig2.domain})
    assert tree_err(m, {**{k: t2n(v) for k, v in ma.items()}, **{k: t2n(v) for k, v in mc.items()}}) < 1e-12
    assert isinstance(both + lh_b, nb.LikelihoodSum) and len((both + lh_b).likelihood_summands) == 3
    with pytest.raises(TypeError):
        lh_a + 3
    # (3) the drivers: an MGVI draw solves (M + 1) x = LSM w_d + w_p, one optimize_kl iteration lowers the KL, the report lists both summands
    pos = both.signal.as_flat(tree)
    n_d = both.signal.target_shape[0]
    wd, wp = rt.asarray(rng.standard_normal(n_d), torch.float64), rt.asarray(rng.standard_normal(both.layout.size), torch.float64)
    res, info = nb.draw_linear_residual(both, pos, 0, cg_kwargs=dict(resnorm=1e-10, maxiter=4 * both.layout.size), _white=(wd, wp))
    lin, _ = both.lin_at(pos)
    assert info == 0 and rel_err(t2n(lin.metric(res, add_identity=True)), t2n(lin.lsm(wd) + wp)) < 1e-8
    kw = dict(n_samples=1, key=5, sample_mode="nonlinear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-8, maxiter=200)),
              nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=30))),
              kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=30))))
    vi = nb.OptimizeVI(both, 1)
    smp, st = nb.optimize_kl(both, pos, n_total_iterations=1, _optimize_vi=vi, **kw)
    e1, _ = vi.kl_value_and_grad(smp.pos, smp.residuals)
    e0, _ = vi.kl_value_and_grad(pos, smp.residuals)
    assert st.nit == 1 and np.isfinite(e1) and e1 < e0
    msg = vi.get_status_message(smp, st, name="OPTIMIZE_KL")
    assert "lh_0" in msg and "lh_1" in msg and "otherxi" in msg


def check_freeze(rt, name="g2d_16x16", frozen=("cfax1fluctuations", "cfax1spectrum")):
    """`lh.freeze(primals=, point_estimates=)` -> (`LikelihoodPartial`, liquid primals) (likelihood.py:386-499; reference check
    test/test_re/test_likelihood.py:53-89) against the oracle's removed-leaf formulation: energy, metric, sqrt-metrics,
    transformation as functions of the liquid leaves only."""
    c, g, lh, olh, lay = _setup(rt, name)
    rng = np.random.default_rng(4)
    pos, tan = {k: 0.3 * v for k, v in lay.random(rng).items()}, lay.random(rng)
    tp = {k: torch.as_tensor(v) for k, v in pos.items()}
    same, p_same = lh.freeze(primals=tp, point_estimates=())
    assert same is lh and p_same is tp
    lp, liquid = lh.freeze(primals=tp, point_estimates=frozen)
    assert isinstance(lp, nb.LikelihoodPartial) and sorted(liquid) == sorted(k for k in pos if k not in frozen)
    assert set(lp.domain) == set(liquid) and set(lp.splitx(tp)[1]) == set(frozen)
    fz = oracle.vi._Frozen(olh, lay.pack(pos), frozen)
    pl = fz.remove(lay.pack(pos))
    tl = {k: torch.as_tensor(tan[k]) for k in liquid}

    def pack_liquid(tree):          # liquid dict -> the oracle's packed vector with the frozen entries removed
        full = {k: (t2n(tree[k]) if k in tree else np.zeros(lay.shapes[k])) for k in lay.keys}
        return fz.remove(lay.pack(full))

    assert abs(lp.energy(liquid) - olh.energy(pos)) <= 1e-10 * abs(olh.energy(pos))
    assert rel_err(pack_liquid(lp.metric(liquid, tl)), fz.metric(pl, fz.remove(lay.pack(tan)))) < 1e-10
    eta = rng.standard_normal(c["shape"])
    assert rel_err(pack_liquid(lp.left_sqrt_metric(liquid, eta)), fz.lsm(pl, eta)) < 1e-10
    assert rel_err(t2n(lp.right_sqrt_metric(liquid, tl)), fz.rsm(pl, fz.remove(lay.pack(tan)))) < 1e-10
    assert rel_err(t2n(lp.transformation(liquid)), fz.trafo(pl)) < 1e-10
    e, gr = lp.energy_and_gradient(liquid)
    assert set(gr) == set(liquid)
    import pytest
    with pytest.raises(ValueError):
        lp.energy(tp)                                   # a full position is not a liquid one
    with pytest.raises(ValueError):
        lh.freeze(primals=tp, point_estimates=("nope",))
