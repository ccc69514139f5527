"""Slab-decomposed 3-D grids on CPU: world_size 2 and 4 with gloo, the kernels on the host
emulator.  Every rank checks its part of signal / energy / metric / left-sqrt-metric against the oracle
evaluated on the GLOBAL grid."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def slab_check(rt, shape, dist_, lh_kind="gauss", seed=3, tol=1e-10):
    """Runs inside an initialised process group; returns the max relative errors found."""
    import nifty_b200 as nb
    import oracle
    c = dict(shape=shape, distances=dist_, offset_mean=0.3, offset_std=(0.2, 0.1), fluctuations=(0.5, 0.1),
             loglogavgslope=(-2.0, 0.3), flexibility=(1.0, 0.5), asperity=(0.5, 0.05))
    ocf = oracle.CorrelatedFieldOracle("cf")
    ocf.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    ocf.add_fluctuations(shape, dist_, c["fluctuations"], c["loglogavgslope"], c["flexibility"], c["asperity"], prefix="ax1",
                         non_parametric_kind="power")
    ocf.finalize()
    osig = oracle.SignalOracle(ocf, "exp")
    lay = oracle.Layout(osig.domain)
    rng = np.random.default_rng(seed)
    pos = {k: 0.5 * v for k, v in lay.random(rng).items()}
    tan = lay.random(rng)
    if lh_kind == "gauss":
        data = osig(pos) + 0.3 * rng.standard_normal(shape)
        olh = oracle.GaussianOracle(data, 1.0 / 0.09, osig)
    else:
        data = rng.poisson(osig(pos)).astype(np.int64)
        olh = oracle.PoissonianOracle(data, osig)
    u = rng.standard_normal(shape)

    cfm = nb.CorrelatedFieldMaker("cf", runtime=rt, comm=True)
    cfm.set_amplitude_total_offset(c["offset_mean"], c["offset_std"])
    cfm.add_fluctuations(shape, dist_, c["fluctuations"], c["loglogavgslope"], c["flexibility"], c["asperity"], prefix="ax1",
                         non_parametric_kind="power")
    cf = cfm.finalize()
    plan = cf.plan
    assert plan.dist
    sig = nb.SignalModel(cf, "exp")
    dloc = plan.scatter_position(data.astype(np.float64))
    if lh_kind == "gauss":
        lh = nb.Gaussian(dloc, noise_cov_inv=1.0 / 0.09).amend(sig)
    else:
        lh = nb.Poissonian(np.rint(dloc.cpu().numpy()).astype(np.int64)).amend(sig)

    def localise(tree):
        d = {k: torch.as_tensor(v) for k, v in tree.items()}
        d["cfxi"] = plan.scatter_latent(tree["cfxi"])
        return sig.layout.pack(d, torch.float64, rt.device)

    def globalise(vec):
        d = sig.layout.unpack(vec)
        out = {k: v.cpu().numpy() for k, v in d.items() if k != "cfxi"}
        out["cfxi"] = plan.gather_latent(d["cfxi"]).cpu().numpy()
        return out

    errs = {}
    pl, tl = localise(pos), localise(tan)
    lin, _ = lh.lin_at(pl)
    s_glob = plan.gather_position(lin.signal()).cpu().numpy()
    ref = osig(pos)
    errs["signal"] = float(np.max(np.abs(s_glob - ref)) / np.max(np.abs(ref)))
    e, oe = lin.energy(), olh.energy(pos)
    errs["energy"] = abs(e - oe) / abs(oe)
    met = globalise(lin.metric(tl))
    omet = olh.metric(pos, tan)
    scale = max(np.max(np.abs(v)) for v in omet.values())
    errs["metric"] = max(float(np.max(np.abs(met[k] - omet[k]))) / scale for k in omet)
    met1 = globalise(lin.metric(tl, add_identity=True))
    errs["metric+1"] = max(float(np.max(np.abs(met1[k] - omet[k] - tan[k]))) / scale for k in omet)
    ls = globalise(lin.lsm(plan.scatter_position(u), scaled=True))
    ols = olh.left_sqrt_metric(pos, u)
    scale = max(np.max(np.abs(v)) for v in ols.values())
    errs["lsm"] = max(float(np.max(np.abs(ls[k] - ols[k]))) / scale for k in ols)
    # right sqrt-metric, transformation and residual live on the local position planes
    rs = plan.gather_position(lin.rsm(tl, scaled=True)).cpu().numpy()
    ors = olh.right_sqrt_metric(pos, tan)
    errs["rsm"] = float(np.max(np.abs(rs - ors)) / np.max(np.abs(ors)))
    tr = plan.gather_position(lin.transformation()).cpu().numpy()
    otr = olh.transformation(pos)
    errs["transformation"] = float(np.max(np.abs(tr - otr)) / np.max(np.abs(otr)))
    nr = plan.gather_position(lin.normalized_residual()).cpu().numpy()
    onr = olh.normalized_residual(pos)
    errs["normalized_residual"] = float(np.max(np.abs(nr - onr)) / np.max(np.abs(onr)))
    ee, gg = lh.energy_and_gradient(pl, add_prior=True)
    oe2, og = olh.energy_and_gradient(pos)
    flat = lay.pack(pos)
    assert abs(ee - (oe2 + 0.5 * flat @ flat)) <= tol * abs(oe2 + 0.5 * flat @ flat)
    gg = globalise(gg)
    scale = max(np.max(np.abs(og[k] + pos[k])) for k in og)
    errs["gradient"] = max(float(np.max(np.abs(gg[k] - og[k] - pos[k]))) / scale for k in og)
    # MGVI sample draw: distributed CG (host recurrences, all-reduced dot products) against the oracle's
    if lh_kind == "gauss":
        wd, wp = rng.standard_normal(shape), lay.random(rng)
        cgkw = dict(absdelta=1e-30, maxiter=8, miniter=9)     # exactly 8 iterations (rounding chaos grows with the count)
        ores, oinfo, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=cgkw)
        res, info = nb.draw_linear_residual(lh, pl, 0, cg_kwargs=cgkw, _white=(plan.scatter_position(wd), localise(wp)))
        got = globalise(res)
        scale = max(np.max(np.abs(v)) for v in ores.values())
        errs["mgvi_draw"] = max(float(np.max(np.abs(got[k] - ores[k]))) / scale for k in ores)
        assert info == oinfo, (info, oinfo, errs)
        # geoVI update of that sample (evi.py:181-255) with distributed reductions: Newton-CG on the pair operator
        mk = dict(xtol=1e-8, maxiter=2, cg_kwargs=dict(absdelta=1e-30, maxiter=6, miniter=7))
        onew, oopt = oracle.nonlinearly_update_residual(olh, pos, ores, wd, wp, 1.0, minimize_kwargs=mk)
        new, opt = nb.nonlinearly_update_residual(lh, pl, res, 0, 1.0, minimize_kwargs=mk,
                                                  _white=(plan.scatter_position(wd), localise(wp)))
        gnew = globalise(new)
        scale = max(np.max(np.abs(v)) for v in onew.values())
        errs["geovi_update"] = max(float(np.max(np.abs(gnew[k] - onew[k]))) / scale for k in onew)
        assert opt.nit == oopt.nit and opt.status == oopt.status, (opt.nit, oopt.nit, opt.status, oopt.status)
        # minisanity on the slab-decomposed field: all-reduced moments equal the global ones
        smp0 = nb.Samples(pos=pl, samples=torch.stack((res, -res)))
        st_p, _ = nb.minisanity(smp0, plan=plan, layout=lh.layout, dist_leaves=("cfxi",))
        pts = [lay.pack(pos) + sgn * lay.pack(ores) for sgn in (1.0, -1.0)]
        xi = np.stack([lay.unpack(p)["cfxi"].reshape(-1) for p in pts])
        rx = (xi * xi).sum(1) / xi.shape[1]
        assert st_p["cfxi"].ndof == xi.shape[1] and abs(st_p["cfxi"].reduced_chisq[0] - rx.mean()) <= 1e-7 * rx.mean()
        st_r, _ = nb.minisanity(smp0, lh.normalized_residual, plan=plan)
        rr = np.stack([olh.normalized_residual(lay.unpack(p)).reshape(-1) for p in pts])
        rr2 = (rr * rr).sum(1) / rr.shape[1]
        assert st_r.ndof == rr.shape[1] and abs(st_r.reduced_chisq[0] - rr2.mean()) <= 1e-7 * rr2.mean()
        # point estimates on the slab-decomposed field: a replicated leaf and (separately) the distributed excitations
        for pe in (("cfax1fluctuations", "cfax1spectrum"), ("cfxi",)):
            ofr, ofi, _ = oracle.draw_linear_residual(olh, pos, wd, wp, cg_kwargs=cgkw, point_estimates=pe)
            rfr, rfi = nb.draw_linear_residual(lh, pl, 0, cg_kwargs=cgkw, point_estimates=pe,
                                               _white=(plan.scatter_position(wd), localise(wp)))
            gfr = globalise(rfr)
            scale = max(np.max(np.abs(v)) for v in ofr.values())
            errs["frozen_draw_" + pe[0]] = max(float(np.max(np.abs(gfr[k] - ofr[k]))) / scale for k in ofr)
            assert rfi == ofi and all(float(np.max(np.abs(gfr[k]))) == 0.0 for k in pe)
        # stochastic draw path (per-rank keys): runs, hyper-parameter leaves identical on all ranks
        r2, _ = nb.draw_linear_residual(lh, pl, 123, cg_kwargs=cgkw)
        hyp = torch.cat((r2[:lh._xi_slice()[0]], r2[lh._xi_slice()[1]:])).to(torch.float64)
        allh = plan.comm.all_gather(hyp)
        assert all(torch.equal(allh[0], h) for h in allh)
        # one KL (Newton-CG) step over the mirrored pair of that sample: energy decreases, replicated leaves stay identical
        vi = nb.OptimizeVI(lh, 1)
        smp = nb.Samples(pos=pl, samples=torch.stack((r2, -r2)), keys=None)
        e0, _ = vi.kl_value_and_grad(pl, smp.residuals)
        opt = vi.kl_minimize(smp, minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=10)))
        e1, _ = vi.kl_value_and_grad(opt.x, smp.residuals)
        assert e1 < e0
        hyp = torch.cat((opt.x[:lh._xi_slice()[0]], opt.x[lh._xi_slice()[1]:])).to(torch.float64)
        allh = plan.comm.all_gather(hyp)
        assert all(torch.equal(allh[0], h) for h in allh)
    if lh_kind == "gauss":
        # whole VI iteration in the reference's default mode (geoVI) on the slab-decomposed field: runs, replicated leaves agree
        smp_kl, st_kl = nb.optimize_kl(lh, pl.clone(), key=11, n_total_iterations=1, n_samples=1,
                                       draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=10)),
                                       nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-3, maxiter=1, cg_kwargs=dict(maxiter=5))),
                                       kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-3, maxiter=2, cg_kwargs=dict(maxiter=5))),
                                       sample_mode="nonlinear_resample")
        assert st_kl.nit == 1 and len(smp_kl) == 2
        hyp = torch.cat((smp_kl.pos[:lh._xi_slice()[0]], smp_kl.pos[lh._xi_slice()[1]:])).to(torch.float64)
        allh = plan.comm.all_gather(hyp)
        assert all(torch.equal(allh[0], h) for h in allh)
        msg = nb.OptimizeVI(lh, 1).get_status_message(smp_kl, st_kl, name="T")
        assert "Likelihood residual(s)" in msg and "cfxi" in msg
    # CG trajectories amplify rounding differences (8 iterations here); the hyper-parameter-only system (frozen excitations)
    # is the worst conditioned of them
    lim = lambda k: 1e-5 if k == "frozen_draw_cfxi" else 1e-7 if k in ("mgvi_draw", "geovi_update") or k.startswith("frozen_draw") else tol
    bad = {k: v for k, v in errs.items() if not v < lim(k)}
    assert not bad, bad
    return errs


def _worker(rank, world, port, shape, chunks, lg_r1=None):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NB200_SLAB_CHUNKS=str(chunks))
    if lg_r1 is not None:          # several lines per CTA in the first pass -> the mirror-quad body (P1MBody)
        os.environ["NB200_LGR1"] = str(lg_r1)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nifty_b200 as nb
    from nifty_b200._capi import CApi
    from emu.build_emu import build
    rt = nb.Runtime(CApi(build()), "cpu")
    slab_check(rt, shape, (0.2, 0.1, 0.05))
    slab_check(rt, shape, 0.1, lh_kind="poisson", seed=5)
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,shape,chunks", [(2, (8, 4, 16), 1), (4, (16, 8, 8), 1), (2, (4, 4, 4), 1),
                                                (2, (16, 4, 16), 2), (4, (32, 4, 16), 3), (2, (8, 8, 8), 4)])
def test_slab_decomposition_matches_global_oracle(world, shape, chunks):
    """chunks > 1: every exchange is pipelined in point-to-point pieces between chunked pass launches."""
    mp.spawn(_worker, args=(world, _free_port(), shape, chunks), nprocs=world, join=True)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,shape,chunks,lg_r1", [(2, (8, 4, 16), 1, 1), (2, (16, 8, 8), 2, 2)])
def test_slab_with_mirror_quad_first_pass(world, shape, chunks, lg_r1):
    """The launch heuristic gives the first pass >= 2 lines per CTA only on large grids; force it here so that the
    mirror-quad body runs on the local slabs of a distributed plan."""
    mp.spawn(_worker, args=(world, _free_port(), shape, chunks, lg_r1), nprocs=world, join=True)
