"""Multi-process path on CPU: world_size-2 gloo, emulator runtime.  The sample-sharded driver must
give the same KL value / gradient / metric and the same set of samples as a single process."""
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


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nifty_b200 as nb
    from nifty_b200._capi import CApi
    from emu.build_emu import build
    import vi_checks as vc
    rt = nb.Runtime(CApi(build()), "cpu")
    c, g, lh, olh, lay = vc._setup(rt, "g2d_16x16")
    pos = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    vi = nb.OptimizeVI(lh, 1, comm=True)
    keys = nb.random_split(123, 4)
    samples, infos = vi.draw_linear_samples(pos, keys, cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    v, gr = vi.kl_value_and_grad(pos, samples.residuals)
    t = lh.layout.random(9, torch.float64, rt.device)
    m = vi.kl_metric(t)
    # one full (sharded) optimize_kl iteration incl. checkpoint gather
    s2, st = nb.optimize_kl(lh, pos, key=5, n_total_iterations=1, n_samples=2, odir=os.path.join(outdir, "ckpt"), comm=True,
                            sample_mode="linear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
                            kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=30))))
    # n_samples == world / 2: every key is drawn by two ranks, the odd one keeps the mirrored sample (optimize_kl.py:403-405)
    sm, _ = vi.draw_linear_samples(pos, nb.random_split(77, 1), cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    vm, gm = vi.kl_value_and_grad(pos, sm.residuals)
    s3, st3 = nb.optimize_kl(lh, pos, key=6, n_total_iterations=1, n_samples=1, odir=os.path.join(outdir, "ckpt_m"), comm=True,
                             sample_mode="nonlinear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
                             nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))),
                             kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=30))))
    # fixed-tree reductions: KL value / gradient / metric independent of the number of ranks, bit for bit
    vt = nb.OptimizeVI(lh, 1, comm=True, kl_reduce="fixed_tree")
    st_smp, _ = vt.draw_linear_samples(pos, nb.random_split(123, 4), cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    v_t, g_t = vt.kl_value_and_grad(pos, st_smp.residuals)
    m_t = vt.kl_metric(t)
    s4, _ = nb.optimize_kl(lh, pos, key=5, n_total_iterations=1, n_samples=2, comm=True, kl_reduce="fixed_tree",
                           sample_mode="linear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=10))))
    np.savez(os.path.join(outdir, f"tree{rank}.npz"), v=v_t, g=g_t.numpy(), m=m_t.numpy(), pos4=s4.pos.numpy())
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), res=samples.residuals.numpy(), v=v, g=gr.numpy(), m=m.numpy(),
             pos2=s2.pos.numpy(), nloc=len(samples), res_m=sm.residuals.numpy(), vm=vm, gm=gm.numpy(), pos3=s3.pos.numpy(),
             res3=s3.residuals.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sample_sharding_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # single-process reference in this process
    sys.path.insert(0, HERE)
    import nifty_b200 as nb
    from nifty_b200._capi import CApi
    from emu.build_emu import build
    import vi_checks as vc
    rt = nb.Runtime(CApi(build()), "cpu")
    c, g, lh, olh, lay = vc._setup(rt, "g2d_16x16")
    pos = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    vi = nb.OptimizeVI(lh, 1)
    keys = nb.random_split(123, 4)
    samples, _ = vi.draw_linear_samples(pos, keys, cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    v, gr = vi.kl_value_and_grad(pos, samples.residuals)
    m = vi.kl_metric(lh.layout.random(9, torch.float64, rt.device))
    assert r[0]["nloc"] == 4 and r[1]["nloc"] == 4            # 2 keys x mirrored pair per rank
    # rank r holds keys r, r+2 -> global order [k0, k1, k2, k3] = [r0[0:2], r1[0:2], r0[2:4], r1[2:4]]
    glob = np.concatenate([r[0]["res"][0:2], r[1]["res"][0:2], r[0]["res"][2:4], r[1]["res"][2:4]])
    np.testing.assert_allclose(glob, samples.residuals.numpy(), rtol=0, atol=1e-13)
    for i in range(world):
        assert abs(r[i]["v"] - v) <= 1e-12 * abs(v)
        np.testing.assert_allclose(r[i]["g"], gr.numpy(), rtol=0, atol=1e-12 * np.abs(gr.numpy()).max())
        np.testing.assert_allclose(r[i]["m"], m.numpy(), rtol=0, atol=1e-12 * np.abs(m.numpy()).max())
    np.testing.assert_array_equal(r[0]["pos2"], r[1]["pos2"])   # replicated position stays bit-identical across ranks
    assert os.path.isfile(tmp_path / "ckpt" / "last.pkl")
    # fixed tree: world 2 == world 1, bit for bit
    vt = nb.OptimizeVI(lh, 1, kl_reduce="fixed_tree")
    st_smp, _ = vt.draw_linear_samples(pos, nb.random_split(123, 4), cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    v_t, g_t = vt.kl_value_and_grad(pos, st_smp.residuals)
    m_t = vt.kl_metric(lh.layout.random(9, torch.float64, rt.device))
    s4, _ = nb.optimize_kl(lh, pos, key=5, n_total_iterations=1, n_samples=2, kl_reduce="fixed_tree",
                           sample_mode="linear_resample", draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=10))))
    for i in range(world):
        tr = np.load(tmp_path / f"tree{i}.npz")
        assert float(tr["v"]) == v_t
        np.testing.assert_array_equal(tr["g"], g_t.numpy())
        np.testing.assert_array_equal(tr["m"], m_t.numpy())
        np.testing.assert_array_equal(tr["pos4"], s4.pos.numpy())
    assert abs(v_t - v) <= 1e-12 * abs(v)                 # and it is the same KL as with the all-reduce
    # mirror rule: rank 0 holds +s0, rank 1 holds -s0; the KL equals the single-process KL over [s0, -s0]
    sm, _ = vi.draw_linear_samples(pos, nb.random_split(77, 1), cg_kwargs=dict(absdelta=1e-8, maxiter=60))
    vm, gm = vi.kl_value_and_grad(pos, sm.residuals)
    assert r[0]["res_m"].shape[0] == 1 and r[1]["res_m"].shape[0] == 1
    np.testing.assert_allclose(np.concatenate([r[0]["res_m"], r[1]["res_m"]]), sm.residuals.numpy(), rtol=0, atol=1e-13)
    for i in range(world):
        assert abs(r[i]["vm"] - vm) <= 1e-12 * abs(vm)
        np.testing.assert_allclose(r[i]["gm"], gm.numpy(), rtol=0, atol=1e-12 * np.abs(gm.numpy()).max())
    s3, st3 = nb.optimize_kl(lh, pos, key=6, n_total_iterations=1, n_samples=1, sample_mode="nonlinear_resample",
                             draw_linear_kwargs=dict(cg_kwargs=dict(absdelta=1e-6, maxiter=60)),
                             nonlinearly_update_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(maxiter=20))),
                             kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=3, cg_kwargs=dict(maxiter=30))))
    np.testing.assert_allclose(r[0]["pos3"], s3.pos.numpy(), rtol=0, atol=1e-9 * np.abs(s3.pos.numpy()).max())
    np.testing.assert_allclose(np.concatenate([r[0]["res3"], r[1]["res3"]]), s3.residuals.numpy(), rtol=0,
                               atol=1e-8 * np.abs(s3.residuals.numpy()).max())
    import pickle
    with open(tmp_path / "ckpt_m" / "last.pkl", "rb") as f:
        ck, _ = pickle.load(f)
    assert next(iter(ck.residuals.values())).shape[0] == 2


def _host_composed_run(comm, which="nonpow2"):
    """KL value / gradient / metric over 4 sample points and one optimize_kl iteration on a host-composed field."""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import nifty_b200 as nb
    from nifty_b200._capi import CApi
    from emu.build_emu import build
    import vi_checks as vc
    rt = nb.Runtime(CApi(build()), "cpu")
    lh = vc._host_composed_pair(rt, which, "gauss")[0]
    pos = 0.1 * lh.layout.random(7, torch.float64, rt.device)
    vi = nb.OptimizeVI(lh, 1, comm=comm)
    cgkw = dict(absdelta=1e-30, miniter=8, maxiter=8)
    samples, _ = vi.draw_linear_samples(pos, nb.random_split(123, 2), cg_kwargs=cgkw)
    v, gr = vi.kl_value_and_grad(pos, samples.residuals)
    m = vi.kl_metric(lh.layout.random(9, torch.float64, rt.device))
    s2, _ = nb.optimize_kl(lh, pos, key=5, n_total_iterations=1, n_samples=2, comm=comm, sample_mode="linear_resample",
                           draw_linear_kwargs=dict(cg_kwargs=cgkw),
                           kl_kwargs=dict(minimize_kwargs=dict(xtol=1e-4, maxiter=2, cg_kwargs=dict(absdelta=1e-30, miniter=5, maxiter=5))))
    return dict(v=v, g=gr.numpy(), m=m.numpy(), pos2=s2.pos.numpy(), nloc=len(samples))


def _worker_host_composed(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    np.savez(os.path.join(outdir, f"hc{rank}.npz"), **_host_composed_run(True))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sample_sharding_of_host_composed_fields(tmp_path):
    """Sample sharding over two ranks for a field on a non-power-of-two grid (host-applied operators, the KL metric summed by the
    same all-reduce): same KL value / gradient / metric and the same new position as a single process."""
    world = 2
    mp.spawn(_worker_host_composed, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    one = _host_composed_run(None)
    for i in range(world):
        r = np.load(tmp_path / f"hc{i}.npz")
        assert r["nloc"] == 2 and one["nloc"] == 4
        assert abs(float(r["v"]) - one["v"]) <= 1e-12 * abs(one["v"])
        np.testing.assert_allclose(r["g"], one["g"], rtol=0, atol=1e-12 * np.abs(one["g"]).max())
        np.testing.assert_allclose(r["m"], one["m"], rtol=0, atol=1e-12 * np.abs(one["m"]).max())
        np.testing.assert_allclose(r["pos2"], one["pos2"], rtol=0, atol=1e-8 * np.abs(one["pos2"]).max())
