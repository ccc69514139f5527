"""Multi-GPU check (run under torchrun on N B200s): NCCL sample sharding of draw_linear_samples + KL
value/gradient/metric equals the single-GPU result computed by rank 0.  Prints one line per check."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nifty_b200 as nb  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = (256, 256)
    cfm = nb.CorrelatedFieldMaker("cf")
    cfm.set_amplitude_total_offset(0.0, (1e-3, 1e-4))
    cfm.add_fluctuations(shape, 1.0 / shape[0], fluctuations=(1e-1, 5e-3), loglogavgslope=(-1.0, 1e-2), flexibility=(1.0, 0.5),
                         asperity=(0.5, 0.05), prefix="ax1", non_parametric_kind="power")
    sig = nb.SignalModel(cfm.finalize(), "exp")
    rt = sig.rt
    truth = sig.layout.random(42, torch.float64, rt.device)
    lh0 = nb.Gaussian(torch.zeros(shape, dtype=torch.float64), noise_cov_inv=100.0).amend(sig)
    data = lh0.signal_response(truth) + 0.1 * torch.randn(shape, dtype=torch.float64, device=rt.device,
                                                          generator=torch.Generator(rt.device).manual_seed(43))
    lh = nb.Gaussian(data, noise_cov_inv=100.0).amend(sig)
    pos = 0.1 * sig.layout.random(44, torch.float64, rt.device)
    keys = nb.random_split(7, 2 * world)
    cgkw = dict(cg_kwargs=dict(absdelta=1e-6, maxiter=100))
    vi = nb.OptimizeVI(lh, 1, comm=True)
    samples, _ = vi.draw_linear_samples(pos, keys, **cgkw)
    v, g = vi.kl_value_and_grad(pos, samples.residuals)
    t = sig.layout.random(9, torch.float64, rt.device)
    m = vi.kl_metric(t)
    gathered = [torch.empty_like(samples.residuals) for _ in range(world)]
    dist.all_gather(gathered, samples.residuals.contiguous())
    ok = True
    if rank == 0:
        vi1 = nb.OptimizeVI(lh, 1)
        s1, _ = vi1.draw_linear_samples(pos, keys, **cgkw)
        v1, g1 = vi1.kl_value_and_grad(pos, s1.residuals)
        m1 = vi1.kl_metric(t)
        glob = torch.empty_like(s1.residuals)
        for r in range(world):                      # rank r holds keys r, r+W, ... as interleaved mirrored pairs
            for i in range(len(keys[r::world])):
                k = r + i * world
                glob[2 * k:2 * k + 2] = gathered[r][2 * i:2 * i + 2]
        e_s = float((glob - s1.residuals).abs().max() / s1.residuals.abs().max())
        e_v = abs(v - v1) / abs(v1)
        e_g = float((g - g1).abs().max() / g1.abs().max())
        e_m = float((m - m1).abs().max() / m1.abs().max())
        ok = max(e_s, e_v, e_g, e_m) < 1e-10
        print(f"dist_check world={world}: samples {e_s:.2e} kl {e_v:.2e} grad {e_g:.2e} metric {e_m:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
