import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nifty_b200 as nb, oracle, parity_checks as pc
from golden_util import CASES, load, build_oracle_lh, rel_err
rt = nb.default_runtime()
name = "g3d_8x8x8"
c, g = CASES[name], load(name)
lh = pc.build_product_lh(c, g, rt); olh = build_oracle_lh(c, g); lay = oracle.Layout(olh.domain)
pos_v = lay.pack(g["pos"]); rng = np.random.default_rng(5)
j, x0 = rng.standard_normal(lay.size), rng.standard_normal(lay.size)
mat = lambda v: lay.pack(olh.metric(g["pos"], lay.unpack(v))) + v
lin, _ = lh.lin_at(torch.as_tensor(pos_v))
tj, tx0 = rt.asarray(j, torch.float64), rt.asarray(x0, torch.float64)
for k in (5, 19, 20, 21, 23):
    for ce in (1, 3):
        ores = oracle.cg(mat, j, x0=x0, absdelta=1e-30, maxiter=k, miniter=k)
        x, res = lin.cg_solve(tj, tx0, absdelta=1e-30, maxiter=k, miniter=k, check_every=ce)
        print(k, ce, "oracle", ores.nit, ores.info, ores.nfev, "dev", res.nit, res.info, res.nfev, res.error, rel_err(x.cpu().numpy(), ores.x), res.energy, ores.energies[-1])
