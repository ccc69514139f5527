"""TEST INFRASTRUCTURE: writes tests/golden/solvers.json -- outputs of the reference's unmodified `_cg` / `_newton_cg`
(executed through ref_solvers.py, this container only) on the seeded problems of solver_cases.py.
Run: python tests/golden/make_solver_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_solvers  # noqa: E402
import solver_cases  # noqa: E402


def run_reference():
    cgm, opt = ref_solvers.load()
    out = {"cg": {}, "newton": {}}
    for name, (a, j, x0, kw) in solver_cases.cg_cases().items():
        r = cgm._cg(lambda v, a=a: a @ v, j, x0, **kw)
        out["cg"][name] = dict(x=np.asarray(r.x, dtype=np.float64).tolist(), nit=int(r.nit), nfev=int(r.nfev), info=int(r.info))
    for name, (fg, hp, x0, kw) in solver_cases.newton_cases().items():
        r = opt._newton_cg(None, x0, fun_and_grad=fg, hessp=hp, **kw)
        out["newton"][name] = dict(x=np.asarray(r.x, dtype=np.float64).tolist(), status=int(r.status), nit=int(r.nit), nfev=int(r.nfev),
                                   njev=int(r.njev), nhev=int(r.nhev), fun=float(r.fun))
    return out


if __name__ == "__main__":
    res = run_reference()
    with open(os.path.join(HERE, "solvers.json"), "w") as f:
        json.dump(res, f, indent=1)
    print({k: len(v) for k, v in res.items()})
