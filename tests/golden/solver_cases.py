"""TEST INFRASTRUCTURE: seeded problems on which the oracle's `cg` / `newton_cg` are pinned against the reference's own
`_cg` / `_newton_cg` (see ref_solvers.py).  Shared by the fixture generator and the tests."""
import numpy as np


def spd(n, cond, seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.geomspace(1.0, cond, n)
    return (q * ev) @ q.T


def cg_cases():
    """name -> (matrix, j, x0 or None, kwargs)"""
    cases = {}
    rng = np.random.default_rng(7)
    a8, a40, a64 = spd(8, 30.0, 1), spd(40, 1e3, 2), spd(64, 1e5, 3)
    j8, j40, j64 = rng.standard_normal(8), rng.standard_normal(40), rng.standard_normal(64)
    x40 = rng.standard_normal(40)
    cases["default_tol"] = (a8, j8, None, {})
    cases["absdelta"] = (a40, j40, None, dict(absdelta=1e-6, maxiter=100))
    cases["absdelta_x0"] = (a40, j40, x40, dict(absdelta=1e-8, maxiter=100))
    cases["resnorm_l1"] = (a40, j40, x40, dict(resnorm=1e-3, norm_ord=1, maxiter=80))
    cases["resnorm_inf"] = (a40, j40, None, dict(resnorm=1e-2, norm_ord=np.inf, maxiter=80))
    cases["both"] = (a40, j40, None, dict(resnorm=1e-5, absdelta=1e-5))
    cases["miniter_maxiter"] = (a40, j40, None, dict(absdelta=1e-30, miniter=7, maxiter=7))
    cases["n_reset"] = (a64, j64, None, dict(absdelta=1e-30, miniter=45, maxiter=45))           # crosses iterations 20 and 40
    cases["n_reset_x0"] = (a64, j64, rng.standard_normal(64), dict(resnorm=1e-9, norm_ord=2, maxiter=63))
    cases["zero_rhs"] = (a8, np.zeros(8), None, {})
    ind = np.diag(np.concatenate(([-1.0], 6.0 + rng.standard_normal(7))))
    cases["negcurv_first"] = (np.diag(np.concatenate(([-3.0], -1.0 - rng.random(7)))), j8, None, dict(_raise_nonposdef=False, resnorm=1e-5))
    cases["negcurv_later"] = (ind, j8, None, dict(_raise_nonposdef=False, resnorm=1e-8, maxiter=30))
    # the reference's own test problem (test/test_re/test_ncg.py:112-124): diagonal system, resnorm + absdelta
    for seed in (3637, 12, 42):
        r = np.random.default_rng(seed)
        x, diag = r.standard_normal(3), 6.0 + r.standard_normal(3)
        cases[f"ref_test_cg_{seed}"] = (np.diag(1.0 / diag), x, None, dict(resnorm=1e-5, absdelta=1e-5))
    return cases


def rosenbrock(x):
    return float(np.sum(100.0 * np.diff(x) ** 2 + (1.0 - x[:-1]) ** 2))


def _fd_free_rosen():
    # f = sum 100 (x_{i+1} - x_i)^2 + (1 - x_i)^2   (the reference test's variant, test_ncg.py:19-20)
    def fg(x):
        d = np.diff(x)
        g = np.zeros_like(x)
        g[:-1] += -200.0 * d - 2.0 * (1.0 - x[:-1])
        g[1:] += 200.0 * d
        return rosenbrock(x), g

    def hessp(x, v):
        dv = np.diff(v)
        out = np.zeros_like(x)
        out[:-1] += -200.0 * dv + 2.0 * v[:-1]
        out[1:] += 200.0 * dv
        return out
    return fg, hessp


def newton_cases():
    """name -> (fun_and_grad, hessp, x0, kwargs)"""
    cases = {}
    for seed in (3637, 12, 42):        # test_ncg.py:81-92
        r = np.random.default_rng(seed)
        x = r.standard_normal(3)
        diag = np.array([1.0, 2.0, 3.0])
        cases[f"ref_test_ncg_{seed}"] = ((lambda y, x=x, diag=diag: (float(np.sum(y ** 2 / diag) / 2 - np.dot(x, y)), y / diag - x)),
                                         (lambda y, t, diag=diag: t / diag), x.copy(), dict(maxiter=20, absdelta=1e-6))
    fg, hp = _fd_free_rosen()
    cases["rosen_xtol"] = (fg, hp, np.zeros(2), dict(xtol=1e-6, energy_reduction_factor=None))      # test_ncg.py:232-247
    cases["rosen5_absdelta"] = (fg, hp, np.linspace(-1, 1, 5), dict(absdelta=1e-8, maxiter=50, cg_kwargs=dict(miniter=2)))
    # strictly convex log-cosh problem with a dense coupling (line search + CG on a non-quadratic energy)
    a = np.random.default_rng(5).standard_normal((12, 12)) / 3.0
    b = np.random.default_rng(6).standard_normal(12)

    def fg2(x):
        z = a @ x - b
        return float(np.sum(np.logaddexp(z, -z)) + 0.05 * np.dot(x, x)), a.T @ np.tanh(z) + 0.1 * x

    def hp2(x, v):
        z = a @ x - b
        return a.T @ ((1.0 - np.tanh(z) ** 2) * (a @ v)) + 0.1 * v
    cases["logcosh"] = (fg2, hp2, np.zeros(12), dict(xtol=1e-8, miniter=2, maxiter=30))
    cases["logcosh_old_fval"] = (fg2, hp2, 3.0 * np.ones(12), dict(absdelta=1e-10, old_fval=1e3, maxiter=30, norm_ord=2))
    return cases
