"""CPU tier: the solver half of the oracle (oracle/solvers.py) is pinned against the REFERENCE'S OWN `_cg` / `_newton_cg`.

tests/golden/solvers.json holds the outputs of the unmodified reference routines
(`/root/reference/nifty/re/conjugate_gradient.py:77-214`, `optimize.py:271-411`, executed through
tests/golden/ref_solvers.py by make_solver_golden.py) on the seeded problems of tests/golden/solver_cases.py.
The oracle must reproduce iteration counts, `info` / `status`, function-evaluation counts exactly and the
solutions to 1e-12.  Where the reference tree is present the fixtures are regenerated live and compared too.
Also restated: the reference's known-answer tests (test/test_re/test_ncg.py:81-124, 232-247)."""
import json
import os
import sys

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import ref_solvers  # noqa: E402
import solver_cases  # noqa: E402

with open(os.path.join(GOLD, "solvers.json")) as _f:
    FIX = json.load(_f)


@pytest.mark.parametrize("name", sorted(solver_cases.cg_cases()))
def test_cg_matches_reference(name):
    a, j, x0, kw = solver_cases.cg_cases()[name]
    want = FIX["cg"][name]
    res = oracle.cg(lambda v: a @ v, j, x0=x0, **kw)
    assert (res.nit, res.nfev, res.info) == (want["nit"], want["nfev"], want["info"]), name
    np.testing.assert_allclose(res.x, np.array(want["x"]), rtol=1e-12, atol=1e-12 * max(1.0, float(np.max(np.abs(want["x"])))))


@pytest.mark.parametrize("name", sorted(solver_cases.newton_cases()))
def test_newton_cg_matches_reference(name):
    fg, hp, x0, kw = solver_cases.newton_cases()[name]
    want = FIX["newton"][name]
    res = oracle.newton_cg(x0, fg, hp, **kw)
    assert (res.status, res.nit, res.nfev, res.njev, res.nhev) == (want["status"], want["nit"], want["nfev"], want["njev"], want["nhev"]), name
    np.testing.assert_allclose(res.x, np.array(want["x"]), rtol=1e-12, atol=1e-12)
    assert abs(res.fun - want["fun"]) <= 1e-12 * max(1.0, abs(want["fun"]))


@pytest.mark.skipif(not ref_solvers.available(), reason="reference tree not present (GPU box): fixtures only")
def test_fixtures_are_what_the_reference_computes():
    import make_solver_golden
    live = make_solver_golden.run_reference()
    for kind in ("cg", "newton"):
        assert sorted(live[kind]) == sorted(FIX[kind])
        for name, want in FIX[kind].items():
            got = live[kind][name]
            for k, v in want.items():
                if k == "x":
                    np.testing.assert_allclose(got[k], v, rtol=0, atol=0)
                else:
                    assert got[k] == v, (kind, name, k)


# --- the reference's own known-answer tests, restated on the oracle ------------------------------------
@pytest.mark.parametrize("seed", (3637, 12, 42))
def test_cg_known_answer(seed):            # test/test_re/test_ncg.py:112-124
    r = np.random.default_rng(seed)
    x, diag = r.standard_normal(3), 6.0 + r.standard_normal(3)
    res = oracle.cg(lambda v: v / diag, x, resnorm=1e-5, absdelta=1e-5)
    np.testing.assert_allclose(res.x, diag * x, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("seed", (3637, 12, 42))
def test_cg_non_pos_def_failure(seed):     # test/test_re/test_ncg.py:158-171
    r = np.random.default_rng(seed)
    x = r.standard_normal(4)
    diag = np.concatenate(([-1.0], 6.0 + r.standard_normal(3)))
    with pytest.raises(ValueError):
        oracle.cg(lambda v: v / diag, x, resnorm=1e-5, absdelta=1e-5)


@pytest.mark.parametrize("seed", (3637, 12, 42))
def test_ncg_known_answer(seed):           # test/test_re/test_ncg.py:81-92
    r = np.random.default_rng(seed)
    x, diag = r.standard_normal(3), np.array([1.0, 2.0, 3.0])
    res = oracle.newton_cg(x, lambda y: (float(np.sum(y ** 2 / diag) / 2 - np.dot(x, y)), y / diag - x), lambda y, t: t / diag,
                           maxiter=20, absdelta=1e-6)
    np.testing.assert_allclose(res.x, diag * x, rtol=1e-4, atol=1e-4)


def test_minimize_ncg_vs_scipy_trust_ncg():     # test/test_re/test_ncg.py:232-247 (rosenbrock from zeros(2))
    from scipy.optimize import minimize as opt_minimize
    fg, hp = solver_cases._fd_free_rosen()
    res = oracle.newton_cg(np.zeros(2), fg, hp, xtol=1e-6, energy_reduction_factor=None)
    ref = opt_minimize(lambda x: fg(x)[0], np.zeros(2), jac=lambda x: fg(x)[1], hessp=hp, method="trust-ncg").x
    np.testing.assert_allclose(res.x, ref, rtol=2e-6, atol=2e-5)
