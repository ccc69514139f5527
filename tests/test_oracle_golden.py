"""Pin the oracle against the reference: fixtures produced by the unmodified ``nifty.cl``
(tests/golden/make_golden.py).  Mirrors test/test_re/test_correlated_field.py:116-192 of the
reference (re == cl field values) and extends it to JVP, VJP, energy, gradient and metric."""
import numpy as np
import pytest

from golden_util import CASES, OUTER_CASES, build_oracle, build_oracle_lh, build_outer, load, rel_err

TOL = 1e-11


@pytest.mark.parametrize("name", sorted(CASES))
def test_field_jvp_vjp(name):
    c, g = CASES[name], load(name)
    cf = build_oracle(c)
    assert rel_err(cf(g["pos"]), g["field"]) < TOL
    assert rel_err(cf.jvp(g["pos"], g["tan"]), g["field_jvp"]) < TOL
    vjp = cf.vjp(g["pos"], g["cot"])
    assert set(vjp) == set(g["field_vjp"])
    for k in vjp:
        assert vjp[k].shape == g["field_vjp"][k].shape, k
        assert rel_err(vjp[k], g["field_vjp"][k]) < TOL, k


@pytest.mark.parametrize("name", sorted(CASES))
def test_energy_grad_metric(name):
    c, g = CASES[name], load(name)
    lh = build_oracle_lh(c, g)
    assert rel_err(lh.signal(g["pos"]), g["signal"]) < TOL
    e, grad = lh.energy_and_gradient(g["pos"])
    assert abs(e - float(g["energy"])) <= TOL * abs(float(g["energy"]))
    # per-leaf error relative to the largest entry of the whole output vector: leaves that are
    # structurally zero (e.g. `spectrum` on a 3x3 grid) only carry cancellation noise
    scale = max(np.max(np.abs(v)) for v in g["grad"].values())
    for k in grad:
        assert np.max(np.abs(grad[k] - g["grad"][k])) < 1e-10 * scale, k
    met = lh.metric(g["pos"], g["tan"])
    scale = max(np.max(np.abs(v)) for v in g["metric"].values())
    for k in met:
        assert np.max(np.abs(met[k] - g["metric"][k])) < 1e-10 * scale, k


@pytest.mark.parametrize("name", sorted(OUTER_CASES))
def test_outer_product_field_jvp_vjp(name):
    """Outer products of two sub-grids: the oracle against the unmodified nifty.cl (the reference's own re == cl check,
    test/test_re/test_correlated_field.py:245-283, extended to the JVP and the VJP)."""
    from oracle import CorrelatedFieldOracle
    c, g = OUTER_CASES[name], load(name)
    cf = build_outer(c, CorrelatedFieldOracle("cf"))
    assert rel_err(cf(g["pos"]), g["field"]) < TOL
    assert rel_err(cf.jvp(g["pos"], g["tan"]), g["field_jvp"]) < TOL
    vjp = cf.vjp(g["pos"], g["cot"])
    assert set(vjp) == set(g["field_vjp"])
    scale = max(np.max(np.abs(v)) for v in g["field_vjp"].values())
    for k in vjp:
        assert np.max(np.abs(vjp[k] - g["field_vjp"][k])) < 1e-10 * scale, k
