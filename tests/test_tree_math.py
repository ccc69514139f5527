"""Host-side tree arithmetic (nifty/re/tree_math semantics) -- no device needed."""
import numpy as np
import pytest
import torch

import nifty_b200 as nb


def _tree(rng):
    return {"b": torch.as_tensor(rng.standard_normal((3, 2))), "a": torch.as_tensor(rng.standard_normal(())),
            "c": torch.as_tensor(rng.standard_normal(5))}


def test_vdot_norm_size_zeros_where():
    rng = np.random.default_rng(0)
    x, y = _tree(rng), _tree(rng)
    flat = lambda t: np.concatenate([np.ravel(t[k].numpy()) for k in sorted(t)])
    assert abs(nb.vdot(x, y) - float(flat(x) @ flat(y))) < 1e-13
    assert nb.size(x) == 12
    # norm = ord-norm of the per-leaf ord-norms (vector_math.py:173-188)
    for o in (1, 2, np.inf):
        per_leaf = [abs(float(x["a"]))] + [np.linalg.norm(np.ravel(x[k].numpy()), ord=o) for k in ("b", "c")]
        assert abs(nb.norm(x, ord=o) - np.linalg.norm(per_leaf, ord=o)) < 1e-13
    assert abs(nb.norm(torch.as_tensor(flat(x)), 2) - np.linalg.norm(flat(x))) < 1e-13
    z = nb.zeros_like(x)
    assert all(float(z[k].abs().sum()) == 0 and z[k].shape == x[k].shape for k in x)
    w = nb.where({k: x[k] > 0 for k in x}, x, 0.0)
    assert all(torch.equal(w[k], torch.where(x[k] > 0, x[k], torch.zeros_like(x[k]))) for k in x)
    w2 = nb.where(True, x, y)
    assert all(torch.equal(w2[k], x[k]) for k in x)


def test_sequential_maps():
    xs = torch.arange(12.0).reshape(4, 3)
    f = lambda row, s: {"sum": row.sum() * s, "row": row + s}
    for m in (nb.smap, nb.lmap, nb.get_map("lmap"), nb.get_map("smap"), nb.get_map("vmap")):
        out = m(f, in_axes=(0, None))(xs, 2.0)
        assert torch.equal(out["sum"], xs.sum(1) * 2.0) and torch.equal(out["row"], xs + 2.0)
    with pytest.raises(ValueError):
        nb.get_map("nope")
    with pytest.raises(TypeError):
        nb.get_map(3)


def test_mean_std_stack():
    rng = np.random.default_rng(1)
    forest = tuple(_tree(rng) for _ in range(5))
    m, sd = nb.mean_and_std(forest)
    for k in forest[0]:
        arr = np.stack([f[k].numpy() for f in forest])
        np.testing.assert_allclose(m[k].numpy(), arr.mean(0), rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(sd[k].numpy(), arr.std(0, ddof=1), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(nb.mean(forest)[k].numpy(), arr.mean(0), rtol=1e-13, atol=1e-15)
    _, sd0 = nb.mean_and_std(forest, correct_bias=False)
    np.testing.assert_allclose(sd0["c"].numpy(), np.stack([f["c"].numpy() for f in forest]).std(0), rtol=1e-10)
    st = nb.stack(forest)
    assert st["b"].shape == (5, 3, 2)
    back = nb.unstack(st)
    assert len(back) == 5 and all(torch.equal(back[i][k], forest[i][k]) for i in range(5) for k in forest[0])


def test_gaussian_callable_covariance_must_be_diagonal():
    """`Gaussian(noise_cov_inv=callable)`: a diagonal operator is recognised (its diagonal is read off, fused path), anything else
    is kept as an operator (likelihood_impl.py:35-80 accepts arbitrary callables) instead of being silently treated as diagonal."""
    import numpy as np
    import pytest
    import torch
    from nifty_b200.likelihood import Gaussian
    d = np.zeros((4, 6))
    w = torch.linspace(1.0, 2.0, 24, dtype=torch.float64).reshape(4, 6)
    lh = Gaussian(d, noise_cov_inv=lambda x: w * x)
    assert lh.w_array is not None and torch.equal(torch.as_tensor(lh.w_array), w)
    lh = Gaussian(d, noise_std_inv=lambda x: 3.0 * x)
    assert torch.allclose(torch.as_tensor(lh.w_array), torch.full((4, 6), 9.0, dtype=torch.float64))
    lh = Gaussian(d, noise_cov_inv=lambda x: x + torch.roll(x, 1, 1))      # couples neighbours: kept as an operator, never as a diagonal
    assert lh.cov_inv_fn is not None and lh.w_array is None
    with pytest.raises(NotImplementedError):
        Gaussian(d, noise_std_inv=lambda x: x + torch.roll(x, 1, 1))      # a non-diagonal square root alone (the reference would assume a diagonal)


def test_vector_and_model_containers():
    """`Vector`, `Model`, `WrappedCall`, `Initializer` (tree_math/vector.py:79-188, model.py:32-340): arithmetic, `@`, evaluation,
    per-leaf key splitting of `init`, shape of `target`."""
    import numpy as np
    import pytest
    import torch
    import nifty_b200 as nb
    a = nb.Vector({"x": torch.arange(3.0), "y": torch.ones(2, 2)})
    b = nb.Vector({"x": torch.ones(3), "y": 2.0 * torch.ones(2, 2)})
    c = 2.0 * a - b / 2.0 + 1.0
    assert torch.equal(c["x"], 2.0 * torch.arange(3.0) - 0.5 + 1.0) and torch.equal(c["y"], torch.full((2, 2), 2.0))
    assert a @ b == pytest.approx(3.0 + 8.0) and len(a) == 2 and a.size == 7
    assert nb.vdot(a, b) == pytest.approx(11.0) and nb.norm(-a) == pytest.approx(nb.norm(a))
    m = nb.Model(lambda p: p["x"].sum() * p["y"], domain={"x": (3,), "y": (2, 2)}, white_init=True)
    assert m.target == (2, 2)
    p0, p1, p0b = m.init(7), m.init(8), m.init(7)
    assert set(p0) == {"x", "y"} and p0["x"].shape == (3,) and p0["y"].shape == (2, 2)
    assert torch.equal(p0["x"], p0b["x"]) and not torch.equal(p0["x"], p1["x"]) and not torch.equal(p0["x"][:2], p0["y"].reshape(-1)[:2])
    assert torch.allclose(m(nb.Vector(p0)), p0["x"].sum() * p0["y"])
    w = nb.WrappedCall(torch.exp, name="z", shape=(4,), white_init=True)
    z = w.init(1)
    assert torch.allclose(w(z), torch.exp(z["z"])) and w.domain == {"z": (4,)}
    init = nb.Initializer({"a": lambda k: torch.zeros(2), "b": lambda k: torch.ones(1)})
    assert torch.equal(init(0)["b"], torch.ones(1))
    with pytest.raises(NotImplementedError):
        nb.Gaussian(np.zeros((2, 2))).amend(m)          # an arbitrary call cannot sit under a likelihood on this path


def test_time_threshold_and_cg_name():
    """`time_threshold` stops the host CG / Newton-CG with `info = i` / `status = i` (conjugate_gradient.py:174-176,
    optimize.py:394-396); the inner CG of Newton-CG is named `name + "CG"` (optimize.py:302)."""
    from datetime import datetime, timedelta
    import torch
    from nifty_b200.conjugate_gradient import _cg
    from nifty_b200.optimize import _newton_cg
    a = torch.diag(torch.linspace(1, 50, 40, dtype=torch.float64))
    j = torch.ones(40, dtype=torch.float64)
    past, future = datetime.now() - timedelta(seconds=1), datetime.now() + timedelta(hours=1)
    r = _cg(lambda v: a @ v, j, absdelta=1e-30, maxiter=100, time_threshold=past)
    assert (r.nit, r.info) == (1, 1)
    r2 = _cg(lambda v: a @ v, j, absdelta=1e-12, maxiter=100, time_threshold=future)
    assert r2.info == 0 and torch.allclose(a @ r2.x, j, atol=1e-5)
    seen = {}

    def spy_cg(mat, jj, **kw):
        seen.update(kw)
        return _cg(mat, jj, **{k: v for k, v in kw.items()})

    fg = lambda x: (float(0.5 * x @ a @ x - j @ x), a @ x - j)      # noqa: E731
    res = _newton_cg(None, x0=torch.zeros(40, dtype=torch.float64), fun_and_grad=fg, hessp=lambda p, v: a @ v, name="N", cg=spy_cg,
                     time_threshold=past, cg_kwargs=dict(maxiter=5))
    assert (res.nit, res.status) == (1, 1) and seen["name"] == "NCG" and seen["time_threshold"] == past


def test_random_like_tree_form():
    """`random_like(key, tree)` (forest_math.py:60-72): one sub-key per leaf in sorted-key order, leaf shapes / dtypes kept,
    reproducible, independent leaves; `Vector` in -> `Vector` out."""
    import torch
    import nifty_b200 as nb
    tree = {"b": (2, 3), "a": torch.zeros(4, dtype=torch.float32)}
    t, t2, t3 = nb.random_like(3, tree), nb.random_like(3, tree), nb.random_like(4, tree)
    assert t["a"].shape == (4,) and t["a"].dtype == torch.float32 and t["b"].shape == (2, 3) and t["b"].dtype == torch.float64
    assert torch.equal(t["a"], t2["a"]) and torch.equal(t["b"], t2["b"]) and not torch.equal(t["b"], t3["b"])
    # the leaf keys follow the sorted order: leaf "a" gets the first sub-key whatever the insertion order
    ka = nb.random_split(3, 2)[0]
    from nifty_b200.evi import random_normal
    assert torch.equal(t["a"], random_normal(ka, (4,), torch.float32, "cpu"))
    assert isinstance(nb.random_like(1, nb.Vector({"x": (2,)})), nb.Vector)


def test_samples_container():
    """`Samples` (evi.py:300-396): residuals around a position, iteration, equality, re-centring."""
    import pytest
    import torch
    import nifty_b200 as nb
    pos = torch.arange(4.0)
    res = torch.stack((torch.ones(4), -torch.ones(4), 2.0 * torch.ones(4)))
    s = nb.Samples(pos=pos, samples=res, keys=[1, 2])
    assert len(s) == 3 and torch.equal(s[1], pos - 1.0) and torch.equal(s.samples, pos[None] + res)
    assert [float(x[0]) for x in s] == [1.0, -1.0, 2.0]
    assert s == nb.Samples(pos=pos.clone(), samples=res.clone()) and s != nb.Samples(pos=pos + 1.0, samples=res) and s != 3
    moved = s.at(pos + 10.0)
    assert torch.equal(moved.residuals, res) and torch.equal(moved[0], pos + 11.0) and moved.keys == [1, 2]
    raw = nb.Samples(pos=None, samples=pos[None] + res)                      # full samples without an offset
    recentred = raw.at(pos, old_pos=pos)
    assert torch.equal(recentred.residuals, res) and recentred == s
    with pytest.raises(ValueError):
        raw.at(pos)
    with pytest.raises(ValueError):
        nb.Samples(pos=pos, samples=None)[0]
    assert len(nb.Samples(pos=pos, samples=None)) == 0


def _rosenbrock(x):
    import torch
    return torch.sum(100.0 * torch.diff(x) ** 2 + (1.0 - x[:-1]) ** 2)


def _matyas(p):
    return 0.26 * (p[0] ** 2 + p[1] ** 2) - 0.48 * p[0] * p[1]


def _eggholder(p):
    import torch
    x, y = p[0], p[1]
    return -(y + 47) * torch.sin(torch.sqrt(torch.abs(x / 2.0 + y + 47.0))) - x * torch.sin(torch.sqrt(torch.abs(x - (y + 47.0))))


def test_minimize_ncg_on_plain_functions():
    """test/test_re/test_ncg.py:225-257 restated on the product's Newton-CG: a scalar function alone (gradient and Hessian-vector
    products from autograd), `maxiter` None / inf, against SciPy's trust-ncg at rtol 2e-6; plus the `fun` + `jac` form of :94-109
    and the `minimize(method=...)` dispatcher (:863-892)."""
    import numpy as np
    import pytest
    import torch
    from scipy.optimize import minimize as opt_minimize
    import nifty_b200 as nb
    for func, x0 in ((_rosenbrock, torch.zeros(2, dtype=torch.float64)), (_matyas, 6.0 * torch.ones(2, dtype=torch.float64)),
                     (_eggholder, 100.0 * torch.ones(2, dtype=torch.float64))):
        def f_np(x):
            return float(func(torch.as_tensor(x, dtype=torch.float64)))

        def g_np(x):
            xr = torch.as_tensor(x, dtype=torch.float64).requires_grad_(True)
            return torch.autograd.grad(func(xr), xr)[0].numpy()

        def h_np(x):
            return torch.autograd.functional.hessian(func, torch.as_tensor(x, dtype=torch.float64)).numpy()

        ref = opt_minimize(f_np, x0.numpy(), jac=g_np, hess=h_np, method="trust-ncg").x
        for maxiter in (np.inf, None):
            res = nb.newton_cg(func, x0, maxiter=maxiter, xtol=1e-6, energy_reduction_factor=None, name="N")
            np.testing.assert_allclose(res.numpy(), ref, rtol=2e-6, atol=2e-5)
    x = torch.as_tensor(np.random.default_rng(12).standard_normal(3))
    diag = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    fun = lambda y: torch.sum(y ** 2 / diag) / 2 - torch.dot(x, y)        # noqa: E731
    grad = lambda y: y / diag - x                                        # noqa: E731
    met = lambda y, t: t / diag                                          # noqa: E731
    for kwargs in ({"fun": fun, "jac": grad}, {"fun_and_grad": lambda y: (float(fun(y)), grad(y))}, {"fun": fun}):
        res = nb.newton_cg(x0=x, hessp=met if "jac" in kwargs or "fun_and_grad" in kwargs else None, maxiter=20, absdelta=1e-6, name="N", **kwargs)
        np.testing.assert_allclose(res.numpy(), (diag * x).numpy(), rtol=1e-4, atol=1e-4)
    opt = nb.minimize(fun, x, method="newton-cg", options=dict(hessp=met, maxiter=20, absdelta=1e-6))
    np.testing.assert_allclose(opt.x.numpy(), (diag * x).numpy(), rtol=1e-4, atol=1e-4)
    with pytest.raises(NotImplementedError):
        nb.minimize(fun, x, method="trust-ncg")
    with pytest.raises(ValueError):
        nb.minimize(fun, x, method="bfgs")
    with pytest.raises(ValueError):
        nb.minimize(fun, x, method="ncg", tol=1e-3)
    with pytest.raises(TypeError):
        nb.minimize(fun, x, args=[1], method="ncg")


def test_solvers_on_latent_trees():
    """test/test_re/test_ncg.py:41-77 restated: Newton-CG on a nested position tree (`Vector` of list / tuple / dict, float32 leaves)
    with a tree-valued metric; and `cg` on a `Vector` right-hand side (:112-124)."""
    import numpy as np
    import torch
    import nifty_b200 as nb
    from nifty_b200.tree_math import ravel
    pos = nb.Vector([torch.tensor(0.0, dtype=torch.float32), (torch.tensor(3.0, dtype=torch.float32),), {"a": torch.tensor(5.0, dtype=torch.float32)}])
    getters = (lambda x: x[0], lambda x: x[1][0], lambda x: x[2]["a"])
    tgt, met = [-10.0, 1.0, 2.0], [10.0, 40.0, 2.0]

    def model_and_grad(p):
        val = sum((get(p) - tgt[i]) ** 2 * met[i] for i, get in enumerate(getters))
        grad = nb.Vector([2 * (p[0] - tgt[0]) * met[0], (2 * (p[1][0] - tgt[1]) * met[1],), {"a": 2 * (p[2]["a"] - tgt[2]) * met[2]}])
        return float(val), grad

    def metric(p, tan):
        return nb.Vector([tan[0] * met[0], (tan[1][0] * met[1],), {"a": tan[2]["a"] * met[2]}])

    res = nb.newton_cg(fun_and_grad=model_and_grad, x0=pos, hessp=metric, maxiter=10, absdelta=1e-6)
    assert isinstance(res, nb.Vector) and isinstance(res.tree[1], tuple) and res[0].dtype == torch.float32
    for i, get in enumerate(getters):
        np.testing.assert_allclose(float(get(res)), tgt[i], atol=1e-6, rtol=1e-5)
    flat, unravel = ravel(pos)
    assert flat.shape == (3,) and torch.equal(ravel(unravel(flat))[0], flat)
    # cg on a Vector (diagonal operator, known answer)
    diag = {"u": torch.tensor([1.0, 2.0], dtype=torch.float64), "v": torch.tensor([[3.0]], dtype=torch.float64)}
    x = nb.Vector({"u": torch.tensor([0.3, -1.2], dtype=torch.float64), "v": torch.tensor([[0.7]], dtype=torch.float64)})
    mat = lambda t: nb.Vector({k: t[k] / diag[k] for k in diag})        # noqa: E731
    sol, info = nb.cg(mat, x, resnorm=1e-8, absdelta=1e-10)
    assert info == 0 and isinstance(sol, nb.Vector)
    for k in diag:
        np.testing.assert_allclose(sol[k].numpy(), (diag[k] * x[k]).numpy(), rtol=1e-6)


def test_tree_math_namespace():
    """The remaining leaf-wise helpers of the reference's `tree_math` (vector_math.py, forest_math.py:21-41, 159-210)."""
    import pytest
    import torch
    from nifty_b200 import tree_math as tm
    t = {"a": torch.tensor([1.0, -2.0]), "b": (torch.tensor([[3.0]]), torch.tensor(4.0))}
    assert tm.sum(t) == 6.0 and tm.max(t) == 4.0 and tm.min(t) == -2.0 and tm.size(t) == 4
    assert tm.any(tm.where(True, t, t)) and not tm.all({"x": torch.tensor([True, False])})
    assert tm.shape(t) == {"a": (2,), "b": ((1, 1), ())} and tm.result_type(t) == torch.float32
    assert tm.dot(t, tm.ones_like(t)) == 6.0 and tm.vdot(t, t) == 30.0 and tm.norm(t) == pytest.approx(30.0 ** 0.5)
    assert tm.has_arithmetics(tm.Vector(t)) and tm.has_arithmetics(torch.ones(2)) and not tm.has_arithmetics(t)
    with pytest.raises(TypeError):
        tm.assert_arithmetics(t)
    forest = tuple({"x": torch.full((2,), float(i))} for i in range(4))
    sq = tm.map_forest(lambda tr: {"y": tr["x"] ** 2})(forest)
    assert len(sq) == 4 and torch.equal(sq[3]["y"], torch.full((2,), 9.0))
    m = tm.map_forest_mean(lambda tr: {"y": tr["x"] ** 2})(forest)
    assert torch.equal(m["y"], torch.full((2,), 3.5))
    with pytest.raises(TypeError):
        tm.map_forest(lambda tr: tr)({"x": torch.zeros(2)})
    assert tm.mean(forest)["x"][0] == 1.5                      # module-level names still work after the tree versions of sum / max


def test_function_style_priors():
    """num/stats_distributions.py:20-134 on torch tensors: round trips, moments of the log-normal, and the push-forward of a
    standard normal against SciPy's distributions through their quantile functions."""
    import numpy as np
    import torch
    from scipy import stats
    import nifty_b200 as nb
    xi = torch.linspace(-3.0, 3.0, 61, dtype=torch.float64)
    q = stats.norm.cdf(xi.numpy())
    np.testing.assert_allclose(nb.normal_invprior(2.0, 0.5)(nb.normal_prior(2.0, 0.5)(xi)).numpy(), xi.numpy(), atol=1e-14)
    np.testing.assert_allclose(nb.lognormal_invprior(2.0, 0.5)(nb.lognormal_prior(2.0, 0.5)(xi)).numpy(), xi.numpy(), atol=1e-13)
    lm, ls = nb.lognormal_moments(2.0, 0.5)
    np.testing.assert_allclose(nb.lognormal_prior(2.0, 0.5)(xi).numpy(), stats.lognorm.ppf(q, s=ls, scale=np.exp(lm)), rtol=1e-10)
    assert abs(stats.lognorm.mean(s=ls, scale=np.exp(lm)) - 2.0) < 1e-12 and abs(stats.lognorm.std(s=ls, scale=np.exp(lm)) - 0.5) < 1e-12
    np.testing.assert_allclose(nb.uniform_prior(-1.0, 3.0)(xi).numpy(), stats.uniform.ppf(q, loc=-1.0, scale=4.0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(nb.laplace_prior(0.7)(xi).numpy(), stats.laplace.ppf(q, scale=0.7), rtol=1e-9, atol=1e-12)
