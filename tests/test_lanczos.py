"""Lanczos / stochastic Lanczos quadrature (nifty_b200/lanczos.py) -- the reference's own checks
(test/test_re/test_lanczos.py:46-70: reconstruction of a positive-definite matrix from the tridiagonal and the basis, zero
padding after a breakdown) and the log-determinant against dense algebra; on the device product in tests/vi_checks.py."""
import numpy as np
import pytest
import torch

from nifty_b200 import lanczos


@pytest.mark.parametrize("seed", [12, 17, 22])
@pytest.mark.parametrize("shape0", [64, 32])
def test_lanczos_tridiag_reconstructs_the_matrix(seed, shape0):
    rng = np.random.default_rng(seed)
    m = rng.normal(size=(shape0,) * 2)
    m = torch.as_tensor(m @ m.T)
    v = torch.as_tensor(rng.integers(0, 2, size=shape0) * 2.0 - 1.0)
    tri, vecs = lanczos.lanczos_tridiag(lambda x: m @ x, v, order=shape0)
    est = vecs.T @ tri @ vecs
    np.testing.assert_allclose(est.numpy(), m.numpy(), atol=1e-9 * float(m.abs().max()), rtol=1e-9)


def test_lanczos_tridiag_pads_after_breakdown_and_preserves_shape():
    v = torch.arange(1.0, 5.0, dtype=torch.float64).reshape(2, 2)
    tri, vecs = lanczos.lanczos_tridiag(lambda x: 2.0 * x, v, order=4)
    assert tri.shape == (4, 4) and vecs.shape == (4, 2, 2)
    np.testing.assert_allclose(tri.numpy(), np.diag([2.0, 0.0, 0.0, 0.0]), atol=1e-14)
    np.testing.assert_allclose(vecs[0].numpy(), (v / torch.linalg.norm(v)).numpy())
    np.testing.assert_allclose(vecs[1:].numpy(), 0.0)


def test_stochastic_lq_logdet_against_slogdet():
    rng = np.random.default_rng(3)
    n = 96
    x = np.tril(rng.normal(size=(n, n)))
    np.fill_diagonal(x, np.diag(x).clip(0.5, None))
    m = x @ x.T + 5.0 * np.eye(n)
    want = np.linalg.slogdet(m)[1]
    got = lanczos.stochastic_lq_logdet(torch.as_tensor(m), order=40, n_samples=200, key=0)
    assert abs(got - want) < 0.03 * abs(want)
    # full order: the quadrature of one probe is exact, |v|^2 e1^T log(T) e1 = v^T log(M) v
    v = torch.as_tensor(rng.normal(size=n))
    tri, _ = lanczos.lanczos_tridiag(lambda u: torch.as_tensor(m) @ u, v, order=n)
    quad = lanczos.stochastic_logdet_from_lanczos(tri[None], 1) * float(v @ v)
    w, q = np.linalg.eigh(m)
    exact = float(v.numpy() @ (q * np.log(w)) @ q.T @ v.numpy())
    assert abs(quad - exact) < 1e-8 * abs(exact)
    with pytest.raises(ValueError):
        lanczos.stochastic_lq_logdet(lambda v: v, 4, 2, 0)


@pytest.mark.parametrize("k", [0, 6])
def test_slq_gauss_radau_brackets_every_probe(k):
    """Gauss-Radau quadratures with a node at either end of the spectrum bracket the quadratic form z^T log(A) z of every probe
    (lanczos.py:211-283, 483-754), also for probes deflated by known eigenvectors; the Gauss estimate is the Hutchinson mean."""
    import nifty_b200 as nb
    rng = np.random.default_rng(0)
    n = 80
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.sort(np.concatenate([np.linspace(1.0, 3.0, n - 6), [10, 20, 40, 80, 160, 320]]))[::-1]
    A = torch.as_tensor((U * ev) @ U.T)
    logA = (U * np.log(ev)) @ U.T
    Q = U[:, :k]
    lam_max = float(ev[k] * (1.05 if k == 0 else 1.0))
    out = nb.slq_gauss_radau(A, torch.log, 8, 12, 5, deflate_eigvecs=Q, lam_min=1.0, lam_max=lam_max, compute_radau=True)
    r2 = np.random.default_rng(5)                     # the same probes
    exact = []
    for _ in range(12):
        z = r2.integers(0, 2, size=n) * 2.0 - 1.0
        z = z - Q @ (Q.T @ z)
        exact.append(z @ logA @ z)
    exact = np.array(exact)
    lo, hi = out["per_probe_radau"]
    assert np.all(np.minimum(lo, hi) <= exact + 1e-9) and np.all(exact <= np.maximum(lo, hi) + 1e-9)
    assert min(out["radau_lo"], out["radau_hi"]) <= exact.mean() <= max(out["radau_lo"], out["radau_hi"])
    assert abs(out["estimate"] - np.sum(np.log(ev[k:]))) < 4.0 * out["stochastic_se"] + 1e-6
    assert out["quadrature_width"] >= 0.0
    with pytest.raises(ValueError):
        nb.slq_gauss_radau(A, torch.log, 8, 2, 5, lam_min=1.0)
    with pytest.raises(ValueError):
        nb.slq_gauss_radau(A, torch.log, 8, 2, 5, compute_radau=True)
    with pytest.raises(ValueError):
        nb.slq_gauss_radau(lambda v: A @ v, torch.log, 8, 2, 5)
