"""Lanczos tridiagonalisation and stochastic Lanczos quadrature on top of the device product
(mirror of the public functions of ``nifty/re/num/lanczos.py:15-123``: ``lanczos_tridiag``,
``stochastic_logdet_from_lanczos``, ``stochastic_lq_logdet``).

``mat`` is any callable on flat torch vectors -- typically ``lambda v: lin.metric(v, add_identity=True)``, one fused
metric-vector product per Lanczos step; everything else (the three-term recurrence, full re-orthogonalisation against the
stored basis, the small eigenproblems of the quadrature) is call sequencing on the same device.  ``slq_gauss_radau`` is the
core of the reference's ``_slq_gauss_radau`` (:483-754): Hutchinson probes with optional deflation of known eigenvectors, the
Gauss estimate with its stochastic standard error, and the two Gauss-Radau quadratures with a prescribed node at either end
of the spectrum (:211-283) that bracket every probe's quadratic form; the reference's further knobs (extra functions, partial
re-orthogonalisation, probe micro-batches, clipping) are not reproduced.
"""
from __future__ import annotations

from typing import Callable, Optional, Union

import numpy as np
import torch

__all__ = ["lanczos_tridiag", "stochastic_logdet_from_lanczos", "stochastic_lq_logdet", "slq_gauss_radau"]


def _lanczos(matvec: Callable, v1: torch.Tensor, order: int, eps: float):
    """Lanczos recurrence with full re-orthogonalisation for one normalised flat start vector (lanczos.py:286-416,
    ``reorth_mode=2``): diagonal ``alpha`` (order), off-diagonal ``off`` (order - 1), basis (order, n); zero padded after a
    breakdown (residual norm <= eps)."""
    n = v1.numel()
    dt, dev = v1.dtype, v1.device
    alpha = torch.zeros(order, dtype=dt, device=dev)
    beta = torch.zeros(order, dtype=dt, device=dev)
    basis = torch.zeros((order, n), dtype=dt, device=dev)
    basis[0] = v1
    v_prev = torch.zeros_like(v1)
    v_curr = v1
    for i in range(order):
        w = matvec(v_curr)
        a = torch.dot(v_curr, w)
        w = w - a * v_curr
        if i > 0:
            w = w - beta[i - 1] * v_prev
        vecs = basis[:i + 1]
        w = w - (vecs @ w) @ vecs                  # full re-orthogonalisation against the stored basis
        b = float(torch.linalg.norm(w))
        alpha[i] = a
        if not b > eps:                             # breakdown: the Krylov space is exhausted, pad with zeros
            break
        beta[i] = b
        v_prev, v_curr = v_curr, w / b
        if i + 1 < order:
            basis[i + 1] = v_curr
    return alpha, beta[:-1], basis


def lanczos_tridiag(mat: Callable, v: torch.Tensor, *, order: int, tol: float = 1e-12):
    """Lanczos tridiagonal (order x order) and its orthonormal basis ``(order,) + v.shape`` (lanczos.py:15-54); padded with
    zeros after an early breakdown."""
    if order < 1:
        raise ValueError("order must be >= 1")
    v = torch.as_tensor(v)
    shape = v.shape

    def flat_matvec(x):
        r = torch.as_tensor(mat(x.reshape(shape)))
        if r.shape != shape:
            raise ValueError(f"shape of `mat(v)` {tuple(r.shape)!r} incompatible with {tuple(shape)!r}")
        return r.reshape(-1)

    v1 = (v / torch.linalg.norm(v)).reshape(-1)
    alpha, off, basis = _lanczos(flat_matvec, v1, order, tol)
    tri = torch.diag(alpha) + torch.diag(off, 1) + torch.diag(off, -1)
    return tri, basis.reshape((order,) + tuple(shape))


def _gauss_unit(alpha, off, func, discard_eigs_below):
    """``e1^T f(T) e1`` from the eigendecomposition of the tridiagonal (lanczos.py:155-208, 433-461); eigenvalues below
    the threshold are discarded (``nansum`` in the reference)."""
    tri = torch.diag(alpha) + torch.diag(off, 1) + torch.diag(off, -1)
    evals, evecs = torch.linalg.eigh(tri)
    w = evecs[0, :] ** 2
    keep = evals >= discard_eigs_below
    return torch.sum(torch.where(keep, w * func(torch.where(keep, evals, torch.ones_like(evals))), torch.zeros_like(w)))


def stochastic_logdet_from_lanczos(tridiag_stack, matrix_shape0: int, func: Callable = torch.log, *, tol=1e-14):
    """Trace estimate ``n * mean_s e1^T f(T_s) e1`` from a stack of Lanczos tridiagonals (lanczos.py:57-83)."""
    ts = torch.as_tensor(tridiag_stack)
    if ts.ndim != 3 or ts.shape[-2] != ts.shape[-1]:
        raise ValueError("tridiag_stack must have shape (num_samples, order, order)")
    est = torch.stack([_gauss_unit(torch.diagonal(t), torch.diagonal(t, 1), func, tol) for t in ts])
    return float(matrix_shape0) * float(est.mean())


def stochastic_lq_logdet(mat: Union[torch.Tensor, Callable], order: int, n_samples: int, key, *, shape0: Optional[int] = None,
                         dtype=torch.float64, device=None) -> float:
    """Log-determinant by stochastic Lanczos quadrature (lanczos.py:86-123): ``n_samples`` Rademacher probes, ``order``
    Lanczos steps each (one device product per step), Gauss quadrature of ``log``.  ``key``: an int seed or a numpy
    Generator (the probes are drawn on the host: PRNG draws stay outside the boundary)."""
    if callable(mat) and shape0 is None:
        raise ValueError("shape0 must be provided if `mat` is callable or has no shape attribute")
    if not callable(mat):
        m = torch.as_tensor(mat)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("mat must be a square matrix")
        if shape0 is not None and shape0 != m.shape[0]:
            raise ValueError("shape0 does not match the matrix dimension")
        shape0, dtype, device = m.shape[0], m.dtype, m.device
        mat = lambda x, m=m: m @ x       # noqa: E731
    rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
    tris = []
    for _ in range(n_samples):
        v = torch.as_tensor(rng.integers(0, 2, size=shape0) * 2.0 - 1.0, dtype=dtype, device=device)
        tri, _ = lanczos_tridiag(mat, v, order=min(order, shape0))
        tris.append(tri)
    return stochastic_logdet_from_lanczos(torch.stack(tris), shape0)


def _tridiag(alpha, off):
    return torch.diag(alpha) + torch.diag(off, 1) + torch.diag(off, -1)


def _radau_unit(alpha, off, mu: float, func: Callable, eps: float):
    """``e1^T f(T_hat) e1`` for the Gauss-Radau modification ``T_hat`` of the Lanczos tridiagonal that makes ``mu`` a quadrature
    node (lanczos.py:211-283): the last diagonal entry becomes ``mu + beta^2 e_last^T (T_lead - mu)^-1 e_last`` with ``T_lead``
    the leading block and ``beta`` the last off-diagonal.  NaN when ``mu`` is not separated from the spectrum of ``T_lead``."""
    m = alpha.numel()
    if m == 1:
        return func(torch.as_tensor(mu, dtype=alpha.dtype, device=alpha.device))
    evals, evecs = torch.linalg.eigh(_tridiag(alpha[:-1], off[:-1]))
    den = evals - mu
    if bool(torch.any(den.abs() <= eps * (1.0 + abs(mu)))):
        return torch.as_tensor(float("nan"), dtype=alpha.dtype, device=alpha.device)
    g = torch.sum(evecs[-1, :] ** 2 / den)
    a_hat = alpha.clone()
    a_hat[-1] = mu + off[-1] ** 2 * g
    ev, vec = torch.linalg.eigh(_tridiag(a_hat, off))
    return torch.sum(vec[0, :] ** 2 * func(ev))


def slq_gauss_radau(mat: Union[torch.Tensor, Callable], func: Callable, order: int, n_samples: int, key, *, shape0: Optional[int] = None,
                    deflate_eigvecs=None, lam_min: Optional[float] = None, lam_max: Optional[float] = None, compute_radau: bool = False,
                    extra_fns: Optional[dict] = None, eps: float = 1e-12, dtype=torch.float64, device=None) -> dict:
    """``tr f(A)`` of a symmetric positive definite ``A`` by stochastic Lanczos quadrature with optional Gauss-Radau bounds
    (lanczos.py:483-754).  ``deflate_eigvecs`` (n, p), orthonormal: the probes are projected onto their orthogonal complement,
    ``z <- z - Q Q^T z``, so the estimate is the trace over that complement (an invariant subspace when the columns are
    eigenvectors).  ``compute_radau`` needs ``lam_min`` / ``lam_max`` enclosing the spectrum seen by the probes; for every probe the
    two Radau values bracket ``z^T f(A) z`` when the derivatives of ``f`` have constant alternating sign (``log``).

    Returns ``estimate`` / ``gauss_estimate``, ``stochastic_se`` / ``gauss_se`` (NaN for one probe), and with ``compute_radau``
    ``radau_lo``, ``radau_hi`` (means over the probes of the node-at-``lam_min`` / node-at-``lam_max`` quadratures) and
    ``quadrature_width``.  ``extra_fns`` (name -> scalar function): further traces from the SAME tridiagonals, Gauss quadrature
    only, reported as ``extra_{name}_estimate`` / ``extra_{name}_se``."""
    if (lam_min is None) != (lam_max is None):
        raise ValueError("Provide both lam_min and lam_max, or neither.")
    if compute_radau and lam_min is None:
        raise ValueError("compute_radau=True requires lam_min and lam_max.")
    if order < 1:
        raise ValueError("order must be >= 1.")
    if n_samples < 1:
        raise ValueError("num_samples must be >= 1.")
    if extra_fns is not None and not isinstance(extra_fns, dict):
        raise ValueError("extra_fns must be a dict of name -> callable.")
    extras = {name: [] for name in (extra_fns or {})}
    if not callable(mat):
        m = torch.as_tensor(mat)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("mat must be a square matrix")
        shape0, dtype, device = m.shape[0], m.dtype, m.device
        mat = lambda x, m=m: m @ x       # noqa: E731
    elif shape0 is None:
        if deflate_eigvecs is None:
            raise ValueError("If A is callable, provide n=... or deflate_eigvecs.")
        shape0 = int(np.shape(deflate_eigvecs)[0])
    Q = None if deflate_eigvecs is None else torch.as_tensor(np.asarray(deflate_eigvecs), dtype=dtype, device=device)
    if Q is not None and Q.shape[1] == 0:
        Q = None
    rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
    gauss, lo, hi = [], [], []
    for _ in range(n_samples):
        z = torch.as_tensor(rng.integers(0, 2, size=shape0) * 2.0 - 1.0, dtype=dtype, device=device)
        if Q is not None:
            z = z - Q @ (Q.T @ z)
        nz2 = float(torch.dot(z, z))
        if not nz2 > 0.0:
            gauss.append(0.0)
            lo.append(0.0)
            hi.append(0.0)
            for v in extras.values():
                v.append(0.0)
            continue
        alpha, off, _ = _lanczos(mat, z / np.sqrt(nz2), min(order, shape0), eps)
        m_eff = 1
        while m_eff < alpha.numel() and float(off[m_eff - 1]) > eps:      # steps before a breakdown (zero padding afterwards)
            m_eff += 1
        alpha, off = alpha[:m_eff], off[:m_eff - 1]
        ev, vec = torch.linalg.eigh(_tridiag(alpha, off))
        gauss.append(nz2 * float(torch.sum(vec[0, :] ** 2 * func(ev))))
        for name, fn in (extra_fns or {}).items():
            extras[name].append(nz2 * float(torch.sum(vec[0, :] ** 2 * fn(ev))))
        if compute_radau:
            exhausted = m_eff < min(order, shape0)                         # Krylov space exhausted: the Gauss value is exact
            lo.append(gauss[-1] if exhausted else nz2 * float(_radau_unit(alpha, off, float(lam_min), func, eps)))
            hi.append(gauss[-1] if exhausted else nz2 * float(_radau_unit(alpha, off, float(lam_max), func, eps)))
    est = float(np.mean(gauss))
    se = float(np.std(gauss, ddof=1) / np.sqrt(len(gauss))) if len(gauss) > 1 else float("nan")
    out = {"estimate": est, "gauss_estimate": est, "stochastic_se": se, "gauss_se": se, "per_probe": np.asarray(gauss)}
    for name, vals in extras.items():
        out[f"extra_{name}_estimate"] = float(np.mean(vals))
        out[f"extra_{name}_se"] = float(np.std(vals, ddof=1) / np.sqrt(len(vals))) if len(vals) > 1 else float("nan")
    if compute_radau:
        out["radau_lo"], out["radau_hi"] = float(np.mean(lo)), float(np.mean(hi))
        out["quadrature_width"] = abs(out["radau_hi"] - out["radau_lo"])
        out["per_probe_radau"] = np.stack((np.asarray(lo), np.asarray(hi)))
    return out
