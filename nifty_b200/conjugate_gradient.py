"""Conjugate gradient with the call protocol of ``nifty/re/conjugate_gradient.py`` (``cg``:28-41,
``_cg``:77-214, ``static_cg``:45-50).

Two execution paths, same semantics (``info``: 0 converged, i > 0 stopped at iteration i, < 0 failed):

* ``mat`` is a :class:`HamiltonianMetric` (``lh.metric(pos, .) + .`` at a fixed ``pos``, i.e. what
  ``draw_linear_residual`` and ``newton_cg`` pass): the whole solve runs on the device through
  ``nb200_cg_solve`` -- fused step kernels, curvature from the epilogue of the metric kernels, all
  stopping rules evaluated on the GPU, the host polls a status word every few iterations.  This is
  the B200 counterpart of ``static_cg`` (``lax.while_loop``).
* ``mat`` is a :class:`SampleAveragedMetric` (the KL metric over several linearisations / ranks): same device
  solve through ``nb200_cg_solve_multi`` -- the products accumulate in the last-pass epilogues, the all-reduce over the
  ranks is enqueued in-stream by a hook, ``<d, q>`` is a fused reduction.
* any other callable: the generic host loop below (one host sync per scalar, like ``_cg``); used for slab-decomposed
  fields and user-supplied operators.
"""

from __future__ import annotations

import os
from datetime import datetime
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

N_RESET = 20  # conjugate_gradient.py:17


class CGResults(NamedTuple):
    x: torch.Tensor
    nit: int
    nfev: int
    info: int
    success: bool


class HamiltonianMetric:
    """``t -> lh.metric(pos, t) + t`` bound to a cached linearisation (evi.py:83-85 ``_ham_metric``).

    ``other`` (a second linearisation) selects the geoVI operator of evi.py:167-172.  ``frozen``: ranges
    ``[(lo, hi), ...]`` of the flat latent vector held fixed (``point_estimates`` / ``constants``,
    likelihood.py:399-499): the operator of the frozen likelihood, written on full-length vectors -- outputs are
    cleared on the frozen entries, inputs are expected to be zero there."""

    def __init__(self, lin, other=None, likelihood=None, frozen=None):
        self.lin, self.other, self.likelihood = lin, other, likelihood
        self.frozen = list(frozen) if frozen else None

    def _clear(self, v):
        if self.frozen:
            for lo, hi in self.frozen:
                v[lo:hi] = 0
        return v

    @property
    def distributed(self):
        return bool(self.lin.model.plan.dist)

    @property
    def host_composed(self):
        """Linearisation of a host-composed model (outer.HostLin): the operator is applied by the host, CG runs the host loop."""
        return bool(getattr(self.lin, "host_composed", False))

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        if self.other is None:
            return self._clear(self.lin.metric(t, add_identity=True))
        tm = self._clear(self.other.metric_pair(self.lin, t, add_identity=True))
        return self._clear(self.lin.metric_pair(self.other, tm, add_identity=True))


class SampleAveragedMetric:
    """``t -> mean_i(lh.metric(x_i, t) + t)`` over the sample points of ALL ranks (``_kl_met``, optimize_kl.py:117-144):
    the local linearisations ``lins`` are applied and accumulated on the device (``nb200_metric_multi``), ``reduce_fn``
    is the in-stream all-reduce (None on a single rank).  ``_cg`` runs the whole solve on the device
    (``nb200_cg_solve_multi``).  ``n_total``: number of sample points over all ranks; ``identity_here``: this rank adds
    the ``+ t`` (exactly one rank does)."""

    def __init__(self, lins, n_total, identity_here=True, reduce_fn=None, frozen=None, scale_zero=False):
        self.lins, self.n_total, self.identity_here, self.reduce_fn = list(lins), int(n_total), bool(identity_here), reduce_fn
        self.frozen = list(frozen) if frozen else None
        self.scale = 0.0 if scale_zero else 1.0 / self.n_total

    @property
    def host_composed(self):
        return bool(self.lins) and bool(getattr(self.lins[0], "host_composed", False))

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        if self.host_composed:
            out = t.clone() if self.identity_here else torch.zeros_like(t)
            for lin in self.lins:
                out += self.scale * lin.metric(t)
            if self.reduce_fn is not None:
                work = self.reduce_fn(out)
                if work is not None:
                    work.wait()
            if self.frozen:
                for lo, hi in self.frozen:
                    out[lo:hi] = 0
            return out
        from ._runtime import metric_multi
        out = metric_multi(self.lins, t, scale=self.scale, identity_here=self.identity_here, reduce_fn=self.reduce_fn)
        if self.frozen:
            for lo, hi in self.frozen:
                out[lo:hi] = 0
        return out


def _norm(v: torch.Tensor, ord) -> float:
    if ord == 1:
        return float(v.abs().sum())
    if ord == 2:
        return float(torch.linalg.vector_norm(v))
    if ord in (np.inf, float("inf")):
        return float(v.abs().max())
    return float(torch.linalg.vector_norm(v, ord=ord))


def _cg(mat: Callable, j: torch.Tensor, x0: Optional[torch.Tensor] = None, *, absdelta=None, resnorm=None,
        norm_ord=None, tol=1e-5, atol=0.0, miniter=None, maxiter=None, name=None, time_threshold=None,
        _raise_nonposdef=True, check_every=4, vdot=None, vnorm=None) -> CGResults:
    """``vdot`` / ``vnorm``: reductions to use in the host loop (slab-decomposed vectors: all-reduced ones)."""
    if not isinstance(j, torch.Tensor):
        # latent trees / Vectors (the reference's cg acts on pytrees): solve on the raveled vector with `mat` wrapped accordingly
        from .tree_math import ravel
        flat_j, unravel = ravel(j)
        res = _cg(lambda v: ravel(mat(unravel(v)))[0], flat_j, None if x0 is None else ravel(x0)[0], absdelta=absdelta, resnorm=resnorm,
                  norm_ord=norm_ord, tol=tol, atol=atol, miniter=miniter, maxiter=maxiter, name=name, time_threshold=time_threshold,
                  _raise_nonposdef=_raise_nonposdef, check_every=check_every, vdot=vdot, vnorm=vnorm)
        return res._replace(x=unravel(res.x))
    norm_ord = 2 if norm_ord is None else norm_ord
    vdot = (lambda a, b: float(torch.dot(a, b))) if vdot is None else vdot
    vnorm = _norm if vnorm is None else vnorm
    if isinstance(mat, (SampleAveragedMetric, HamiltonianMetric)) and mat.host_composed:
        mat = mat.__call__                       # host-composed operators: the generic loop below
    if isinstance(mat, HamiltonianMetric) and mat.distributed:
        # slab-decomposed field: the recurrence of the loop below with in-place vector updates and ONE all-reduce + host
        # synchronisation per group of reductions
        if mat.likelihood is None:
            raise ValueError("a slab-decomposed HamiltonianMetric needs its likelihood (for the distributed reductions)")
        if norm_ord in (1, 2) and os.environ.get("NB200_SLAB_CG", "1") != "0":
            return _cg_slab(mat, j, x0, absdelta=absdelta, resnorm=resnorm, norm_ord=norm_ord, tol=tol, atol=atol, miniter=miniter,
                            maxiter=maxiter, name=name, _raise_nonposdef=_raise_nonposdef, time_threshold=time_threshold)
        vdot, vnorm = mat.likelihood.vdot, mat.likelihood.vnorm
    elif isinstance(mat, (SampleAveragedMetric, HamiltonianMetric)) and time_threshold is not None:
        # the device solves evaluate their stopping rules on the GPU; a wall-clock rule belongs to the host loop
        raise NotImplementedError("`time_threshold` is not available for the device-resident CG (pass a plain callable `mat` for "
                                  "the host loop, which honours it)")
    elif isinstance(mat, SampleAveragedMetric):
        from ._runtime import cg_solve_multi
        x, res = cg_solve_multi(mat.lins, j, x0, scale=mat.scale, identity_here=mat.identity_here, reduce_fn=mat.reduce_fn,
                                absdelta=absdelta, resnorm=resnorm, norm_ord=norm_ord, tol=tol, atol=atol, miniter=miniter, maxiter=maxiter,
                                raise_nonposdef=_raise_nonposdef, check_every=check_every, frozen=mat.frozen)
        nm = "CG" if name is None else name
        if res.error == 1:
            raise ValueError(f"{nm}: zero curvature")
        if res.error == 2:
            raise ValueError(f"{nm}: negative curvature")
        if res.error == 3:
            raise ValueError(f"{nm}: WARNING: energy increased")
        return CGResults(x, int(res.nit), int(res.nfev), int(res.info), res.info == 0)
    elif isinstance(mat, HamiltonianMetric):
        x, res = mat.lin.cg_solve(j, x0, other=mat.other, absdelta=absdelta, resnorm=resnorm, norm_ord=norm_ord, tol=tol,
                                  atol=atol, miniter=miniter, maxiter=maxiter, raise_nonposdef=_raise_nonposdef,
                                  check_every=check_every, frozen=mat.frozen)
        nm = "CG" if name is None else name
        if res.error == 1:
            raise ValueError(f"{nm}: zero curvature")
        if res.error == 2:
            raise ValueError(f"{nm}: negative curvature")
        if res.error == 3:
            raise ValueError(f"{nm}: WARNING: energy increased")
        return CGResults(x, int(res.nit), int(res.nfev), int(res.info), res.info == 0)

    # generic host loop, statement by statement the recurrence of conjugate_gradient.py:107-214
    maxiter_fallback = 20 * j.numel()
    if miniter is None:
        miniter = min(6, maxiter if maxiter is not None else maxiter_fallback)
    if maxiter is None:
        maxiter = max(min(200, maxiter_fallback), miniter)
    if absdelta is None and resnorm is None:
        resnorm = max(tol * vnorm(j, norm_ord), atol)
    fi = torch.finfo(j.dtype)
    eps, tiny = 6.0 * fi.eps, 6.0 * fi.tiny
    nm = "CG" if name is None else name
    if x0 is None:
        pos = torch.zeros_like(j)
        r = -j
        d = r.clone()
        energy, nfev = 0.0, 0
    else:
        pos = x0.clone()
        r = mat(pos) - j
        d = r.clone()
        energy = vdot((r - j) / 2, pos)
        nfev = 1
    previous_gamma = vdot(r, r)
    if previous_gamma == 0:
        return CGResults(pos, 0, nfev, 0, True)
    info, i = -1, 0
    for i in range(1, maxiter + 1):
        q = mat(d)
        nfev += 1
        curv = vdot(d, q)
        if curv == 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: zero curvature")
            info = 0
            break
        if curv < 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: negative curvature")
            if i == 1:
                pos = previous_gamma / (-curv) * (-j)
            info = 0
            break
        alpha = previous_gamma / curv
        pos = pos - alpha * d
        if i % N_RESET == 0:
            r = mat(pos) - j
            nfev += 1
        else:
            r = r - q * alpha
        gamma = vdot(r, r)
        if time_threshold is not None and datetime.now() > time_threshold:      # conjugate_gradient.py:174-176
            info = i
            break
        if 0.0 <= gamma <= tiny:
            info = 0
            break
        if resnorm is not None and vnorm(r, norm_ord) < resnorm and i >= miniter:
            info = 0
            break
        new_energy = vdot((r - j) / 2, pos)
        energy_diff = energy - new_energy
        if energy_diff < -eps * abs(new_energy):
            if _raise_nonposdef:
                raise ValueError(f"{nm}: WARNING: energy increased")
            info = i
            break
        if absdelta is not None and energy_diff < absdelta and i >= miniter:
            info = 0
            break
        energy = new_energy
        d = d * max(0.0, gamma / previous_gamma) + r
        previous_gamma = gamma
    info = i if info == -1 else info
    return CGResults(pos, i, nfev, info, info == 0)


class _SlabReductions:
    """Dot products / 1-norms of slab-decomposed latent vectors in groups: the excitation rows are reduced locally, ALL
    partial results of a group travel in one all-reduce, the replicated hyper-parameter entries are reduced from the
    replicated data (identical on every rank, so every rank takes the same control-flow decisions) -- one collective and
    one host synchronisation per group instead of one per scalar (likelihood.vdot / vnorm)."""

    def __init__(self, likelihood):
        self.lo, self.hi = likelihood._xi_slice()
        self.comm = likelihood._plan.comm

    def _rep(self, v):
        return v[:self.lo], v[self.hi:]

    def __call__(self, dots=(), norms1=()):
        lo, hi = self.lo, self.hi
        loc = [torch.dot(a[lo:hi], b[lo:hi]) for a, b in dots] + [torch.linalg.vector_norm(v[lo:hi], ord=1) for v in norms1]
        t = torch.stack(loc).to(torch.float64)
        self.comm.allreduce_sum(t)
        rep = []
        for a, b in dots:
            (a0, a1), (b0, b1) = self._rep(a), self._rep(b)
            rep.append(torch.dot(a0, b0) + torch.dot(a1, b1))
        for v in norms1:
            v0, v1 = self._rep(v)
            rep.append(torch.linalg.vector_norm(v0, ord=1) + torch.linalg.vector_norm(v1, ord=1))
        return (t + torch.stack(rep).to(torch.float64)).tolist()          # (the one host synchronisation of the group)


def _cg_slab(mat, j, x0, *, absdelta, resnorm, norm_ord, tol, atol, miniter, maxiter, name, _raise_nonposdef, time_threshold=None) -> CGResults:
    """The recurrence of ``_cg`` (conjugate_gradient.py:107-214, same stopping rules and ``info`` codes) for slab-decomposed
    fields: vector updates in place (``add_`` with a scalar multiplier, no temporaries), reductions in groups
    (:class:`_SlabReductions`).  Per iteration: one product, one curvature all-reduce, one all-reduce of
    {<r, r>, <r - j, pos>, |r|_1}."""
    red = _SlabReductions(mat.likelihood)
    n_glob = mat.likelihood.global_size()
    maxiter_fallback = 20 * n_glob
    if miniter is None:
        miniter = min(6, maxiter if maxiter is not None else maxiter_fallback)
    if maxiter is None:
        maxiter = max(min(200, maxiter_fallback), miniter)
    if absdelta is None and resnorm is None:
        resnorm = max(tol * mat.likelihood.vnorm(j, norm_ord), atol)
    fi = torch.finfo(j.dtype)
    eps, tiny = 6.0 * fi.eps, 6.0 * fi.tiny
    nm = "CG" if name is None else name
    tmp = torch.empty_like(j)
    if x0 is None:
        pos = torch.zeros_like(j)
        r = -j
        d = r.clone()
        energy, nfev = 0.0, 0
        (previous_gamma,) = red(dots=[(r, r)])
    else:
        pos = x0.clone()
        r = mat(pos).sub_(j)
        d = r.clone()
        torch.sub(r, j, out=tmp)
        previous_gamma, e2 = red(dots=[(r, r), (tmp, pos)])
        energy, nfev = 0.5 * e2, 1
    if previous_gamma == 0:
        return CGResults(pos, 0, nfev, 0, True)
    info, i = -1, 0
    for i in range(1, maxiter + 1):
        q = mat(d)
        nfev += 1
        (curv,) = red(dots=[(d, q)])
        if curv == 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: zero curvature")
            info = 0
            break
        if curv < 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: negative curvature")
            if i == 1:
                pos = previous_gamma / (-curv) * (-j)
            info = 0
            break
        alpha = previous_gamma / curv
        pos.add_(d, alpha=-alpha)
        if i % N_RESET == 0:
            r = mat(pos).sub_(j)
            nfev += 1
        else:
            r.add_(q, alpha=-alpha)
        torch.sub(r, j, out=tmp)
        want_norm = resnorm is not None and norm_ord == 1
        vals = red(dots=[(r, r), (tmp, pos)], norms1=[r] if want_norm else ())
        gamma, new_energy = vals[0], 0.5 * vals[1]
        if time_threshold is not None and datetime.now() > time_threshold:
            info = i
            break
        if 0.0 <= gamma <= tiny:
            info = 0
            break
        if resnorm is not None and i >= miniter:
            rn = vals[2] if want_norm else float(np.sqrt(gamma))
            if rn < resnorm:
                info = 0
                break
        energy_diff = energy - new_energy
        if energy_diff < -eps * abs(new_energy):
            if _raise_nonposdef:
                raise ValueError(f"{nm}: WARNING: energy increased")
            info = i
            break
        if absdelta is not None and energy_diff < absdelta and i >= miniter:
            info = 0
            break
        energy = new_energy
        torch.add(r, d, alpha=max(0.0, gamma / previous_gamma), out=d)
        previous_gamma = gamma
    info = i if info == -1 else info
    return CGResults(pos, i, nfev, info, info == 0)


def cg(mat, j, x0=None, *args, **kwargs):
    """``jft.cg``: returns ``(x, info)`` (conjugate_gradient.py:28-41)."""
    res = _cg(mat, j, x0, *args, **kwargs)
    return res.x, res.info


def static_cg(mat, j, x0=None, *args, **kwargs):
    """``jft.static_cg``: same contract; with a :class:`HamiltonianMetric` the loop already runs on the device."""
    res = _cg(mat, j, x0, *args, **kwargs)
    return res.x, res.info
