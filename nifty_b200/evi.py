"""MGVI / geoVI sampling with the call protocol of ``nifty/re/evi.py``.

``draw_linear_residual`` (:88-150), ``nonlinearly_update_residual`` (:181-255), ``draw_residual``
(:258-297) and the ``Samples`` container (:300-396) on flat device vectors.  Every arithmetic step
on the latent / data space runs in libniftyb200.so; Python only sequences the calls.

PRNG keys: the reference splits ``jax.random`` keys (evi.py:121-123); JAX is not available here, so
a key is a ``numpy.random.SeedSequence`` (or an int) and ``random_split`` mirrors ``random.split``.
The white-noise draws happen *outside* the boundary exactly as in the reference, so a caller who
has the reference's draws can pass them in through ``_white=(data_shaped, latent_shaped)``.
"""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import conjugate_gradient
from .conjugate_gradient import HamiltonianMetric
from .likelihood import LikelihoodWithModel
from .optimize import OptimizeResults, _newton_cg


# ---- keys ----------------------------------------------------------------------------------------
def as_key(key):
    return key if isinstance(key, np.random.SeedSequence) else np.random.SeedSequence(int(key))


def random_split(key, n: int = 2):
    """``jax.random.split`` stand-in: n statistically independent child keys, reproducible."""
    key = as_key(key)
    return [np.random.SeedSequence(entropy=key.entropy, spawn_key=tuple(key.spawn_key) + (i,)) for i in range(n)]


def _seed_of(key) -> int:
    return int(as_key(key).generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array([1, 2**32], dtype=np.uint64) % (2**63 - 1))


def random_normal(key, shape, dtype, device) -> torch.Tensor:
    gen = torch.Generator(device=device)
    gen.manual_seed(_seed_of(key))
    return torch.randn(shape, dtype=dtype, device=device, generator=gen)


def random_like(key, lh, dtype=torch.float64, device="cpu"):
    """``jft.random_like(key, primals)`` (tree_math/forest_math.py:60-72).  With a latent TREE (dict / ``Vector`` of arrays or
    shapes) as second argument: the key is split into one sub-key per leaf in sorted-key (pytree) order and every leaf is a
    standard-normal draw of its shape -- the reference's semantics (the draws themselves come from this package's key
    scheme, not from threefry).  With a likelihood: the flat standard-normal latent vector of its model; slab-decomposed
    fields: the replicated hyper-parameter leaves come from the shared key, the local excitation rows from a per-rank child
    key; padding rows are zero."""
    if not isinstance(lh, LikelihoodWithModel):
        from .tree_math import Vector, _leaves
        from .model import _map_domain, _shape_of
        tree = lh.tree if isinstance(lh, Vector) else lh
        leaves_sorted = _leaves(_map_domain(lambda s: (s,), tree))          # one entry per leaf, sorted-key order
        keys = iter(random_split(key, len(leaves_sorted)))

        def draw(s):
            dt = s.dtype if isinstance(s, torch.Tensor) else dtype
            dv = s.device if isinstance(s, torch.Tensor) else device
            return random_normal(next(keys), _shape_of(s), dt, dv)

        def walk(t):
            if isinstance(t, dict):
                return {k: walk(t[k]) for k in sorted(t)}
            if isinstance(t, (tuple, list)) and not all(isinstance(i, (int, np.integer)) for i in t):
                return type(t)(walk(v) for v in t)
            return draw(t)

        out = walk(tree)
        return Vector(out) if isinstance(lh, Vector) else out
    plan = lh.signal.cf.plan
    if not plan.dist:
        return random_normal(key, (lh.layout.size,), lh.dtype, lh.rt.device)
    lo, hi = lh._xi_slice()
    L = lh.layout.size
    keys = random_split(key, plan.comm.world + 1)
    v = torch.empty(L, dtype=lh.dtype, device=lh.rt.device)
    # replicated leaves: a draw whose SIZE is the same on every rank (device generators are size dependent)
    rep = random_normal(keys[0], (L - (hi - lo),), lh.dtype, lh.rt.device)
    v[:lo] = rep[:lo]
    v[hi:] = rep[lo:]
    v[lo:hi] = random_normal(keys[plan.comm.rank + 1], (hi - lo,), lh.dtype, lh.rt.device)
    lh.zero_padding(v)
    return v


# ---- Samples -------------------------------------------------------------------------------------------
class Samples:
    """Residual samples around an expansion point (evi.py:300-396): ``samples = pos + residuals``."""

    def __init__(self, *, pos: Optional[torch.Tensor], samples: Optional[torch.Tensor], keys=None):
        self._pos, self._samples, self._keys = pos, samples, keys

    @property
    def pos(self):
        return self._pos

    @property
    def keys(self):
        return self._keys

    @property
    def samples(self):
        """pos + residuals, leading sample axis (evi.py:334-343)."""
        if self._samples is None:
            raise ValueError(f"{self.__class__.__name__} has no samples")
        return self._samples if self._pos is None else self._pos[None] + self._samples

    @property
    def residuals(self):
        return self._samples

    def __len__(self):
        return 0 if self._samples is None else int(self._samples.shape[0])

    def __getitem__(self, i):
        if self._samples is None:
            raise ValueError(f"{self.__class__.__name__} has no samples")
        r = self._samples[i]
        return r if self._pos is None else self._pos + r

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __eq__(self, other) -> bool:
        if not isinstance(other, self.__class__):
            return False
        if self._samples is None or other._samples is None:
            return self._samples is None and other._samples is None
        a, b = self.samples, other.samples
        return a.shape == b.shape and bool(torch.equal(a, b))

    __hash__ = None

    def at(self, pos, old_pos=None):
        """New offset for all samples (evi.py:365-376): the same residuals around ``pos``; with ``old_pos`` that offset is first
        subtracted from the full samples."""
        if self._pos is not None and old_pos is None:
            smpls = self._samples
        elif old_pos is not None:
            smpls = self.samples - old_pos[None]
        else:
            raise ValueError("invalid combination of `pos` and `old_pos`")
        return Samples(pos=pos, samples=smpls, keys=self._keys)

    def squeeze(self):
        return Samples(pos=self._pos, samples=None if self._samples is None else self._samples.reshape((-1,) + tuple(self._samples.shape[2:])), keys=self._keys)


def concatenate_zip(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Interleave along the sample axis: [a0, b0, a1, b1, ...] (evi.py:53-57)."""
    return torch.stack((a, b), dim=1).reshape((-1,) + tuple(a.shape[1:]))


# ---- MGVI -------------------------------------------------------------------------------------------------
def sample_likelihood(likelihood: LikelihoodWithModel, primals, key, _white_data=None):
    """evi.py:77-80: ``left_sqrt_metric(primals, N(0,1)[data shape])``."""
    lin, _ = likelihood.lin_at(primals)
    plan = likelihood.signal.cf.plan
    if plan.dist:      # independent white noise per rank for its local planes
        key = random_split(key, plan.comm.world + 1)[plan.comm.rank + 1]
    white = _white_data if _white_data is not None else random_normal(key, likelihood.signal.target_shape, likelihood.dtype,
                                                                    likelihood.rt.device)
    return lin.lsm(white, scaled=True)


def draw_linear_residual(likelihood: LikelihoodWithModel, pos, key, *, from_inverse: bool = True, point_estimates=(),
                         cg=conjugate_gradient.cg, cg_name=None, cg_kwargs: Optional[dict] = None,
                         _raise_nonposdef: bool = False, _white=None):
    """One MGVI residual sample at ``pos`` (evi.py:88-150); returns ``(residual, info)``.

    ``point_estimates``: leaves frozen at ``pos`` (:109-111, 149): the draw happens in the space of the other leaves
    and the returned residual is zero on the frozen ones (``_process_point_estimate(..., insert=True)``)."""
    frozen = likelihood.frozen_ranges(point_estimates)
    pos = likelihood.signal.as_flat(pos)
    lin, _ = likelihood.lin_at(pos)
    k_nll, k_prr = random_split(key, 2)
    w_data, w_prior = (None, None) if _white is None else _white
    nll_smpl = sample_likelihood(likelihood, pos, k_nll, _white_data=w_data)
    prr_smpl = random_like(k_prr, likelihood) if w_prior is None else likelihood.signal.as_flat(w_prior).clone()
    if frozen:
        likelihood.clear_frozen(nll_smpl, frozen)
        likelihood.clear_frozen(prr_smpl, frozen)
    smpl = nll_smpl + prr_smpl
    info = 0
    if from_inverse:
        smpl, info = cg(HamiltonianMetric(lin, likelihood=likelihood, frozen=frozen), smpl, x0=prr_smpl, name=cg_name,
                        _raise_nonposdef=_raise_nonposdef, **(cg_kwargs or {}))
        if info is not None and info < 0:
            raise ValueError("conjugate gradient failed")
    return smpl, info


# ---- geoVI --------------------------------------------------------------------------------------------------
def nonlinearly_update_residual(likelihood: LikelihoodWithModel, pos, residual_sample, metric_sample_key,
                                metric_sample_sign=1.0, *, point_estimates=(), minimize=_newton_cg,
                                minimize_kwargs: Optional[dict] = None, _raise_notconverged: bool = False, _white=None):
    """geoVI update of one residual sample (evi.py:181-255); returns ``(residual, OptimizeResults|None)``.

    ``point_estimates``: the update runs in the space of the non-frozen leaves (:153-178 on ``likelihood.freeze``);
    on full-length vectors this means the frozen entries of the sample stay at ``pos`` and every cotangent /
    operator output is cleared there."""
    frozen = likelihood.frozen_ranges(point_estimates)
    clr = (lambda v: likelihood.clear_frozen(v, frozen)) if frozen else (lambda v: v)
    dist = bool(likelihood.signal.cf.plan.dist)
    # slab-decomposed fields: latent dot products / norms and the data-space sum are all-reduced (replicated leaves once)
    vdot = likelihood.vdot if dist else (lambda a, b: float(torch.dot(a, b)))

    def pos_sumsq(u):
        t = (u * u).sum().to(torch.float64).reshape(1)
        return float(likelihood.signal.cf.plan.comm.allreduce_sum(t)) if dist else float(t)

    e = likelihood.signal.as_flat(pos)
    sample = e + clr(likelihood.signal.as_flat(residual_sample).clone())
    ms, _ = draw_linear_residual(likelihood, e, metric_sample_key, from_inverse=False, point_estimates=point_estimates, _white=_white)
    ms = metric_sample_sign * ms
    mk = dict(minimize_kwargs or {})
    if isinstance(mk.get("maxiter", None), int) and mk["maxiter"] == 0:
        return sample - e, None
    lin_e = likelihood.new_lin()
    lin_e.update(e)
    trafo_at_p = lin_e.transformation()
    lin_x = likelihood.new_lin()

    def residual_vg(x):  # evi.py:153-164
        lin_x.update(x)
        t = lin_x.transformation() - trafo_at_p
        g = x - e + clr(lin_e.lsm(t, scaled=True))
        r = ms - g
        val = 0.5 * vdot(r, r)
        ngrad = r + clr(lin_x.lsm(lin_e.rsm(r, scaled=True), scaled=True))
        return val, -ngrad

    def metric_at(x):  # evi.py:167-172 at the point of the last residual_vg evaluation (== x)
        return HamiltonianMetric(lin_x, other=lin_e, likelihood=likelihood, frozen=frozen)

    def sampnorm(natgrad):  # evi.py:175-178
        fpp = lin_e.rsm(natgrad, scaled=True)
        return float(np.sqrt(vdot(natgrad, natgrad) + pos_sumsq(fpp)))

    # Newton-CG evaluates hessp at `pos` right after fun_and_grad(pos) accepted it, so lin_x is current;
    # after a rejected line-search trial it is re-linearised by the wrapper below.
    state = {"x": None}

    def fg(x):
        state["x"] = x
        return residual_vg(x)

    def op_at(x):
        if state["x"] is None or state["x"].data_ptr() != x.data_ptr():
            lin_x.update(x)
            state["x"] = x
        return metric_at(x)

    if frozen or dist:       # xtol * size counts the liquid entries of the whole model (the reference minimises over the liquid vector)
        mk.setdefault("_size", likelihood.global_size(frozen))
    if dist:
        mk.setdefault("vdot", likelihood.vdot)
        mk.setdefault("vnorm", likelihood.vnorm)
    opt = minimize(None, x0=sample, fun_and_grad=fg, hessp_at=op_at, custom_gradnorm=sampnorm, **mk)
    if _raise_notconverged and (opt.status is None or opt.status < 0):
        raise ValueError("S: failed to invert map")
    return opt.x - e, opt


def draw_residual(likelihood: LikelihoodWithModel, pos, key, *, point_estimates=(), cg=conjugate_gradient.cg, cg_name=None,
                  cg_kwargs=None, minimize=_newton_cg, minimize_kwargs=None, _raise_nonposdef=False, _raise_notconverged=False):
    """Antithetic pair of (optionally non-linearly updated) residuals for one key (evi.py:258-297)."""
    residual, _ = draw_linear_residual(likelihood, pos, key, point_estimates=point_estimates, cg=cg, cg_name=cg_name,
                                       cg_kwargs=cg_kwargs, _raise_nonposdef=_raise_nonposdef)
    out, states = [], []
    for sign in (1.0, -1.0):
        r, st = nonlinearly_update_residual(likelihood, pos, sign * residual, key, sign, point_estimates=point_estimates,
                                            minimize=minimize, minimize_kwargs=minimize_kwargs,
                                            _raise_notconverged=_raise_notconverged)
        out.append(r)
        states.append(st)
    return torch.stack(out), states


# ---- Wiener filter ------------------------------------------------------------------------------------------
def wiener_filter_posterior(likelihood: LikelihoodWithModel, position=None, *, key, n_samples: int = 0, residual_map="lmap",
                            draw_linear_kwargs: Optional[dict] = None, jit=True, model_is_linear: Optional[bool] = True,
                            signal_space: Optional[bool] = True, noise_covariance=None):
    """Wiener-filter posterior of the model linearised at ``position`` (evi.py:399-517); returns
    ``(Samples, (post_info, samples_info))``.

    Signal-space branch (:453-476): ``j = J^T M (d - f(p) + J p)``, ``mean = (J^T M J + 1)^-1 j`` by conjugate
    gradient on the fused metric-vector product, then ``n_samples`` MGVI draws at the mean, mirrored
    (:502-511).  For a model that is linear in the reference's sense ``d - f(p) + J p = d``, so the same formula
    serves ``model_is_linear=True`` (the reference then transposes ``forward`` itself).  Data-space branch
    (``signal_space=False``, :477-497, Gaussian likelihoods): ``(J J^T + N) x = d'``, ``mean = J^T x``.
    ``residual_map`` / ``jit`` are accepted for call compatibility (samples are drawn one after the other)."""
    if not isinstance(likelihood, LikelihoodWithModel):
        raise TypeError(f"likelihood must be of LikelihoodWithModel type; got {likelihood}")
    if not model_is_linear and position is None:
        raise ValueError("For nonlinear models a position to linearize must be specified.")
    if likelihood.signal.cf.plan.dist:
        raise NotImplementedError("wiener_filter_posterior on slab-decomposed fields is not supported yet")
    kw = dict(draw_linear_kwargs or {})
    sig = likelihood.signal
    pos = torch.zeros(sig.layout.size, dtype=likelihood.dtype, device=likelihood.rt.device) if position is None else sig.as_flat(position)
    lin, _ = likelihood.lin_at(pos)
    cg = kw.get("cg", conjugate_gradient.cg)
    if signal_space:
        d_lin = lin.normalized_residual() + lin.rsm(pos, scaled=True)      # M^(1/2) (d - f(p) + J p), M diagonal
        j = lin.lsm(d_lin, scaled=True)
        post_mean, post_info = cg(HamiltonianMetric(lin, likelihood=likelihood), j, name=kw.get("cg_name", None), **kw.get("cg_kwargs", {}))
    else:
        # data-space branch (:477-497): (J J^T + N) x = d', mean = J^T x, with J = l^-1 RSM and J^T = LSM l^-1
        # (l = N^(-1/2), diagonal); the solve runs on data-shaped vectors in the host CG loop
        if noise_covariance is None:
            raise ValueError("To use the Wiener filter in data space, please set the noise_covariance")
        lk = likelihood.likelihood
        if lk.kind != 0:
            raise NotImplementedError("the data-space Wiener filter needs a Gaussian likelihood")
        shape = tuple(sig.target_shape)
        w = lk.w_scalar if lk.w_array is None else likelihood.rt.asarray(lk.w_array, likelihood.dtype)
        l = float(np.sqrt(w)) if lk.w_array is None else torch.sqrt(w)

        def jac(t):
            return lin.rsm(t, scaled=True) / l

        def jac_t(u):
            return lin.lsm(u / l, scaled=True)

        def post_dspace_cov_inv(u_flat):
            u = u_flat.reshape(shape)
            return (jac(jac_t(u)) + noise_covariance(u)).reshape(-1)

        d_lin = lin.normalized_residual() / l + jac(pos)
        x, post_info = cg(post_dspace_cov_inv, d_lin.reshape(-1), name=kw.get("cg_name", None), **kw.get("cg_kwargs", {}))
        post_mean = jac_t(x.reshape(shape))
    if post_info is not None and post_info < 0:
        raise ValueError("conjugate gradient failed")
    if n_samples > 0:
        ks = random_split(key, n_samples)
        drawn = [draw_linear_residual(likelihood, post_mean, k, **kw) for k in ks]
        smpls = torch.stack([d[0] for d in drawn])
        smpls_info = [d[1] for d in drawn]
        out = Samples(pos=post_mean, samples=concatenate_zip(smpls, -smpls), keys=ks)
    else:
        out, smpls_info = Samples(pos=post_mean, samples=None), None
    return out, (post_info, smpls_info)
