"""Oracle (TEST INFRASTRUCTURE ONLY): MGVI / geoVI sampling and the sample-averaged KL.

Restates

* ``draw_linear_residual``          nifty/re/evi.py:88-150  (``sample_likelihood`` :77-80,
                                    ``_ham_metric`` :83-85)
* ``nonlinearly_update_residual``   nifty/re/evi.py:181-255 (residual functions :153-178)
* ``_kl_vg`` / ``_kl_met``          nifty/re/optimize_kl.py:90-144 (``_StandardHamiltonian`` :67-87)

The reference draws its white-noise inputs with ``jax.random`` *outside* of the
arithmetic restated here (evi.py:121-123); the oracle takes those draws as
explicit arrays so that oracle and GPU path can be fed identical inputs.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .solvers import cg, newton_cg


class Layout:
    """Flat-vector view of a parameter dict in sorted-key (= JAX pytree) order."""

    def __init__(self, domain: Dict[str, tuple]):
        self.keys = sorted(domain)
        self.shapes = {k: tuple(domain[k]) for k in self.keys}
        self.offsets = {}
        off = 0
        for k in self.keys:
            self.offsets[k] = off
            off += int(np.prod(self.shapes[k], dtype=np.int64))
        self.size = off

    def pack(self, tree) -> np.ndarray:
        out = np.empty(self.size, dtype=np.float64)
        for k in self.keys:
            n = int(np.prod(self.shapes[k], dtype=np.int64))
            out[self.offsets[k]:self.offsets[k] + n] = np.ravel(np.asarray(tree[k], dtype=np.float64))
        return out

    def unpack(self, vec) -> Dict[str, np.ndarray]:
        out = {}
        for k in self.keys:
            n = int(np.prod(self.shapes[k], dtype=np.int64))
            out[k] = np.reshape(vec[self.offsets[k]:self.offsets[k] + n], self.shapes[k])
        return out

    def random(self, rng: np.random.Generator):
        """Leaf-by-leaf N(0,1) draws in sorted-key order."""
        return {k: rng.standard_normal(self.shapes[k]) for k in self.keys}


def tree_size(domain) -> int:
    return Layout(domain).size


def _flat_ops(lh):
    lay = Layout(lh.domain)

    def metric(pos_v, t_v):
        return lay.pack(lh.metric(lay.unpack(pos_v), lay.unpack(t_v)))

    def ham_metric(pos_v, t_v):  # evi.py:83-85
        return metric(pos_v, t_v) + t_v

    def lsm(pos_v, eta):
        return lay.pack(lh.left_sqrt_metric(lay.unpack(pos_v), eta))

    def rsm(pos_v, t_v):
        return lh.right_sqrt_metric(lay.unpack(pos_v), lay.unpack(t_v))

    def trafo(pos_v):
        return lh.transformation(lay.unpack(pos_v))

    return lay, metric, ham_metric, lsm, rsm, trafo


class _Frozen:
    """``Likelihood.freeze(point_estimates=..., primals=pos)`` (likelihood.py:399-499, ``_parse_point_estimates``
    :57-92) on packed vectors: the operators of the likelihood restricted to the non-frozen ("liquid") leaves, the
    frozen ones inserted at their value in ``pos`` (``partial_insert_and_remove`` :119-177).  Liquid vectors are the
    packed vector with the frozen entries REMOVED, exactly as the reference minimises / solves over them."""

    def __init__(self, lh, pos_v, point_estimates):
        self.lay, self.metric_full, _, self.lsm_full, self.rsm_full, self.trafo_full = _flat_ops(lh)
        keys = [point_estimates] if isinstance(point_estimates, str) else list(point_estimates)
        frozen = np.zeros(self.lay.size, dtype=bool)
        for k in keys:
            n = int(np.prod(self.lay.shapes[k], dtype=np.int64))
            frozen[self.lay.offsets[k]:self.lay.offsets[k] + n] = True
        self.liquid = np.flatnonzero(~frozen)
        self.pos_v = np.array(pos_v, dtype=np.float64)

    def remove(self, v):
        return np.asarray(v)[self.liquid]

    def insert(self, vl, fill=None):
        out = np.array(self.pos_v if fill is None else np.full(self.lay.size, fill, dtype=np.float64))
        out[self.liquid] = vl
        return out

    def metric(self, xl, tl):
        return self.remove(self.metric_full(self.insert(xl), self.insert(tl, 0.0)))

    def lsm(self, xl, eta):
        return self.remove(self.lsm_full(self.insert(xl), eta))

    def rsm(self, xl, tl):
        return self.rsm_full(self.insert(xl), self.insert(tl, 0.0))

    def trafo(self, xl):
        return self.trafo_full(self.insert(xl))


def draw_linear_residual(lh, pos, white_data, white_prior, *, from_inverse=True, cg_kwargs=None,
                         _raise_nonposdef=False, point_estimates=()):
    """One MGVI sample; returns ``(residual dict, info, CGResult|None)``.

    ``white_data`` (data shape) and ``white_prior`` (latent dict) are the N(0,1)
    draws of evi.py:122-123.
    """
    lay, _, ham_metric, lsm, _, _ = _flat_ops(lh)
    pos_v = lay.pack(pos)
    if point_estimates:
        # evi.py:109-111, 121-123, 149: everything happens on the liquid leaves, zeros are inserted at the end
        fz = _Frozen(lh, pos_v, point_estimates)
        pl = fz.remove(pos_v)
        nll_smpl = fz.lsm(pl, white_data)
        prr = fz.remove(lay.pack(white_prior))
        smpl = nll_smpl + prr
        info, res = 0, None
        if from_inverse:
            kw = dict(cg_kwargs or {})
            kw.pop("name", None)
            res = cg(lambda v: fz.metric(pl, v) + v, smpl, x0=prr, _raise_nonposdef=_raise_nonposdef, **kw)
            smpl, info = res.x, res.info
            if info < 0:
                raise ValueError("conjugate gradient failed")
        return lay.unpack(fz.insert(smpl, 0.0)), info, res
    nll_smpl = lsm(pos_v, white_data)
    prr = lay.pack(white_prior)
    smpl = nll_smpl + prr
    info, res = 0, None
    if from_inverse:
        kw = dict(cg_kwargs or {})
        kw.pop("name", None)
        res = cg(lambda v: ham_metric(pos_v, v), smpl, x0=prr,
                 _raise_nonposdef=_raise_nonposdef, **kw)
        smpl, info = res.x, res.info
        if info < 0:
            raise ValueError("conjugate gradient failed")
    return lay.unpack(smpl), info, res


def wiener_filter_posterior_mean(lh, pos, *, cg_kwargs=None):
    """Posterior mean of the model linearised at ``pos``; signal-space branch of
    ``wiener_filter_posterior`` (evi.py:453-476 with the linearised data of :457-458):
    ``j = J^T M (d - f(pos) + J pos)``, ``mean = (J^T M J + 1)^-1 j`` by conjugate gradient.
    Returns ``(mean dict, info, CGResult)``."""
    lay, _, ham_metric, lsm, rsm, _ = _flat_ops(lh)
    pos_v = lay.pack(pos)
    d_lin = lh.normalized_residual(pos) + rsm(pos_v, pos_v)      # M^(1/2) (d - f(pos) + J pos), M diagonal
    j = lsm(pos_v, d_lin)
    kw = dict(cg_kwargs or {})
    kw.pop("name", None)
    res = cg(lambda v: ham_metric(pos_v, v), j, **kw)
    if res.info < 0:
        raise ValueError("conjugate gradient failed")
    return lay.unpack(res.x), res.info, res


def nonlinearly_update_residual(lh, pos, residual_sample, white_data, white_prior,
                                metric_sample_sign=1.0, *, minimize_kwargs=None, point_estimates=()):
    """geoVI update of one residual sample (evi.py:181-255); returns ``(residual, NewtonResult)``."""
    lay, _, _, lsm, rsm, trafo = _flat_ops(lh)
    e = lay.pack(pos)
    sample = e + lay.pack(residual_sample)
    ms, _, _ = draw_linear_residual(lh, pos, white_data, white_prior, from_inverse=False, point_estimates=point_estimates)
    ms = metric_sample_sign * lay.pack(ms)
    mk = dict(minimize_kwargs or {})
    mk.pop("name", None)
    if isinstance(mk.get("maxiter", None), int) and mk["maxiter"] == 0:
        return lay.unpack(sample - e), None
    if point_estimates:
        # evi.py:224-254: sample, metric sample and expansion point with the frozen leaves removed; the frozen
        # likelihood (:153-178) supplies transformation / sqrt-metrics; zeros are inserted into the result
        fz = _Frozen(lh, e, point_estimates)
        full_e = e
        sample, ms, e = fz.remove(sample), fz.remove(ms), fz.remove(e)
        lsm, rsm, trafo = fz.lsm, fz.rsm, fz.trafo
    trafo_at_p = trafo(e)

    def residual_vg(x):  # evi.py:153-164
        t = trafo(x) - trafo_at_p
        g = x - e + lsm(e, t)
        r = ms - g
        val = 0.5 * float(np.vdot(r, r))
        ngrad = r + lsm(x, rsm(e, r))
        return val, -ngrad

    def metric(x, t):  # evi.py:167-172
        tm = lsm(e, rsm(x, t)) + t
        return lsm(x, rsm(e, tm)) + tm

    def sampnorm(natgrad):  # evi.py:175-178
        fpp = rsm(e, natgrad)
        return float(np.sqrt(np.vdot(natgrad, natgrad) + np.vdot(fpp, fpp)))

    opt = newton_cg(sample, residual_vg, metric, custom_gradnorm=sampnorm, **mk)
    if point_estimates:
        return lay.unpack(fz.insert(opt.x - e, 0.0)), opt
    return lay.unpack(opt.x - e), opt


def kl_minimize(lh, pos, residuals, *, constants=(), minimize_kwargs=None):
    """``OptimizeVI.kl_minimize`` (optimize_kl.py:540-591): Newton-CG on the sample-averaged Hamiltonian; with
    ``constants`` over the non-constant leaves only (:553-573), the constant ones re-inserted into ``x`` (:581-590).
    Returns ``(position dict, NewtonResult)``."""
    lay = Layout(lh.domain)
    p = lay.pack(pos)
    mk = dict(minimize_kwargs or {})
    mk.pop("name", None)

    def vg_full(x_v):
        val, g = kl_value_and_grad(lh, lay.unpack(x_v), residuals)
        return val, lay.pack(g)

    def met_full(x_v, t_v):
        return lay.pack(kl_metric(lh, lay.unpack(x_v), lay.unpack(t_v), residuals))

    if not constants:
        opt = newton_cg(p, vg_full, met_full, **mk)
        return lay.unpack(opt.x), opt
    fz = _Frozen(lh, p, constants)

    def vg(xl):
        val, g = vg_full(fz.insert(xl))
        return val, fz.remove(g)

    def met(xl, tl):
        return fz.remove(met_full(fz.insert(xl), fz.insert(tl, 0.0)))

    opt = newton_cg(fz.remove(p), vg, met, **mk)
    return lay.unpack(fz.insert(opt.x)), opt


def _ham_vg(lh, lay, x_v):
    e, g = lh.energy_and_gradient(lay.unpack(x_v))
    return e + 0.5 * float(np.vdot(x_v, x_v)), lay.pack(g) + x_v


def kl_value_and_grad(lh, pos, residuals: Sequence[dict]):
    """Mean over samples of value_and_grad of the standard Hamiltonian (optimize_kl.py:90-114)."""
    lay = Layout(lh.domain)
    p = lay.pack(pos)
    if len(residuals) == 0:
        v, g = _ham_vg(lh, lay, p)
        return v, lay.unpack(g)
    vals, grads = [], []
    for r in residuals:
        v, g = _ham_vg(lh, lay, p + lay.pack(r))
        vals.append(v)
        grads.append(g)
    return float(np.mean(vals)), lay.unpack(np.mean(np.stack(grads), axis=0))


def kl_metric(lh, pos, tangents, residuals: Sequence[dict]):
    """Mean over samples of ``metric(pos + r_i, t) + t`` (optimize_kl.py:117-144)."""
    lay = Layout(lh.domain)
    p, t = lay.pack(pos), lay.pack(tangents)
    if len(residuals) == 0:
        return lay.unpack(lay.pack(lh.metric(pos, tangents)) + t)
    outs = []
    for r in residuals:
        x = lay.unpack(p + lay.pack(r))
        outs.append(lay.pack(lh.metric(x, tangents)) + t)
    return lay.unpack(np.mean(np.stack(outs), axis=0))
