"""Correlated fields on OUTER PRODUCTS of several sub-grids (``nifty/re/correlated_field.py:856-912``: one amplitude spectrum
per sub-grid combined with ``tensordot(axes=0)``, one Hartley transform per sub-grid over its own axes) -- the space x frequency /
space x time models.  SURVEY.md section 8(f)1, first version.

What runs where.  The N-sized transform is the device Hartley transform of this library (``nb200_hartley`` on the JOINT grid);
the composite ``H_2 H_1`` is obtained from the joint transform ``J`` by the reflect-combine identity

    H_2 H_1 x = 1/2 ( J x + (J x) o f_1 + (J x) o f_2 - (J x) o f_1 o f_2 ),      f_i : k -> -k along the axes of sub-grid i

(both Hartley conventions; linear and self-adjoint like its factors).  The O(K_i) amplitude spectra, the outer product, the
reflections and the pointwise likelihood are torch operations on the same device, and derivatives come from torch autograd
with the composite transform as a self-adjoint autograd function -- this is a HOST-COMPOSED path: correct and checked against
the oracle, but not fused into the pass kernels the single-grid hot path uses and not roofline-grade (every operator application
makes ~10 N-sized elementwise passes besides the transform).  It provides the model (``cf(p)``, ``domain``, ``init``,
``normalized_amplitudes``, ``target_grids``) and the operator-level likelihood interface (energy, gradient, metric,
sqrt-metrics, a CG solve and MGVI sample draws in the host loop, the sample-averaged KL and an MGVI driver ``mgvi``: linear
samples + Newton-CG on the KL); geoVI updates and the ``optimize_kl`` state machine are not wired up for such models.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ._runtime import Plan
from .model import LazyModel
from .tree import Layout


def _amplitude(spec, tabs, p, prefix):
    """``NonParametricAmplitude.__call__`` (correlated_field.py:481-516) with ``integrated_wiener_process``
    (gauss_markov.py:102-114) and the priors of num/stats_distributions.py:42-98 as differentiable torch operations."""
    ell, mult, dt, V = tabs["ell"], tabs["mult"], tabs["dt"], tabs["V"]
    flu = spec["flu"](p[prefix + "fluctuations"]) if spec["flu"] is not None else torch.ones((), dtype=ell.dtype, device=ell.device)
    slope = spec["slp"](p[prefix + "loglogavgslope"])
    u = slope * ell
    if spec["has_dev"]:
        xi = p[prefix + "spectrum"]
        sig = spec["flx"](p[prefix + "flexibility"])
        asp = spec["asp"](p[prefix + "asperity"]) if spec["asp"] is not None else torch.zeros((), dtype=ell.dtype, device=ell.device)
        sd = sig * torch.sqrt(dt)
        q = torch.sqrt(dt * dt / 12.0 + asp)
        r1 = sd * xi[:, 1]
        r0 = sd * xi[:, 0] * q + 0.5 * dt * r1
        zero = torch.zeros(1, dtype=ell.dtype, device=ell.device)
        y = torch.cat((zero, torch.cumsum(r1, 0)))
        x = torch.cumsum(torch.cat((zero, r0 + dt * y[:-1])), 0)
        tw = torch.cat((zero, x))
        u = u + tw - tw[-1] * (ell / ell[-1])
    P = torch.exp(u)
    if spec["kind"] == "power":
        S, shape_fn = torch.sum(mult[1:] * P[1:]), torch.sqrt(P)
    else:
        S, shape_fn = torch.sum(mult[1:] * P[1:] ** 2), P
    amp = flu * (np.sqrt(V) / (torch.sqrt(S) / np.sqrt(V))) * shape_fn
    return torch.cat((torch.full((1,), V, dtype=amp.dtype, device=amp.device), amp[1:]))


class _Prior:
    """(log-)normal reparametrisation of a scalar leaf on torch scalars."""

    def __init__(self, prior):
        self.a, self.b = prior.ab()
        self.log = prior.kind == "lognormal"

    def __call__(self, xi):
        v = self.a + self.b * torch.as_tensor(xi).reshape(())
        return torch.exp(v) if self.log else v


class OuterCorrelatedField(LazyModel):
    """The finalised model of a :class:`~nifty_b200.correlated_field.CorrelatedFieldMaker` with several ``add_fluctuations``."""

    def __init__(self, prefix, offset_mean, azm_prior, flucts, *, dtype, convention, runtime):
        self.prefix, self.offset_mean, self.dtype = prefix, float(offset_mean), dtype
        self._azm = _Prior(azm_prior)
        if len(flucts) != 2:
            raise NotImplementedError("outer products of exactly two sub-grids are supported")
        shape, dists, self._axes, self._subs = (), (), [], []
        for f in flucts:
            if f.get("matern"):
                raise NotImplementedError("Matern amplitudes inside outer products are not supported")
            sub_plan = Plan(f["shape"], f["distances"], dtype=dtype, hartley_convention=convention, runtime=runtime)
            n0 = len(shape)
            shape += tuple(f["shape"])
            d = f["distances"]
            dists += tuple(float(x) for x in (d if np.ndim(d) else (d,) * len(f["shape"])))
            self._axes.append(tuple(range(n0, len(shape))))
            self._subs.append((f, sub_plan))
        if len(shape) > 3:
            raise NotImplementedError("outer products with more than three axes in total are not supported")
        self.plan = Plan(shape, dists, dtype=dtype, hartley_convention=convention, runtime=runtime)     # the JOINT transform
        self.rt = self.plan.rt
        dev = self.rt.device
        self.shape = shape
        domain = {prefix + "zeromode": ()}
        self._tabs, self._specs = [], []
        for f, sp in self._subs:
            pf = prefix + f["prefix"]
            has_dev = f["flx"] is not None and sp.K > 2
            spec = dict(kind=f["kind"], has_dev=has_dev, flu=None if f["flu"] is None else _Prior(f["flu"]), slp=_Prior(f["slp"]),
                        flx=_Prior(f["flx"]) if has_dev else None, asp=_Prior(f["asp"]) if (has_dev and f["asp"] is not None) else None,
                        pf=pf)
            if f["flu"] is not None:
                domain[pf + "fluctuations"] = ()
            domain[pf + "loglogavgslope"] = ()
            if has_dev:
                domain[pf + "flexibility"] = ()
                if f["asp"] is not None:
                    domain[pf + "asperity"] = ()
                domain[pf + "spectrum"] = (sp.K - 2, 2)
            t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dtype, device=dev)      # noqa: E731
            self._tabs.append(dict(ell=t(sp.relative_log_mode_lengths), mult=t(sp.mode_multiplicity), dt=t(sp.log_volume),
                                   V=float(sp.total_volume),
                                   pd=torch.as_tensor(np.asarray(sp.power_distributor, dtype=np.int64), device=dev)))
            self._specs.append(spec)
        domain[prefix + "xi"] = shape
        self.domain = dict(sorted(domain.items()))
        self.layout = Layout(self.domain)
        self.target_shape = shape
        self._vol = float(np.prod([tb["V"] for tb in self._tabs]))
        outer = self

        class SepHartley(torch.autograd.Function):
            """x -> H_n ... H_1 x through the joint device transform and the reflect-combine identity; self-adjoint."""

            @staticmethod
            def forward(ctx, x):
                return outer._sep_hartley(x)

            @staticmethod
            def backward(ctx, g):
                return SepHartley.apply(g)

        self._sep = SepHartley

    # -- transform ---------------------------------------------------------------------------------------------------
    def _sep_hartley(self, x: torch.Tensor) -> torch.Tensor:
        """The composite transform of two sub-grids by the identity of the module docstring."""
        J = self.plan.hartley(x.detach().contiguous())
        a1, a2 = self._axes

        def refl(t, axes):
            return torch.roll(torch.flip(t, dims=axes), shifts=[1] * len(axes), dims=axes)

        return 0.5 * (J + refl(J, a1) + refl(J, a2) - refl(J, a1 + a2))

    # -- model surface -----------------------------------------------------------------------------------------------
    @property
    def target(self):
        return self.target_shape

    def init(self, seed):
        return self.layout.unpack(self.layout.random(seed, self.dtype, self.rt.device))

    def _tree(self, pos):
        pos = getattr(pos, "tree", pos)
        if isinstance(pos, torch.Tensor):
            return self.layout.unpack(pos)
        return {k: torch.as_tensor(v, dtype=self.dtype, device=self.rt.device) if not isinstance(v, torch.Tensor) else v.to(self.rt.device)
                for k, v in pos.items()}

    def _normalized(self, p):
        z = self._azm(p[self.prefix + "zeromode"])
        nas = []
        for spec, tb in zip(self._specs, self._tabs):
            a = _amplitude(spec, tb, p, spec["pf"])
            nas.append(torch.cat((a[:1], a[1:] / z)))                    # get_normalized_amplitudes (:807-821)
        return z, nas

    @property
    def normalized_amplitudes(self):
        return tuple((lambda pos, i=i: self._normalized(self._tree(pos))[1][i]) for i in range(len(self._specs)))

    @property
    def target_grids(self):
        from .correlated_field import _grid_record
        return tuple(_grid_record(sp) for _, sp in self._subs)

    def __call__(self, pos) -> torch.Tensor:
        """correlated_field.py:889-912; differentiable with respect to every leaf (torch autograd)."""
        p = self._tree(pos)
        z, nas = self._normalized(p)
        ea = None
        for na, tb, axes in zip(nas, self._tabs, self._axes):
            shp = [1] * len(self.shape)
            for ax in axes:
                shp[ax] = self.shape[ax]
            e = na[tb["pd"]].reshape(shp)                                   # expanded amplitude, broadcast over the other sub-grids
            ea = e if ea is None else ea * e                                # tensordot(axes=0) (:900-907)
        h = z * ea * p[self.prefix + "xi"]
        return self.offset_mean + self._sep.apply(h) / self._vol


class OuterLikelihood:
    """Operator-level likelihood interface (``likelihood.py:599-633``) of ``Gaussian`` / ``Poissonian`` data on
    ``signal = exp(cf)`` (or ``cf`` itself) for an :class:`OuterCorrelatedField`: energy, gradient, metric, sqrt-metrics, all
    through torch autograd around the device transform.  Trees in, trees out."""

    def __init__(self, likelihood, cf: OuterCorrelatedField, nonlinearity="exp"):
        if nonlinearity not in ("exp", "identity") and not callable(nonlinearity):
            raise ValueError(f"unsupported nonlinearity {nonlinearity!r}")
        self.cf, self.kind, self.rt, self.dtype = cf, likelihood.kind, cf.rt, cf.dtype
        self.nl = torch.exp if nonlinearity == "exp" else ((lambda f: f) if nonlinearity == "identity" else nonlinearity)
        dev = cf.rt.device
        self.data = torch.as_tensor(np.asarray(likelihood.data), dtype=cf.dtype, device=dev) if not isinstance(likelihood.data, torch.Tensor) \
            else likelihood.data.to(dtype=cf.dtype, device=dev)
        if self.kind == 0:
            w = likelihood.w_array if likelihood.w_array is not None else likelihood.w_scalar
            self.w = torch.as_tensor(w, dtype=cf.dtype, device=dev)
        self.layout = cf.layout

    def signal_response(self, pos):
        with torch.no_grad():
            return self.nl(self.cf(pos))

    def _lh_energy(self, s):
        if self.kind == 0:                                   # Gaussian (likelihood_impl.py:124-126)
            r = s - self.data
            return 0.5 * torch.sum(self.w * r * r)
        return torch.sum(s) - torch.sum(self.data * torch.log(s))      # Poissonian (:238-240)

    def _lh_metric_weight(self, s):
        return self.w * torch.ones_like(s) if self.kind == 0 else 1.0 / s      # :131-132, :245-246

    def _leaves(self, pos, grad=True):
        p = self.cf._tree(pos)
        return {k: v.detach().clone().requires_grad_(grad) for k, v in p.items()}

    def energy(self, pos) -> float:
        with torch.no_grad():
            return float(self._lh_energy(self.nl(self.cf(pos))))

    def energy_and_gradient(self, pos):
        p = self._leaves(pos)
        e = self._lh_energy(self.nl(self.cf(p)))
        keys = sorted(p)
        g = torch.autograd.grad(e, [p[k] for k in keys], allow_unused=True)
        return float(e.detach()), {k: (torch.zeros_like(p[k]) if gi is None else gi.detach()) for k, gi in zip(keys, g)}

    def _jvp(self, p, tan):
        """J t of signal = nl(cf(p)) by the double-backward construction (the transform is its own adjoint)."""
        keys = sorted(p)
        s = self.nl(self.cf(p))
        u = torch.zeros_like(s, requires_grad=True)
        g = torch.autograd.grad(s, [p[k] for k in keys], grad_outputs=u, create_graph=True, allow_unused=True)
        acc = sum(torch.sum(gi * tan[k]) for k, gi in zip(keys, g) if gi is not None)
        (jt,) = torch.autograd.grad(acc, u)
        return s.detach(), jt.detach()

    def _vjp(self, p, c):
        keys = sorted(p)
        s = self.nl(self.cf(p))
        g = torch.autograd.grad(s, [p[k] for k in keys], grad_outputs=c, allow_unused=True)
        return {k: (torch.zeros_like(p[k]) if gi is None else gi.detach()) for k, gi in zip(keys, g)}

    def metric(self, pos, tan):
        """``J^T M J t`` (likelihood.py:613-621)."""
        p = self._leaves(pos)
        t = {k: torch.as_tensor(v, dtype=self.dtype, device=self.rt.device) for k, v in getattr(tan, "tree", tan).items()}
        s, jt = self._jvp(p, t)
        return self._vjp(p, self._lh_metric_weight(s) * jt)

    def right_sqrt_metric(self, pos, tan):
        p = self._leaves(pos)
        t = {k: torch.as_tensor(v, dtype=self.dtype, device=self.rt.device) for k, v in getattr(tan, "tree", tan).items()}
        s, jt = self._jvp(p, t)
        return torch.sqrt(self._lh_metric_weight(s)) * jt

    def left_sqrt_metric(self, pos, u):
        p = self._leaves(pos)
        with torch.no_grad():
            s = self.nl(self.cf(p))
        u = torch.as_tensor(u, dtype=self.dtype, device=self.rt.device)
        return self._vjp(p, torch.sqrt(self._lh_metric_weight(s)) * u)

    def cg_on_metric(self, pos, j, **cg_kwargs):
        """``cg(metric + 1, j)`` on flat vectors in the host loop (conjugate_gradient.py:77-214)."""
        from .conjugate_gradient import _cg
        lay = self.layout

        def mat(v):
            return lay.pack(self.metric(pos, lay.unpack(v)), self.dtype, self.rt.device) + v

        return _cg(mat, lay.pack(getattr(j, "tree", j), self.dtype, self.rt.device) if not isinstance(j, torch.Tensor) else j, **cg_kwargs)

    def draw_linear_residual(self, pos, key, *, from_inverse: bool = True, cg_kwargs: Optional[dict] = None, _raise_nonposdef=False,
                             _white=None):
        """One MGVI residual sample at ``pos`` (evi.py:88-150): ``left_sqrt_metric(pos, N(0,1)) + N(0,1)`` as the right-hand side
        and ``x0 = `` the prior draw of a CG on ``metric + 1`` (host loop).  Returns ``(residual tree, info)``; the mirrored
        sample is its negative (evi.py:53-57)."""
        from .conjugate_gradient import _cg
        from .evi import random_normal, random_split
        lay, dev = self.layout, self.rt.device
        k_nll, k_prr = random_split(key, 2)
        w_data, w_prior = (None, None) if _white is None else _white
        white = random_normal(k_nll, self.cf.target_shape, self.dtype, dev) if w_data is None else torch.as_tensor(w_data, dtype=self.dtype, device=dev)
        prr = random_normal(k_prr, (lay.size,), self.dtype, dev) if w_prior is None else \
            lay.pack(getattr(w_prior, "tree", w_prior), self.dtype, dev) if not isinstance(w_prior, torch.Tensor) else w_prior.to(dev)
        smpl = lay.pack(self.left_sqrt_metric(pos, white), self.dtype, dev) + prr
        info = 0
        if from_inverse:
            def mat(v):
                return lay.pack(self.metric(pos, lay.unpack(v)), self.dtype, dev) + v
            res = _cg(mat, smpl, x0=prr.clone(), _raise_nonposdef=_raise_nonposdef, **(cg_kwargs or {}))
            smpl, info = res.x, res.info
            if info is not None and info < 0:
                raise ValueError("conjugate gradient failed")
        return lay.unpack(smpl), info

    # -- sample-averaged KL and MGVI iterations (optimize_kl.py:67-144, 391-444, 540-591), on flat vectors ------------------
    def _flat(self, pos):
        pos = getattr(pos, "tree", pos)
        return pos.to(dtype=self.dtype, device=self.rt.device) if isinstance(pos, torch.Tensor) else self.layout.pack(pos, self.dtype, self.rt.device)

    def kl_value_and_grad(self, pos, residuals):
        """``_kl_vg``: mean over ``pos + residuals`` of value and gradient of the standard Hamiltonian ``lh(x) + <x, x> / 2``."""
        p, lay = self._flat(pos), self.layout
        pts = [p] if residuals is None or len(residuals) == 0 else [p + self._flat(r) for r in residuals]
        val, grad = 0.0, torch.zeros_like(p)
        for x in pts:
            e, g = self.energy_and_gradient(lay.unpack(x))
            val += e + 0.5 * float(torch.dot(x, x))
            grad += lay.pack(g, self.dtype, self.rt.device) + x
        return val / len(pts), grad / len(pts)

    def kl_metric(self, pos, tangents, residuals):
        """``_kl_met``: mean over the sample points of ``metric(x, t) + t``."""
        p, t, lay = self._flat(pos), self._flat(tangents), self.layout
        pts = [p] if residuals is None or len(residuals) == 0 else [p + self._flat(r) for r in residuals]
        out = torch.zeros_like(p)
        for x in pts:
            out += lay.pack(self.metric(lay.unpack(x), lay.unpack(t)), self.dtype, self.rt.device) + t
        return out / len(pts)

    def mgvi(self, pos, *, key, n_total_iterations: int, n_samples: int, draw_linear_kwargs=None, kl_kwargs=None, _whites=None):
        """MGVI (``optimize_kl(..., sample_mode="linear_resample")``, optimize_kl.py:672-729) on the host-composed operators:
        per iteration ``n_samples`` residual draws (mirrored: ``[s, -s]`` interleaved, evi.py:53-57) and one Newton-CG
        minimisation of the sample-averaged KL.  Returns ``(position tree, residuals [2 n_samples, L], list of
        OptimizeResults)``."""
        from .evi import random_split
        from .optimize import _newton_cg
        lay = self.layout
        p = self._flat(pos).clone()
        dkw = dict(draw_linear_kwargs or {})
        mk = dict((kl_kwargs or {}).get("minimize_kwargs", {}))
        states, res = [], None
        for it in range(n_total_iterations):
            key, sk = random_split(key, 2)
            ks = random_split(sk, n_samples)
            rows = []
            for i, k in enumerate(ks):
                w = None if _whites is None else _whites[it * n_samples + i]
                r, _ = self.draw_linear_residual(lay.unpack(p), k, _white=w, **dkw)
                r = lay.pack(r, self.dtype, self.rt.device)
                rows += [r, -r]
            res = torch.stack(rows)
            opt = _newton_cg(None, x0=p, fun_and_grad=lambda x: self.kl_value_and_grad(x, res),
                             hessp=lambda x, t: self.kl_metric(x, t, res), **mk)
            p = opt.x
            states.append(opt._replace(x=None, jac=None))
        return lay.unpack(p), res, states

