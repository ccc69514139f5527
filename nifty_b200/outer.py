"""Correlated fields on OUTER PRODUCTS of several sub-grids (``nifty/re/correlated_field.py:856-912``: one amplitude spectrum
per sub-grid combined with ``tensordot(axes=0)``, one Hartley transform per sub-grid over its own axes) -- the space x frequency /
space x time models.  SURVEY.md section 8(f)1, first version.

What runs where.  The N-sized transform is the device Hartley transform of this library (``nb200_hartley`` on the JOINT grid);
the composite ``H_2 H_1`` is obtained from the joint transform ``J`` by the reflect-combine identity

    H_2 H_1 x = 1/2 ( J x + (J x) o f_1 + (J x) o f_2 - (J x) o f_1 o f_2 ),      f_i : k -> -k along the axes of sub-grid i

and its recursion for three sub-grids (three 1-D grids: a device plan has at most three axes; with more axes in total every
sub-grid is transformed on its own axes, one device call per slice of the other axes)

(both Hartley conventions; linear and self-adjoint like its factors).  The O(K_i) amplitude spectra, the outer product, the
reflections and the pointwise likelihood are torch operations on the same device -- a HOST-COMPOSED path: correct and checked
against the oracle, but not fused into the pass kernels the single-grid hot path uses and not roofline-grade.

Linearisations are explicit (:class:`_FieldLin`): one forward evaluation per position, then ``J t`` and ``J^T c`` cost ONE device
transform each plus a handful of N-sized elementwise passes; the O(K) amplitude chains are differentiated with ``torch.func``.
:class:`HostLin` gives such a linearisation the flat-vector interface of the device ``Lin`` and :class:`OuterLikelihood` is a
``LikelihoodWithModel``, so ``draw_linear_residual``, ``nonlinearly_update_residual`` (geoVI), ``optimize_kl`` (state machine,
point estimates, checkpoints), ``wiener_filter_posterior`` and the minisanity message run unchanged on such models; their CG
solves use the host loop (``conjugate_gradient._cg``).  The model itself (``cf(p)``) stays differentiable through torch
autograd with the composite transform as a self-adjoint autograd function.  The non-power-of-two fields of ``bluestein.py``
share all of this.
"""
from __future__ import annotations

import numpy as np
import torch

from ._runtime import Plan
from .likelihood import LikelihoodWithModel
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


def _matern_amplitude(spec, tabs, p, prefix):
    """``MaternAmplitude.__call__`` (correlated_field.py:361-395) as differentiable torch operations."""
    um, mult, V = tabs["um"], tabs["mult"], tabs["V"]
    scl = spec["scl"](p[prefix + "scale"])
    ctf = spec["ctf"](p[prefix + "cutoff"])
    slp = spec["slp"](p[prefix + "loglogslope"])
    spectrum = torch.exp(0.25 * slp * torch.log1p((um / ctf) ** 2))
    norm = 1.0
    if spec["renorm"]:
        S = torch.sum(mult[1:] * spectrum[1:] ** 2) if spec["kind"] == "amplitude" else torch.sum(mult[1:] * spectrum[1:])
        norm = torch.sqrt(S) / np.sqrt(V)
    if spec["kind"] == "power":
        spectrum = torch.sqrt(spectrum)
    spectrum = scl * (np.sqrt(V) / norm) * spectrum
    return torch.cat((torch.full((1,), V, dtype=spectrum.dtype, device=spectrum.device), spectrum[1:]))


def _amplitude_spec(f, pf, K):
    """(spec, latent leaves {name: shape}) of one sub-grid's amplitude model from the maker's record ``f``
    (``add_fluctuations`` :572-659 / ``add_fluctuations_matern`` :661-755)."""
    if f.get("matern"):
        spec = dict(matern=True, kind=f["kind"], renorm=f["renorm"], scl=_Prior(f["scl"]), ctf=_Prior(f["ctf"]), slp=_Prior(f["slp"]), pf=pf)
        return spec, {pf + "scale": (), pf + "cutoff": (), pf + "loglogslope": ()}
    has_dev = f["flx"] is not None and K > 2
    spec = dict(matern=False, kind=f["kind"], has_dev=has_dev, flu=None if f["flu"] is None else _Prior(f["flu"]), slp=_Prior(f["slp"]),
                flx=_Prior(f["flx"]) if has_dev else None, asp=_Prior(f["asp"]) if (has_dev and f["asp"] is not None) else None, pf=pf)
    leaves = {}
    if f["flu"] is not None:
        leaves[pf + "fluctuations"] = ()
    leaves[pf + "loglogavgslope"] = ()
    if has_dev:
        leaves[pf + "flexibility"] = ()
        if f["asp"] is not None:
            leaves[pf + "asperity"] = ()
        leaves[pf + "spectrum"] = (K - 2, 2)
    return spec, leaves


def _eval_amplitude(spec, tabs, p):
    return _matern_amplitude(spec, tabs, p, spec["pf"]) if spec["matern"] else _amplitude(spec, tabs, p, spec["pf"])


def _device_tables(tb, dtype, dev):
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dtype, device=dev)      # noqa: E731
    return dict(ell=t(tb["relative_log_mode_lengths"]), mult=t(tb["mode_multiplicity"]), dt=t(tb["log_volume"]), um=t(tb["mode_lengths"]),
                V=float(tb["total_volume"]), pd=torch.as_tensor(np.asarray(tb["power_distributor"], dtype=np.int64), device=dev))


class _FieldLin:
    """The field of a host-composed model linearised at one position: ``J t`` and ``J^T c`` with ONE device transform each
    (what ``jax.linearize`` / ``jax.linear_transpose`` hand the reference, likelihood.py:613-621).  The field is
    ``offset + T(prod_i B_i[pd_i] * xi) / V`` with O(K_i) factor tables ``B_i`` of the non-xi leaves (``cf._factor_tables``):
    their derivatives are O(K) torch.func calls, the N-sized part is written out -- tangent ``T(ea dxi + d(ea) xi) / V``,
    cotangent ``g = T(c) / V``, ``xi_bar = ea g``, ``B_i_bar = `` segment sum over the bins of ``xi g prod_{j != i} B_j[pd_j]``."""

    def __init__(self, cf, p):
        self.cf, self.xi_key = cf, cf.prefix + "xi"
        self.keys = sorted(k for k in p if k != self.xi_key)
        self.vals = tuple(p[k].detach() for k in self.keys)
        self.xi = p[self.xi_key].detach()
        self._fn = lambda *v: tuple(cf._factor_tables(dict(zip(self.keys, v))))
        self.tabs, self._tab_vjp = torch.func.vjp(self._fn, *self.vals)
        self.exp = [self._expand(i, t) for i, t in enumerate(self.tabs)]
        self.ea = self._prod(self.exp)
        self.field = cf.offset_mean + cf._raw_transform(self.ea * self.xi) / cf._vol

    def _expand(self, i, table):
        pd, axes = self.cf._factors[i]
        shp = [1] * len(self.cf.target_shape)
        for ax in axes:
            shp[ax] = self.cf.target_shape[ax]
        return table[pd].reshape(shp)

    @staticmethod
    def _prod(ts):
        out = ts[0]
        for t in ts[1:]:
            out = out * t
        return out

    def jvp(self, t):
        _, dts = torch.func.jvp(self._fn, self.vals, tuple(t[k].to(v.dtype).reshape(v.shape) for k, v in zip(self.keys, self.vals)))
        dea = None
        for i, dt in enumerate(dts):
            term = self._prod([self._expand(i, dt)] + [e for j, e in enumerate(self.exp) if j != i])
            dea = term if dea is None else dea + term
        return self.cf._raw_transform(self.ea * t[self.xi_key] + dea * self.xi) / self.cf._vol

    def vjp(self, c):
        g = self.cf._raw_transform(c) / self.cf._vol
        w = self.xi * g
        bars = []
        for i, tab in enumerate(self.tabs):
            pd, axes = self.cf._factors[i]
            wi = self._prod([w] + [e for j, e in enumerate(self.exp) if j != i])
            other = tuple(ax for ax in range(w.ndim) if ax not in axes)
            if other:
                wi = wi.sum(dim=other)
            bars.append(torch.zeros_like(tab).index_add_(0, pd.reshape(-1), wi.reshape(-1)))
        out = dict(zip(self.keys, self._tab_vjp(tuple(bars))))
        out[self.xi_key] = self.ea * g
        return out


class _HostLin:
    """``signal = nl(field)`` under a likelihood, linearised at one position: metric and sqrt-metrics on trees."""

    def __init__(self, lh, pos):
        self.lh = lh
        p = lh._tree(pos)
        self.fl = _FieldLin(lh.cf, {k: p[k] for k in lh.cf.domain})
        f = self.fl.field
        if lh.nl_name == "exp":
            self.s = torch.exp(f)
            self.d = self.s
        elif lh.nl_name == "identity":
            self.s, self.d = f, None
        else:                                                  # pointwise callable: derivative from autograd of the sum
            with torch.enable_grad():
                fr = f.detach().requires_grad_(True)
                sr = lh.nl(fr)
                if sr.shape != fr.shape:
                    raise ValueError("the non-linearity must be a pointwise map of the correlated field (same shape in and out)")
                (self.d,) = torch.autograd.grad(sr.sum(), fr)
            self.s = sr.detach()
        self.sc_b = None
        if lh.scaling is not None:                             # signal = scaling(p) * exp(field), scaling log-normal of shape (1,)
            sc = lh.scaling(p[lh.scaling_key])
            self.s = sc * self.s
            self.d = self.s
            self.sc_b = lh.scaling.b                           # d scaling / d leaf = b * scaling
        self.M, self.Msqrt = lh._metric_ops(self.s)          # the likelihood metric and its square root at this signal, as operators

    def _t(self, tan):
        lh = self.lh
        return {k: torch.as_tensor(v, dtype=lh.dtype, device=lh.rt.device) for k, v in getattr(tan, "tree", tan).items()}

    def jvp(self, tan):
        t = self._t(tan)
        jt = self.fl.jvp(t)
        jt = jt if self.d is None else self.d * jt
        if self.sc_b is not None:
            jt = jt + self.s * (self.sc_b * t[self.lh.scaling_key].reshape(()))
        return jt

    def vjp(self, c):
        out = self.fl.vjp(c if self.d is None else self.d * c)
        if self.sc_b is not None:
            out[self.lh.scaling_key] = (self.sc_b * torch.sum(c * self.s)).reshape(1)
        return out

    def metric(self, tan):
        return self.vjp(self.M(self.jvp(tan)))

    def metric_flat(self, v):
        lay, lh = self.lh.layout, self.lh
        return lay.pack(self.metric(lay.unpack(v)), lh.dtype, lh.rt.device)

    def right_sqrt_metric(self, tan):
        return self.Msqrt(self.jvp(tan))

    def left_sqrt_metric(self, u):
        u = torch.as_tensor(u, dtype=self.lh.dtype, device=self.lh.rt.device)
        return self.vjp(self.Msqrt(u))


class OuterCorrelatedField(LazyModel):
    """The finalised model of a :class:`~nifty_b200.correlated_field.CorrelatedFieldMaker` with several ``add_fluctuations``."""

    def __init__(self, prefix, offset_mean, azm_prior, flucts, *, dtype, convention, runtime):
        self.prefix, self.offset_mean, self.dtype = prefix, float(offset_mean), dtype
        self._azm = _Prior(azm_prior)
        if len(flucts) < 2:
            raise ValueError("an outer product needs at least two sub-grids")
        from .bluestein import BluesteinHartley, grid_tables
        shape, dists, self._axes, self._subs = (), (), [], []
        for f in flucts:
            if len(f["shape"]) > 3:
                raise NotImplementedError("sub-grids with more than three axes are not supported")
            n0 = len(shape)
            shape += tuple(int(v) for v in f["shape"])
            d = f["distances"]
            dists += tuple(float(x) for x in (d if np.ndim(d) else (d,) * len(f["shape"])))
            self._axes.append(tuple(range(n0, len(shape))))
            # mode tables of the sub-grid (device plan for power-of-two extents, host NumPy otherwise)
            self._subs.append((f, grid_tables(f["shape"], f["distances"], dtype=dtype, convention=convention, runtime=runtime)))
        pow2 = lambda shp: all(n >= 2 and not (n & (n - 1)) for n in shp)      # noqa: E731
        self._per_sub = None
        if len(shape) > 3:
            # more axes than one device plan holds (e.g. the (3, 3) x (3, 3) case of the reference's product test): every sub-grid is
            # transformed on its own axes, slice by slice over the other axes -- functional, one device call per slice
            self._per_sub = []
            for f in flucts:
                shp = tuple(int(v) for v in f["shape"])
                if pow2(shp):
                    pl = Plan(shp, 1.0, dtype=dtype, hartley_convention=convention, runtime=runtime)
                    self._per_sub.append((pl.hartley, pl))
                else:
                    bh = BluesteinHartley(shp, dtype=dtype, convention=convention, runtime=runtime)
                    self._per_sub.append((bh, bh.plan))
            self.plan = self._per_sub[0][1]
        elif pow2(shape):
            self.plan = Plan(shape, dists, dtype=dtype, hartley_convention=convention, runtime=runtime)     # the JOINT transform
            self._joint = self.plan.hartley
        else:                     # e.g. the (3, 3) x (6,) grids of the reference's own product test (test_correlated_field.py:238-283)
            self._joint = BluesteinHartley(shape, dtype=dtype, convention=convention, runtime=runtime)
            self.plan = self._joint.plan
        self.rt = self.plan.rt
        dev = self.rt.device
        self.shape = shape
        domain = {prefix + "zeromode": ()}
        self._tabs, self._specs = [], []
        for f, tb in self._subs:
            spec, leaves = _amplitude_spec(f, prefix + f["prefix"], tb["mode_lengths"].size)
            domain.update(leaves)
            self._tabs.append(_device_tables(tb, dtype, dev))
            self._specs.append(spec)
        domain[prefix + "xi"] = shape
        self.domain = dict(sorted(domain.items()))
        self.layout = Layout(self.domain)
        self.target_shape = shape
        self._vol = float(np.prod([tb["V"] for tb in self._tabs]))
        self._coef = self._reflect_coefficients(len(flucts))
        self._factors = [(tb["pd"], axes) for tb, axes in zip(self._tabs, self._axes)]     # (bin table, axes) per sub-grid (_FieldLin)
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
    @staticmethod
    def _reflect_coefficients(m):
        """``prod_i cas(t_i) = sum_s c_s cas(sum_i s_i t_i)`` over the sign patterns ``s`` in {+1, -1}^m, from
        ``cas(x) cas(y) = (cas(x + y) + cas(x - y) + cas(-x + y) - cas(-x - y)) / 2`` applied m - 1 times."""
        coef = {(1,): 1.0}
        for _ in range(m - 1):
            nxt = {}
            for sg, c in coef.items():
                neg = tuple(-v for v in sg)
                for key, w in ((sg + (1,), 0.5), (sg + (-1,), 0.5), (neg + (1,), 0.5), (neg + (-1,), -0.5)):
                    nxt[key] = nxt.get(key, 0.0) + w * c
            coef = nxt
        return {k: v for k, v in coef.items() if v != 0.0}

    def _sep_hartley(self, x: torch.Tensor) -> torch.Tensor:
        """The composite transform ``H_m ... H_1`` from the joint transform and its reflections (module docstring)."""
        if self._per_sub is not None:
            return self._sliced_hartley(x.detach())
        J = self._joint(x.detach().contiguous())
        out = None
        for sg, c in self._coef.items():
            axes = sum((self._axes[i] for i, v in enumerate(sg) if v < 0), ())
            term = torch.roll(torch.flip(J, dims=axes), shifts=[1] * len(axes), dims=axes) if axes else J
            out = c * term if out is None else out + c * term
        return out

    def _sliced_hartley(self, x: torch.Tensor) -> torch.Tensor:
        """``H_m ... H_1`` for grids with more than three axes in total: sub-grid by sub-grid on its own axes."""
        nd = len(self.shape)
        for (fn, _), axes in zip(self._per_sub, self._axes):
            rest = [a for a in range(nd) if a not in axes]
            perm = rest + list(axes)
            xp = x.permute(perm).contiguous()
            sub = tuple(self.shape[a] for a in axes)
            flat = xp.reshape((-1,) + sub)
            out = torch.stack([fn(flat[b]) for b in range(flat.shape[0])])
            inv = [perm.index(a) for a in range(nd)]
            x = out.reshape(xp.shape).permute(inv)
        return x.contiguous()

    def _raw_transform(self, x):
        return self._sep_hartley(x)

    def _factor_tables(self, small):
        """O(K_i) tables whose expanded product is the amplitude of every mode: ``z na_1`` and ``na_2``."""
        z, nas = self._normalized(small)
        return (z * nas[0],) + tuple(nas[1:])

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
            a = _eval_amplitude(spec, tb, p)
            nas.append(torch.cat((a[:1], a[1:] / z)))                    # get_normalized_amplitudes (:807-821)
        return z, nas

    def amplitude(self, pos):
        """correlated_field.py:824-833: with more than one spectrum only the relative scales are determined."""
        raise NotImplementedError("If more than one spectrum is present in the model, no unique set of amplitudes exist because only the "
                                  "relative scale is determined.")

    power_spectrum = amplitude

    @property
    def normalized_amplitudes(self):
        return tuple((lambda pos, i=i: self._normalized(self._tree(pos))[1][i]) for i in range(len(self._specs)))

    @property
    def target_grids(self):
        from .bluestein import grid_record
        return tuple(grid_record(tb) for _, tb in self._subs)

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


class HostLin:
    """Counterpart of :class:`~nifty_b200._runtime.Lin` for host-composed models: the same flat-vector interface (``update``,
    ``energy``, ``metric``, ``metric_pair``, sqrt-metrics, ``transformation``, ``normalized_residual``) on the explicit
    linearisation :class:`_HostLin`, so that the sample draws, geoVI updates, the KL and ``optimize_kl`` of this package run
    unchanged on outer-product and non-power-of-two fields.  CG solves on such operators use the host loop."""

    host_composed = True

    class _ModelStub:
        def __init__(self, plan):
            self.plan = plan

    def __init__(self, lh):
        self.lh, self.rt = lh, lh.rt
        self.model = HostLin._ModelStub(lh.signal.cf.plan)
        self._l, self._energy = None, None

    def _pack(self, tree):
        return self.lh.layout.pack(tree, self.lh.dtype, self.rt.device)

    def update(self, pos, want_grad=False, add_prior=False):
        lh = self.lh
        pos = lh.signal.as_flat(pos)
        with torch.no_grad():
            self._l = _HostLin(lh, pos)
            s = self._l.s
            self._energy = float(lh._lh_energy(s))
            if not want_grad:
                return None
            dE = self._l.M(s - lh.data) if lh.kind == 0 else 1.0 - lh.data / s         # likelihood_impl.py:124-126, :238-240
            g = self._pack(self._l.vjp(dE))
        return g + pos if add_prior else g

    def energy(self) -> float:
        return self._energy

    def signal(self):
        return self._l.s

    def metric(self, t, add_identity=False, out=None):
        with torch.no_grad():
            r = self._l.metric_flat(t)
            if add_identity:
                r += t
        if out is not None:
            out.copy_(r)
            return out
        return r

    def rsm(self, t, scaled=True):
        with torch.no_grad():
            tt = self.lh.layout.unpack(t)
            return self._l.right_sqrt_metric(tt) if scaled else self._l.fl.jvp(tt)

    def lsm(self, u, scaled=True):
        with torch.no_grad():
            u = torch.as_tensor(u, dtype=self.lh.dtype, device=self.rt.device)
            return self._pack(self._l.left_sqrt_metric(u) if scaled else self._l.fl.vjp(u))

    def metric_pair(self, other: "HostLin", t, add_identity=False):
        """``lsm_self(rsm_other(t)) (+ t)``: the two halves of the geoVI operator (evi.py:167-172)."""
        r = self.lsm(other.rsm(t))
        return r + t if add_identity else r

    def transformation(self):
        lh, s = self.lh, self._l.s
        return self._l.Msqrt(s) if lh.kind == 0 else 2.0 * torch.sqrt(s)               # likelihood_impl.py:134-138, :248-251

    def normalized_residual(self):
        lh, s = self.lh, self._l.s
        return self._l.Msqrt(lh.data - s) if lh.kind == 0 else (lh.data - s) / torch.sqrt(s)


class OuterLikelihood(LikelihoodWithModel):
    """``Gaussian`` / ``Poissonian`` data on ``signal = nl(cf)`` for a host-composed field (:class:`OuterCorrelatedField`,
    ``bluestein.BluesteinCorrelatedField``): the interface of :class:`~nifty_b200.likelihood.LikelihoodWithModel`
    (likelihood.py:599-633 of the reference) with :class:`HostLin` linearisations, so ``draw_linear_residual``,
    ``nonlinearly_update_residual``, ``optimize_kl``, ``wiener_filter_posterior`` accept it.  The N-sized transforms run on the
    device; amplitude chains, pointwise likelihood and vector algebra are torch operations on the same device."""

    def __init__(self, likelihood, cf, nonlinearity="exp", scaling=None, scaling_key="scaling"):
        from .likelihood import SignalModel
        if nonlinearity not in ("exp", "identity") and not callable(nonlinearity):
            raise ValueError(f"unsupported nonlinearity {nonlinearity!r}")
        self.likelihood, self.signal = likelihood, SignalModel(cf, nonlinearity, scaling=scaling, scaling_key=scaling_key)
        self.cf, self.kind, self.rt, self.dtype = cf, likelihood.kind, cf.rt, cf.dtype
        self.layout, self.domain = self.signal.layout, self.signal.domain
        self.scaling = None if self.signal.scaling is None else _Prior(self.signal.scaling)
        self.scaling_key = scaling_key
        if tuple(np.shape(likelihood.data)) != tuple(cf.target_shape):
            raise ValueError(f"data shape {np.shape(likelihood.data)} does not match the model target {cf.target_shape}")
        self.nl = torch.exp if nonlinearity == "exp" else ((lambda f: f) if nonlinearity == "identity" else nonlinearity)
        self.nl_name = nonlinearity if isinstance(nonlinearity, str) else "callable"
        dev = cf.rt.device
        self.data = torch.as_tensor(np.asarray(likelihood.data), dtype=cf.dtype, device=dev) if not isinstance(likelihood.data, torch.Tensor) \
            else likelihood.data.to(dtype=cf.dtype, device=dev)
        self._cov_fn, self._std_fn = getattr(likelihood, "cov_inv_fn", None), getattr(likelihood, "std_inv_fn", None)
        if self.kind == 0:
            w = likelihood.w_array if likelihood.w_array is not None else likelihood.w_scalar
            self.w = torch.as_tensor(w, dtype=cf.dtype, device=dev)
        self._lins, self._max_lins = [], 3
        self._kl_cache = None

    def new_lin(self) -> HostLin:
        return HostLin(self)

    def lin_at(self, pos, want_grad=False, add_prior=False):
        """Same contract (and cache) as ``LikelihoodWithModel.lin_at``: ``(HostLin, gradient or None)``."""
        flat = self.signal.as_flat(pos)
        key = self._key(flat)
        if not want_grad:
            for k, lin in self._lins:
                if k == key:
                    return lin, None
        lin = HostLin(self)
        grad = lin.update(flat, want_grad=want_grad, add_prior=add_prior)
        lin._pos_ref = flat
        self._lins.append((key, lin))
        if len(self._lins) > self._max_lins:
            self._lins.pop(0)
        return lin, grad

    def _lh_energy(self, s):
        if self.kind == 0:                                   # Gaussian (likelihood_impl.py:124-126)
            r = s - self.data
            return 0.5 * torch.sum(r * self._metric_ops(s)[0](r))
        return torch.sum(s) - torch.sum(self.data * torch.log(s))      # Poissonian (:238-240)

    def _metric_ops(self, s):
        """``(M, M^(1/2))`` of the likelihood at signal ``s`` as callables on data-shaped arrays (:131-138, :245-251): Gaussian with
        diagonal weights, Gaussian with a non-diagonal (symmetric) operator pair, Poissonian."""
        if self.kind != 0:
            return (lambda x: x / s), (lambda x: x / torch.sqrt(s))
        if self._cov_fn is not None:
            def std(x):
                if self._std_fn is None:
                    raise NotImplementedError("this operation needs `noise_std_inv` (the square root of the non-diagonal inverse covariance)")
                return torch.as_tensor(self._std_fn(x), dtype=self.dtype, device=self.rt.device)
            return (lambda x: torch.as_tensor(self._cov_fn(x), dtype=self.dtype, device=self.rt.device)), std
        sw = torch.sqrt(self.w)
        return (lambda x: self.w * x), (lambda x: sw * x)

    def _tree(self, pos):
        """Latent position (flat vector in the order of ``self.layout``, dict or ``Vector``) as a dict of device tensors."""
        pos = getattr(pos, "tree", pos)
        if isinstance(pos, torch.Tensor):
            return self.layout.unpack(pos.to(dtype=self.dtype, device=self.rt.device))
        return {k: torch.as_tensor(v, dtype=self.dtype, device=self.rt.device) if not isinstance(v, torch.Tensor)
                else v.to(dtype=self.dtype, device=self.rt.device) for k, v in pos.items()}

    def cg_on_metric(self, pos, j, **cg_kwargs):
        """``cg(metric + 1, j)`` on flat vectors in the host loop (conjugate_gradient.py:77-214)."""
        from .conjugate_gradient import HamiltonianMetric, _cg
        lin, _ = self.lin_at(pos)
        return _cg(HamiltonianMetric(lin, likelihood=self), self.signal.as_flat(j), **cg_kwargs)

    def draw_linear_residual(self, pos, key, *, _white=None, **kwargs):
        """One MGVI residual sample at ``pos`` (evi.py:88-150) as a tree: ``evi.draw_linear_residual`` on this likelihood."""
        from .evi import draw_linear_residual
        r, info = draw_linear_residual(self, pos, key, _white=_white, **kwargs)
        return self.layout.unpack(r), info

    # -- sample-averaged KL and MGVI iterations (optimize_kl.py:67-144, 391-444, 540-591), on flat vectors ------------------
    def _flat(self, pos):
        pos = getattr(pos, "tree", pos)
        return pos.to(dtype=self.dtype, device=self.rt.device) if isinstance(pos, torch.Tensor) else self.layout.pack(pos, self.dtype, self.rt.device)

    def kl_value_and_grad(self, pos, residuals):
        """``_kl_vg``: mean over ``pos + residuals`` of value and gradient of the standard Hamiltonian ``lh(x) + <x, x> / 2``."""
        p = self._flat(pos)
        pts = [p] if residuals is None or len(residuals) == 0 else [p + self._flat(r) for r in residuals]
        lins = [HostLin(self) for _ in pts]
        val, grad = 0.0, torch.zeros_like(p)
        for lin, x in zip(lins, pts):
            grad += lin.update(x, want_grad=True, add_prior=True)
            val += lin.energy() + 0.5 * float(torch.dot(x, x))
        self._kl_cache = (p.clone(), residuals, lins)
        return val / len(pts), grad / len(pts)

    def kl_metric(self, pos, tangents, residuals):
        """``_kl_met``: mean over the sample points of ``metric(x, t) + t``."""
        p, t = self._flat(pos), self._flat(tangents)
        c = self._kl_cache                   # the CG of one Newton step applies the metric at one (position, samples) many times
        if c is None or c[1] is not residuals or not torch.equal(c[0], p):
            self.kl_value_and_grad(p, residuals)
            c = self._kl_cache
        out = torch.zeros_like(p)
        for lin in c[2]:
            out += lin.metric(t, add_identity=True)
        return out / len(c[2])

    def mgvi(self, pos, *, key, n_total_iterations: int, n_samples: int, draw_linear_kwargs=None, kl_kwargs=None, _whites=None):
        """MGVI (``optimize_kl(..., sample_mode="linear_resample")``, optimize_kl.py:672-729) with explicit white noise per draw
        (``_whites``: test hook; ``optimize_kl`` itself is the general driver): per iteration ``n_samples`` residual draws
        (mirrored: ``[s, -s]`` interleaved, evi.py:53-57) and one Newton-CG minimisation of the sample-averaged KL.  Returns
        ``(position tree, residuals [2 n_samples, L], list of OptimizeResults)``."""
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
                r, _ = self.draw_linear_residual(p, k, _white=w, **dkw)
                r = lay.pack(r, self.dtype, self.rt.device)
                rows += [r, -r]
            res = torch.stack(rows)
            opt = _newton_cg(None, x0=p, fun_and_grad=lambda x: self.kl_value_and_grad(x, res),
                             hessp=lambda x, t: self.kl_metric(x, t, res), **mk)
            p = opt.x
            states.append(opt._replace(x=None, jac=None))
        return lay.unpack(p), res, states
