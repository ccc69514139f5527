"""Likelihoods amended with the correlated-field signal model.

Host-side mirror of ``nifty/re/likelihood.py`` (``Likelihood``, ``LikelihoodWithModel``:546-658) and
``nifty/re/likelihood_impl.py`` (``Gaussian``:83-138, ``Poissonian``:203-251) for the model family of
the hot path: ``signal = [scaling *] nl(correlated_field)``, ``nl`` in {exp, identity}
(demos/re/0_intro.py:39-57, misc/re/paper/minimal_benchmark.py:89).  Method names, argument order
and error behaviour follow the reference; the arithmetic runs in libniftyb200.so:

* ``metric(pos, t)`` is ONE fused sequence (JVP -> likelihood metric -> VJP with the middle axis pass
  shared) against a cached linearisation, instead of ``jax.linearize`` + ``linear_transpose`` on
  every call (likelihood.py:613-621).
* latent positions may be dicts (pytrees, as in the reference) or flat tensors in sorted-key order.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ._runtime import Lin, ModelHandle
from .correlated_field import CorrelatedField
from .prior import LogNormalPrior, _as_prior
from .tree import Layout


from .model import LazyModel  # noqa: E402

class SignalModel(LazyModel):
    """``signal(p) = scaling(p) * nl(cf(p))`` with an optional log-normal ``scaling`` leaf of shape (1,)."""

    def __init__(self, cf: CorrelatedField, nonlinearity="exp", scaling=None, scaling_key: str = "scaling"):
        # `nonlinearity`: "exp", "identity", or any torch-callable POINTWISE map of the field (`Model(lambda x: f(cf(x)))` of
        # the reference, nifty/re/model.py:146-181): f and f' are evaluated on the host side at every linearisation
        # (torch autograd) and handed to the library as tables (nb200_lin_set_pointwise); the products never see f.
        self.nl_fn = None
        if callable(nonlinearity):
            self.nl_fn, nonlinearity = nonlinearity, "tabulated"
            if getattr(cf.plan, "dist", False):
                raise NotImplementedError("custom non-linearities are not available on slab-decomposed fields")
        elif nonlinearity not in ("exp", "identity"):
            raise ValueError(f"unsupported nonlinearity {nonlinearity!r}")
        self.cf, self.nonlinearity = cf, nonlinearity
        self.scaling = _as_prior(scaling, LogNormalPrior, "scaling", optional=True)
        if self.scaling is not None and nonlinearity != "exp":
            raise ValueError("`scaling` requires the exp non-linearity")
        self.scaling_key = scaling_key
        domain = dict(cf.domain)
        if self.scaling is not None:
            domain[scaling_key] = (1,)
        self.domain = dict(sorted(domain.items()))
        self.layout = Layout(self.domain)
        self.rt, self.dtype, self.target_shape = cf.rt, cf.dtype, cf.target_shape

    @property
    def target(self):
        return self.target_shape

    def init(self, seed):
        return self.layout.unpack(self.layout.random(seed, self.dtype, self.rt.device))

    def as_flat(self, pos) -> torch.Tensor:
        pos = getattr(pos, "tree", pos)          # jft.Vector
        if isinstance(pos, torch.Tensor):
            if pos.numel() != self.layout.size:
                raise ValueError(f"latent vector has {pos.numel()} entries, expected {self.layout.size}")
            return self.rt.asarray(pos.reshape(-1), self.dtype)
        return self.layout.pack(pos, self.dtype, self.rt.device)

    def like(self, template, vec):
        if isinstance(template, torch.Tensor):
            return vec
        if hasattr(template, "tree") and not isinstance(template, dict):      # jft.Vector in -> Vector out
            from .tree_math import Vector
            return Vector(self.layout.unpack(vec))
        return self.layout.unpack(vec)

    def _new_handle(self) -> ModelHandle:
        return ModelHandle(self.cf.plan, self.cf._descriptor(self.layout, self.scaling, self.scaling_key))


class Likelihood:
    """Data-space likelihood before ``amend``; see :class:`Gaussian`, :class:`Poissonian`."""
    kind = -1

    def amend(self, signal, **kwargs):
        from .bluestein import BluesteinCorrelatedField
        from .outer import OuterCorrelatedField, OuterLikelihood
        composed = (OuterCorrelatedField, BluesteinCorrelatedField)      # host-composed models (outer.py, bluestein.py)
        if isinstance(signal, composed):
            return OuterLikelihood(self, signal, "identity")
        if isinstance(signal, SignalModel) and isinstance(signal.cf, composed):
            return OuterLikelihood(self, signal.cf, signal.nl_fn if signal.nl_fn is not None else signal.nonlinearity,
                                   scaling=signal.scaling, scaling_key=signal.scaling_key)
        if isinstance(signal, CorrelatedField):
            signal = SignalModel(signal, "identity")
        if isinstance(signal, SignalModel) and getattr(self, "cov_inv_fn", None) is not None:
            return OperatorLikelihood(self, signal)
        if not isinstance(signal, SignalModel):
            raise NotImplementedError(
                "the B200 path implements likelihoods on a correlated field followed by a pointwise map: amend with a "
                "`CorrelatedField`, a `SignalModel(cf, 'exp' | 'identity' | <torch callable>, scaling=...)` or "
                "`Model.pointwise(cf, fn)`; an arbitrary `Model(call, ...)` cannot be differentiated here (no tracing compiler)")
        return LikelihoodWithModel(self, signal)


class Gaussian(Likelihood):
    """``jft.Gaussian(data, noise_cov_inv=None, noise_std_inv=None)`` (likelihood_impl.py:35-138).

    ``noise_cov_inv`` may be a scalar, an array, or (as in the reference) a callable.  Diagonal callables ``x -> w * x`` are
    recognised by probing and run on the fused path; any other (symmetric) operator is applied on the host side between the
    device JVP and VJP of the model (:class:`OperatorLikelihood`); its ``noise_std_inv`` is needed for the sqrt-metrics,
    ``transformation`` and ``normalized_residual`` (sample draws), not for energy / gradient / metric.
    """
    kind = 0

    def __init__(self, data, noise_cov_inv=None, noise_std_inv=None):
        self.data = data
        shape = tuple(np.shape(data)) if not isinstance(data, torch.Tensor) else tuple(data.shape)
        if noise_cov_inv is None and noise_std_inv is None:
            w = 1.0
        elif noise_cov_inv is not None:
            w = noise_cov_inv
        else:
            w = noise_std_inv
        self.cov_inv_fn = self.std_inv_fn = None
        if callable(w):
            # a callable is probed: the diagonal is read off a vector of ones and checked on two random probes.  `x -> diag * x`
            # takes the fused path (weights inside the axis-0 pass); anything else is kept as an OPERATOR and applied between the
            # device JVP and VJP of the model (OperatorLikelihood below) -- never silently treated as diagonal
            fn = w
            w = fn(torch.ones(shape, dtype=torch.float64))
            gen = torch.Generator().manual_seed(12345)
            diagonal = True
            for _ in range(2):
                probe = torch.randn(shape, dtype=torch.float64, generator=gen)
                got = torch.as_tensor(fn(probe), dtype=torch.float64)
                want = torch.as_tensor(w, dtype=torch.float64) * probe
                if got.shape != want.shape or not torch.allclose(got, want, rtol=1e-10, atol=1e-12 * float(want.abs().max() + 1e-300)):
                    diagonal = False
            if not diagonal:
                if noise_cov_inv is None:
                    raise NotImplementedError("Gaussian: a non-diagonal `noise_std_inv` needs `noise_cov_inv` as well (the reference would "
                                              "assume a diagonal covariance here, likelihood_impl.py:46-55)")
                if noise_std_inv is not None and not callable(noise_std_inv):
                    raise NotImplementedError("Gaussian: a non-diagonal `noise_cov_inv` needs a callable `noise_std_inv` (or none)")
                self.cov_inv_fn, self.std_inv_fn = noise_cov_inv, noise_std_inv
                self.w_scalar, self.w_array = 1.0, None
                return
        if noise_cov_inv is None and noise_std_inv is not None:
            w = w * w
        if np.ndim(w) == 0 if not isinstance(w, torch.Tensor) else w.ndim == 0:
            self.w_scalar, self.w_array = float(w), None
        else:
            self.w_scalar, self.w_array = 1.0, w


class Poissonian(Likelihood):
    """``jft.Poissonian(data)``: integer, non-negative counts (likelihood_impl.py:226-233)."""
    kind = 1

    def __init__(self, data):
        arr = data.cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
        if not np.issubdtype(arr.dtype, np.integer):
            raise TypeError("`data` of invalid type")
        if np.any(arr < 0):
            raise ValueError("`data` must not be negative")
        self.data = arr
        self.w_scalar, self.w_array = 1.0, None


class LikelihoodWithModel:
    """``likelihood.amend(signal)``; energy / metric / sqrt-metrics of likelihood.py:599-633."""

    def __init__(self, likelihood: Likelihood, signal: SignalModel):
        self.likelihood, self.signal = likelihood, signal
        self.rt, self.dtype = signal.rt, signal.dtype
        self.layout, self.domain = signal.layout, signal.domain
        if tuple(np.shape(likelihood.data)) != tuple(signal.target_shape):
            raise ValueError(f"data shape {np.shape(likelihood.data)} does not match the model target {signal.target_shape}")
        self.handle = signal._new_handle()
        self.handle.nl_fn = signal.nl_fn
        self.handle.set_likelihood(likelihood.kind, {"identity": 0, "exp": 1, "tabulated": 2}[signal.nonlinearity], likelihood.data,
                                   likelihood.w_scalar, likelihood.w_array)
        self._lins = []     # small cache of linearisations: [(key, Lin)]
        self._max_lins = 3

    def __add__(self, other):
        """``lh_a + lh_b`` (likelihood.py:383-384): the sum of amended likelihoods on the union of their latent domains."""
        return LikelihoodSum(self, other)

    def init(self, seed):
        """White latent position of the amended model (``LikelihoodWithModel`` forwards ``init`` of its model, likelihood.py:546-598)."""
        return self.signal.init(seed)

    @property
    def left_sqrt_metric_tangents_shape(self):
        """Shape of the data-space tangents of ``left_sqrt_metric`` (likelihood.py:352-363)."""
        return tuple(self.signal.target_shape)

    lsm_tangents_shape = left_sqrt_metric_tangents_shape

    @property
    def right_sqrt_metric_tangents_shape(self):
        """The latent domain (likelihood.py:365-376)."""
        return self.domain

    rsm_tangents_shape = right_sqrt_metric_tangents_shape

    def freeze(self, *, primals, point_estimates):
        """``(likelihood with the named leaves inserted at their value in primals, remaining liquid primals)``
        (likelihood.py:386-393); no point estimates: ``(self, primals)``."""
        if not point_estimates:
            return self, primals
        lp = LikelihoodPartial(self, primals=primals, point_estimates=point_estimates)
        return lp, lp.splitx(primals)[0]

    # -- linearisation cache --------------------------------------------------------------------
    @staticmethod
    def _key(flat: torch.Tensor):
        return (flat.data_ptr(), flat._version, flat.numel())

    def lin_at(self, pos, want_grad=False, add_prior=False):
        """Linearisation at ``pos`` (cached while the tensor is neither replaced nor modified in place)."""
        flat = self.signal.as_flat(pos)
        key = self._key(flat)
        if not want_grad:
            for k, lin in self._lins:
                if k == key:
                    return lin, None
        lin = None
        for i, (k, l) in enumerate(self._lins):
            if k == key:
                lin = l
                self._lins.pop(i)
                break
        if lin is None:
            lin = self._lins.pop(0)[1] if len(self._lins) >= self._max_lins else Lin(self.handle)
        grad = lin.update(flat, want_grad=want_grad, add_prior=add_prior)
        lin._pos_ref = flat          # keep the buffer alive so the (ptr, version) key stays unique
        self._lins.append((key, lin))
        return lin, grad

    def new_lin(self) -> Lin:
        return Lin(self.handle)

    # -- point estimates / constants (likelihood.py:399-499, _parse_point_estimates :57-92) ----------------------
    def frozen_ranges(self, point_estimates):
        """Flat-vector ranges ``[(lo, hi), ...]`` of the leaves named in ``point_estimates`` (a collection of keys of
        the latent domain, or a dict ``{key: bool}``); merged and sorted.  The reference freezes those leaves at their
        current value (``Likelihood.freeze``) and runs every solve in the space of the remaining ones; here the same
        solves run on full-length vectors whose frozen entries are kept at zero / at the expansion point."""
        if not point_estimates:
            return []
        if isinstance(point_estimates, dict):
            keys = [k for k, v in point_estimates.items() if v]
        elif isinstance(point_estimates, str):
            keys = [point_estimates]
        else:
            keys = list(point_estimates)
        unknown = [k for k in keys if k not in self.layout.offsets]
        if unknown:
            raise ValueError(f"point_estimates {unknown!r} are not leaves of the latent domain {self.layout.keys!r}")
        rs = sorted((self.layout.offsets[k], self.layout.offsets[k] + self.layout.numel(k)) for k in set(keys))
        merged = []
        for lo, hi in rs:
            if merged and merged[-1][1] == lo:
                merged[-1] = (merged[-1][0], hi)
            else:
                merged.append((lo, hi))
        return merged

    @staticmethod
    def clear_frozen(v: torch.Tensor, ranges):
        for lo, hi in ranges:
            v[..., lo:hi] = 0
        return v

    # -- vector algebra on latent vectors (tree_math.vdot / norm); slab-decomposed: xi block summed over ranks ------
    @property
    def _plan(self):
        return self.signal.cf.plan

    def _xi_slice(self):
        key = self.signal.cf.prefix + "xi"
        o = self.layout.offsets[key]
        return o, o + self.layout.numel(key)

    def vdot(self, a: torch.Tensor, b: torch.Tensor) -> float:
        if not self._plan.dist:
            return float(torch.dot(a, b))
        lo, hi = self._xi_slice()
        d_xi = torch.dot(a[lo:hi], b[lo:hi]).to(torch.float64).reshape(1)
        # replicated hyper-parameter leaves: counted once, from the replicated entries only so that every
        # rank gets bit-identical scalars (and the replicated leaves stay bit-identical through CG)
        d_rep = float(torch.dot(a[:lo], b[:lo])) + float(torch.dot(a[hi:], b[hi:]))
        return float(self._plan.comm.allreduce_sum(d_xi)) + d_rep

    def vnorm(self, v: torch.Tensor, ord=2) -> float:
        if not self._plan.dist:
            from .conjugate_gradient import _norm
            return _norm(v, ord)
        lo, hi = self._xi_slice()
        rep = torch.cat((v[:lo], v[hi:]))
        if ord == 2:
            return float(np.sqrt(self.vdot(v, v)))
        if ord == 1:
            t = v[lo:hi].abs().sum().to(torch.float64).reshape(1)
            return float(self._plan.comm.allreduce_sum(t)) + float(rep.abs().sum())
        if ord in (np.inf, float("inf")):
            t = v[lo:hi].abs().max().to(torch.float64).reshape(1)
            self._plan.comm.dist.all_reduce(t, op=self._plan.comm.dist.ReduceOp.MAX, group=self._plan.comm.group)
            return max(float(t), float(rep.abs().max()) if rep.numel() else 0.0)
        raise ValueError(f"unsupported norm order {ord!r} on slab-decomposed vectors")

    def global_size(self, frozen=()) -> int:
        """Number of latent degrees of freedom of the whole model (slab-decomposed: hyper-parameters + the GLOBAL grid),
        minus the entries of the ``frozen`` ranges -- what ``xtol * size`` of the reference's Newton-CG refers to;
        identical on every rank."""
        lo, hi = self._xi_slice()
        n = self.layout.size
        for a, b in frozen:
            n -= b - a
        if self._plan.dist:
            n += int(self._plan.N) - (hi - lo)
            if any(a <= lo and hi <= b for a, b in frozen):      # the excitations are frozen: none of the global grid counts
                n -= int(self._plan.N) - (hi - lo)
        return n

    def zero_padding(self, v: torch.Tensor) -> torch.Tensor:
        """Zero the padding rows of the xi block of a slab-decomposed latent vector (no-op otherwise)."""
        if self._plan.dist:
            lo, hi = self._xi_slice()
            blk = v[lo:hi].view(self._plan.local_shape)
            blk[torch.as_tensor(self._plan.row_map < 0, device=v.device)] = 0
        return v

    # -- reference API ------------------------------------------------------------------------------
    def energy(self, pos) -> float:
        lin, _ = self.lin_at(pos)
        return lin.energy()

    __call__ = energy

    def energy_and_gradient(self, pos, add_prior=False):
        flat = self.signal.as_flat(pos)
        lin, grad = self.lin_at(flat, want_grad=True, add_prior=add_prior)
        e = lin.energy()
        if add_prior:
            e += 0.5 * self.vdot(flat, flat)
        return e, self.signal.like(pos, grad)

    def metric(self, pos, tangents):
        lin, _ = self.lin_at(pos)
        return self.signal.like(tangents, lin.metric(self.signal.as_flat(tangents)))

    def left_sqrt_metric(self, pos, tangents):
        lin, _ = self.lin_at(pos)
        return self.signal.like(pos, lin.lsm(tangents, scaled=True))

    def right_sqrt_metric(self, pos, tangents):
        lin, _ = self.lin_at(pos)
        return lin.rsm(self.signal.as_flat(tangents), scaled=True)

    def transformation(self, pos):
        lin, _ = self.lin_at(pos)
        return lin.transformation()

    def normalized_residual(self, pos):
        lin, _ = self.lin_at(pos)
        return lin.normalized_residual()

    def signal_response(self, pos):
        lin, _ = self.lin_at(pos)
        return lin.signal()


class OpLin:
    """Linearisation of ``Gaussian(non-diagonal N^-1)`` on ``signal = nl(cf)``: the device :class:`Lin` supplies the signal and
    the field-level JVP / VJP (``nb200_rsm`` / ``nb200_lsm`` with ``scaled = 0``), the covariance operators are applied between
    them on the host side (likelihood_impl.py:124-138 through likelihood.py:599-633).  Same flat-vector interface as ``Lin``; CG
    solves on such operators run the host loop."""

    host_composed = True

    def __init__(self, lh: "OperatorLikelihood"):
        self.lh, self.rt = lh, lh.rt
        self.dev = Lin(lh.handle)
        self.model = self.dev.model
        self.s = self._nr = self._energy = None

    def _js(self, t):
        jf = self.dev.rsm(t, scaled=False)
        return self.s * jf if self.lh.exp else jf

    def _jst(self, c):
        return self.dev.lsm((self.s * c if self.lh.exp else c).contiguous(), scaled=False)

    def update(self, pos, want_grad=False, add_prior=False):
        pos = self.lh.signal.as_flat(pos)
        self.dev.update(pos)
        self.s = self.dev.signal()
        r = self.s - self.lh.data_t
        self._nr = self.lh.cov_inv(r)
        self._energy = 0.5 * float(torch.sum(r * self._nr))
        if not want_grad:
            return None
        g = self._jst(self._nr)
        return g + pos if add_prior else g

    def energy(self) -> float:
        return self._energy

    def signal(self):
        return self.s

    def metric(self, t, add_identity=False, out=None):
        r = self._jst(self.lh.cov_inv(self._js(t)))
        if add_identity:
            r = r + t
        if out is not None:
            out.copy_(r)
            return out
        return r

    def rsm(self, t, scaled=True):
        return self.lh.std_inv(self._js(t)) if scaled else self.dev.rsm(t, scaled=False)

    def lsm(self, u, scaled=True):
        u = self.rt.asarray(u, self.lh.dtype)
        return self._jst(self.lh.std_inv(u)) if scaled else self.dev.lsm(u, scaled=False)

    def metric_pair(self, other: "OpLin", t, add_identity=False):
        r = self.lsm(other.rsm(t))
        return r + t if add_identity else r

    def transformation(self):
        return self.lh.std_inv(self.s)

    def normalized_residual(self):
        return self.lh.std_inv(self.lh.data_t - self.s)


class OperatorLikelihood(LikelihoodWithModel):
    """``Gaussian(data, noise_cov_inv=<operator>, noise_std_inv=<operator>).amend(signal)`` for covariances that are not diagonal."""

    def __init__(self, likelihood: Gaussian, signal: SignalModel):
        if signal.nonlinearity not in ("exp", "identity") or signal.scaling is not None:
            raise NotImplementedError("non-diagonal noise covariances: `signal` must be exp(cf) or cf (no scaling leaf, no custom map)")
        if getattr(signal.cf.plan, "dist", False):
            raise NotImplementedError("non-diagonal noise covariances are not available on slab-decomposed fields")
        super().__init__(likelihood, signal)             # device model with unit weights: supplies the signal and the field JVP / VJP
        self.exp = signal.nonlinearity == "exp"
        self.data_t = self.rt.asarray(likelihood.data, self.dtype).reshape(signal.target_shape)
        self._cov_fn, self._std_fn = likelihood.cov_inv_fn, likelihood.std_inv_fn

    def cov_inv(self, x):
        return torch.as_tensor(self._cov_fn(x), dtype=self.dtype, device=self.rt.device)

    def std_inv(self, x):
        if self._std_fn is None:
            raise NotImplementedError("this operation needs `noise_std_inv` (the square root of the non-diagonal inverse covariance)")
        return torch.as_tensor(self._std_fn(x), dtype=self.dtype, device=self.rt.device)

    def new_lin(self) -> OpLin:
        return OpLin(self)

    def lin_at(self, pos, want_grad=False, add_prior=False):
        flat = self.signal.as_flat(pos)
        key = self._key(flat)
        if not want_grad:
            for k, lin in self._lins:
                if k == key:
                    return lin, None
        lin = OpLin(self)
        grad = lin.update(flat, want_grad=want_grad, add_prior=add_prior)
        lin._pos_ref = flat
        self._lins.append((key, lin))
        if len(self._lins) > self._max_lins:
            self._lins.pop(0)
        return lin, grad


class _SumSignal:
    """What the drivers ask of ``likelihood.signal`` for a sum of likelihoods: the joint latent layout and a 1-D data space (the
    summands' data vectors one after the other)."""

    class _CfStub:
        def __init__(self, plan, prefix):
            self.plan, self.prefix = plan, prefix

    def __init__(self, layout: Layout, n_data: int, rt, dtype, plan, prefix):
        self.layout, self.domain = layout, dict(layout.shapes)
        self.target_shape, self.rt, self.dtype = (int(n_data),), rt, dtype
        self.cf = _SumSignal._CfStub(plan, prefix)
        self.scaling, self.nl_fn, self.nonlinearity = None, None, "sum"

    target = property(lambda self: self.target_shape)

    def as_flat(self, pos) -> torch.Tensor:
        pos = getattr(pos, "tree", pos)
        if isinstance(pos, torch.Tensor):
            if pos.numel() != self.layout.size:
                raise ValueError(f"latent vector has {pos.numel()} entries, expected {self.layout.size}")
            return self.rt.asarray(pos.reshape(-1), self.dtype)
        return self.layout.pack(pos, self.dtype, self.rt.device)

    def init(self, seed):
        return self.layout.unpack(self.layout.random(seed, self.dtype, self.rt.device))

    def like(self, template, vec):
        if isinstance(template, torch.Tensor):
            return vec
        if hasattr(template, "tree") and not isinstance(template, dict):
            from .tree_math import Vector
            return Vector(self.layout.unpack(vec))
        return self.layout.unpack(vec)


class SumLin:
    """Linearisation of a sum of amended likelihoods: one child linearisation per summand, latent vectors gathered from / scattered
    into the joint layout, data-space vectors concatenated.  Same interface as ``Lin``; CG solves run the host loop."""

    host_composed = True

    def __init__(self, lh: "LikelihoodSum"):
        self.lh, self.rt = lh, lh.rt
        self.children = [c.new_lin() for c in lh.likelihood_summands]
        self.model = _SumSignal._CfStub(lh.signal.cf.plan, "")
        self.model.plan = lh.signal.cf.plan

    def _gather(self, i, v):
        return v[self.lh._index[i]]

    def _scatter_add(self, out, i, v):
        out.index_add_(0, self.lh._index[i], v.to(out.dtype))
        return out

    def _split(self, u):
        u = self.rt.asarray(u, self.lh.dtype).reshape(-1)
        return [u[a:b].reshape(shp) for (a, b), shp in zip(self.lh._data_ranges, self.lh._data_shapes)]

    def update(self, pos, want_grad=False, add_prior=False):
        pos = self.lh.signal.as_flat(pos)
        grad = torch.zeros_like(pos) if want_grad else None
        for i, c in enumerate(self.children):
            g = c.update(self._gather(i, pos).contiguous(), want_grad=want_grad, add_prior=False)
            if want_grad:
                self._scatter_add(grad, i, g)
        if want_grad and add_prior:
            grad += pos
        return grad

    def energy(self) -> float:
        return float(sum(c.energy() for c in self.children))

    def signal(self):
        return torch.cat([c.signal().reshape(-1) for c in self.children])

    def metric(self, t, add_identity=False, out=None):
        r = t.clone() if add_identity else torch.zeros_like(t)
        for i, c in enumerate(self.children):
            self._scatter_add(r, i, c.metric(self._gather(i, t).contiguous()))
        if out is not None:
            out.copy_(r)
            return out
        return r

    def rsm(self, t, scaled=True):
        return torch.cat([c.rsm(self._gather(i, t).contiguous(), scaled=scaled).reshape(-1) for i, c in enumerate(self.children)])

    def lsm(self, u, scaled=True):
        r = torch.zeros(self.lh.layout.size, dtype=self.lh.dtype, device=self.rt.device)
        for i, (c, ui) in enumerate(zip(self.children, self._split(u))):
            self._scatter_add(r, i, c.lsm(ui.contiguous(), scaled=scaled))
        return r

    def metric_pair(self, other: "SumLin", t, add_identity=False):
        r = self.lsm(other.rsm(t))
        return r + t if add_identity else r

    def transformation(self):
        return torch.cat([c.transformation().reshape(-1) for c in self.children])

    def normalized_residual(self):
        return torch.cat([c.normalized_residual().reshape(-1) for c in self.children])


class LikelihoodSum(LikelihoodWithModel):
    """``jft.LikelihoodSum`` (likelihood.py:661-757) for amended likelihoods of this package: energies and metrics add on the union of
    the latent domains (shared leaves -- e.g. two data sets seen through ONE correlated field -- are shared), the data spaces are kept
    side by side (``right_sqrt_metric`` / ``transformation`` / ``normalized_residual`` return ``{"lh_0": ..., "lh_1": ...}``,
    ``left_sqrt_metric`` takes such a dict).  Every summand keeps its own fused device products."""

    def __init__(self, *likelihood_summands, _key_template="lh_{index}"):
        flat = []
        for i, lh in enumerate(likelihood_summands):
            if not isinstance(lh, LikelihoodWithModel):
                raise TypeError(f"object at position {i} which to add to this instance is of invalid type {type(lh)!r}")
            flat += list(lh.likelihood_summands) if isinstance(lh, LikelihoodSum) else [lh]
        self.likelihood_summands = tuple(flat)
        self._key_template = _key_template
        first = flat[0]
        self.rt, self.dtype = first.rt, first.dtype
        domain = {}
        for lh in flat:
            if lh.rt is not first.rt or lh.dtype != first.dtype:
                raise ValueError("summands must share runtime and dtype")
            if getattr(lh.signal.cf.plan, "dist", False):
                raise NotImplementedError("sums of likelihoods on slab-decomposed fields are not supported")
            for k, shp in lh.layout.shapes.items():
                if k in domain and tuple(domain[k]) != tuple(shp):
                    raise ValueError(f"leaf {k!r} has different shapes in the summands: {domain[k]} vs {shp}")
                domain[k] = tuple(shp)
        self.layout = Layout(domain)
        self.domain = dict(self.layout.shapes)
        dev = self.rt.device
        self._index = []                                   # joint flat index of every entry of a summand's flat vector
        for lh in flat:
            idx = torch.cat([torch.arange(self.layout.offsets[k], self.layout.offsets[k] + self.layout.numel(k)) for k in lh.layout.keys])
            self._index.append(idx.to(dev))
        self._data_shapes = [tuple(lh.signal.target_shape) for lh in flat]
        sizes = [int(np.prod(s)) for s in self._data_shapes]
        ends = np.cumsum(sizes)
        self._data_ranges = [(int(e - n), int(e)) for e, n in zip(ends, sizes)]
        self.signal = _SumSignal(self.layout, int(ends[-1]), self.rt, self.dtype, first.signal.cf.plan, first.signal.cf.prefix)
        self.likelihood = None
        self._lins, self._max_lins = [], 3

    def _items(self):
        for i, lh in enumerate(self.likelihood_summands):
            yield self._key_template.format(index=i, likelihood=lh), lh

    def __add__(self, other):
        return LikelihoodSum(*self.likelihood_summands, other)

    def new_lin(self) -> SumLin:
        return SumLin(self)

    def lin_at(self, pos, want_grad=False, add_prior=False):
        flat = self.signal.as_flat(pos)
        key = self._key(flat)
        if not want_grad:
            for k, lin in self._lins:
                if k == key:
                    return lin, None
        lin = SumLin(self)
        grad = lin.update(flat, want_grad=want_grad, add_prior=add_prior)
        lin._pos_ref = flat
        self._lins.append((key, lin))
        if len(self._lins) > self._max_lins:
            self._lins.pop(0)
        return lin, grad

    # data-space quantities as dicts keyed like the reference's (:706-754)
    def _as_dict(self, vec):
        return {key: vec[a:b].reshape(shp) for (key, _), (a, b), shp in zip(self._items(), self._data_ranges, self._data_shapes)}

    def right_sqrt_metric(self, pos, tangents):
        return self._as_dict(super().right_sqrt_metric(pos, tangents))

    def transformation(self, pos):
        return self._as_dict(super().transformation(pos))

    def normalized_residual(self, pos):
        return self._as_dict(super().normalized_residual(pos))

    def left_sqrt_metric(self, pos, tangents):
        if isinstance(tangents, dict):
            tangents = torch.cat([self.rt.asarray(tangents[key], self.dtype).reshape(-1) for key, _ in self._items()])
        return super().left_sqrt_metric(pos, tangents)



class LikelihoodPartial:
    """``jft.LikelihoodPartial`` (likelihood.py:399-499, ``_parse_point_estimates`` :57-92, ``partial_insert_and_remove`` :119-177):
    the likelihood as a function of the LIQUID leaves only, the frozen ones inserted at their value in ``primals``.  Positions and
    tangents are dicts of the liquid leaves; every operator inserts (frozen values for positions, zeros for tangents), applies the
    operator of the full likelihood -- the fused device products -- and removes the frozen leaves from latent-space results.  The
    solvers of this package reach the same restriction through frozen ranges of full-length vectors instead (``frozen_ranges``)."""

    def __init__(self, likelihood: LikelihoodWithModel, *, primals, point_estimates):
        self.likelihood, self.point_estimates = likelihood, point_estimates
        ranges = likelihood.frozen_ranges(point_estimates)           # validates the keys
        lay = likelihood.layout
        self.frozen_keys = [k for k in lay.keys if any(lo <= lay.offsets[k] < hi for lo, hi in ranges)]
        self.liquid_keys = [k for k in lay.keys if k not in self.frozen_keys]
        full = likelihood.layout.unpack(likelihood.signal.as_flat(primals))
        self.primals_frozen = {k: full[k].clone() for k in self.frozen_keys}
        self.domain = {k: lay.shapes[k] for k in self.liquid_keys}
        self.rt, self.dtype = likelihood.rt, likelihood.dtype

    def splitx(self, primals):
        """``(liquid part, frozen part)`` of a full position."""
        full = self.likelihood.layout.unpack(self.likelihood.signal.as_flat(primals))
        return {k: full[k] for k in self.liquid_keys}, {k: full[k] for k in self.frozen_keys}

    def _insert(self, liquid, zeros=False):
        liquid = getattr(liquid, "tree", liquid)
        if set(liquid) != set(self.liquid_keys):
            raise ValueError(f"expected the liquid leaves {self.liquid_keys!r}, got {sorted(liquid)!r}")
        tree = {k: torch.as_tensor(v, dtype=self.dtype, device=self.rt.device) if not isinstance(v, torch.Tensor) else v for k, v in liquid.items()}
        for k, v in self.primals_frozen.items():
            tree[k] = torch.zeros_like(v) if zeros else v
        return tree

    def _remove(self, tree):
        tree = getattr(tree, "tree", tree)
        return {k: tree[k] for k in self.liquid_keys}

    def energy(self, primals) -> float:
        return self.likelihood.energy(self._insert(primals))

    __call__ = energy

    def energy_and_gradient(self, primals):
        e, g = self.likelihood.energy_and_gradient(self._insert(primals))
        return e, self._remove(g)

    def metric(self, primals, tangents):
        return self._remove(self.likelihood.metric(self._insert(primals), self._insert(tangents, zeros=True)))

    def left_sqrt_metric(self, primals, tangents):
        return self._remove(self.likelihood.left_sqrt_metric(self._insert(primals), tangents))

    def right_sqrt_metric(self, primals, tangents):
        return self.likelihood.right_sqrt_metric(self._insert(primals), self._insert(tangents, zeros=True))

    def transformation(self, primals):
        return self.likelihood.transformation(self._insert(primals))

    def normalized_residual(self, primals):
        return self.likelihood.normalized_residual(self._insert(primals))
