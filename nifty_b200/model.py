"""``Model`` / ``LazyModel`` / ``WrappedCall`` / ``Initializer`` -- the container types of ``nifty/re/model.py:32-340`` on torch
trees, so that code written against the reference's model API (``cf.domain``, ``cf.init``, ``cf.target``, ``Model(call,
domain=, init=)``, ``WrappedCall(call, name=, shape=)``) keeps working.

What they are here: HOST-side containers.  A model can always be evaluated (``model(p)``) and initialised
(``model.init(key)``); it can sit under a likelihood (``Likelihood.amend``) when it has one of the structures the device
path implements -- a :class:`~nifty_b200.correlated_field.CorrelatedField`, optionally followed by a pointwise map
(``"exp"``, ``"identity"`` or any torch callable, see :meth:`Model.pointwise`) and an optional log-normal scaling leaf
(:class:`~nifty_b200.likelihood.SignalModel`).  There is no tracing compiler on this path, so an arbitrary ``call`` cannot be
differentiated; ``amend`` says so instead of guessing.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .tree_math import Vector, _leaves, _map


def _is_shape(x):
    return hasattr(x, "shape") or (isinstance(x, (tuple, list)) and all(isinstance(i, (int, np.integer)) for i in x))


def _map_domain(f, dom):
    """Map over a domain tree whose leaves are shapes (tuples of ints) or arrays."""
    if _is_shape(dom):
        return f(dom)
    if isinstance(dom, dict):
        return {k: _map_domain(f, v) for k, v in dom.items()}
    if isinstance(dom, (tuple, list)):
        return type(dom)(_map_domain(f, v) for v in dom)
    return f(dom)


def _shape_of(x):
    if hasattr(x, "shape"):
        return tuple(x.shape)
    return tuple(x) if isinstance(x, (tuple, list)) else ()


class Initializer:
    """Tree of leaf initialisers ``f(key) -> array`` (model.py:299-340): ``init(key)`` splits the key into one sub-key per
    leaf in sorted-key (pytree) order, exactly as ``random_like`` does (forest_math.py:60-72)."""

    def __init__(self, call_or_struct):
        self._tree = call_or_struct

    def __call__(self, key, *args, **kwargs):
        from .evi import random_split
        tree = self._tree
        if callable(tree):
            return tree(key, *args, **kwargs)
        leaves = _leaves(tree)
        keys = iter(random_split(key, len(leaves)))
        return _map(lambda f: f(next(keys), *args, **kwargs), tree)

    def __getitem__(self, k):
        return Initializer(self._tree[k])

    def __or__(self, other):
        a = self._tree if isinstance(self._tree, dict) else {}
        b = other._tree if isinstance(other, Initializer) else other
        return Initializer({**a, **b})


class LazyModel:
    """Base of all models (model.py:32-143): a callable on latent trees with ``domain`` / ``target`` / ``init``."""
    domain = None
    _target = None
    _init = None

    def __call__(self, *args, **kwargs):
        raise NotImplementedError

    @property
    def target(self):
        return self._target

    @property
    def init(self):
        return self._init


class Model(LazyModel):
    """``Model(call, *, domain=, target=, init=, white_init=False)`` (model.py:146-296).  ``domain``: tree of shapes (tuples
    or arrays); ``init``: an :class:`Initializer`, a tree of per-leaf callables, or omitted with ``white_init=True`` for
    standard-normal leaves of the domain's shapes."""

    def __init__(self, call: Optional[Callable] = None, *, domain=None, target=None, init=None, white_init: bool = False,
                 dtype=torch.float64, device="cpu"):
        self._call = call
        if domain is None and isinstance(init, dict):
            raise ValueError("`domain` is required (this path has no `eval_shape`: shapes cannot be traced from `init`)")
        self.domain = domain
        self._target = target
        self._dtype, self._device = dtype, device
        if init is None and white_init and domain is not None:
            from .evi import random_normal
            init = _map_domain(lambda s: (lambda key, s=s: random_normal(key, _shape_of(s), dtype, device)), domain)
        self._init = init if isinstance(init, Initializer) or init is None else Initializer(init)

    def __call__(self, *args, **kwargs):
        if self._call is None:
            raise NotImplementedError("this Model has no `call`")
        args = tuple(a.tree if isinstance(a, Vector) else a for a in args)
        return self._call(*args, **kwargs)

    @property
    def target(self):
        if self._target is None and self.domain is not None and self._call is not None:
            # the reference traces the output shape (model.py:171-181); here: one evaluation at zeros
            zeros = _map_domain(lambda s: torch.zeros(_shape_of(s), dtype=self._dtype, device=self._device), self.domain)
            self._target = _map(lambda x: _shape_of(x), self(zeros))
        return self._target

    @staticmethod
    def pointwise(cf, fn="exp", **kwargs):
        """The model ``x -> fn(cf(x))`` in the form the device path implements (``Model(lambda x: fn(cf(x)), domain=cf.domain,
        init=cf.init)`` of the reference): a :class:`~nifty_b200.likelihood.SignalModel`."""
        from .likelihood import SignalModel
        return SignalModel(cf, fn, **kwargs)


class WrappedCall(Model):
    """``WrappedCall(call, *, name=, shape=, dtype=, white_init=)`` (model.py:197-296): applies ``call`` to the leaf ``name``
    of the latent tree (to the whole input if ``name`` is None)."""

    def __init__(self, call: Callable, *, name=None, shape=(), dtype=torch.float64, white_init: bool = False, target=None, device="cpu"):
        leaf_call = call if name is None else (lambda p, **kw: call(p[name], **kw))
        domain = tuple(shape) if name is None else {name: tuple(shape)}
        super().__init__(leaf_call, domain=domain, target=target, white_init=white_init, dtype=dtype, device=device)
        self.name = name


class VModel(LazyModel):
    """``VModel(model, axis_size, in_axes=0, out_axes=0)`` (model.py:370-417): the model mapped over a leading sample axis of
    (some of) its input leaves.  ``in_axes``: an axis index (every leaf), a leaf name / names (those leaves only, e.g.
    ``VModel(cf, n, in_axes="cfxi")``: many excitation fields, one amplitude model -- test_empirical_power_spectrum.py:39) or a
    dict leaf -> 0 / None.  A :class:`~nifty_b200.correlated_field.CorrelatedField` with only its excitations mapped runs as ONE
    batched device call (``nb200_cf_apply_batch``, shared amplitude table); everything else is an in-order loop."""

    def __init__(self, model, axis_size: int, in_axes=0, out_axes=0):
        if not isinstance(model, LazyModel):
            raise ValueError(f"Model {model} of invalid type")
        if not isinstance(axis_size, int) or axis_size <= 0:
            raise ValueError(f"invalid axis size {axis_size}")
        if not isinstance(out_axes, int):
            raise ValueError(f"invalid `out_axes` {out_axes!r}")
        self.model, self.axis_size = model, axis_size
        dom = model.domain
        if isinstance(in_axes, int):
            axes = {k: in_axes for k in dom}
        else:
            if isinstance(in_axes, str):
                in_axes = (in_axes,)
            if not isinstance(in_axes, dict):
                in_axes = {k: (0 if k in in_axes else None) for k in dom}
            if set(in_axes) - set(dom):
                raise ValueError(f"Model domain structure {sorted(dom)} does not match axis structure {sorted(in_axes)}")
            axes = {k: in_axes.get(k) for k in dom}
        self.in_axes, self.out_axes = axes, out_axes

        def with_axis(shape, a):
            shape = tuple(shape)
            if a is None:
                return shape
            if not -len(shape) - 1 <= a <= len(shape):
                raise ValueError(f"sample axis {a} out of range for a leaf of shape {shape}")
            a = a % (len(shape) + 1)
            return shape[:a] + (axis_size,) + shape[a:]
        self.domain = {k: with_axis(_shape_of(s), axes[k]) for k, s in dom.items()}

    def init(self, key):
        """Mapped leaves: one draw per sample from split keys, stacked (model.py:394-409); the others as in the model."""
        from .evi import random_split
        keys = random_split(key, self.axis_size + 1)
        base = self.model.init(keys[0])
        draws = [self.model.init(k) for k in keys[1:]]
        return {k: (torch.stack([d[k] for d in draws], dim=self.in_axes[k]) if self.in_axes[k] is not None else base[k]) for k in base}

    @property
    def target(self):
        t = self.model.target
        if not _is_shape(t):
            return t
        t = tuple(t)
        o = self.out_axes % (len(t) + 1)
        return t[:o] + (self.axis_size,) + t[o:]

    def __call__(self, x):
        x = x.tree if isinstance(x, Vector) else x
        mapped = [k for k, a in self.in_axes.items() if a is not None]
        from .correlated_field import CorrelatedField
        if (isinstance(self.model, CorrelatedField) and mapped == [self.model.prefix + "xi"] and not self.model.plan.dist
                and self.in_axes[mapped[0]] == 0 and self.out_axes == 0):
            cf = self.model
            shared = {k: v for k, v in x.items() if k not in mapped}
            shared[mapped[0]] = torch.zeros(cf.domain[mapped[0]], dtype=cf.dtype, device=cf.rt.device)
            amp = cf.amplitude(shared)                 # azm * normalized amplitude, amp[0] = zeromode * V (correlated_field.py:889-912)
            return cf.plan.cf_apply_batch(amp, x[mapped[0]], cf.offset_mean)
        outs = []
        for i in range(self.axis_size):
            xi = {k: (v.select(self.in_axes[k], i) if self.in_axes.get(k) is not None else v) for k, v in x.items()}
            outs.append(self.model(xi))
        return torch.stack(outs, dim=self.out_axes)
