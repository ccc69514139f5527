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
