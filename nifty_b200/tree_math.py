"""Tree arithmetic helpers with the semantics of ``nifty/re/tree_math`` on this package's two latent representations:
dicts of tensors (the reference's pytrees, sorted-key leaf order) and flat device vectors (``Layout.pack``).

``vdot`` (vector_math.py:221-223), ``norm`` (:173-188: the ord-norm of the per-leaf ord-norms -- for a flat vector
that is the plain vector norm), ``size`` (:141-144), ``zeros_like`` (:160), ``where`` (:249-281), and the sequential
maps ``smap`` / ``lmap`` / ``get_map`` (custom_map.py:106-160, forest_math.py:136-156): on this path every sample is
a sequence of device launches, so all maps are the same in-order Python loop with ``in_axes`` semantics.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch

builtins_max, builtins_min, builtins_all, builtins_any, builtins_sum = max, min, all, any, sum    # the module defines tree versions of these names below


class Vector:
    """``jft.Vector`` (tree_math/vector.py:79-188): a latent tree with leaf-wise arithmetic.  ``Vector(tree)``, ``.tree``,
    ``+ - * /`` with vectors / scalars, unary ``-``, ``abs``, ``@`` (= :func:`vdot`), ``len`` (number of leaves), ``size``."""
    __slots__ = ("tree",)

    def __init__(self, tree):
        self.tree = tree.tree if isinstance(tree, Vector) else tree

    @staticmethod
    def _t(x):
        return x.tree if isinstance(x, Vector) else x

    def _bin(self, other, f):
        o = self._t(other)
        if isinstance(o, (dict, tuple, list)):
            return Vector(_map(f, self.tree, o))
        return Vector(_map(lambda a: f(a, o), self.tree))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return Vector(_map(lambda a: -a, self.tree))
    def __abs__(self): return Vector(_map(lambda a: abs(a), self.tree))
    def __matmul__(self, o): return vdot(self.tree, self._t(o))
    def __len__(self): return len(_leaves(self.tree))
    def __getitem__(self, k): return self.tree[k]
    def __iter__(self): return iter(self.tree)

    @property
    def size(self): return size(self.tree)

    def __repr__(self): return f"Vector({self.tree!r})"


def _leaves(tree):
    if isinstance(tree, Vector):
        return _leaves(tree.tree)
    if isinstance(tree, dict):
        return [l for k in sorted(tree) for l in _leaves(tree[k])]
    if isinstance(tree, (tuple, list)):
        return [l for t in tree for l in _leaves(t)]
    return [tree]


def _map(f, *trees):
    trees = tuple(t.tree if isinstance(t, Vector) else t for t in trees)
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: _map(f, *(t[k] if isinstance(t, dict) else t for t in trees)) for k in t0}
    if isinstance(t0, (tuple, list)):
        return type(t0)(_map(f, *(t[i] if isinstance(t, (tuple, list)) else t for t in trees)) for i in range(len(t0)))
    return f(*trees)


def ravel(tree):
    """``(flat, unravel)``: the leaves of a latent tree concatenated into one vector (sorted dict keys, as everywhere in this
    package) and the inverse map; what ``jax.flatten_util.ravel_pytree`` is to the reference's solvers, which work on pytrees
    directly.  ``unravel`` restores structure, shapes, dtypes and the ``Vector`` wrapper."""
    wrap = isinstance(tree, Vector)
    t = tree.tree if wrap else tree
    leaves = [torch.as_tensor(l) for l in _leaves(t)]
    dt = leaves[0].dtype
    for l in leaves[1:]:
        dt = torch.promote_types(dt, l.dtype)
    flat = torch.cat([l.reshape(-1).to(dt) for l in leaves]) if leaves else torch.zeros(0)
    shapes, dtypes = [tuple(l.shape) for l in leaves], [l.dtype for l in leaves]

    def unravel(v):
        it, off = [], 0
        for shp, d in zip(shapes, dtypes):
            n = int(np.prod(shp)) if shp else 1
            it.append(v[off:off + n].reshape(shp).to(d))
            off += n
        pieces = iter(it)

        def build(node):
            if isinstance(node, dict):
                return {k: build(node[k]) for k in sorted(node)}
            if isinstance(node, (tuple, list)):
                return type(node)(build(c) for c in node)
            return next(pieces)
        out = build(t)
        return Vector(out) if wrap else out
    return flat, unravel


def vdot(a, b) -> float:
    return float(builtins_sum(torch.dot(torch.as_tensor(x).reshape(-1), torch.as_tensor(y).reshape(-1)) for x, y in zip(_leaves(a), _leaves(b))))


def norm(tree, ord=2) -> float:
    def el(x):
        x = torch.as_tensor(x)
        return float(x.abs()) if x.ndim == 0 else float(torch.linalg.vector_norm(x.reshape(-1).to(torch.float64), ord=ord))
    return float(np.linalg.norm(np.array([el(x) for x in _leaves(tree)]), ord=ord))


def size(tree) -> int:
    return int(builtins_sum(int(np.prod(np.shape(x), dtype=np.int64)) if not isinstance(x, torch.Tensor) else x.numel() for x in _leaves(tree)))


def zeros_like(tree):
    return _map(lambda x: torch.zeros_like(torch.as_tensor(x)), tree)


def where(condition, x, y):
    """Leaf-wise ``torch.where``; a non-tree ``condition`` / ``x`` / ``y`` is broadcast over the leaves of the largest tree."""
    trees = [t for t in (condition, x, y) if isinstance(t, (dict, tuple, list))]
    if not trees:
        return torch.where(torch.as_tensor(condition), torch.as_tensor(x), torch.as_tensor(y))
    ref = builtins_max(trees, key=lambda t: len(_leaves(t)))
    return _map(lambda _, c, a, b: torch.where(torch.as_tensor(c), torch.as_tensor(a), torch.as_tensor(b)), ref, condition, x, y)


def stack(arrays, axis=0):
    """forest_math.py:115-116."""
    return _map(lambda *el: torch.stack([torch.as_tensor(e) for e in el], dim=axis), *arrays)


def unstack(stacked, axis=0):
    """forest_math.py:119-127: a tuple of trees, one per entry along ``axis``."""
    n = _leaves(stacked)[0].shape[axis]
    return tuple(_map(lambda t: torch.as_tensor(t).select(axis, i), stacked) for i in range(n))


def mean(forest):
    """forest_math.py:213-223: leaf-wise mean of a sequence of trees (e.g. ``tuple(signal(s) for s in samples)``)."""
    n = len(forest)
    return _map(lambda *ts: builtins_sum(torch.as_tensor(t) for t in ts) / n, *forest)


def mean_and_std(forest, correct_bias=True):
    """forest_math.py:226-242: ``sqrt(<x^2> - <x>^2)``, times ``sqrt(n / (n - 1))`` if ``correct_bias``."""
    n = len(forest)
    m = mean(forest)
    msq = mean(tuple(_map(lambda t: torch.as_tensor(t) ** 2, f) for f in forest))
    scl = float(np.sqrt(n / (n - 1))) if correct_bias else 1.0
    return m, _map(lambda a, b: scl * torch.sqrt(a - b ** 2), msq, m)


def _sequential_map(fun: Callable, in_axes=0, out_axes=0):
    """``vmap``-like call signature, evaluated one batch element after the other (custom_map.py:106-160)."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _leaves(a)[0].shape[ax]
                break
        if n is None:
            raise ValueError("at least one argument must be mapped")
        outs = [fun(*(a if ax is None else _map(lambda t: torch.as_tensor(t).select(ax, i), a) for a, ax in zip(args, axes))) for i in range(n)]
        return _map(lambda *ts: torch.stack([torch.as_tensor(t) for t in ts], dim=out_axes), *outs)
    return mapped


def smap(fun, in_axes=0, out_axes=0, *, unroll=1):
    return _sequential_map(fun, in_axes, out_axes)


def lmap(fun, in_axes=0, out_axes=0):
    return _sequential_map(fun, in_axes, out_axes)


def get_map(map) -> Callable:
    """forest_math.py:136-156; "vmap" / "pmap" have no separate meaning here and select the same sequential map."""
    if isinstance(map, str):
        if map in ("vmap", "v", "pmap", "p", "lmap", "l", "smap", "s"):
            return lmap if map[0] in "lvp" else smap
        raise ValueError(f"unknown `map` {map!r}")
    if callable(map):
        return map
    raise TypeError(f"invalid `map` {map!r}; expected string or callable")


# ---- further leaf-wise helpers of the reference's tree_math namespace (vector_math.py, forest_math.py) ----------------------
def ones_like(tree):
    return _map(lambda t: torch.ones_like(torch.as_tensor(t)), tree)


def shape(tree):
    """Tree of leaf shapes (vector_math.py ``shape``)."""
    return _map(lambda t: tuple(torch.as_tensor(t).shape), tree)


tree_shape = shape


def result_type(*trees):
    dt = None
    for t in trees:
        for leaf in _leaves(t):
            d = torch.as_tensor(leaf).dtype
            dt = d if dt is None else torch.promote_types(dt, d)
    return dt


def dot(a, b) -> float:
    """Sum over the leaves of the leaf-wise dot products (no conjugation; equals :func:`vdot` for real trees)."""
    return vdot(a, b)


matmul = dot


def sum(tree) -> float:        # noqa: A001  (name of the reference)
    total = 0.0
    for leaf in _leaves(tree):
        total += float(torch.as_tensor(leaf).sum())
    return total


def max(tree) -> float:        # noqa: A001
    return float(builtins_max(float(torch.as_tensor(leaf).max()) for leaf in _leaves(tree)))


def min(tree) -> float:        # noqa: A001
    return float(builtins_min(float(torch.as_tensor(leaf).min()) for leaf in _leaves(tree)))


def all(tree) -> bool:         # noqa: A001
    return builtins_all(bool(torch.as_tensor(leaf).all()) for leaf in _leaves(tree))


def any(tree) -> bool:         # noqa: A001
    return builtins_any(bool(torch.as_tensor(leaf).any()) for leaf in _leaves(tree))


def conj(tree):
    return _map(lambda t: torch.conj(torch.as_tensor(t)), tree)


conjugate = conj


def has_arithmetics(obj, additional_methods=()) -> bool:
    """forest_math.py:21-41: does the object support ``+ - * /`` and unary minus (``Vector``s and arrays do, bare dicts do not)."""
    desired = ("__add__", "__sub__", "__mul__", "__truediv__", "__neg__") + tuple(additional_methods)
    return builtins_all(hasattr(obj, m) for m in desired)


def assert_arithmetics(obj, *args, **kwargs):
    if not has_arithmetics(obj, *args, **kwargs):
        raise TypeError("input of invalid type: object must support arithmetic operations; wrap trees in `Vector`")


def map_forest(f: Callable, in_axes=0, out_axes=0, tree_transpose_output: bool = True, map="vmap", **kwargs) -> Callable:  # noqa: A002
    """forest_math.py:159-199: ``f`` mapped over a forest (a tuple / list of trees) given for exactly one argument."""
    if out_axes != 0:
        raise NotImplementedError("`out_axis` not yet supported")
    in_axes = in_axes if isinstance(in_axes, tuple) else (in_axes,)
    mapped = [idx for idx, el in enumerate(in_axes) if el is not None]
    if not mapped:
        raise ValueError("must map over at least one axis")
    if len(mapped) > 1:
        raise NotImplementedError("mapping over more than one axis is not yet supported")
    i = mapped[0]
    map_f = get_map(map)(f, in_axes=in_axes, out_axes=out_axes)

    def apply(*xs):
        if not isinstance(xs[i], (list, tuple)):
            raise TypeError(f"expected mapped axes to be a tuple; got {type(xs[i])}")
        out = map_f(*xs[:i], stack(xs[i]), *xs[i + 1:])
        return unstack(out) if tree_transpose_output else out
    return apply


def map_forest_mean(method, map="vmap", *args, **kwargs) -> Callable:        # noqa: A002
    """forest_math.py:202-210: the mean over the forest of ``method`` applied to every tree."""
    method_map = map_forest(method, *args, tree_transpose_output=False, map=map, **kwargs)

    def meaned_apply(*xs):
        return _map(lambda t: torch.as_tensor(t).mean(dim=0), method_map(*xs))
    return meaned_apply
