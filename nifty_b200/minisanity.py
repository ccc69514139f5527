"""Per-iteration diagnostics with the interface of ``nifty/re/minisanity.py``.

``reduced_residual_stats`` (:30-82) and ``minisanity`` (:110-129): mean, reduced chi-squared and number of
degrees of freedom of every leaf, for one position or averaged (mean, std) over the samples.  The two
moments of a leaf are one fused reduction on the device (``nb200_vec_stats``); Python only forms the
sample statistics and the report, whose text layout follows the reference so that ``minisanity.txt``
reads the same.
"""

from __future__ import annotations

from typing import Any, NamedTuple, Optional

import numpy as np
import torch

from .evi import Samples
from .tree import Layout


class ChiSqStats(NamedTuple):
    mean: Any
    reduced_chisq: Any
    ndof: Any


def _leaf_stats(plan, x: torch.Tensor, n_global=None):
    """minisanity.py:17-21 for one real leaf: (mean, reduced chi^2, ndof).  ``n_global``: the leaf is the local part of a
    slab-decomposed array (padding entries are zero): the two moments are all-reduced and divided by the global count."""
    n = int(x.numel())
    s1, s2 = plan.vec_stats(x.contiguous())
    if n_global is not None:
        t = torch.tensor([s1, s2], dtype=torch.float64, device=x.device)
        s1, s2 = (float(v) for v in plan.comm.allreduce_sum(t))
        n = int(n_global)
    return s1 / n, s2 / n, n


def _as_tree(v, layout: Optional[Layout]):
    if isinstance(v, dict):
        return v
    if layout is not None and isinstance(v, torch.Tensor) and v.dim() == 1 and v.numel() == layout.size:
        return layout.unpack(v)
    return v


def reduced_residual_stats(position_or_samples, func=None, *, plan, layout: Optional[Layout] = None, map="lmap",
                           dist_leaves=()):
    """Tree of :class:`ChiSqStats` (one per leaf): ``mean`` and ``reduced_chisq`` are ``[sample mean, sample std]``.

    ``position_or_samples``: a flat latent vector, a dict of leaves or :class:`Samples`; ``func`` (optional) is
    applied to every position first (e.g. ``likelihood.normalized_residual``).  ``layout`` splits flat latent
    vectors into their leaves; ``plan`` supplies the device reduction.  On a slab-decomposed plan ``dist_leaves`` names
    the latent leaves of which every rank holds a part (the excitations); ``func`` outputs are position-space parts."""
    if isinstance(position_or_samples, Samples) and len(position_or_samples) > 0:
        points = [position_or_samples[i] for i in range(len(position_or_samples))]
    else:
        p = position_or_samples.pos if isinstance(position_or_samples, Samples) else position_or_samples
        points = [p]
    per_sample = []
    for p in points:
        v = func(p) if func is not None else p
        tree = _as_tree(v, layout if func is None else None)
        ng = int(plan.N) if getattr(plan, "dist", False) else None
        if isinstance(tree, dict):
            per_sample.append({k: _leaf_stats(plan, t, ng if (func is not None or k in dist_leaves) else None) for k, t in tree.items()})
        else:
            per_sample.append(_leaf_stats(plan, tree, ng if func is not None else None))

    def combine(rows):
        m = np.array([r[0] for r in rows])
        rx = np.array([r[1] for r in rows])
        return ChiSqStats(np.array([m.mean(), m.std()]), np.array([rx.mean(), rx.std()]), rows[0][2])

    if isinstance(per_sample[0], dict):
        return {k: combine([s[k] for s in per_sample]) for k in per_sample[0]}
    return combine(per_sample)


def _pretty(x: ChiSqStats) -> str:
    rsq = x.reduced_chisq
    return f"reduced Chi²:{rsq[0]:8.2}±{rsq[1]:8.2}, avg:{x.mean[0]:+9.2}±{x.mean[1]:8.2}, #dof:{int(x.ndof):7d}"


def minisanity(position_or_samples, func=None, *, plan, layout: Optional[Layout] = None, map="lmap", dist_leaves=()):
    """``(stat_tree, report)`` (minisanity.py:110-129); one ``key:: reduced Chi² …`` line per leaf."""
    stats = reduced_residual_stats(position_or_samples, func, plan=plan, layout=layout, map=map, dist_leaves=dist_leaves)
    if isinstance(stats, dict):
        msg = "".join(f"{k:24s}:: {_pretty(v)}\n" for k, v in stats.items())
    else:
        msg = _pretty(stats) + "\n"
    return stats, msg
