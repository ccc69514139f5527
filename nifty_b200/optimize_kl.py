"""``OptimizeVI`` / ``optimize_kl`` with the call protocol of ``nifty/re/optimize_kl.py``.

* ``_kl_vg`` (:90-114) / ``_kl_met`` (:117-144): sample-averaged value+gradient and metric of the
  standard Hamiltonian (:67-87).  One linearisation per sample point is cached for the whole
  Newton step, so every KL-CG iteration is ``2*n_samples`` fused metric products.
* sample sharding (:317-320, :397-476): instead of a JAX device mesh, the independent sample solves
  are distributed over the ranks of a ``torch.distributed`` process group (one process per GPU, NCCL
  over NVLink on a B200 box, gloo in the CPU tests).  ``pos`` is replicated, rank r draws the keys
  ``r, r+W, r+2W, ...``; the KL value / gradient / metric-vector products are summed with one
  all-reduce each (the only collective of the path) and divided by the global count.
* checkpoint / resume: ``last.pkl`` with ``(Samples, OptimizeVIState)`` after every iteration
  (:871-875, :835-839), written by rank 0 with every rank's residuals gathered.
"""

from __future__ import annotations

import inspect
import logging
import os
import pickle
from functools import partial
from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from .conjugate_gradient import HamiltonianMetric, _cg
from .evi import Samples, as_key, concatenate_zip, draw_linear_residual, nonlinearly_update_residual, random_split
from .likelihood import LikelihoodWithModel
from .optimize import OptimizeResults, _newton_cg

SAMPLE_MODES = ("linear_sample", "linear_resample", "nonlinear_sample", "nonlinear_resample", "nonlinear_update")


class OptimizeVIState(NamedTuple):
    nit: int
    key: object
    sample_state: object = None
    minimization_state: object = None
    config: dict = {}


def _getitem_at_nit(config, key, nit):
    """Most kwargs may be callables of the iteration index (optimize_kl.py:166-170): only one-argument callables are
    evaluated, exactly as in the reference."""
    c = config[key]
    if callable(c) and len(inspect.getfullargspec(c).args) == 1:
        return c(nit)
    return c


class _Comm:
    """Thin view of a torch.distributed process group (None = single process)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        if group is None or not dist.is_available() or not dist.is_initialized():
            self.group, self.rank, self.world = None, 0, 1
        else:
            self.group = None if group is True else group
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_sum_async(self, t: torch.Tensor):
        """Start the in-place SUM all-reduce of `t` and return its work handle (None on a single rank)."""
        if self.world > 1:
            return self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        return None

    def tree_sum(self, local: dict, n_global: int, owner) -> torch.Tensor:
        """Sum of one vector per GLOBAL sample index in a fixed binary tree over the indices (level s: index g, a multiple of
        2 s, absorbs g + s as left + right) -- the result does not depend on how many ranks hold the samples, bit for bit (the
        pattern of the reference's MPI reduction, nifty/cl/utilities.py:349-415).  `local`: {global index: vector} of this
        rank, `owner(g)`: rank holding index g.  Pairs on different ranks exchange point to point, every rank walks the same
        (level, index) sequence, so the blocking sends / receives cannot deadlock.  Returns the total on every rank."""
        vals = dict(local)
        s = 1
        while s < n_global:
            for g in range(0, n_global, 2 * s):
                h = g + s
                if h >= n_global:
                    continue
                og, oh = owner(g), owner(h)
                if og == self.rank and oh == self.rank:
                    vals[g] = vals[g] + vals.pop(h)
                elif oh == self.rank:
                    self.dist.send(vals.pop(h).contiguous(), dst=self._global_rank(og), group=self.group)
                elif og == self.rank:
                    buf = torch.empty_like(vals[g])
                    self.dist.recv(buf, src=self._global_rank(oh), group=self.group)
                    vals[g] = vals[g] + buf
            s *= 2
        root = owner(0)
        if self.world == 1:
            return vals[0]
        any_vec = next(iter(local.values())) if local else None
        shape_src = vals[0] if root == self.rank else any_vec
        total = vals[0] if root == self.rank else torch.empty_like(shape_src)
        self.dist.broadcast(total, src=self._global_rank(root), group=self.group)
        return total

    def _global_rank(self, r: int) -> int:
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def allgather(self, t: torch.Tensor):
        """Per-rank tensors whose leading extent may differ between ranks (uneven sample sharding): the counts are
        exchanged first, every rank pads to the maximum, the result is trimmed again."""
        if self.world == 1:
            return [t]
        cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        self.dist.all_gather(cnts, cnt, group=self.group)
        cnts = [int(c) for c in cnts]
        m = max(cnts)
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad.contiguous(), group=self.group)
        return [o[:c] for o, c in zip(out, cnts)]


def get_status_message(samples, state, residual=None, *, name="", plan=None, layout=None, map="lmap", dist_leaves=()) -> str:
    """Per-iteration report (optimize_kl.py:40-61): energy, sampling status, KL steps and the minisanity tables of
    the likelihood residuals and of the latent parameters over the samples."""
    from .minisanity import minisanity
    energy = state.minimization_state.fun
    msg_smpl = ""
    ss = state.sample_state
    if isinstance(ss, OptimizeResults):
        ss = [ss]
    if isinstance(ss, (list, tuple)) and len(ss) > 0 and all(isinstance(el, OptimizeResults) for el in ss):
        nlsi = tuple(int(n) for el in ss for n in np.atleast_1d(el.nit))       # one result per sample on this path
        msg_smpl = f"\n{name}: #(Nonlinear sampling steps) {nlsi}"
    elif isinstance(ss, (np.ndarray, list, tuple)) and len(ss) > 0:
        nlsi = tuple(int(el) for el in ss)
        msg_smpl = f"\n{name}: Linear sampling status {nlsi}"
    mini_res = ""
    if residual is not None:
        _, mini_res = minisanity(samples, residual, plan=plan, map=map)
    _, mini_pr = minisanity(samples, plan=plan, layout=layout, map=map, dist_leaves=dist_leaves)
    return (f"{name}: Iteration {state.nit:04d} E:{energy:+2.4e}"
            f"{msg_smpl}"
            f"\n{name}: #(KL minimization steps) {state.minimization_state.nit}"
            f"\n{name}: Likelihood residual(s):\n{mini_res}"
            f"\n{name}: Prior residual(s):\n{mini_pr}"
            f"\n")


class OptimizeVI:
    """State-less MGVI / geoVI driver (optimize_kl.py:173-741)."""

    def __init__(self, likelihood: LikelihoodWithModel, n_total_iterations: int, *, comm=None,
                 jit=True, linear_minimizer_jit=True, nonlinear_minimizer_jit=False, kl_map=None, residual_map="lmap",
                 kl_reduce=None, mirror_samples=True, devices=None,
                 _kl_value_and_grad: Optional[Callable] = None, _kl_metric: Optional[Callable] = None,
                 _draw_linear_residual: Callable = draw_linear_residual,
                 _nonlinearly_update_residual: Callable = nonlinearly_update_residual,
                 _get_status_message: Optional[Callable] = None):
        # jit / *_map / kl_reduce select how JAX traces and batches the per-sample work (optimize_kl.py:228-246); here every
        # sample point is a sequence of device launches and the reduction over samples is the mean, so they are accepted
        # for call compatibility only.  The two options that would change results are refused.
        # kl_reduce="fixed_tree": the sums over the sample points of the KL (value, gradient, metric) are taken in a fixed binary
        # tree over the GLOBAL sample indices, so the results do not depend on the number of ranks, bit for bit (slower: one
        # vector per sample travels point to point, the KL-CG runs in the host loop).  Default: all-reduce of per-rank sums.
        if kl_reduce not in (None, "fixed_tree") and not callable(kl_reduce):
            raise ValueError("kl_reduce must be None, 'fixed_tree' or (ignored, for call compatibility) a callable")
        self.fixed_tree = kl_reduce == "fixed_tree"
        if not mirror_samples:
            raise NotImplementedError("mirror_samples=False is not supported on the B200 path")
        if devices is not None:
            raise NotImplementedError("`devices` is a JAX mesh; pass a torch.distributed group as `comm` (one process per GPU)")
        self.likelihood = likelihood
        if _get_status_message is None:        # optimize_kl.py:376-389
            plan = likelihood.signal.cf.plan
            dl = (likelihood.signal.cf.prefix + "xi",) if plan.dist else ()     # slab-decomposed: moments all-reduced
            _get_status_message = partial(get_status_message, residual=likelihood.normalized_residual, plan=plan,
                                          layout=likelihood.layout, dist_leaves=dl)
        self.get_status_message = _get_status_message
        self.n_total_iterations = n_total_iterations
        self.comm = _Comm(comm)
        plan = likelihood.signal.cf.plan
        if self.comm.world > 1 and not plan.dist:
            # KL metric over several ranks: the last pass hands its result to the all-reduce in pieces (NVLink transfer of
            # one piece beside the computation of the next)
            plan.set_reduce_chunks(int(os.environ.get("NB200_REDUCE_CHUNKS", "1")))
        self._draw_linear_residual = _draw_linear_residual
        self._nonlinearly_update_residual = _nonlinearly_update_residual
        self._kl_vg_override, self._kl_met_override = _kl_value_and_grad, _kl_metric
        self._lins = []      # one linearisation per local sample point, reused across KL evaluations

    # -- sample bookkeeping -----------------------------------------------------------------------
    def _mirror_split(self, n_keys) -> bool:
        """The reference's special case `n_samples == mesh.size / 2` (optimize_kl.py:403-405, 437-439): every key is drawn by
        TWO devices and the odd one keeps the mirrored sample, so that the 2 n_samples points of the KL land one per device."""
        return self.comm.world > 1 and 2 * n_keys == self.comm.world

    def _local_keys(self, keys):
        if keys is not None and self._mirror_split(len(keys)):
            return keys[self.comm.rank // 2:self.comm.rank // 2 + 1]
        return keys[self.comm.rank::self.comm.world]

    def _n_global(self, n_local_points):
        t = torch.tensor([float(n_local_points)], dtype=torch.float64, device=self.likelihood.rt.device)
        return int(round(float(self.comm.allreduce_sum(t))))

    # -- sampling ------------------------------------------------------------------------------------
    def draw_linear_samples(self, primals, keys, **kwargs):
        """optimize_kl.py:391-444: local keys only; mirrored pairs interleaved [s0, -s0, s1, -s1, ...]."""
        res, infos = [], []
        for k in self._local_keys(keys):
            r, info = self._draw_linear_residual(self.likelihood, primals, k, **kwargs)
            res.append(r)
            infos.append(info)
        L = self.likelihood.layout.size
        smpls = torch.stack(res) if res else torch.empty((0, L), dtype=self.likelihood.dtype, device=self.likelihood.rt.device)
        if self._mirror_split(len(keys)):
            smpls = smpls if self.comm.rank % 2 == 0 else -smpls          # `_special_mirror_samples` (:421-423)
        else:
            smpls = concatenate_zip(smpls, -smpls)
        return Samples(pos=primals, samples=smpls, keys=keys), infos

    def nonlinearly_update_samples(self, samples: Samples, **kwargs):
        """optimize_kl.py:446-476: the same key for both signs, metric_sample_sign = +1 / -1."""
        local = self._local_keys(samples.keys)
        out, states = [], []
        if self._mirror_split(len(samples.keys)):      # one point per rank: the key's sample on even ranks, its mirror on odd ones
            assert len(samples) == 1 and len(local) == 1
            sign = 1.0 if self.comm.rank % 2 == 0 else -1.0
            r, st = self._nonlinearly_update_residual(self.likelihood, samples.pos, samples.residuals[0], local[0], sign, **kwargs)
            return Samples(pos=samples.pos, samples=torch.stack([r]), keys=samples.keys), [st]
        assert len(samples) == 2 * len(local)
        for i, k in enumerate(local):
            for s, sign in enumerate((1.0, -1.0)):
                r, st = self._nonlinearly_update_residual(self.likelihood, samples.pos, samples.residuals[2 * i + s], k, sign, **kwargs)
                out.append(r)
                states.append(st)
        smpls = torch.stack(out) if out else samples.residuals
        return Samples(pos=samples.pos, samples=smpls, keys=samples.keys), states

    def draw_samples(self, samples: Samples, *, key, sample_mode: str, n_samples: int, point_estimates=(),
                     draw_linear_kwargs=None, nonlinearly_update_kwargs=None, **kwargs):
        """State machine of optimize_kl.py:478-538."""
        if sample_mode not in SAMPLE_MODES:
            raise ValueError(f"`sample_mode` must be one of {SAMPLE_MODES}; got {sample_mode!r}")
        draw_linear_kwargs = dict(draw_linear_kwargs or {})
        nonlinearly_update_kwargs = dict(nonlinearly_update_kwargs or {})
        st_smpls = None
        n_keys = 0 if samples.keys is None else len(samples.keys)
        if n_samples != n_keys and sample_mode.lower() == "nonlinear_update":
            sample_mode = "nonlinear_resample"       # :490-497
        elif n_samples != n_keys and sample_mode.lower().endswith("_sample"):
            sample_mode = sample_mode.replace("_sample", "_resample")
        if n_samples == 0:
            return samples, 0          # "Do nothing for MAP" (:530-531): the samples object passes through unchanged
        if sample_mode.lower() in ("linear_resample", "nonlinear_resample"):
            k_smpls = random_split(key, n_samples)       # :507
            samples, st_smpls = self.draw_linear_samples(samples.pos, k_smpls, point_estimates=point_estimates, **draw_linear_kwargs)
            if sample_mode.lower() == "nonlinear_resample":
                samples, st_smpls = self.nonlinearly_update_samples(samples, point_estimates=point_estimates, **nonlinearly_update_kwargs)
        elif sample_mode.lower() in ("linear_sample", "nonlinear_sample"):
            samples, st_smpls = self.draw_linear_samples(samples.pos, samples.keys, point_estimates=point_estimates, **draw_linear_kwargs)
            if sample_mode.lower() == "nonlinear_sample":
                samples, st_smpls = self.nonlinearly_update_samples(samples, point_estimates=point_estimates, **nonlinearly_update_kwargs)
        else:   # nonlinear_update
            samples, st_smpls = self.nonlinearly_update_samples(samples, point_estimates=point_estimates, **nonlinearly_update_kwargs)
        return samples, st_smpls

    # -- KL ----------------------------------------------------------------------------------------------
    def _ensure_lins(self, n):
        while len(self._lins) < n:
            self._lins.append(self.likelihood.new_lin())

    def kl_value_and_grad(self, pos: torch.Tensor, residuals: Optional[torch.Tensor]):
        """``_kl_vg``: mean over all sample points of value_and_grad of the Hamiltonian; also
        (re)linearises the cached per-sample linearisations used by :meth:`kl_metric`."""
        lh = self.likelihood
        pts = [pos] if residuals is None or (len(residuals) == 0 and self.comm.world == 1) else [pos + r for r in residuals]
        if residuals is not None and len(residuals) == 0 and self.comm.world > 1:
            pts = []
        self._ensure_lins(len(pts))
        if self.fixed_tree:
            return self._kl_value_and_grad_tree(pos, pts)
        acc = torch.zeros(lh.layout.size + 2, dtype=torch.float64, device=lh.rt.device)
        for lin, x in zip(self._lins, pts):
            g = lin.update(x, want_grad=True, add_prior=True)
            acc[2:] += g.to(torch.float64)
            acc[0] += lin.energy() + 0.5 * lh.vdot(x, x)
            acc[1] += 1.0
        self._n_active = len(pts)
        for lin in self._lins[:len(pts)]:
            lin._ever_updated = True
        if len(pts) == 0 and not getattr(self._lins[0] if self._lins else None, "_ever_updated", False):
            self._ensure_lins(1)                        # a rank without sample points keeps one linearisation (scale 0) for the device operator
            self._lins[0].update(pos, want_grad=False)
            self._lins[0]._ever_updated = True
        acc = self.comm.allreduce_sum(acc)
        n = float(acc[1])
        self._n_total = int(round(n))
        return float(acc[0]) / n, (acc[2:] / n).to(lh.dtype)

    # -- fixed-tree reductions (kl_reduce="fixed_tree") ------------------------------------------------
    def _global_indices(self, n_local: int):
        """Global sample index of every local sample point and the owner map: rank r holds the mirrored pairs of the keys
        r, r + W, ... (one point per rank when every rank holds exactly one: the `n_samples == W / 2` layout or a MAP point)."""
        W, r = self.comm.world, self.comm.rank
        cnt = torch.tensor([n_local], dtype=torch.int64, device=self.likelihood.rt.device)
        cnts = [int(c[0]) for c in self.comm.allgather(cnt)]
        n = sum(cnts)
        if all(c == 1 for c in cnts):
            return [r], n, (lambda g: g)
        return [2 * (r + (i // 2) * W) + (i % 2) for i in range(n_local)], n, (lambda g: (g // 2) % W)

    def _kl_value_and_grad_tree(self, pos, pts):
        lh = self.likelihood
        local = {}
        idx, n, owner = self._global_indices(len(pts))
        for lin, x, g in zip(self._lins, pts, idx):
            grad = lin.update(x, want_grad=True, add_prior=True)
            e = lin.energy() + 0.5 * lh.vdot(x, x)
            local[g] = torch.cat([torch.tensor([e], dtype=torch.float64, device=lh.rt.device), grad.to(torch.float64)])
            lin._ever_updated = True
        self._n_active, self._n_total = len(pts), n
        self._tree_idx, self._tree_owner = idx, owner
        tot = self.comm.tree_sum(local, n, owner)
        return float(tot[0]) / n, (tot[1:] / n).to(lh.dtype)

    def _kl_metric_tree(self, tangents, frozen=None):
        local = {g: lin.metric(tangents, add_identity=True).to(torch.float64) for lin, g in zip(self._lins[:self._n_active], self._tree_idx)}
        out = (self.comm.tree_sum(local, self._n_total, self._tree_owner) / self._n_total).to(self.likelihood.dtype)
        for lo, hi in frozen or ():
            out[lo:hi] = 0
        return out

    def _kl_operator(self, frozen=None):
        """The sample-averaged metric at the points of the last :meth:`kl_value_and_grad` as a device operator."""
        if self.fixed_tree:
            return lambda t: self._kl_metric_tree(t, frozen)
        from .conjugate_gradient import SampleAveragedMetric
        lins = self._lins[:self._n_active]
        none_here = len(lins) == 0                     # more ranks than sample points: contribute zero
        if none_here:
            self._ensure_lins(1)
            lins = self._lins[:1]
            if not getattr(lins[0], "_ever_updated", False):
                raise RuntimeError("kl_metric: a rank without sample points needs one initialised linearisation")
        reduce_fn = (lambda t: self.comm.allreduce_sum_async(t)) if self.comm.world > 1 else None
        return SampleAveragedMetric(lins, self._n_total, identity_here=self.comm.rank == 0, reduce_fn=reduce_fn, frozen=frozen,
                                    scale_zero=none_here)

    def kl_metric(self, tangents: torch.Tensor) -> torch.Tensor:
        """``_kl_met`` at the points of the last :meth:`kl_value_and_grad`: mean of metric(x_i, t) + t, accumulated on the
        device (``nb200_metric_multi``); the all-reduce over the ranks is enqueued in-stream."""
        if self.fixed_tree:
            return self._kl_metric_tree(tangents)
        if self.likelihood.signal.cf.plan.dist:         # slab-decomposed fields: phase-wise products, host accumulation
            lh = self.likelihood
            acc = torch.zeros(lh.layout.size + 1, dtype=torch.float64, device=lh.rt.device)
            for lin in self._lins[:self._n_active]:
                acc[1:] += lin.metric(tangents, add_identity=True).to(torch.float64)
                acc[0] += 1.0
            acc = self.comm.allreduce_sum(acc)
            return (acc[1:] / float(acc[0])).to(lh.dtype)
        return self._kl_operator()(tangents)

    def kl_minimize(self, samples: Samples, minimize: Callable = _newton_cg, minimize_kwargs=None, constants=(), **kwargs) -> OptimizeResults:
        """optimize_kl.py:540-591.  ``constants``: leaves held at their current value during the minimisation (:553-573):
        the value / gradient / metric are evaluated at full positions, the gradient and the metric output are restricted
        to the other leaves (cleared on the constant ones), so the Newton steps never move them."""
        res = samples.residuals
        state = {"x": None}
        frozen = self.likelihood.frozen_ranges(constants)
        clr = (lambda v: self.likelihood.clear_frozen(v, frozen)) if frozen else (lambda v: v)

        def fg(x):
            state["x"] = x
            val, grad = self.kl_value_and_grad(x, res)
            return val, clr(grad)

        def hessp(x, t):
            if state["x"] is None or state["x"].data_ptr() != x.data_ptr():
                fg(x)
            return clr(self.kl_metric(t))

        def hessp_at(x):
            # device operator for the inner CG of Newton-CG (optimize.py:335): linearisations of the last fg(x)
            if state["x"] is None or state["x"].data_ptr() != x.data_ptr():
                fg(x)
            return self._kl_operator(frozen=frozen)

        kw = dict(minimize_kwargs or {})
        if frozen:
            kw.setdefault("_size", self.likelihood.global_size(frozen))
        if self.likelihood.signal.cf.plan.dist:      # slab-decomposed latent vectors: all-reduced reductions
            kw.setdefault("vdot", self.likelihood.vdot)
            kw.setdefault("vnorm", self.likelihood.vnorm)
            kw.setdefault("_size", self.likelihood.global_size())
        if not self.likelihood.signal.cf.plan.dist and minimize is _newton_cg:
            kw.setdefault("hessp_at", hessp_at)          # KL-CG on the device (nb200_cg_solve_multi)
        return minimize(None, x0=samples.pos, fun_and_grad=fg, hessp=hessp, **kw)

    # -- driver -------------------------------------------------------------------------------------------
    def init_state(self, key, *, n_samples, draw_linear_kwargs=None, nonlinearly_update_kwargs=None, kl_kwargs=None,
                   sample_mode="nonlinear_resample", point_estimates=(), constants=()) -> OptimizeVIState:
        self.likelihood.frozen_ranges(point_estimates)      # validates the keys (and the unsupported slab case) early
        self.likelihood.frozen_ranges(constants)
        config = dict(n_samples=n_samples, sample_mode=sample_mode, point_estimates=point_estimates, constants=constants,
                      draw_linear_kwargs=draw_linear_kwargs or dict(cg_name="SL", cg_kwargs=dict()),
                      nonlinearly_update_kwargs=nonlinearly_update_kwargs or dict(minimize_kwargs=dict(name="SN", cg_kwargs=dict(name="SNCG"))),
                      kl_kwargs=kl_kwargs or dict(minimize_kwargs=dict(name="M", cg_kwargs=dict(name="MCG"))))
        return OptimizeVIState(0, as_key(key), None, None, config)

    def update(self, samples: Samples, state: OptimizeVIState, **kwargs):
        """One VI iteration (optimize_kl.py:672-729)."""
        nit = state.nit                                # schedules see the PRE-increment index (:691-701): 0 on the first iteration
        cfg = state.config
        key, sk = random_split(state.key, 2)           # :703, ticks every iteration
        kw = {k: _getitem_at_nit(cfg, k, nit) for k in ("n_samples", "sample_mode", "point_estimates", "draw_linear_kwargs",
                                                      "nonlinearly_update_kwargs", "kl_kwargs")}
        kl_kwargs = dict(kw.pop("kl_kwargs"))
        constants = _getitem_at_nit(cfg, "constants", nit)
        samples, st_smpls = self.draw_samples(samples, key=sk, **kw)
        kl_opt = self.kl_minimize(samples, constants=constants, **kl_kwargs)
        samples = samples.at(kl_opt.x)
        state = state._replace(nit=nit + 1, key=key, sample_state=st_smpls, minimization_state=kl_opt._replace(x=None, jac=None))
        return samples, state

    def run(self, samples: Samples, *, key, **kwargs):
        state = self.init_state(key, **kwargs)
        for _ in range(self.n_total_iterations):
            samples, state = self.update(samples, state)
        return samples, state


def _samples_to_checkpoint(samples: Samples, likelihood, comm) -> Samples:
    """``Samples`` with NumPy pytrees (the layout of the reference's pickled object, evi.py:385-396: ``pos`` a dict of
    leaves, ``samples`` the residual leaves with a leading sample axis, ``keys``) holding the residuals of ALL ranks in
    global sample order (rank r draws the keys r, r + W, ...; two mirrored residuals per key)."""
    lay = likelihood.layout
    pos = {k: v.detach().cpu().numpy() for k, v in lay.unpack(samples.pos).items()}
    res = None
    if samples.residuals is not None:
        parts = comm.allgather(samples.residuals)
        n_keys = 0 if samples.keys is None else len(samples.keys)
        per_key = 2 if n_keys and sum(p.shape[0] for p in parts) == 2 * n_keys else 1
        total = sum(p.shape[0] for p in parts)
        full = np.empty((total, lay.size), dtype=parts[0].cpu().numpy().dtype)
        mirror_split = comm.world > 1 and 2 * n_keys == comm.world and all(p.shape[0] == 1 for p in parts)
        for r, p in enumerate(parts):
            pn = p.detach().cpu().numpy()
            if mirror_split:                      # rank r holds global sample r (key r // 2, sign r % 2)
                full[r] = pn[0]
                continue
            for i in range(pn.shape[0] // per_key):
                g = (r + i * comm.world) * per_key
                full[g:g + per_key] = pn[i * per_key:(i + 1) * per_key]
        res = {k: np.stack([lay.unpack_numpy(full[i])[k] for i in range(total)]) for k in pos}
    return Samples(pos=pos, samples=res, keys=samples.keys)


def _samples_from_checkpoint(ck: Samples, likelihood, rank, world) -> Samples:
    """Inverse of :func:`_samples_to_checkpoint` for THIS run's sharding (the checkpoint holds all samples, so the world
    size may differ from the run that wrote it)."""
    lay, dev, dt = likelihood.layout, likelihood.rt.device, likelihood.dtype
    pos = torch.as_tensor(lay.pack_numpy(ck.pos), dtype=dt, device=dev)
    res = None
    if ck.residuals is not None:
        total = next(iter(ck.residuals.values())).shape[0]
        full = np.stack([lay.pack_numpy({k: v[i] for k, v in ck.residuals.items()}) for i in range(total)])
        n_keys = 0 if ck.keys is None else len(ck.keys)
        per_key = 2 if n_keys and total == 2 * n_keys else 1
        if world > 1 and per_key == 2 and 2 * n_keys == world:
            idx = [rank]                          # the `n_samples == n_devices / 2` layout: one point per rank
        else:
            idx = [g * per_key + s for g in range(rank, total // per_key, world) for s in range(per_key)]
        res = torch.as_tensor(full[idx], dtype=dt, device=dev)
    return Samples(pos=pos, samples=res, keys=ck.keys)


def optimize_kl(likelihood: LikelihoodWithModel, position_or_samples, *, key, n_total_iterations: int, n_samples,
                point_estimates=(), constants=(), draw_linear_kwargs=None, nonlinearly_update_kwargs=None, kl_kwargs=None,
                sample_mode="nonlinear_resample", resume=False, callback: Optional[Callable] = None, odir: Optional[str] = None,
                comm=None, jit=True, linear_minimizer_jit=False, nonlinear_minimizer_jit=False, kl_map=None, residual_map="lmap",
                kl_reduce=None, mirror_samples=True, devices=None, _optimize_vi=None, _optimize_vi_state=None):
    """``jft.optimize_kl`` (optimize_kl.py:744-879): returns ``(Samples, OptimizeVIState)``.  ``resume`` may be ``True``
    (continue from ``odir/last.pkl``) or the path of a checkpoint (:806-839)."""
    opt_vi = _optimize_vi if _optimize_vi is not None else OptimizeVI(
        likelihood, n_total_iterations, comm=comm, jit=jit, linear_minimizer_jit=linear_minimizer_jit,
        nonlinear_minimizer_jit=nonlinear_minimizer_jit, kl_map=kl_map, residual_map=residual_map, kl_reduce=kl_reduce,
        mirror_samples=mirror_samples, devices=devices)
    comm = opt_vi.comm
    rank, world = comm.rank, comm.world
    plan = likelihood.signal.cf.plan
    if plan.dist and world > 1:
        raise ValueError("a slab-decomposed field cannot be combined with sample sharding (`comm`): the slab ranks hold "
                         "parts of ONE latent vector, not independent samples")
    if plan.dist and odir is not None:
        raise NotImplementedError("checkpointing (`odir`) is not available for slab-decomposed fields: every rank holds only "
                                  "its slab of the excitations (gather them with plan.gather_latent and save explicitly)")
    last_fn = os.path.join(odir, "last.pkl") if odir is not None else None
    # `resume` may be True/False or a path (optimize_kl.py:821-826)
    resume_fn = resume if isinstance(resume, str) and os.path.isfile(resume) else last_fn
    samples = None
    if isinstance(position_or_samples, Samples):
        samples = position_or_samples
    else:
        samples = Samples(pos=likelihood.signal.as_flat(position_or_samples), samples=None, keys=None)
    state = None
    if resume and resume_fn is not None and os.path.isfile(resume_fn):
        with open(resume_fn, "rb") as f:
            ck_samples, state = pickle.load(f)
        samples = _samples_from_checkpoint(ck_samples, likelihood, rank, world)
    fresh = opt_vi.init_state(key, n_samples=n_samples, draw_linear_kwargs=draw_linear_kwargs,
                              nonlinearly_update_kwargs=nonlinearly_update_kwargs, kl_kwargs=kl_kwargs, sample_mode=sample_mode,
                              point_estimates=point_estimates, constants=constants)
    state = _optimize_vi_state if _optimize_vi_state is not None else state      # an explicit state wins (:852)
    state = fresh if state is None else state
    if not state.config:                                                            # resumed / supplied states carry no config (:854-855)
        state = state._replace(config=fresh.config)
    if odir is not None and rank == 0:
        os.makedirs(odir, exist_ok=True)
    sanity_fn = os.path.join(odir, "minisanity.txt") if odir is not None else None       # optimize_kl.py:803, 827
    if not resume and sanity_fn is not None and rank == 0:
        open(sanity_fn, "w").close()
    logger = logging.getLogger("nifty_b200")
    for i in range(state.nit, n_total_iterations):
        logger.info(f"OPTIMIZE_KL: Starting {i + 1:04d}")
        samples, state = opt_vi.update(samples, state)
        msg = opt_vi.get_status_message(samples, state, name="OPTIMIZE_KL")      # :866-870 (rank-local samples)
        logger.info(msg)
        if sanity_fn is not None and rank == 0:
            with open(sanity_fn, "a") as f:
                f.write("\n" + msg)
        if last_fn is not None:
            ck = _samples_to_checkpoint(samples, likelihood, comm)
            if rank == 0:
                with open(last_fn, "wb") as f:   # (Samples, OptimizeVIState) as in the reference; config is not pickled (:871-875)
                    pickle.dump((ck, state._replace(config={})), f)
        if callback is not None:
            callback(samples, state)
    return samples, state
