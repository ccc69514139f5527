"""Handles around the C ABI: plans, models, linearisations; torch tensors supply device memory.

PyTorch is plumbing here (allocator, streams, torch.distributed); every arithmetic operation of
the hot path runs in libniftyb200.so.  The default runtime requires a CUDA device and the sm_100a
library and raises otherwise -- there is no CPU fallback.  (The test-suite may construct a
``Runtime`` around the host emulator of tests/emu with ``device="cpu"``.)
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from ._capi import CApi, CgOpts, CgResult, ModelDesc, NB200Error, default_api

_DTYPES = {torch.float32: 0, torch.float64: 1}


class Runtime:
    def __init__(self, api: Optional[CApi] = None, device=None):
        if api is None:
            if not torch.cuda.is_available():
                raise NB200Error("nifty_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            api = default_api()
            device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            if device.type != "cuda":
                raise NB200Error("nifty_b200 runs on CUDA devices only")
        self.api = api
        self.device = torch.device(device)

    # -- helpers --------------------------------------------------------------------------------
    def stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def device_index(self) -> int:
        return self.device.index or 0 if self.device.type == "cuda" else 0

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def asarray(self, x, dtype):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device).contiguous()

    @staticmethod
    def ptr(t: Optional[torch.Tensor]):
        if t is None:
            return None
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())

    def launch_count(self) -> int:
        return int(self.api.lib.nb200_launch_count())

    def timing_begin(self):
        self.api.call("nb200_timing_begin")

    def timing_end(self):
        """-> {kernel body name: (launches, total_ms)} since timing_begin()"""
        buf = C.create_string_buffer(1 << 16)
        self.api.call("nb200_timing_end", buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out


_default_rt = None


def default_runtime() -> Runtime:
    global _default_rt
    if _default_rt is None:
        _default_rt = Runtime()
    return _default_rt


class SlabComm:
    """The collectives the slab-decomposed path needs, on a torch.distributed process group."""

    def __init__(self, group=True):
        import torch.distributed as dist
        self.dist = dist
        if not dist.is_available() or not dist.is_initialized():
            self.group, self.rank, self.world = None, 0, 1
        else:
            self.group = None if group is True else group
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)

    def all_to_all(self, out, inp, out_splits, in_splits):
        n_out, n_in = sum(out_splits), sum(in_splits)
        self.dist.all_to_all_single(out[:n_out], inp[:n_in], out_splits, in_splits, group=self.group)

    def p2p_exchange(self, out, inp, recv, send):
        """Asynchronous personalised exchange: peer p gets inp[send[p]], out[recv[p]] is filled from peer p.
        (offset, count) pairs in elements of the tensors; the own block is a device copy."""
        ops = []
        for p in range(self.world):
            (so, sn), (ro, rn) = send[p], recv[p]
            if p == self.rank:
                if sn:
                    out[ro:ro + rn].copy_(inp[so:so + sn])
                continue
            if sn:
                ops.append(self.dist.P2POp(self.dist.isend, inp[so:so + sn], self._global_rank(p), group=self.group))
            if rn:
                ops.append(self.dist.P2POp(self.dist.irecv, out[ro:ro + rn], self._global_rank(p), group=self.group))
        return self.dist.batch_isend_irecv(ops) if ops else []

    def _global_rank(self, p):
        return p if self.group is None else self.dist.get_global_rank(self.group, p)

    def allreduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    def all_gather_obj(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


class Plan:
    """``nb200_plan``: regular Fourier grid + transform plan."""

    def __init__(self, shape: Sequence[int], distances, dtype=torch.float64, hartley_convention="non_canonical_hartley",
                 runtime: Optional[Runtime] = None, comm=None):
        """``comm``: a ``torch.distributed`` process group (or True for the default group) -> the 3-D grid is
        slab-decomposed over its ranks (one process per GPU); None -> the whole grid lives on this GPU."""
        self.rt = runtime or default_runtime()
        self.comm = SlabComm(comm) if comm is not None else None
        self.dist = self.comm is not None and self.comm.world > 1
        self.shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
        self.distances = tuple(float(d) for d in np.broadcast_to(distances, (len(self.shape),)))
        if dtype not in _DTYPES:
            raise TypeError(f"unsupported dtype {dtype}")
        if hartley_convention not in ("non_canonical_hartley", "canonical_hartley"):
            raise ValueError(f"invalid hartley convention {hartley_convention!r}")
        self.dtype = dtype
        self.convention = hartley_convention
        self._h = C.c_void_p()
        shp = (C.c_int64 * len(self.shape))(*self.shape)
        dst = (C.c_double * len(self.shape))(*self.distances)
        conv = 0 if hartley_convention == "non_canonical_hartley" else 1
        if self.dist:
            self.rt.api.call("nb200_plan_create_dist", C.byref(self._h), self.rt.device_index(), len(self.shape), shp, dst,
                             _DTYPES[dtype], conv, self.comm.rank, self.comm.world)
        else:
            self.rt.api.call("nb200_plan_create", C.byref(self._h), self.rt.device_index(), len(self.shape), shp, dst,
                             _DTYPES[dtype], conv)
        lib = self.rt.api.lib
        self.K = int(lib.nb200_plan_num_modes(self._h))
        self.N = int(lib.nb200_plan_size(self._h))
        self.total_volume = float(lib.nb200_plan_total_volume(self._h))
        self.local_shape = self.shape              # latent xi block held by this rank
        self.local_pos_shape = self.shape          # position-space block (API layout)
        if self.dist:
            self._init_dist()

    # -- slab decomposition -----------------------------------------------------------------------
    def _init_dist(self):
        W = self.comm.world
        info = (C.c_int64 * (9 + 4 * W))()
        self.rt.api.call("nb200_plan_dist_info", self._h, info, len(info))
        v = list(info)
        self.rows0, self.planes2, self.scratch_elems = int(v[2]), int(v[3]), int(v[4])
        n0, n1, n2 = int(v[5]), int(v[6]), int(v[7])
        npad0, npad2 = v[9:9 + W], v[9 + W:9 + 2 * W]
        cA0, cA2 = v[9 + 2 * W:9 + 3 * W], v[9 + 3 * W:9 + 4 * W]
        me = self.comm.rank
        self.local_shape = (self.rows0, n1, n2)
        self.local_pos_shape = (self.planes2, n1, n0)      # internal reversed-axis layout [x2][x1][x0]
        # exchange 1 (after the middle-axis pass): to rank q the k2 planes of q's half-range piece
        self.x1_send = [2 * int(cA2[q] * n1 * npad0[me]) for q in range(W)]
        self.x1_recv = [2 * int(cA2[me] * n1 * npad0[p]) for p in range(W)]
        # exchange 2 (after the second middle-axis pass): to rank r the k0 planes of r's piece
        self.x2_send = [2 * int(cA0[r] * n1 * npad2[me]) for r in range(W)]
        self.x2_recv = [2 * int(cA0[me] * n1 * npad2[q]) for q in range(W)]
        self._s0 = self.rt.zeros((2 * self.scratch_elems,), self.dtype)
        self._s1 = self.rt.zeros((2 * self.scratch_elems,), self.dtype)
        self._s2 = self.rt.zeros((2 * self.scratch_elems,), self.dtype)
        self.rt.api.call("nb200_plan_set_scratch", self._h, self.rt.ptr(self._s0), self.rt.ptr(self._s1), self.rt.ptr(self._s2))
        self._dist_tabs = (n1, [int(x) for x in npad0], [int(x) for x in npad2], [int(x) for x in cA0], [int(x) for x in cA2])
        self.nchunks = 1
        nch = int(os.environ.get("NB200_SLAB_CHUNKS", "1"))
        if nch > 1:
            self.set_chunks(nch)
        rows = np.zeros(self.rows0, dtype=np.int32)
        planes = np.zeros(self.planes2, dtype=np.int32)
        self.rt.api.call("nb200_plan_local_map", self._h, 0, rows.ctypes.data_as(C.c_void_p))
        self.rt.api.call("nb200_plan_local_map", self._h, 2, planes.ctypes.data_as(C.c_void_p))
        self.row_map, self.plane_map = rows, planes          # local -> global index, -1 = zero padding

    def exchange(self, which: int):
        """All-to-all between the local passes: 1 = S1 -> S0 (before the axis-0 pass), 2 = S0 -> S1."""
        if which == 1:
            self.comm.all_to_all(self._s0, self._s1, self.x1_recv, self.x1_send)
        else:
            self.comm.all_to_all(self._s1, self._s0, self.x2_recv, self.x2_send)

    def set_reduce_chunks(self, nchunks: int):
        """Sample-averaged products (`metric_multi`): run the last pass in `nchunks` launches and hand every finished
        range of the result to the reduction hook right away, so that its all-reduce overlaps the remaining launches."""
        self.rt.api.call("nb200_plan_set_reduce_chunks", self._h, int(nchunks))

    def set_chunks(self, nchunks: int):
        """Pipeline every exchange in `nchunks` pieces (point-to-point sends of chunk c overlap the passes
        that produce chunk c+1 and consume chunk c-1)."""
        self.rt.api.call("nb200_plan_set_chunks", self._h, int(nchunks))
        self.nchunks = int(nchunks)
        n1, npad0, npad2, cA0, cA2 = self._dist_tabs
        W, me = self.comm.world, self.comm.rank

        def split(start, count, c):
            return start + (count * c) // nchunks

        def tables(cA, npad):
            a0 = [sum(cA[:q]) for q in range(W)]
            roff = [sum(cA[me] * n1 * npad[p2] for p2 in range(p)) for p in range(W)]
            send, recv = [], []
            for c in range(nchunks):
                sc, rc = [], []
                for q in range(W):
                    s, e = split(a0[q], cA[q], c), split(a0[q], cA[q], c + 1)
                    sc.append((2 * s * n1 * npad[me], 2 * (e - s) * n1 * npad[me]))
                    ls, le = split(0, cA[me], c), split(0, cA[me], c + 1)
                    rc.append((2 * (roff[q] + ls * n1 * npad[q]), 2 * (le - ls) * n1 * npad[q]))
                send.append(sc)
                recv.append(rc)
            return send, recv
        # exchange 1 moves k2 planes (ownership of the last axis), rows padded per the axis-0 decomposition
        self._x1 = tables(cA2, npad0)
        self._x2 = tables(cA0, npad2)

    def exchange_chunk(self, which: int, c: int):
        """Start chunk c of exchange `which`; returns a handle for :meth:`wait`."""
        src, dst = (self._s1, self._s0) if which == 1 else (self._s0, self._s1)
        send, recv = self._x1 if which == 1 else self._x2
        return self.comm.p2p_exchange(dst, src, recv[c], send[c])

    @staticmethod
    def wait(handle):
        for w in handle:
            w.wait()

    def scatter_latent(self, xi_global) -> torch.Tensor:
        """Rows of a global natural-order grid array owned by this rank (zero padding rows)."""
        g = self.rt.asarray(xi_global, self.dtype)
        out = self.rt.zeros(self.local_shape, self.dtype)
        idx = torch.as_tensor(self.row_map.astype(np.int64), device=self.rt.device)
        ok = idx >= 0
        out[ok] = g[idx[ok]]
        return out

    def scatter_position(self, pos_global) -> torch.Tensor:
        """Planes x2 of a global natural-order position array owned by this rank, in [x2][x1][x0] layout."""
        g = self.rt.asarray(pos_global, self.dtype).permute(2, 1, 0)
        out = self.rt.zeros(self.local_pos_shape, self.dtype)
        idx = torch.as_tensor(self.plane_map.astype(np.int64), device=self.rt.device)
        ok = idx >= 0
        out[ok] = g[idx[ok]]
        return out.contiguous()

    def gather_latent(self, local) -> torch.Tensor:
        """Inverse of scatter_latent on every rank (all-gather; for tests and small grids)."""
        return self._gather(local.reshape(self.local_shape), 0)

    def gather_position(self, local) -> torch.Tensor:
        return self._gather(local.reshape(self.local_pos_shape), 2).permute(2, 1, 0).contiguous()

    def _gather(self, local, axis):
        W = self.comm.world
        n = self.shape[axis]
        nloc = torch.tensor([local.shape[0]], dtype=torch.int64, device=self.rt.device)
        sizes = [int(x) for x in self.comm.all_gather_obj(int(local.shape[0]))]
        maps = self.comm.all_gather_obj((self.row_map if axis == 0 else self.plane_map).tolist())
        mx = max(sizes)
        pad = self.rt.zeros((mx,) + tuple(local.shape[1:]), self.dtype)
        pad[:local.shape[0]] = local
        parts = self.comm.all_gather(pad)
        out = self.rt.zeros((n,) + tuple(local.shape[1:]), self.dtype)
        for p in range(W):
            m = torch.as_tensor(np.asarray(maps[p], dtype=np.int64), device=self.rt.device)
            ok = m >= 0
            out[m[ok]] = parts[p][:sizes[p]][ok]
        return out

    def __del__(self):
        try:
            if self._h:
                self.rt.api.lib.nb200_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _table(self, name, n, dtype):
        out = np.zeros(n, dtype=dtype)
        self.rt.api.call(name, self._h, out.ctypes.data_as(C.c_void_p))
        return out

    @property
    def mode_lengths(self):
        return self._table("nb200_plan_mode_lengths", self.K, np.float64)

    @property
    def mode_multiplicity(self):
        return self._table("nb200_plan_mode_multiplicity", self.K, np.int64)

    @property
    def relative_log_mode_lengths(self):
        return self._table("nb200_plan_relative_log_mode_lengths", self.K, np.float64)

    @property
    def log_volume(self):
        return self._table("nb200_plan_log_volume", max(self.K - 2, 0), np.float64)

    @property
    def power_distributor(self):
        return self._table("nb200_plan_power_distributor", self.N, np.int32).reshape(self.shape)

    # -- raw operators --------------------------------------------------------------------------
    def vec_stats(self, x: torch.Tensor):
        """(sum x, sum x^2) of a contiguous device vector of this plan's dtype (``nb200_vec_stats``)."""
        x = self.rt.asarray(x, self.dtype).reshape(-1)
        out = (C.c_double * 2)()
        self.rt.api.call("nb200_vec_stats", self._h, self.rt.stream(), int(x.numel()), self.rt.ptr(x), out)
        return float(out[0]), float(out[1])

    def hartley(self, x: torch.Tensor) -> torch.Tensor:
        x = self.rt.asarray(x, self.dtype)
        if tuple(x.shape) != self.shape:
            raise ValueError(f"shape mismatch: {tuple(x.shape)} vs {self.shape}")
        out = torch.empty_like(x)
        self.rt.api.call("nb200_hartley", self._h, self.rt.stream(), self.rt.ptr(x), self.rt.ptr(out))
        return out

    def cf_apply(self, amp: torch.Tensor, xi: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
        amp, xi = self.rt.asarray(amp, self.dtype), self.rt.asarray(xi, self.dtype)
        out = torch.empty_like(xi)
        self.rt.api.call("nb200_cf_apply", self._h, self.rt.stream(), self.rt.ptr(amp), self.rt.ptr(xi), float(offset),
                         self.rt.ptr(out))
        return out

    def cf_apply_batch(self, amp: torch.Tensor, xi: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
        """`jax.vmap` form of :meth:`cf_apply`: xi [batch, *grid]; amp [K] (shared by the batch) or [batch, K]."""
        amp, xi = self.rt.asarray(amp, self.dtype).contiguous(), self.rt.asarray(xi, self.dtype).contiguous()
        batch = xi.shape[0]
        if amp.ndim == 2 and amp.shape[0] != batch:
            raise ValueError("cf_apply_batch: amp must be [K] or [batch, K]")
        out = torch.empty_like(xi)
        self.rt.api.call("nb200_cf_apply_batch", self._h, self.rt.stream(), self.rt.ptr(amp), self.K if amp.ndim == 2 else 0,
                         self.rt.ptr(xi), float(offset), self.rt.ptr(out), int(batch))
        return out

    def cf_apply_adjoint_batch(self, amp, xi, cot):
        """`jax.vmap` form of :meth:`cf_apply_adjoint`: cot / xi [batch, *grid]; amp [K] or [batch, K]; returns
        (xi_bar [batch, *grid], amp_bar [batch, K] or None when xi is None)."""
        amp, cot = self.rt.asarray(amp, self.dtype).contiguous(), self.rt.asarray(cot, self.dtype).contiguous()
        batch = cot.shape[0]
        xi_bar = torch.empty_like(cot)
        amp_bar, xi_p = None, None
        if xi is not None:
            xi = self.rt.asarray(xi, self.dtype).contiguous()
            amp_bar = self.rt.empty((batch, self.K), self.dtype)
            xi_p = self.rt.ptr(xi)
        self.rt.api.call("nb200_cf_apply_adjoint_batch", self._h, self.rt.stream(), self.rt.ptr(amp), self.K if amp.ndim == 2 else 0,
                         xi_p, self.rt.ptr(cot), self.rt.ptr(xi_bar), self.rt.ptr(amp_bar) if amp_bar is not None else None,
                         self.K, int(batch))
        return xi_bar, amp_bar

    def cf_apply_adjoint(self, amp, xi, cot):
        amp, cot = self.rt.asarray(amp, self.dtype), self.rt.asarray(cot, self.dtype)
        xi_bar = torch.empty_like(cot)
        if xi is None:
            self.rt.api.call("nb200_cf_apply_adjoint", self._h, self.rt.stream(), self.rt.ptr(amp), None, self.rt.ptr(cot),
                             self.rt.ptr(xi_bar), None)
            return xi_bar, None
        xi = self.rt.asarray(xi, self.dtype)
        amp_bar = self.rt.empty((self.K,), self.dtype)
        self.rt.api.call("nb200_cf_apply_adjoint", self._h, self.rt.stream(), self.rt.ptr(amp), self.rt.ptr(xi),
                         self.rt.ptr(cot), self.rt.ptr(xi_bar), self.rt.ptr(amp_bar))
        return xi_bar, amp_bar


class ModelHandle:
    """``nb200_model``"""

    def __init__(self, plan: Plan, desc: ModelDesc):
        self.plan, self.rt, self.desc = plan, plan.rt, desc
        self.L = int(desc.latent_size)
        self._h = C.c_void_p()
        self.rt.api.call("nb200_model_create", C.byref(self._h), plan._h, C.byref(desc))
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                self.rt.api.lib.nb200_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_likelihood(self, kind: int, nonlinearity: int, data, w_scalar: float, w_array):
        d = None if data is None else self.rt.asarray(data, self.plan.dtype)
        w = None if w_array is None else self.rt.asarray(w_array, self.plan.dtype)
        self.rt.api.call("nb200_model_set_likelihood", self._h, self.rt.stream(), kind, nonlinearity, self.rt.ptr(d),
                         float(w_scalar), self.rt.ptr(w))

    def cf_forward(self, pos: torch.Tensor) -> torch.Tensor:
        if self.plan.dist:
            raise NB200Error("cf_forward is not available on slab-decomposed plans; linearise and read the signal")
        out = self.rt.empty(self.plan.shape, self.plan.dtype)
        self.rt.api.call("nb200_cf_forward", self._h, self.rt.stream(), self.rt.ptr(pos), self.rt.ptr(out))
        return out

    def cf_amplitude(self, pos: torch.Tensor) -> torch.Tensor:
        """``[azm V, amp_1, ..., amp_{K-1}]`` at ``pos`` (``nb200_cf_amplitude``); replicated on slab-decomposed plans."""
        out = self.rt.empty((self.plan.K,), self.plan.dtype)
        self.rt.api.call("nb200_cf_amplitude", self._h, self.rt.stream(), self.rt.ptr(pos), self.rt.ptr(out))
        return out


class Lin:
    """``nb200_lin``: cached linearisation of a model at one latent position."""

    def __init__(self, model: ModelHandle):
        self.model, self.rt = model, model.rt
        self._h = C.c_void_p()
        self.rt.api.call("nb200_lin_create", C.byref(self._h), model._h)

    def __del__(self):
        try:
            if self._h:
                self.rt.api.lib.nb200_lin_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _vec(self):
        if self.model.plan.dist:      # padding rows of the local xi block are never written by the kernels
            return self.rt.zeros((self.model.L,), self.model.plan.dtype)
        return self.rt.empty((self.model.L,), self.model.plan.dtype)

    def _pos(self):
        return self.rt.empty(self.model.plan.local_pos_shape, self.model.plan.dtype)

    # -- slab-decomposed sequences: local phases (nb200_dist_phase) + host collectives ------------------
    def _phase(self, code, other=None, inp=None, out=None, flag=0, chunk=-1):
        self.rt.api.call("nb200_dist_phase", self._h, None if other is None else other._h, self.rt.stream(), int(code), int(chunk),
                         self.rt.ptr(inp), self.rt.ptr(out), self.rt.ptr(self._abar), self.rt.ptr(self._xs), int(flag))

    def _pipelined(self, which, produce_code, consume, **produce_kw):
        """chunk c: produce (PCa / PCb) -> start exchange c; then for every chunk: wait -> consume (P3 / P5)."""
        plan = self.model.plan
        handles = []
        for c in range(plan.nchunks):
            self._phase(produce_code, chunk=c, **produce_kw)
            handles.append(plan.exchange_chunk(which, c))
        for c in range(plan.nchunks):
            plan.wait(handles[c])
            consume(c)

    def _dist_buffers(self):
        if not hasattr(self, "_abar"):
            self._abar = self.rt.zeros((self.model.plan.K,), self.model.plan.dtype)
            self._xs = self.rt.zeros((4,), self.model.plan.dtype)

    def _dist_adjoint_tail(self, t, out, add_identity, scaled, grad=False):
        plan = self.model.plan
        tin = t if add_identity else None
        f5 = int(add_identity) | (4 if grad else 0)
        if plan.nchunks > 1:      # PCb chunk -> exchange 2 chunk -> P5 chunk, then the bin sums
            self._pipelined(2, 13, lambda c: self._phase(23, inp=tin, out=out, flag=int(add_identity), chunk=c))
            self._phase(24, inp=tin, out=out, flag=f5)
        else:
            plan.exchange(2)
            self._phase(5, inp=tin, out=out, flag=f5)
        plan.comm.allreduce_sum(self._abar)
        plan.comm.allreduce_sum(self._xs)
        self._phase(6, inp=t if add_identity else None, out=out, flag=int(add_identity) | (2 if scaled else 0))
        return out

    def update(self, pos: torch.Tensor, want_grad=False, add_prior=False):
        if self.model.plan.dist:
            self._dist_buffers()
            plan = self.model.plan
            g = int(bool(want_grad))
            if plan.nchunks > 1:
                self._phase(10, inp=pos)
                self._pipelined(1, 11, lambda c: self._phase(12, chunk=c, flag=g))
                self._phase(14)
            else:
                self._phase(0, inp=pos)
                plan.exchange(1)
                self._phase(1, flag=g)
            if not want_grad:
                return None
            if plan.nchunks == 1:
                self._phase(13)                   # PCb of dE/df (code 1 stops after P3)
            grad = self._vec()
            # gradient = J^T dE/ds (+ pos): the adjoint tail with `pos` as the additive term
            return self._dist_adjoint_tail(pos, grad, bool(add_prior), True, grad=True)
        grad = self._vec() if want_grad else None
        nl_fn = getattr(self.model, "nl_fn", None)
        if nl_fn is not None:
            # user-supplied pointwise map of the field: evaluated here, at this position, together with its derivative
            # (torch autograd on the host side of the boundary -- once per linearisation, not per product)
            field = self.model.cf_forward(pos)
            with torch.enable_grad():
                f = field.detach().requires_grad_(True)
                s = nl_fn(f)
                if s.shape != f.shape:
                    raise ValueError("the non-linearity must be a pointwise map of the correlated field (same shape in and out)")
                (ds,) = torch.autograd.grad(s.sum(), f)
            s, ds = s.detach().contiguous(), ds.contiguous()
            self.rt.api.call("nb200_lin_set_pointwise", self._h, self.rt.stream(), self.rt.ptr(s), self.rt.ptr(ds))
        self.rt.api.call("nb200_lin_update", self._h, self.rt.stream(), self.rt.ptr(pos), self.rt.ptr(grad), int(add_prior))
        return grad

    def energy(self) -> float:
        e = C.c_double()
        self.rt.api.call("nb200_lin_energy", self._h, self.rt.stream(), C.byref(e))
        if self.model.plan.dist:
            t = torch.tensor([e.value], dtype=torch.float64, device=self.rt.device)
            return float(self.model.plan.comm.allreduce_sum(t))
        return e.value

    def amplitude(self):
        out = self.rt.empty((self.model.plan.K,), self.model.plan.dtype)
        self.rt.api.call("nb200_lin_amplitude", self._h, self.rt.stream(), self.rt.ptr(out))
        return out

    def signal(self):
        out = self._pos()
        self.rt.api.call("nb200_lin_signal", self._h, self.rt.stream(), self.rt.ptr(out))
        return out

    def metric(self, t, add_identity=False, out=None):
        out = self._vec() if out is None else out
        if self.model.plan.dist:
            plan = self.model.plan
            if plan.nchunks > 1:
                self._phase(20, inp=t)
                self._pipelined(1, 11, lambda c: self._phase(21, inp=t, chunk=c))
            else:
                self._phase(2, inp=t)
                plan.exchange(1)
                self._phase(3, inp=t)
            return self._dist_adjoint_tail(t, out, add_identity, True)
        self.rt.api.call("nb200_metric", self._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out), int(add_identity))
        return out

    def metric_pair(self, other: "Lin", t, add_identity=False):
        out = self._vec()
        if self.model.plan.dist:      # same phases as `metric`: tangent side from `other`, adjoint side from `self`
            plan = self.model.plan
            self._dist_buffers()
            if plan.nchunks > 1:
                self._phase(20, other=other, inp=t)
                self._pipelined(1, 11, lambda c: self._phase(21, other=other, inp=t, chunk=c))
            else:
                self._phase(2, other=other, inp=t)
                plan.exchange(1)
                self._phase(3, other=other, inp=t)
            return self._dist_adjoint_tail(t, out, add_identity, True)
        self.rt.api.call("nb200_metric_pair", self._h, other._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out),
                         int(add_identity))
        return out

    def rsm(self, t, scaled=True):
        out = self._pos()
        if self.model.plan.dist:      # tangent chain + P1 + PCa | exchange 1 | forward-only P3 into the local planes
            plan = self.model.plan
            self._dist_buffers()
            out.zero_()
            if plan.nchunks > 1:
                self._phase(20, inp=t)
                self._pipelined(1, 11, lambda c: self._phase(25, inp=t, out=out, flag=int(scaled), chunk=c))
            else:
                self._phase(2, inp=t)
                plan.exchange(1)
                self._phase(7, inp=t, out=out, flag=int(scaled))
            return out
        self.rt.api.call("nb200_rsm", self._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out), int(scaled))
        return out

    def lsm(self, u, scaled=True):
        out = self._vec()
        u = self.rt.asarray(u, self.model.plan.dtype)
        if self.model.plan.dist:      # u: local planes in the internal [x2][x1][x0] layout
            if self.model.plan.nchunks > 1:
                for c in range(self.model.plan.nchunks):
                    self._phase(22, inp=u, flag=int(scaled), chunk=c)
            else:
                self._phase(4, inp=u, flag=int(scaled))
            return self._dist_adjoint_tail(None, out, False, scaled)
        self.rt.api.call("nb200_lsm", self._h, self.rt.stream(), self.rt.ptr(u), self.rt.ptr(out), int(scaled))
        return out

    def transformation(self):
        out = self._pos()
        self.rt.api.call("nb200_transformation", self._h, self.rt.stream(), self.rt.ptr(out))
        return out

    def normalized_residual(self):
        out = self._pos()
        self.rt.api.call("nb200_normalized_residual", self._h, self.rt.stream(), self.rt.ptr(out))
        return out

    def cg_solve(self, j, x0=None, other: Optional["Lin"] = None, *, absdelta=None, resnorm=None, norm_ord=None,
                 tol=1e-5, atol=0.0, miniter=None, maxiter=None, raise_nonposdef=True, check_every=4, frozen=None):
        """Solve (metric + 1) x = j on the device (``_cg`` semantics); returns (x, CgResult).
        ``frozen``: ``[(lo, hi), ...]`` ranges of the flat latent vector held fixed (point estimates / constants):
        the solve runs in the subspace of the other entries; ``j`` / ``x0`` must be zero on them."""
        o = CgOpts()
        self.rt.api.lib.nb200_cg_default_opts(C.byref(o))
        o.absdelta = -1.0 if absdelta is None else float(absdelta)
        o.resnorm = -1.0 if resnorm is None else float(resnorm)
        o.tol, o.atol = float(tol), float(atol)
        if norm_ord is None:
            norm_ord = 2
        o.norm_ord = {1: 1, 2: 2, np.inf: 0, float("inf"): 0}[norm_ord]
        o.miniter = -1 if miniter is None else int(miniter)
        o.maxiter = -1 if maxiter is None else int(maxiter)
        o.raise_nonposdef = int(bool(raise_nonposdef))
        o.check_every = int(check_every)
        o.x0_is_zero = int(x0 is None)
        fr = None
        if frozen:
            flat = [int(v) for r in frozen for v in r]
            fr = (C.c_int64 * len(flat))(*flat)
            o.n_frozen, o.frozen = len(flat) // 2, fr
        x = self._vec() if x0 is None else x0.clone()
        res = CgResult()
        self.rt.api.call("nb200_cg_solve", self._h, other._h if other is not None else None, self.rt.stream(),
                         self.rt.ptr(j), self.rt.ptr(x), C.byref(o), C.byref(res))
        return x, res


# -- sample-averaged operator of the KL on the device (nb200_metric_multi / nb200_cg_solve_multi) -------------------
class _DevView:
    """`__cuda_array_interface__` view of a device buffer handed to a reduction hook by the library."""

    def __init__(self, ptr, n, dtype):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8" if dtype == torch.float64 else "<f4",
                                         "data": (int(ptr), False), "version": 2}


def make_reduce_hook(rt: "Runtime", dtype, reduce_fn):
    """ctypes callback for `nb200_reduce_hook`: `reduce_fn(tensor)` receives a tensor aliasing one finished range of the
    library's output and starts its in-place SUM all-reduce; it may return a work handle (``async_op=True``: the
    collective runs on the communication stream while the library launches the next piece) -- the handles are waited on
    when the library signals that every range has been handed over.  Returns None if `reduce_fn` is None.  The caller
    keeps the returned object alive for the duration of the call."""
    if reduce_fn is None:
        return None
    from ._capi import REDUCE_HOOK
    pending = []

    def _hook(user, buf, n, stream):
        if not buf or n == 0:                 # every range handed over: complete them in stream order
            for w in pending:
                w.wait()
            pending.clear()
            return
        if rt.device.type == "cuda":
            t = torch.as_tensor(_DevView(buf, n, dtype), device=rt.device)
        else:       # host emulation (tests): the buffer is host memory
            import numpy as np
            ctype = C.c_double if dtype == torch.float64 else C.c_float
            t = torch.from_numpy(np.ctypeslib.as_array((ctype * n).from_address(buf)))
        w = reduce_fn(t)
        if w is not None:
            pending.append(w)

    return REDUCE_HOOK(_hook)


def _lin_array(lins):
    arr = (C.c_void_p * len(lins))(*[l._h for l in lins])
    return arr


def metric_multi(lins, t, *, scale=1.0, identity_here=True, out=None, reduce_fn=None):
    """out = sum_i scale * metric(lin_i, t) (+ t), all-reduced through `reduce_fn` -- `_kl_met` on the device."""
    lin0 = lins[0]
    rt = lin0.rt
    out = lin0._vec() if out is None else out
    hook = make_reduce_hook(rt, lin0.model.plan.dtype, reduce_fn)
    rt.api.call("nb200_metric_multi", _lin_array(lins), len(lins), float(scale), int(identity_here), rt.stream(), rt.ptr(t), rt.ptr(out),
                C.cast(hook, C.c_void_p) if hook is not None else None, None)
    return out


def cg_solve_multi(lins, j, x0=None, *, scale=1.0, identity_here=True, reduce_fn=None, absdelta=None, resnorm=None, norm_ord=None,
                   tol=1e-5, atol=0.0, miniter=None, maxiter=None, raise_nonposdef=True, check_every=4, frozen=None):
    """`_cg` on the sample-averaged operator, on the device (see :meth:`Lin.cg_solve` for the arguments)."""
    lin0 = lins[0]
    rt = lin0.rt
    o = CgOpts()
    rt.api.lib.nb200_cg_default_opts(C.byref(o))
    o.absdelta = -1.0 if absdelta is None else float(absdelta)
    o.resnorm = -1.0 if resnorm is None else float(resnorm)
    o.tol, o.atol = float(tol), float(atol)
    if norm_ord is None:
        norm_ord = 2
    o.norm_ord = {1: 1, 2: 2, np.inf: 0, float("inf"): 0}[norm_ord]
    o.miniter = -1 if miniter is None else int(miniter)
    o.maxiter = -1 if maxiter is None else int(maxiter)
    o.raise_nonposdef = int(bool(raise_nonposdef))
    o.check_every = int(check_every)
    o.x0_is_zero = int(x0 is None)
    fr = None
    if frozen:
        flat = [int(v) for r in frozen for v in r]
        fr = (C.c_int64 * len(flat))(*flat)
        o.n_frozen, o.frozen = len(flat) // 2, fr
    x = lin0._vec() if x0 is None else x0.clone()
    res = CgResult()
    hook = make_reduce_hook(rt, lin0.model.plan.dtype, reduce_fn)
    rt.api.call("nb200_cg_solve_multi", _lin_array(lins), len(lins), float(scale), int(identity_here), rt.stream(), rt.ptr(j), rt.ptr(x),
                C.byref(o), C.byref(res), C.cast(hook, C.c_void_p) if hook is not None else None, None)
    return x, res
