"""Handles around the C ABI: plans, models, linearisations; torch tensors supply device memory.

PyTorch is plumbing here (allocator, streams, torch.distributed); every arithmetic operation of
the hot path runs in libniftyb200.so.  The default runtime requires a CUDA device and the sm_100a
library and raises otherwise -- there is no CPU fallback.  (The test-suite may construct a
``Runtime`` around the host emulator of tests/emu with ``device="cpu"``.)
"""

from __future__ import annotations

import ctypes as C
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


class Plan:
    """``nb200_plan``: regular Fourier grid + transform plan."""

    def __init__(self, shape: Sequence[int], distances, dtype=torch.float64, hartley_convention="non_canonical_hartley",
                 runtime: Optional[Runtime] = None):
        self.rt = runtime or default_runtime()
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
        self.rt.api.call("nb200_plan_create", C.byref(self._h), self.rt.device_index(), len(self.shape), shp, dst,
                         _DTYPES[dtype], 0 if hartley_convention == "non_canonical_hartley" else 1)
        lib = self.rt.api.lib
        self.K = int(lib.nb200_plan_num_modes(self._h))
        self.N = int(lib.nb200_plan_size(self._h))
        self.total_volume = float(lib.nb200_plan_total_volume(self._h))

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
        out = self.rt.empty(self.plan.shape, self.plan.dtype)
        self.rt.api.call("nb200_cf_forward", self._h, self.rt.stream(), self.rt.ptr(pos), self.rt.ptr(out))
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
        return self.rt.empty((self.model.L,), self.model.plan.dtype)

    def _pos(self):
        return self.rt.empty(self.model.plan.shape, self.model.plan.dtype)

    def update(self, pos: torch.Tensor, want_grad=False, add_prior=False):
        grad = self._vec() if want_grad else None
        self.rt.api.call("nb200_lin_update", self._h, self.rt.stream(), self.rt.ptr(pos), self.rt.ptr(grad), int(add_prior))
        return grad

    def energy(self) -> float:
        e = C.c_double()
        self.rt.api.call("nb200_lin_energy", self._h, self.rt.stream(), C.byref(e))
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
        self.rt.api.call("nb200_metric", self._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out), int(add_identity))
        return out

    def metric_pair(self, other: "Lin", t, add_identity=False):
        out = self._vec()
        self.rt.api.call("nb200_metric_pair", self._h, other._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out),
                         int(add_identity))
        return out

    def rsm(self, t, scaled=True):
        out = self._pos()
        self.rt.api.call("nb200_rsm", self._h, self.rt.stream(), self.rt.ptr(t), self.rt.ptr(out), int(scaled))
        return out

    def lsm(self, u, scaled=True):
        out = self._vec()
        u = self.rt.asarray(u, self.model.plan.dtype)
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
                 tol=1e-5, atol=0.0, miniter=None, maxiter=None, raise_nonposdef=True, check_every=4):
        """Solve (metric + 1) x = j on the device (``_cg`` semantics); returns (x, CgResult)."""
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
        x = self._vec() if x0 is None else x0.clone()
        res = CgResult()
        self.rt.api.call("nb200_cg_solve", self._h, other._h if other is not None else None, self.rt.stream(),
                         self.rt.ptr(j), self.rt.ptr(x), C.byref(o), C.byref(res))
        return x, res
