"""Correlated fields on grids whose extents are NOT powers of two (the reference's own parity test runs (3, 3),
``test/test_re/test_correlated_field.py:123-124``; its published benchmark 3618^2 ... 10000^2).

The fused sm_100a passes of this library are power-of-two line transforms.  An n-point DFT with arbitrary n is a convolution
with a chirp (Bluestein): with ``c_m = exp(i pi m^2 / n)``

    X_k = conj(c_k) * sum_j (x_j conj(c_j)) c_{k-j} ,

and the convolution is cyclic on any padded length ``M >= 2 n - 1`` -- a power of two, i.e. a transform this library has.  In
several dimensions chirps and filter are outer products of per-axis tables and the padded transform is the n-D one.  A complex
transform of the padded grid is two real device Hartley transforms recombined with their reflections.  The whole sequence --
pad + chirp, two Hartley transforms, spectrum product on mirror pairs, two Hartley transforms, crop + chirp -- is ONE C-ABI call
(``nb200_hartley_chirpz``, three streaming kernels of ``csrc/nb_bluestein.cuh`` around the existing passes).  Cost: four
power-of-two transforms of the padded grid, (M / n)^d ~ 4 ... 16 times the points of the logical grid in 2-D; a mixed-radix line
FFT inside the pass bodies would avoid the padding and is the roofline-grade answer.  Extents up to
4096 per axis in float64 (padded 8192, the longest line the passes hold in shared memory), 8192 in float32.

The transform is linear and self-adjoint, so it enters torch autograd as one function; the O(K) amplitude chain and the
pointwise likelihood are torch operations on the same device, and model / likelihood operators / CG / MGVI come from the
host-composed machinery of the outer-product fields (``outer.py``).  Checked against the oracle and the nifty.cl fixture
``g2d_3x3`` on both test tiers.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._runtime import Plan
from .model import LazyModel
from .outer import _Prior, _amplitude_spec, _device_tables, _eval_amplitude
from .tree import Layout


def fourier_mode_tables(shape, distances, uniqueness_rtol=1e-12):
    """``get_fourier_mode_distributor`` + ``_unique_mode_distributor`` + ``_log_modes`` + ``make_grid`` of the reference
    (nifty/re/correlated_field.py:134-176, 55-67, 228-235, 238-265) for arbitrary extents, host NumPy like there.  Returns a
    dict: power_distributor, mode_lengths, mode_multiplicity, relative_log_mode_lengths, log_volume, total_volume."""
    shape = tuple(int(s) for s in shape)
    distances = tuple(np.broadcast_to(distances, (len(shape),)).astype(np.float64))
    step = 1.0 / (np.array(shape) * np.array(distances))
    # the length of mode i along an axis is min(i, n - i) * step: every value of the full grid occurs on the FOLDED index range
    # [0, n // 2] per axis, so lengths, unique values and bins are computed there (same arithmetic, a fraction of the elements) and
    # expanded by indexing with the folded index of every grid point
    fold = [np.minimum(np.arange(n), n - np.arange(n)) for n in shape]
    m = np.arange(shape[0] // 2 + 1) * step[0]
    if len(shape) != 1:
        m = m * m
        for i in range(1, len(shape)):
            t = np.arange(shape[i] // 2 + 1) * step[i]
            m = np.expand_dims(m, axis=-1) + t * t
        m = np.sqrt(m)
    um = np.unique(m)
    tol = uniqueness_rtol * um[-1]
    um = um[np.diff(np.append(um, 2 * um[-1])) > tol]
    idx = np.searchsorted(0.5 * (um[:-1] + um[1:]), m)[np.ix_(*fold)]
    cnt = np.bincount(idx.ravel(), minlength=um.size)
    if np.any(cnt == 0) or um.shape != cnt.shape:
        raise RuntimeError("invalid harmonic mode(s) encountered")
    rel = um.copy()
    rel[1:] = np.log(rel[1:])
    rel[1:] -= rel[1]
    return dict(power_distributor=idx, mode_lengths=um, mode_multiplicity=cnt, relative_log_mode_lengths=rel,
                log_volume=rel[2:] - rel[1:-1], total_volume=float(np.prod(np.array(shape) * np.array(distances))))


def grid_tables(shape, distances, *, dtype=torch.float64, convention="non_canonical_hartley", runtime=None):
    """Mode tables of one (sub-)grid as a dict (keys of :func:`fourier_mode_tables` + shape, distances): from the device plan
    (``csrc/nb_plan.cuh``, host C++) when every extent is a power of two, from the NumPy restatement otherwise."""
    shape = tuple(int(s) for s in shape)
    dists = tuple(float(x) for x in np.broadcast_to(distances, (len(shape),)))
    if all(n >= 2 and not (n & (n - 1)) for n in shape):
        pl = Plan(shape, dists, dtype=dtype, hartley_convention=convention, runtime=runtime)
        tb = dict(power_distributor=np.asarray(pl.power_distributor), mode_lengths=np.asarray(pl.mode_lengths),
                  mode_multiplicity=np.asarray(pl.mode_multiplicity), relative_log_mode_lengths=np.asarray(pl.relative_log_mode_lengths),
                  log_volume=np.asarray(pl.log_volume), total_volume=float(pl.total_volume))
    else:
        tb = fourier_mode_tables(shape, dists)
    tb["shape"], tb["distances"] = shape, dists
    return tb


def grid_record(tb):
    """``RegularCartesianGrid`` with its harmonic grid (correlated_field.py:238-265) from a :func:`grid_tables` dict."""
    from .correlated_field import RegularCartesianGrid, RegularFourierGrid
    hg = RegularFourierGrid(shape=tb["shape"], power_distributor=tb["power_distributor"], mode_multiplicity=tb["mode_multiplicity"],
                            mode_lengths=tb["mode_lengths"], relative_log_mode_lengths=tb["relative_log_mode_lengths"],
                            log_volume=tb["log_volume"])
    return RegularCartesianGrid(shape=tb["shape"], total_volume=tb["total_volume"], distances=tb["distances"], harmonic_grid=hg)


class BluesteinHartley:
    """Unnormalised n-D Hartley transform (``hartley``, correlated_field.py:24-30) of a grid with arbitrary extents through
    power-of-two device transforms of the padded grid: one ``nb200_hartley_chirpz`` call (csrc/nb_bluestein.cuh)."""

    def __init__(self, shape, *, dtype=torch.float64, convention="non_canonical_hartley", runtime=None):
        self.shape = tuple(int(s) for s in shape)
        if not 1 <= len(self.shape) <= 3:
            raise NotImplementedError("1 to 3 axes")
        self.pad = tuple(max(2, 1 << int(np.ceil(np.log2(2 * n - 1)))) for n in self.shape)
        max_line = 8192 if dtype == torch.float64 else 16384       # one complex line of the passes must fit into shared memory (227 KB)
        if max(self.pad) > max_line:
            raise NotImplementedError(f"non-power-of-two extents above {max_line // 2} are not supported in {dtype} (shape {self.shape} needs "
                                      f"padded lines of {max(self.pad)} points; the passes hold one line of at most {max_line} in shared memory)")
        self.plan = Plan(self.pad, 1.0, dtype=dtype, hartley_convention=convention, runtime=runtime)
        self.rt, self.dtype = self.plan.rt, dtype
        lead = 3 - len(self.shape)
        chirps, filters = [np.ones(1, dtype=np.complex128)] * lead, [np.ones(1, dtype=np.complex128)] * lead
        for n, M in zip(self.shape, self.pad):
            j = np.arange(n, dtype=np.int64)
            c = np.exp(1j * np.pi * ((j * j) % (2 * n)).astype(np.float64) / n)     # pi j^2 / n reduced mod 2 pi exactly
            g = np.zeros(M, dtype=np.complex128)
            g[:n] = c
            g[M - n + 1:] = c[1:][::-1]                                            # m = -(n-1) .. -1
            chirps.append(np.conj(c))
            filters.append(np.fft.fft(g))
        filters[-1] = filters[-1] / float(np.prod(self.pad))                       # the 1 / prod(M) of the inverse transform
        tab = np.concatenate(chirps + filters)
        tab = np.stack((tab.real, tab.imag), axis=-1).reshape(-1)
        self._tab = torch.as_tensor(tab, dtype=dtype, device=self.rt.device).contiguous()
        self._n = (ctypes.c_int64 * len(self.shape))(*self.shape)
        self._work = None

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        x = self.rt.asarray(x, self.dtype)
        if tuple(x.shape) != self.shape:
            raise ValueError(f"shape mismatch: {tuple(x.shape)} vs {self.shape}")
        if self._work is None:
            self._work = torch.empty(2 * int(np.prod(self.pad)), dtype=self.dtype, device=self.rt.device)
        out = torch.empty_like(x)
        self.rt.api.call("nb200_hartley_chirpz", self.plan._h, self.rt.stream(), self._n, self.rt.ptr(self._tab), self.rt.ptr(x),
                         self.rt.ptr(self._work), self.rt.ptr(out))
        return out


class BluesteinCorrelatedField(LazyModel):
    """The finalised single-sub-grid correlated field on a grid with non-power-of-two extents (non-parametric or Matern amplitude)."""

    def __init__(self, prefix, offset_mean, azm_prior, f, *, dtype, convention, runtime):
        self.prefix, self.offset_mean, self.dtype = prefix, float(offset_mean), dtype
        self._azm = _Prior(azm_prior)
        self.shape = tuple(int(s) for s in f["shape"])
        self._ht = BluesteinHartley(self.shape, dtype=dtype, convention=convention, runtime=runtime)
        self.rt, self.plan = self._ht.rt, self._ht.plan          # `plan`: the PADDED power-of-two plan behind the transform
        dev = self.rt.device
        tb = fourier_mode_tables(self.shape, f["distances"])
        tb["shape"], tb["distances"] = self.shape, tuple(float(x) for x in np.broadcast_to(f["distances"], (len(self.shape),)))
        self._grid = tb
        self._spec, leaves = _amplitude_spec(f, prefix + f["prefix"], tb["mode_lengths"].size)
        domain = {prefix + "zeromode": (), prefix + "xi": self.shape}
        domain.update(leaves)
        self.domain = dict(sorted(domain.items()))
        self.layout = Layout(self.domain)
        self.target_shape = self.shape
        self._tabs = _device_tables(tb, dtype, dev)
        self._vol = tb["total_volume"]
        self._factors = [(self._tabs["pd"], tuple(range(len(self.shape))))]              # one (bin table, axes) pair (outer._FieldLin)
        field = self

        class Transform(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                return field._ht(x.detach())

            @staticmethod
            def backward(ctx, g):
                return Transform.apply(g)

        self._transform = Transform

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

    def _raw_transform(self, x):
        return self._ht(x)

    def _factor_tables(self, small):
        z, na = self._normalized(small)
        return (z * na,)

    def _normalized(self, p):
        z = self._azm(p[self.prefix + "zeromode"])
        a = _eval_amplitude(self._spec, self._tabs, p)
        return z, torch.cat((a[:1], a[1:] / z))

    def amplitude(self, pos) -> torch.Tensor:
        """``CorrelatedFieldMaker.amplitude`` (:824-838): ``[azm V, amp_1, ..., amp_{K-1}]``."""
        p = self._tree(pos)
        z = self._azm(p[self.prefix + "zeromode"])
        a = _eval_amplitude(self._spec, self._tabs, p)
        return torch.cat((a[:1] * z, a[1:]))

    def power_spectrum(self, pos) -> torch.Tensor:
        """``CorrelatedFieldMaker.power_spectrum`` (:840-845): ``amplitude(p) ** 2``."""
        return self.amplitude(pos) ** 2

    @property
    def normalized_amplitudes(self):
        return ((lambda pos: self._normalized(self._tree(pos))[1]),)

    @property
    def target_grids(self):
        return (grid_record(self._grid),)

    def __call__(self, pos) -> torch.Tensor:
        """correlated_field.py:889-912 for one sub-grid; differentiable with respect to every leaf (torch autograd)."""
        p = self._tree(pos)
        z, na = self._normalized(p)
        h = z * na[self._tabs["pd"]] * p[self.prefix + "xi"]
        return self.offset_mean + self._transform.apply(h) / self._tabs["V"]
