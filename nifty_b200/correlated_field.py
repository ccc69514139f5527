"""``CorrelatedFieldMaker`` -- host-side mirror of ``nifty/re/correlated_field.py:519-920``.

Same builder protocol (``set_amplitude_total_offset``, ``add_fluctuations``, ``finalize``), same
leaf names (``<prefix>xi``, ``<prefix>zeromode``, ``<prefix><sub>fluctuations`` ...), same error
behaviour for malformed priors.  ``finalize()`` returns a :class:`CorrelatedField` whose forward /
JVP / VJP run in the sm_100a kernels of libniftyb200.so; nothing is evaluated in Python.

Scope: one Fourier sub-grid (1-3 axes, power-of-two extents), non-parametric (``kind`` "amplitude" or "power")
or Matern amplitude on the fused device path; several sub-grids (at most three axes in total) give the host-composed
outer-product model of ``outer.py``, extents that are not powers of two the chirp-z model of ``bluestein.py``.
``harmonic_type="spherical"`` raises ``NotImplementedError`` (out of scope).
"""

from __future__ import annotations

from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from ._capi import ModelDesc
from ._runtime import Lin, ModelHandle, Plan, Runtime, default_runtime
from .prior import LogNormalPrior, NormalPrior, _as_prior
from .tree import Layout


RegularCartesianGrid = namedtuple("RegularCartesianGrid", ("shape", "total_volume", "distances", "harmonic_grid"), defaults=(None,))
RegularFourierGrid = namedtuple("RegularFourierGrid", ("shape", "power_distributor", "mode_multiplicity", "mode_lengths",
                                                       "relative_log_mode_lengths", "log_volume"))


from .model import LazyModel  # noqa: E402

class CorrelatedField(LazyModel):
    """The finalised model: ``cf(pos)`` evaluates the field; mirrors ``jft.Model`` attributes."""

    def __init__(self, plan: Plan, domain: dict, prefix: str, sub_prefix: str, desc_fields: dict, offset_mean: float,
                 matern: bool = False):
        self.plan, self.rt = plan, plan.rt
        self.matern = matern
        self.prefix, self.sub_prefix = prefix, sub_prefix
        self.offset_mean = float(offset_mean)
        self._desc_fields = dict(desc_fields)
        self.domain = dict(sorted(domain.items()))
        self.layout = Layout(self.domain)
        self.target_shape = plan.local_pos_shape
        self.dtype = plan.dtype
        self._handle: Optional[ModelHandle] = None

    # -- jft.Model surface --------------------------------------------------------------------
    @property
    def target(self):
        return self.target_shape

    def init(self, seed):
        """Random latent position (dict of tensors), N(0,1) per leaf as ``Model.init`` does."""
        return self.layout.unpack(self.layout.random(seed, self.dtype, self.rt.device))

    def _descriptor(self, layout: Layout, scaling=None, scaling_key="scaling") -> ModelDesc:
        d = ModelDesc()
        for k, v in self._desc_fields.items():
            setattr(d, k, v)
        p, sp = self.prefix, self.prefix + self.sub_prefix
        off = lambda key: layout.offsets.get(key, -1)
        d.off_xi, d.off_zeromode = off(p + "xi"), off(p + "zeromode")
        d.off_fluct, d.off_slope = off(sp + "fluctuations"), off(sp + "loglogavgslope")
        d.off_cutoff = -1
        if self.matern:
            d.off_fluct, d.off_slope, d.off_cutoff = off(sp + "scale"), off(sp + "loglogslope"), off(sp + "cutoff")
        d.off_flex, d.off_asp, d.off_spectrum = off(sp + "flexibility"), off(sp + "asperity"), off(sp + "spectrum")
        d.has_scaling, d.off_scaling = 0, -1
        if scaling is not None:
            a, b = scaling.ab()
            d.has_scaling, d.scaling_a, d.scaling_b, d.off_scaling = 1, a, b, off(scaling_key)
        d.latent_size = layout.size
        d.offset_mean = self.offset_mean
        return d

    def handle(self) -> ModelHandle:
        if self._handle is None:
            self._handle = ModelHandle(self.plan, self._descriptor(self.layout))
        return self._handle

    def as_flat(self, pos) -> torch.Tensor:
        pos = getattr(pos, "tree", pos)          # jft.Vector
        if isinstance(pos, torch.Tensor):
            return self.rt.asarray(pos.reshape(-1), self.dtype)
        return self.layout.pack(pos, self.dtype, self.rt.device)

    def __call__(self, pos) -> torch.Tensor:
        return self.handle().cf_forward(self.as_flat(pos))

    # -- amplitude spectra (correlated_field.py:807-845) ----------------------------------------------------------
    def _azm_value(self, flat: torch.Tensor) -> float:
        a, b = self._desc_fields["zeromode_a"], self._desc_fields["zeromode_b"]
        return float(torch.exp(a + b * flat[self.layout.offsets[self.prefix + "zeromode"]].to(torch.float64)))

    def amplitude(self, pos) -> torch.Tensor:
        """``CorrelatedFieldMaker.amplitude`` (:824-838): the un-normalised amplitude with the zero mode scaled by the
        amplitude of the total offset, ``[azm V, amp_1, ..., amp_{K-1}]`` -- exactly the table the device kernels gather
        from (evaluated by the O(K) amplitude chain on the device)."""
        return self.handle().cf_amplitude(self.as_flat(pos))

    def power_spectrum(self, pos) -> torch.Tensor:
        """``CorrelatedFieldMaker.power_spectrum`` (:840-845): ``amplitude(p) ** 2``."""
        return self.amplitude(pos) ** 2

    @property
    def normalized_amplitudes(self):
        """``cf.normalized_amplitudes`` (:807-821, 918): one callable per sub-grid, ``amp(p).at[1:] / azm(p)``, ``[0] = V``."""
        def normed_amplitude(pos):
            flat = self.as_flat(pos)
            amp = self.amplitude(flat).clone()
            z = self._azm_value(flat)
            amp /= z
            return amp
        return (normed_amplitude,)

    @property
    def target_grids(self):
        """``cf.target_grids`` (:919): the position-space grid with its harmonic partner (tables built by the plan with the
        reference's arithmetic, :134-176, 228-265)."""
        return (_grid_record(self.plan),)


def _grid_record(pl: Plan):
    hg = RegularFourierGrid(shape=tuple(pl.shape), power_distributor=pl.power_distributor, mode_multiplicity=pl.mode_multiplicity,
                            mode_lengths=pl.mode_lengths, relative_log_mode_lengths=pl.relative_log_mode_lengths,
                            log_volume=pl.log_volume)
    return RegularCartesianGrid(shape=tuple(pl.shape), total_volume=pl.total_volume, distances=tuple(pl.distances), harmonic_grid=hg)


def get_fourier_mode_distributor(shape, distances):
    """``(power_distributor, unique mode lengths, mode multiplicity)`` of a regular grid (correlated_field.py:134-176), host NumPy
    like there, any extents.  (The plans of the fused path build the same tables in C++ on the folded index range.)"""
    from .bluestein import fourier_mode_tables
    shape = (int(shape),) if np.ndim(shape) == 0 else tuple(int(s) for s in shape)
    tb = fourier_mode_tables(shape, distances)
    return tb["power_distributor"], tb["mode_lengths"], tb["mode_multiplicity"]


def make_grid(shape, distances, harmonic_type: str = "fourier"):
    """``RegularCartesianGrid`` with its harmonic partner (correlated_field.py:238-265)."""
    if harmonic_type.lower() != "fourier":
        if harmonic_type.lower() == "spherical":
            raise NotImplementedError("harmonic_type='spherical' is outside the B200 hot path")
        raise ValueError(f"invalid `harmonic_type` {harmonic_type!r}")
    from .bluestein import fourier_mode_tables, grid_record
    shape = (int(shape),) if np.ndim(shape) == 0 else tuple(int(s) for s in shape)
    tb = fourier_mode_tables(shape, distances)
    tb["shape"], tb["distances"] = shape, tuple(float(x) for x in np.broadcast_to(distances, (len(shape),)))
    return grid_record(tb)


_HARTLEY_PLANS = {}


def hartley(p, axes=None, *, hartley_convention: str = "non_canonical_hartley", runtime: Optional[Runtime] = None) -> torch.Tensor:
    """``hartley(p, axes)`` of the reference (correlated_field.py:24-30): ``Re(fftn(p, axes)) +/- Im(fftn(p, axes))``, unnormalised,
    on the device.  ``axes=None``: all axes; otherwise the transform runs over the named axes (at most three) for every index of
    the others, one device call per slice.  Power-of-two extents use the fused passes (``nb200_hartley``), other extents the
    chirp-z call (``nb200_hartley_chirpz``); plans are cached per (shape, dtype, convention, runtime)."""
    from ._runtime import default_runtime
    rt = runtime if runtime is not None else default_runtime()
    x = p if isinstance(p, torch.Tensor) else torch.as_tensor(np.asarray(p))
    dtype = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float64
    x = rt.asarray(x, dtype)
    nd = x.ndim
    axes = tuple(range(nd)) if axes is None else tuple(sorted(int(a) % nd for a in ((axes,) if np.ndim(axes) == 0 else axes)))
    if len(set(axes)) != len(axes) or not 1 <= len(axes) <= 3:
        raise NotImplementedError("hartley: between one and three distinct axes")
    sub = tuple(int(x.shape[a]) for a in axes)
    key = (sub, dtype, hartley_convention, id(rt))
    fn = _HARTLEY_PLANS.get(key)
    if fn is None:
        if all(n >= 2 and not (n & (n - 1)) for n in sub):
            fn = Plan(sub, 1.0, dtype=dtype, hartley_convention=hartley_convention, runtime=rt).hartley
        else:
            from .bluestein import BluesteinHartley
            fn = BluesteinHartley(sub, dtype=dtype, convention=hartley_convention, runtime=rt)
        _HARTLEY_PLANS[key] = fn
    rest = [a for a in range(nd) if a not in axes]
    perm = rest + list(axes)
    xp = x.permute(perm).contiguous()
    flat = xp.reshape((-1,) + sub)
    out = torch.stack([fn(flat[b]) for b in range(flat.shape[0])]).reshape(xp.shape)
    return out.permute([perm.index(a) for a in range(nd)]).contiguous()


class CorrelatedFieldMaker:
    """Builder with the call protocol of ``jft.CorrelatedFieldMaker`` (correlated_field.py:519-920)."""

    def __init__(self, prefix: str, *, dtype=torch.float64, hartley_convention="non_canonical_hartley",
                 runtime: Optional[Runtime] = None, comm=None):
        """``comm``: torch.distributed group over which a 3-D grid is slab-decomposed (one process per GPU)."""
        self._prefix = prefix
        self._dtype, self._conv, self._rt, self._comm = dtype, hartley_convention, runtime, comm
        self._offset_mean = None
        self._azm = None
        self._fluct = []
        self._parameter_tree = {}

    def set_amplitude_total_offset(self, offset_mean, offset_std):
        """correlated_field.py:583-659; ``offset_std`` is a ``(mean, std)`` log-normal prior."""
        if offset_std is None or not isinstance(offset_std, (tuple, list, LogNormalPrior)):
            raise TypeError(f"`offset_std` of invalid type {type(offset_std)!r}")
        self._azm = _as_prior(offset_std, LogNormalPrior, "offset_std")
        self._offset_mean = float(offset_mean)
        self._parameter_tree[self._prefix + "zeromode"] = ()

    def add_fluctuations(self, shape, distances, fluctuations, loglogavgslope, flexibility=None, asperity=None,
                         prefix: str = "", harmonic_type: str = "fourier", non_parametric_kind: str = "amplitude"):
        """correlated_field.py:661-755."""
        if harmonic_type.lower() != "fourier":
            if harmonic_type.lower() == "spherical":
                raise NotImplementedError("harmonic_type='spherical' is outside the B200 hot path")
            raise ValueError(f"invalid `harmonic_type` {harmonic_type!r}")
        kind = non_parametric_kind.lower()
        if kind not in ("amplitude", "power"):
            raise ValueError(f"invalid `non_parametric_kind` {non_parametric_kind!r}")
        if self._fluct and self._comm is not None:
            raise NotImplementedError("outer products of several sub-grids are not available on slab-decomposed fields")
        shape = (int(shape),) if np.ndim(shape) == 0 else tuple(int(s) for s in shape)
        flu = _as_prior(fluctuations, LogNormalPrior, "fluctuations", optional=True)
        slp = _as_prior(loglogavgslope, NormalPrior, "loglogavgslope")
        flx = _as_prior(flexibility, LogNormalPrior, "flexibility", optional=True)
        asp = _as_prior(asperity, LogNormalPrior, "asperity", optional=True)
        self._fluct.append(dict(shape=shape, distances=distances, flu=flu, slp=slp, flx=flx, asp=asp, prefix=prefix, kind=kind))

    def add_fluctuations_matern(self, shape, distances, scale, cutoff, loglogslope, renormalize_amplitude: bool,
                                prefix: str = "", harmonic_type: str = "fourier", non_parametric_kind: str = "amplitude"):
        """correlated_field.py:661-755 (``MaternAmplitude`` :302-395): leaves ``<prefix>scale`` (log-normal),
        ``<prefix>cutoff`` (log-normal), ``<prefix>loglogslope`` (normal)."""
        if harmonic_type.lower() != "fourier":
            if harmonic_type.lower() == "spherical":
                raise NotImplementedError("harmonic_type='spherical' is outside the B200 hot path")
            raise ValueError(f"invalid `harmonic_type` {harmonic_type!r}")
        kind = non_parametric_kind.lower()
        if kind not in ("amplitude", "power"):
            raise ValueError(f"invalid `non_parametric_kind` {non_parametric_kind!r}")
        if self._fluct and self._comm is not None:
            raise NotImplementedError("outer products of several sub-grids are not available on slab-decomposed fields")
        shape = (int(shape),) if np.ndim(shape) == 0 else tuple(int(s) for s in shape)
        scl = _as_prior(scale, LogNormalPrior, "scale")
        ctf = _as_prior(cutoff, LogNormalPrior, "cutoff")
        slp = _as_prior(loglogslope, NormalPrior, "loglogslope")
        self._fluct.append(dict(shape=shape, distances=distances, matern=True, scl=scl, ctf=ctf, slp=slp,
                                renorm=bool(renormalize_amplitude), prefix=prefix, kind=kind))

    # -- amplitude accessors of the maker (correlated_field.py:800-845); available after finalize() ----------------
    def _finalized(self) -> "CorrelatedField":
        if getattr(self, "_cf", None) is None:
            raise ValueError("call finalize() first: the amplitude chain lives in the finalised device model")
        return self._cf

    @property
    def amplitude(self):
        return self._finalized().amplitude

    @property
    def power_spectrum(self):
        return self._finalized().power_spectrum

    def get_normalized_amplitudes(self):
        return self._finalized().normalized_amplitudes

    @property
    def amplitude_total_offset(self):
        """correlated_field.py:785-791: the log-normal amplitude of the total offset as a callable of the latent position."""
        if self._azm is None:
            raise NotImplementedError("You need to set the `amplitude_total_offset` first")
        a, b = self._azm.ab()
        key = self._prefix + "zeromode"

        def azm(pos):
            pos = getattr(pos, "tree", pos)
            if isinstance(pos, torch.Tensor):
                pos = self._finalized().layout.unpack(pos)
            return torch.exp(a + b * torch.as_tensor(pos[key], dtype=torch.float64).reshape(()))
        return azm

    @property
    def azm(self):
        """Alias for `amplitude_total_offset` (:793-796)."""
        return self.amplitude_total_offset

    @property
    def fluctuations(self):
        """correlated_field.py:798-805: the un-normalised amplitudes, one callable per sub-grid (``[0] = V``); evaluated by the
        finalised model (its amplitude chain lives on the device), so available after ``finalize()``."""
        nas, azm = self._finalized().normalized_amplitudes, self.amplitude_total_offset

        def make(na):
            def fluct(pos):
                a = na(pos)
                z = azm(pos).to(device=a.device, dtype=a.dtype)
                return torch.cat((a[:1], a[1:] * z))
            return fluct
        return tuple(make(na) for na in nas)

    def finalize(self) -> CorrelatedField:
        self._cf = self._finalize()
        return self._cf

    def _finalize(self) -> CorrelatedField:
        """correlated_field.py:850-920: builds the grid tables (on the C side) and the model."""
        if self._azm is None:
            raise ValueError("set_amplitude_total_offset must be called before finalize")
        if not self._fluct:
            raise ValueError("add_fluctuations must be called before finalize")
        if len(self._fluct) > 1:
            # outer product of sub-grids (correlated_field.py:856-912): host-composed model around the joint device transform
            from .outer import OuterCorrelatedField
            return OuterCorrelatedField(self._prefix, self._offset_mean, self._azm, self._fluct, dtype=self._dtype,
                                        convention=self._conv, runtime=self._rt)
        f = self._fluct[0]
        if any(n < 2 or (n & (n - 1)) for n in f["shape"]):
            # extents that are not powers of two: host-composed model around power-of-two device transforms (Bluestein)
            if self._comm is not None:
                raise NotImplementedError("slab-decomposed fields need power-of-two extents")
            from .bluestein import BluesteinCorrelatedField
            return BluesteinCorrelatedField(self._prefix, self._offset_mean, self._azm, f, dtype=self._dtype, convention=self._conv,
                                            runtime=self._rt)
        plan = Plan(f["shape"], f["distances"], dtype=self._dtype, hartley_convention=self._conv, runtime=self._rt,
                    comm=self._comm)
        sp = self._prefix + f["prefix"]
        if f.get("matern"):
            domain = dict(self._parameter_tree)
            domain[sp + "scale"] = ()
            domain[sp + "cutoff"] = ()
            domain[sp + "loglogslope"] = ()
            domain[self._prefix + "xi"] = plan.local_shape
            fields = dict(kind_power=int(f["kind"] == "power"), has_fluctuations=1, has_deviations=0, has_asperity=0,
                          amplitude_type=1, renormalize_amplitude=int(f["renorm"]))
            fields["zeromode_a"], fields["zeromode_b"] = self._azm.ab()
            fields["fluct_a"], fields["fluct_b"] = f["scl"].ab()
            fields["slope_a"], fields["slope_b"] = f["slp"].ab()
            fields["cutoff_a"], fields["cutoff_b"] = f["ctf"].ab()
            return CorrelatedField(plan, domain, self._prefix, f["prefix"], fields, self._offset_mean, matern=True)
        has_dev = f["flx"] is not None and plan.K > 2
        domain = dict(self._parameter_tree)
        if f["flu"] is not None:
            domain[sp + "fluctuations"] = ()
        domain[sp + "loglogavgslope"] = ()
        if has_dev:
            domain[sp + "flexibility"] = ()
            if f["asp"] is not None:
                domain[sp + "asperity"] = ()
            domain[sp + "spectrum"] = (plan.K - 2, 2)
        domain[self._prefix + "xi"] = plan.local_shape     # the rows this rank owns when slab-decomposed
        fields = dict(kind_power=int(f["kind"] == "power"), has_fluctuations=int(f["flu"] is not None),
                      has_deviations=int(has_dev), has_asperity=int(has_dev and f["asp"] is not None))
        fields["zeromode_a"], fields["zeromode_b"] = self._azm.ab()
        if f["flu"] is not None:
            fields["fluct_a"], fields["fluct_b"] = f["flu"].ab()
        fields["slope_a"], fields["slope_b"] = f["slp"].ab()
        if has_dev:
            fields["flex_a"], fields["flex_b"] = f["flx"].ab()
            if f["asp"] is not None:
                fields["asp_a"], fields["asp_b"] = f["asp"].ab()
        return CorrelatedField(plan, domain, self._prefix, f["prefix"], fields, self._offset_mean)
