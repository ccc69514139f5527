"""Oracle (TEST INFRASTRUCTURE ONLY): correlated-field forward model, JVP and VJP.

NumPy float64 restatement of ``nifty/re/correlated_field.py`` (reference tree
``/root/reference``); derivatives are hand-derived (the reference obtains
them from JAX AD) and are pinned against finite differences, adjointness and
``nifty.cl``'s ``Linearization`` in ``tests/``.

Parameter trees are plain ``dict[str, np.ndarray]`` with the reference's key
names (``<prefix>xi``, ``<prefix>zeromode``, ``<prefix><sub>fluctuations`` ...).
"""

from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np
import scipy.fft

_NTHREADS = 1


def set_nthreads(n: int) -> None:
    """Worker threads used by scipy.fft inside the oracle (CPU baseline only)."""
    global _NTHREADS
    _NTHREADS = int(n)


def hartley(p, axes=None, convention="non_canonical_hartley"):
    """Unnormalised Hartley transform as ``Re(fftn) +/- Im(fftn)``.

    Follows nifty/re/correlated_field.py:24-30 (sign chosen by
    nifty/config.py:44 ``hartley_convention``).
    """
    f = scipy.fft.fftn(np.asarray(p), axes=axes, workers=_NTHREADS)
    if convention == "non_canonical_hartley":
        return f.real + f.imag
    if convention == "canonical_hartley":
        return f.real - f.imag
    raise ValueError(f"invalid hartley convention {convention!r}")


def fourier_mode_distributor(shape, distances, uniqueness_rtol=1e-12):
    """Mode-length binning: (bin index per mode, unique lengths, multiplicity).

    Follows nifty/re/correlated_field.py:134-176 (mode lengths) and :55-67
    (unique-with-tolerance, mid-point binning, empty-bin check).
    """
    shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
    distances = np.broadcast_to(np.asarray(distances, dtype=np.float64), (len(shape),))
    step = 1.0 / (np.array(shape) * distances)
    ax0 = np.arange(shape[0])
    length = np.minimum(ax0, shape[0] - ax0) * step[0]
    if len(shape) > 1:
        length = length * length
        for i in range(1, len(shape)):
            axi = np.arange(shape[i])
            li = np.minimum(axi, shape[i] - axi) * step[i]
            length = length[..., None] + li * li
        length = np.sqrt(length)
    um = np.unique(length)
    tol = uniqueness_rtol * um[-1]
    keep = np.diff(np.append(um, 2 * um[-1])) > tol
    um = um[keep]
    bounds = 0.5 * (um[:-1] + um[1:])
    idx = np.searchsorted(bounds, length)
    count = np.bincount(idx.ravel(), minlength=um.size)
    if np.any(count == 0) or um.shape != count.shape:
        raise RuntimeError("invalid harmonic mode(s) encountered")
    return idx, um, count


@dataclasses.dataclass
class FourierGrid:
    """Static tables of one sub-grid (nifty/re/correlated_field.py:179-200, 238-265)."""

    shape: tuple
    distances: tuple
    total_volume: float
    power_distributor: np.ndarray
    mode_multiplicity: np.ndarray
    mode_lengths: np.ndarray
    relative_log_mode_lengths: np.ndarray
    log_volume: np.ndarray


def make_fourier_grid(shape, distances) -> FourierGrid:
    """nifty/re/correlated_field.py:238-265 with ``_log_modes`` :228-235."""
    shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
    distances = tuple(float(d) for d in np.broadcast_to(distances, (len(shape),)))
    totvol = float(np.prod(np.array(shape) * np.array(distances)))
    idx, um, count = fourier_mode_distributor(shape, distances)
    rel = um.copy()
    rel[1:] = np.log(rel[1:])
    rel[1:] -= rel[1]
    assert rel[0] == 0.0
    log_vol = rel[2:] - rel[1:-1]
    return FourierGrid(shape, distances, totvol, idx, count, um, rel, log_vol)


def lognormal_moments(mean, std):
    """nifty/re/num/stats_distributions.py:62-73."""
    if mean <= 0.0:
        raise ValueError(f"`mean` must be greater zero; got {mean!r}")
    if std <= 0.0:
        raise ValueError(f"`std` must be greater zero; got {std!r}")
    logstd = np.sqrt(np.log1p((std / mean) ** 2))
    logmean = np.log(mean) - 0.5 * logstd**2
    return float(logmean), float(logstd)


class _Prior:
    """Standard-normal -> (log)normal reparametrisation of one scalar leaf.

    nifty/re/num/stats_distributions.py:42-98 (``normal_prior`` / ``lognormal_prior``).
    """

    def __init__(self, kind, mean, std):
        self.kind = kind
        self.mean, self.std = float(mean), float(std)
        if kind == "lognormal":
            self.a, self.b = lognormal_moments(mean, std)
        elif kind == "normal":
            self.a, self.b = float(mean), float(std)
        else:
            raise ValueError(kind)

    def __call__(self, xi):
        v = self.a + self.b * np.asarray(xi, dtype=np.float64)
        return np.exp(v) if self.kind == "lognormal" else v

    def deriv(self, xi):
        """d value / d xi."""
        if self.kind == "lognormal":
            return self.b * self(xi)
        return self.b * np.ones_like(np.asarray(xi, dtype=np.float64))


def _as_prior(spec, kind, name, optional=False):
    if spec is None:
        if optional:
            return None
        raise TypeError(f"invalid `{name}` specified; got '{type(spec)}'")
    if isinstance(spec, (tuple, list)):
        if len(spec) != 2:
            raise TypeError(f"invalid `{name}` specified; got {spec!r}")
        return _Prior(kind, *spec)
    raise TypeError(f"invalid `{name}` specified; got '{type(spec)}'")


class NonParametricAmplitudeOracle:
    """nifty/re/correlated_field.py:398-516 + gauss_markov.py:102-114 (IWP)."""

    def __init__(self, grid: FourierGrid, fluctuations, loglogavgslope, flexibility=None,
                 asperity=None, prefix="", kind="amplitude"):
        self.grid = grid
        self.kind = kind.lower()
        if self.kind not in ("amplitude", "power"):
            raise ValueError(f"Invalid kind specified {self.kind!r}")
        self.prefix = prefix
        self.flu = _as_prior(fluctuations, "lognormal", "fluctuations", optional=True)
        self.slp = _as_prior(loglogavgslope, "normal", "loglogavgslope")
        self.flx = _as_prior(flexibility, "lognormal", "flexibility", optional=True)
        self.asp = _as_prior(asperity, "lognormal", "asperity", optional=True)
        self.has_dev = self.flx is not None and grid.log_volume.size > 0
        if not self.has_dev:
            self.flx = None
            self.asp = None
        self.domain = {}
        if self.flu is not None:
            self.domain[prefix + "fluctuations"] = ()
        self.domain[prefix + "loglogavgslope"] = ()
        if self.has_dev:
            self.domain[prefix + "flexibility"] = ()
            if self.asp is not None:
                self.domain[prefix + "asperity"] = ()
            self.domain[prefix + "spectrum"] = (grid.log_volume.size, 2)

    # -- forward with all intermediates -------------------------------------------------
    def _forward(self, p):
        g, pf = self.grid, self.prefix
        ell = g.relative_log_mode_lengths
        mult = g.mode_multiplicity.astype(np.float64)
        V = g.total_volume
        st = {}
        st["flu"] = 1.0 if self.flu is None else float(self.flu(p[pf + "fluctuations"]))
        st["slope"] = float(self.slp(p[pf + "loglogavgslope"]))
        u = st["slope"] * ell
        if self.has_dev:
            dt = g.log_volume
            xi = np.asarray(p[pf + "spectrum"], dtype=np.float64)
            sig = float(self.flx(p[pf + "flexibility"]))
            asp = 0.0 if self.asp is None else float(self.asp(p[pf + "asperity"]))
            sd = sig * np.sqrt(dt)
            q = np.sqrt(dt**2 / 12.0 + asp)
            r1 = sd * xi[:, 1]
            r0 = sd * xi[:, 0] * q + 0.5 * dt * r1
            y = np.concatenate(([0.0], np.cumsum(r1)))
            x0 = np.concatenate(([0.0], r0 + dt * y[:-1]))
            x = np.cumsum(x0)
            tw = np.concatenate(([0.0], x))
            dev = tw - tw[-1] * (ell / ell[-1])
            u = u + dev
            st.update(sig=sig, asp=asp, sd=sd, q=q, xi=xi, dt=dt)
        P = np.exp(u)
        if self.kind == "power":
            S = float(np.sum(mult[1:] * P[1:]))
            shape_fn = np.sqrt(P)
        else:
            S = float(np.sum(mult[1:] * P[1:] ** 2))
            shape_fn = P
        norm = np.sqrt(S) / np.sqrt(V)
        amp = st["flu"] * (np.sqrt(V) / norm) * shape_fn
        amp[0] = V
        st.update(P=P, S=S, amp=amp, ell=ell, mult=mult)
        return amp, st

    def __call__(self, p):
        return self._forward(p)[0]

    def jvp(self, p, dp):
        amp, st = self._forward(p)
        pf, ell, mult, P, S = self.prefix, st["ell"], st["mult"], st["P"], st["S"]
        dflu_rel = 0.0
        if self.flu is not None:
            dflu_rel = float(self.flu.deriv(p[pf + "fluctuations"]) * dp[pf + "fluctuations"]) / st["flu"]
        dslope = float(self.slp.deriv(p[pf + "loglogavgslope"]) * dp[pf + "loglogavgslope"])
        du = dslope * ell
        if self.has_dev:
            dt, sd, q, xi = st["dt"], st["sd"], st["q"], st["xi"]
            dxi = np.asarray(dp[pf + "spectrum"], dtype=np.float64)
            dsig = float(self.flx.deriv(p[pf + "flexibility"]) * dp[pf + "flexibility"])
            dasp = 0.0
            if self.asp is not None:
                dasp = float(self.asp.deriv(p[pf + "asperity"]) * dp[pf + "asperity"])
            dsd = dsig * np.sqrt(dt)
            dq = dasp / (2.0 * q)
            dr1 = dsd * xi[:, 1] + sd * dxi[:, 1]
            dr0 = dsd * xi[:, 0] * q + sd * dxi[:, 0] * q + sd * xi[:, 0] * dq + 0.5 * dt * dr1
            dy = np.concatenate(([0.0], np.cumsum(dr1)))
            dx0 = np.concatenate(([0.0], dr0 + dt * dy[:-1]))
            dtw = np.concatenate(([0.0], np.cumsum(dx0)))
            du = du + dtw - dtw[-1] * (ell / ell[-1])
        if self.kind == "power":
            dS = float(np.sum(mult[1:] * P[1:] * du[1:]))
            damp = amp * (dflu_rel + 0.5 * du - 0.5 * dS / S)
        else:
            dS = float(np.sum(2.0 * mult[1:] * P[1:] ** 2 * du[1:]))
            damp = amp * (dflu_rel + du - 0.5 * dS / S)
        damp[0] = 0.0
        return amp, damp

    def vjp(self, p, abar):
        """Cotangent ``abar`` (K,) on the amplitude -> dict of parameter cotangents."""
        amp, st = self._forward(p)
        pf, ell, mult, P, S = self.prefix, st["ell"], st["mult"], st["P"], st["S"]
        g = np.asarray(abar, dtype=np.float64) * amp
        g[0] = 0.0
        out = {}
        if self.flu is not None:
            out[pf + "fluctuations"] = np.asarray(
                np.sum(g) / st["flu"] * self.flu.deriv(p[pf + "fluctuations"]))
        Sbar = -0.5 * np.sum(g) / S
        if self.kind == "power":
            ubar = 0.5 * g
            ubar[1:] += mult[1:] * P[1:] * Sbar
        else:
            ubar = g.copy()
            ubar[1:] += 2.0 * mult[1:] * P[1:] ** 2 * Sbar
        out[pf + "loglogavgslope"] = np.asarray(
            np.sum(ubar * ell) * self.slp.deriv(p[pf + "loglogavgslope"]))
        if self.has_dev:
            dt, sd, q, xi = st["dt"], st["sd"], st["q"], st["xi"]
            twbar = ubar.copy()
            twbar[-1] -= np.sum(ubar * ell) / ell[-1]
            xbar = twbar[1:]  # (K-1)
            x0bar = np.cumsum(xbar[::-1])[::-1]
            r0bar = x0bar[1:].copy()
            ybar = np.concatenate((dt * x0bar[1:], [0.0]))  # (K-1)
            # y_i = sum_{j<i} r1_j  ->  r1bar_j = sum_{i>j} ybar_i
            suffix = np.cumsum(ybar[::-1])[::-1]
            r1bar = suffix[1:] + 0.5 * dt * r0bar
            sdbar = r0bar * xi[:, 0] * q + r1bar * xi[:, 1]
            spec = np.empty_like(xi)
            spec[:, 0] = r0bar * sd * q
            spec[:, 1] = r1bar * sd
            out[pf + "spectrum"] = spec
            sigbar = np.sum(sdbar * np.sqrt(dt))
            out[pf + "flexibility"] = np.asarray(sigbar * self.flx.deriv(p[pf + "flexibility"]))
            if self.asp is not None:
                aspbar = np.sum(r0bar * sd * xi[:, 0] / (2.0 * q))
                out[pf + "asperity"] = np.asarray(aspbar * self.asp.deriv(p[pf + "asperity"]))
        return out


class MaternAmplitudeOracle:
    """nifty/re/correlated_field.py:302-395."""

    def __init__(self, grid: FourierGrid, scale, cutoff, loglogslope, renormalize_amplitude,
                 prefix="", kind="amplitude"):
        self.grid = grid
        self.kind = kind.lower()
        if self.kind not in ("amplitude", "power"):
            raise ValueError(f"Invalid kind specified {self.kind!r}")
        self.prefix = prefix
        self.scale = _as_prior(scale, "lognormal", "scale")
        self.cutoff = _as_prior(cutoff, "lognormal", "cutoff")
        self.slope = _as_prior(loglogslope, "normal", "loglogslope")
        self.renorm = bool(renormalize_amplitude)
        self.domain = {prefix + "scale": (), prefix + "cutoff": (), prefix + "loglogslope": ()}

    def _forward(self, p):
        g, pf = self.grid, self.prefix
        k = g.mode_lengths
        mult = g.mode_multiplicity.astype(np.float64)
        V = g.total_volume
        scl = float(self.scale(p[pf + "scale"]))
        ctf = float(self.cutoff(p[pf + "cutoff"]))
        slp = float(self.slope(p[pf + "loglogslope"]))
        lg = np.log1p((k / ctf) ** 2)
        u = 0.25 * slp * lg
        P = np.exp(u)
        S = None
        norm = 1.0
        if self.renorm:
            S = float(np.sum(mult[1:] * (P[1:] ** 2 if self.kind == "amplitude" else P[1:])))
            norm = np.sqrt(S) / np.sqrt(V)
        shape_fn = np.sqrt(P) if self.kind == "power" else P
        amp = scl * (np.sqrt(V) / norm) * shape_fn
        amp[0] = V
        return amp, dict(scl=scl, ctf=ctf, slp=slp, lg=lg, P=P, S=S, k=k, mult=mult, amp=amp)

    def __call__(self, p):
        return self._forward(p)[0]

    def _du(self, st):
        k, ctf = st["k"], st["ctf"]
        du_dslp = 0.25 * st["lg"]
        # d/dctf log1p((k/ctf)^2) = -2 k^2 / ctf^3 / (1 + (k/ctf)^2)
        du_dctf = 0.25 * st["slp"] * (-2.0 * k**2 / ctf**3) / (1.0 + (k / ctf) ** 2)
        return du_dslp, du_dctf

    def jvp(self, p, dp):
        amp, st = self._forward(p)
        pf = self.prefix
        dscl_rel = float(self.scale.deriv(p[pf + "scale"]) * dp[pf + "scale"]) / st["scl"]
        dctf = float(self.cutoff.deriv(p[pf + "cutoff"]) * dp[pf + "cutoff"])
        dslp = float(self.slope.deriv(p[pf + "loglogslope"]) * dp[pf + "loglogslope"])
        a, b = self._du(st)
        du = a * dslp + b * dctf
        fac = 0.5 if self.kind == "power" else 1.0
        damp = amp * (dscl_rel + fac * du)
        if self.renorm:
            P, mult = st["P"], st["mult"]
            if self.kind == "power":
                dS = np.sum(mult[1:] * P[1:] * du[1:])
            else:
                dS = np.sum(2.0 * mult[1:] * P[1:] ** 2 * du[1:])
            damp -= amp * 0.5 * dS / st["S"]
        damp[0] = 0.0
        return amp, damp

    def vjp(self, p, abar):
        amp, st = self._forward(p)
        pf = self.prefix
        g = np.asarray(abar, dtype=np.float64) * amp
        g[0] = 0.0
        fac = 0.5 if self.kind == "power" else 1.0
        ubar = fac * g
        if self.renorm:
            P, mult = st["P"], st["mult"]
            Sbar = -0.5 * np.sum(g) / st["S"]
            if self.kind == "power":
                ubar[1:] += mult[1:] * P[1:] * Sbar
            else:
                ubar[1:] += 2.0 * mult[1:] * P[1:] ** 2 * Sbar
        a, b = self._du(st)
        return {
            pf + "scale": np.asarray(np.sum(g) / st["scl"] * self.scale.deriv(p[pf + "scale"])),
            pf + "cutoff": np.asarray(np.sum(ubar * b) * self.cutoff.deriv(p[pf + "cutoff"])),
            pf + "loglogslope": np.asarray(np.sum(ubar * a) * self.slope.deriv(p[pf + "loglogslope"])),
        }


class CorrelatedFieldOracle:
    """nifty/re/correlated_field.py:519-920 (``CorrelatedFieldMaker`` + ``finalize``).

    ``field = offset_mean + prod_i (1/V_i) hartley_i( azm * outer_i(namp_i[idx_i]) * xi )``
    """

    def __init__(self, prefix: str, hartley_convention="non_canonical_hartley"):
        self.prefix = prefix
        self.convention = hartley_convention
        self.offset_mean = None
        self.azm: Optional[_Prior] = None
        self.amps = []
        self.grids = []
        self.domain = {}
        self._final = False

    def set_amplitude_total_offset(self, offset_mean, offset_std):
        if offset_std is None or not isinstance(offset_std, (tuple, list)) or len(offset_std) != 2:
            raise TypeError(f"`offset_std` of invalid type {type(offset_std)!r}")
        self.offset_mean = offset_mean
        self.azm = _Prior("lognormal", *offset_std)
        self.domain[self.prefix + "zeromode"] = ()

    def add_fluctuations(self, shape, distances, fluctuations, loglogavgslope, flexibility=None,
                         asperity=None, prefix="", harmonic_type="fourier",
                         non_parametric_kind="amplitude"):
        if harmonic_type.lower() != "fourier":
            raise ValueError(f"invalid `harmonic_type` {harmonic_type!r}")
        grid = make_fourier_grid(shape, distances)
        amp = NonParametricAmplitudeOracle(grid, fluctuations, loglogavgslope, flexibility,
                                           asperity, prefix=self.prefix + prefix,
                                           kind=non_parametric_kind)
        self.amps.append(amp)
        self.grids.append(grid)
        self.domain.update(amp.domain)

    def add_fluctuations_matern(self, shape, distances, scale, cutoff, loglogslope,
                                renormalize_amplitude, prefix="", harmonic_type="fourier",
                                non_parametric_kind="amplitude"):
        if harmonic_type.lower() != "fourier":
            raise ValueError(f"invalid `harmonic_type` {harmonic_type!r}")
        grid = make_fourier_grid(shape, distances)
        amp = MaternAmplitudeOracle(grid, scale, cutoff, loglogslope, renormalize_amplitude,
                                    prefix=self.prefix + prefix, kind=non_parametric_kind)
        self.amps.append(amp)
        self.grids.append(grid)
        self.domain.update(amp.domain)

    def finalize(self):
        shape = ()
        self._axes = []
        for g in self.grids:
            n0 = len(shape)
            shape += g.shape
            self._axes.append(tuple(range(n0, len(shape))))
        self.shape = shape
        self.domain[self.prefix + "xi"] = shape
        self.domain = dict(sorted(self.domain.items()))
        self._final = True
        return self

    # -- helpers -------------------------------------------------------------------------
    def _bcast(self, i, a):
        """Broadcast a sub-grid shaped array to the full excitation shape."""
        shp = [1] * len(self.shape)
        for ax in self._axes[i]:
            shp[ax] = self.shape[ax]
        return np.reshape(a, shp)

    def _transform(self, x):
        out = x
        for g, axes in zip(self.grids, self._axes):
            out = (1.0 / g.total_volume) * hartley(out, axes=axes, convention=self.convention)
        return out

    def _expanded(self, namps):
        e = [self._bcast(i, na[g.power_distributor]) for i, (na, g) in enumerate(zip(namps, self.grids))]
        return e

    def normalized_amplitudes(self, p):
        z = float(self.azm(p[self.prefix + "zeromode"]))
        res = []
        for amp in self.amps:
            a = amp(p).copy()
            a[1:] *= 1.0 / z
            res.append(a)
        return res

    def amplitude(self, p):
        """Amplitude with zero mode (single sub-grid only), correlated_field.py:822-838."""
        if len(self.amps) != 1:
            raise NotImplementedError
        a = self.amps[0](p).copy()
        a[0] *= float(self.azm(p[self.prefix + "zeromode"]))
        return a

    def __call__(self, p):
        # memo of the last evaluation point (same leaf objects): metric = forward + JVP + VJP would
        # otherwise pay the forward transform three times (nifty.re pays it once per call)
        key = tuple((k, id(v)) for k, v in sorted(p.items()))
        memo = getattr(self, "_memo", None)
        if memo is not None and memo[0] == key:
            return memo[2]
        z = float(self.azm(p[self.prefix + "zeromode"]))
        e = self._expanded(self.normalized_amplitudes(p))
        ea = e[0]
        for x in e[1:]:
            ea = ea * x
        h = z * ea * np.asarray(p[self.prefix + "xi"], dtype=np.float64)
        out = self.offset_mean + self._transform(h)
        self._memo = (key, [v for _, v in sorted(p.items())], out)   # hold the leaves so ids stay unique
        return out

    def jvp(self, p, dp):
        pf = self.prefix
        z = float(self.azm(p[pf + "zeromode"]))
        dz = float(self.azm.deriv(p[pf + "zeromode"]) * dp[pf + "zeromode"])
        namps, dnamps = [], []
        for amp in self.amps:
            a, da = amp.jvp(p, dp)
            na = a.copy()
            na[1:] /= z
            dna = np.zeros_like(a)
            dna[1:] = da[1:] / z - a[1:] * dz / z**2
            namps.append(na)
            dnamps.append(dna)
        e = self._expanded(namps)
        de = self._expanded(dnamps)
        xi = np.asarray(p[pf + "xi"], dtype=np.float64)
        dxi = np.asarray(dp[pf + "xi"], dtype=np.float64)
        prod = np.ones(())
        for x in e:
            prod = prod * x
        dprod = np.zeros(())
        for i in range(len(e)):
            term = de[i]
            for j in range(len(e)):
                if j != i:
                    term = term * e[j]
            dprod = dprod + term
        dh = dz * prod * xi + z * dprod * xi + z * prod * dxi
        return self._transform(dh)

    def vjp(self, p, c):
        pf = self.prefix
        z = float(self.azm(p[pf + "zeromode"]))
        amps = [amp(p) for amp in self.amps]
        namps = []
        for a in amps:
            na = a.copy()
            na[1:] /= z
            namps.append(na)
        e = self._expanded(namps)
        prod = np.ones(())
        for x in e:
            prod = prod * x
        xi = np.asarray(p[pf + "xi"], dtype=np.float64)
        g = self._transform(np.asarray(c, dtype=np.float64))  # transform is symmetric
        out = {pf + "xi": z * prod * g}
        G = g * xi
        zbar = float(np.sum(prod * G))
        for i, (amp, grid) in enumerate(zip(self.amps, self.grids)):
            other = np.ones(())
            for j in range(len(e)):
                if j != i:
                    other = other * e[j]
            ebar = z * other * G
            other_axes = tuple(ax for ax in range(len(self.shape)) if ax not in self._axes[i])
            ebar = ebar.sum(axis=other_axes) if other_axes else ebar
            nabar = np.bincount(grid.power_distributor.ravel(), weights=ebar.ravel(),
                                minlength=grid.mode_lengths.size)
            abar = nabar / z
            abar[0] = 0.0
            zbar += -float(np.sum(nabar[1:] * amps[i][1:])) / z**2
            out.update(amp.vjp(p, abar))
        out[pf + "zeromode"] = np.asarray(zbar * self.azm.deriv(p[pf + "zeromode"]))
        return out
