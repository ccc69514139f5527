"""Oracle (TEST INFRASTRUCTURE ONLY): signal model and amended likelihoods.

Restates, for the model family of the hot path (``signal = [scaling *] nl(cf)``,
``nl`` in {identity, exp}; demos/re/0_intro.py:39-57,
misc/re/paper/minimal_benchmark.py:89), what ``nifty.re`` computes through

* ``Gaussian`` / ``Poissonian``        nifty/re/likelihood_impl.py:83-138, 203-251
* ``LikelihoodWithModel`` chain rule   nifty/re/likelihood.py:599-633
* ``Likelihood.metric = LSM o RSM``    nifty/re/likelihood.py:263-332

with the JAX AD calls (``linearize``/``linear_transpose``/``vjp``) replaced by
the hand-derived ``jvp``/``vjp`` of :class:`SignalOracle`.
"""

from __future__ import annotations

import numpy as np

from .correlated_field import CorrelatedFieldOracle, _Prior


class SignalOracle:
    """``signal(p) = scaling(p) * nl(cf(p))`` (scaling optional, log-normal, shape (1,))."""

    def __init__(self, cf: CorrelatedFieldOracle, nonlinearity="exp", scaling=None,
                 scaling_key="scaling"):
        # a pair (fn, dfn) of NumPy callables: an arbitrary POINTWISE map of the field and its derivative (what jax.jvp / vjp
        # give for `Model(lambda x: fn(cf(x)))`, nifty/re/model.py:146-181)
        self.nl_pair = None
        if isinstance(nonlinearity, tuple):
            self.nl_pair, nonlinearity = nonlinearity, "custom"
            if scaling is not None:
                raise ValueError("scaling + custom nonlinearity")
        elif nonlinearity not in ("exp", "identity"):
            raise ValueError(nonlinearity)
        self.cf = cf
        self.nl = nonlinearity
        self.scaling = None if scaling is None else _Prior("lognormal", *scaling)
        self.scaling_key = scaling_key
        self.domain = dict(cf.domain)
        if self.scaling is not None:
            self.domain[scaling_key] = (1,)
        self.domain = dict(sorted(self.domain.items()))
        self.target_shape = cf.shape

    def _scal(self, p):
        if self.scaling is None:
            return 1.0
        return float(np.ravel(self.scaling(p[self.scaling_key]))[0])

    def __call__(self, p):
        f = self.cf(p)
        if self.nl_pair is not None:
            return self.nl_pair[0](f)
        y = np.exp(f) if self.nl == "exp" else f
        return self._scal(p) * y

    def jvp(self, p, dp):
        f = self.cf(p)
        df = self.cf.jvp(p, dp)
        if self.nl_pair is not None:
            return self.nl_pair[1](f) * df
        sc = self._scal(p)
        if self.nl == "exp":
            y = np.exp(f)
            dy = sc * y * df
        else:
            y = f
            dy = sc * df
        if self.scaling is not None:
            dsc = float(np.ravel(self.scaling.deriv(p[self.scaling_key]) * dp[self.scaling_key])[0])
            dy = dy + dsc * y
        return dy

    def vjp(self, p, c):
        f = self.cf(p)
        sc = self._scal(p)
        c = np.asarray(c, dtype=np.float64)
        if self.nl_pair is not None:
            return self.cf.vjp(p, self.nl_pair[1](f) * c)
        if self.nl == "exp":
            y = np.exp(f)
            fbar = sc * y * c
        else:
            y = f
            fbar = sc * c
        out = self.cf.vjp(p, fbar)
        if self.scaling is not None:
            scbar = float(np.sum(y * c))
            out[self.scaling_key] = np.reshape(
                scbar * np.ravel(self.scaling.deriv(p[self.scaling_key])), (1,))
        return out


class _AmendedLikelihood:
    """Chain rule of nifty/re/likelihood.py:599-633 for a data-space diagonal likelihood."""

    signal: SignalOracle

    # data-space pieces, to be provided by subclasses
    def _energy(self, y): raise NotImplementedError
    def _denergy(self, y): raise NotImplementedError
    def _metric_diag(self, y): raise NotImplementedError
    def _lsm_diag(self, y): raise NotImplementedError
    def _trafo(self, y): raise NotImplementedError

    @property
    def domain(self):
        return self.signal.domain

    @property
    def data_shape(self):
        return self.signal.target_shape

    def energy(self, p):
        return float(self._energy(self.signal(p)))

    def energy_and_gradient(self, p):
        y = self.signal(p)
        return float(self._energy(y)), self.signal.vjp(p, self._denergy(y))

    def metric(self, p, t):
        y = self.signal(p)
        return self.signal.vjp(p, self._metric_diag(y) * self.signal.jvp(p, t))

    def left_sqrt_metric(self, p, eta):
        y = self.signal(p)
        return self.signal.vjp(p, self._lsm_diag(y) * np.asarray(eta, dtype=np.float64))

    def right_sqrt_metric(self, p, t):
        y = self.signal(p)
        return self._lsm_diag(y) * self.signal.jvp(p, t)

    def transformation(self, p):
        return self._trafo(self.signal(p))


class GaussianOracle(_AmendedLikelihood):
    """``Gaussian(data, noise_cov_inv).amend(signal)`` with a diagonal ``noise_cov_inv``.

    nifty/re/likelihood_impl.py:124-138; ``noise_cov_inv`` is the scalar or array
    ``w`` of ``lambda x: w * x``; ``noise_std_inv = sqrt(w)`` (:35-80).
    """

    def __init__(self, data, noise_cov_inv, signal: SignalOracle):
        self.data = np.asarray(data, dtype=np.float64)
        self.w = np.asarray(noise_cov_inv, dtype=np.float64)
        self.signal = signal

    def _energy(self, y):
        r = self.data - y
        return 0.5 * np.vdot(r, self.w * r)

    def _denergy(self, y):
        return self.w * (y - self.data)

    def _metric_diag(self, y):
        return self.w * np.ones_like(y)

    def _lsm_diag(self, y):
        return np.sqrt(self.w) * np.ones_like(y)

    def _trafo(self, y):
        return np.sqrt(self.w) * y

    def normalized_residual(self, p):
        return np.sqrt(self.w) * (self.data - self.signal(p))


class PoissonianOracle(_AmendedLikelihood):
    """``Poissonian(data).amend(signal)``; nifty/re/likelihood_impl.py:226-251."""

    def __init__(self, data, signal: SignalOracle):
        data = np.asarray(data)
        if not np.issubdtype(data.dtype, np.integer):
            raise TypeError("`data` of invalid type")
        if np.any(data < 0):
            raise ValueError("`data` must not be negative")
        self.data = data
        self.signal = signal

    def _energy(self, y):
        return np.sum(y) - np.vdot(np.log(y), self.data)

    def _denergy(self, y):
        return 1.0 - self.data / y

    def _metric_diag(self, y):
        return 1.0 / y

    def _lsm_diag(self, y):
        return 1.0 / np.sqrt(y)

    def _trafo(self, y):
        return 2.0 * np.sqrt(y)

    def normalized_residual(self, p):
        y = self.signal(p)
        return (self.data - y) / np.sqrt(y)
