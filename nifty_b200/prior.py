"""Standard-normal -> (log)normal reparametrisations of scalar hyper-parameters.

Host-side mirror of ``nifty/re/num/stats_distributions.py:42-98`` and ``nifty/re/prior.py:46-86``.
Only the parametrisation constants are computed here; the device kernels evaluate
``a + b*xi`` / ``exp(a + b*xi)`` (nb_amp.cuh).
"""

from __future__ import annotations

import math


def lognormal_moments(mean, std):
    """(logmean, logstd) of a log-normal with the given mean / std (stats_distributions.py:62-73)."""
    mean, std = float(mean), float(std)
    if mean <= 0.0:
        raise ValueError(f"`mean` must be greater zero; got {mean!r}")
    if std <= 0.0:
        raise ValueError(f"`std` must be greater zero; got {std!r}")
    logstd = math.sqrt(math.log1p((std / mean) ** 2))
    logmean = math.log(mean) - 0.5 * logstd**2
    return logmean, logstd


class _ScalarPrior:
    kind = ""

    def __init__(self, mean, std, name="", shape=()):
        self.mean, self.std, self.name, self.shape = float(mean), float(std), name, tuple(shape)

    def ab(self):
        raise NotImplementedError

    def __call__(self, xi):
        a, b = self.ab()
        v = a + b * xi
        return v.exp() if self.kind == "lognormal" and hasattr(v, "exp") else (math.exp(v) if self.kind == "lognormal" else v)


class NormalPrior(_ScalarPrior):
    """``jft.NormalPrior(mean, std)``: value = mean + std * xi."""
    kind = "normal"

    def ab(self):
        return self.mean, self.std


class LogNormalPrior(_ScalarPrior):
    """``jft.LogNormalPrior(mean, std)``: value = exp(logmean + logstd * xi)."""
    kind = "lognormal"

    def ab(self):
        return lognormal_moments(self.mean, self.std)


def _as_prior(spec, cls, name, optional=False):
    """Accept the reference's ``(mean, std)`` tuples (correlated_field.py:623-646 error behaviour)."""
    if spec is None:
        if optional:
            return None
        raise TypeError(f"invalid `{name}` specified; got '{type(spec)}'")
    if isinstance(spec, _ScalarPrior):
        return spec
    if isinstance(spec, (tuple, list)):
        if len(spec) != 2:
            raise TypeError(f"invalid `{name}` specified; got {spec!r}")
        return cls(*spec, name=name)
    raise TypeError(f"invalid `{name}` specified; got '{type(spec)}'")
