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


# ---- the function-style reparametrisations of num/stats_distributions.py:20-134 on torch tensors ---------------------------------
def normal_prior(mean, std):
    """Standard normal -> N(mean, std) (stats_distributions.py:42-50)."""
    return lambda xi: mean + std * xi


def normal_invprior(mean, std):
    """Inverse of :func:`normal_prior` (:53-59)."""
    return lambda y: (y - mean) / std


def lognormal_prior(mean, std, *, _log_mean=None, _log_std=None):
    """Standard normal -> log-normal with the given mean / std (:76-93)."""
    import torch
    if _log_mean is None and _log_std is None:
        _log_mean, _log_std = lognormal_moments(mean, std)
    return lambda xi: torch.exp(_log_mean + _log_std * torch.as_tensor(xi))


def lognormal_invprior(mean, std, *, _log_mean=None, _log_std=None):
    """Inverse of :func:`lognormal_prior` (:96-104)."""
    import torch
    if _log_mean is None and _log_std is None:
        _log_mean, _log_std = lognormal_moments(mean, std)
    return lambda y: (torch.log(torch.as_tensor(y)) - _log_mean) / _log_std


def uniform_prior(a_min=0.0, a_max=1.0):
    """Standard normal -> uniform on [a_min, a_max] through the normal CDF (:107-134)."""
    import torch
    scale = a_max - a_min
    return lambda xi: a_min + scale * torch.special.ndtr(torch.as_tensor(xi))


def laplace_prior(alpha):
    """Standard normal -> Laplace, ``P(x|a) = exp(-|x| / a) / (2 a)`` (:20-39)."""
    import torch

    def standard_to_laplace(xi):
        xi = torch.as_tensor(xi)
        res = (xi < 0) * (torch.special.log_ndtr(xi) + math.log(2.0))
        res = res - (xi > 0) * (torch.special.log_ndtr(-xi) + math.log(2.0))
        return res * alpha
    return standard_to_laplace
