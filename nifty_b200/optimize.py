"""Newton-CG with the call protocol of ``nifty/re/optimize.py`` (``_newton_cg``:271-411,
``_line_search_successive_halving``:583-654, ``OptimizeResults``:31-72) on flat device vectors."""

from __future__ import annotations

from typing import Callable, NamedTuple, Optional

from datetime import datetime

import numpy as np
import torch

from .conjugate_gradient import _cg, _norm


class OptimizeResults(NamedTuple):
    x: torch.Tensor
    success: bool
    status: int
    fun: float
    jac: torch.Tensor
    nit: int
    nfev: int
    njev: int
    nhev: int


def _prepare_fun_vag_hessp(fun, jac, hessp, fun_and_grad):
    """``(fun, fun_and_grad, hessp)`` from whatever the caller supplied (optimize.py:75-111).  The reference fills the gaps with
    ``jax.value_and_grad`` / ``jvp(grad)``; here a scalar torch function of a flat tensor is differentiated with torch autograd
    (Hessian-vector products by double backward).  The operators of the hot path always arrive as ``fun_and_grad`` / ``hessp``."""
    if fun_and_grad is None:
        if fun is not None and jac is not None:
            def fun_and_grad(x):
                return fun(x), jac(x)
        elif fun is not None:
            def fun_and_grad(x):
                with torch.enable_grad():
                    xr = x.detach().clone().requires_grad_(True)
                    v = fun(xr)
                    (g,) = torch.autograd.grad(v, xr)
                return float(v.detach()), g.detach()
        else:
            raise ValueError("no function specified")
    if hessp is None:
        if jac is None and fun is None:
            raise ValueError("`hessp` (or `fun` / `jac` to differentiate) is required")

        def hessp(primals, tangents):
            with torch.enable_grad():
                xr = primals.detach().clone().requires_grad_(True)
                if jac is not None:
                    g = jac(xr)
                else:
                    (g,) = torch.autograd.grad(fun(xr), xr, create_graph=True)
                (hv,) = torch.autograd.grad(g, xr, grad_outputs=tangents.detach(), allow_unused=True)
            return torch.zeros_like(primals) if hv is None else hv.detach()
    if fun is None:
        def fun(primals):
            return fun_and_grad(primals)[0]
    return fun, fun_and_grad, hessp


def _newton_cg(fun=None, x0=None, *, miniter=None, maxiter=None, energy_reduction_factor=0.1, old_fval=None, absdelta=None,
               norm_ord=None, xtol=1e-5, jac: Optional[Callable] = None, fun_and_grad: Optional[Callable] = None,
               hessp: Optional[Callable] = None,
               cg=_cg, name=None, time_threshold=None, cg_kwargs=None, custom_gradnorm: Optional[Callable] = None,
               hessp_at: Optional[Callable] = None, vdot: Optional[Callable] = None,
               vnorm: Optional[Callable] = None, _size: Optional[int] = None) -> OptimizeResults:
    """``hessp(pos, v)`` as in the reference; ``hessp_at(pos)`` may instead return an operator object
    (e.g. a :class:`~nifty_b200.conjugate_gradient.HamiltonianMetric`) so that the inner CG runs on the device."""
    if x0 is not None and not isinstance(x0, torch.Tensor):
        # latent trees / Vectors (the reference's solvers act on pytrees, test/test_re/test_ncg.py:41-77): solve on the raveled vector
        from .tree_math import ravel
        flat0, unravel = ravel(x0)
        fl = lambda t: ravel(t)[0]                                  # noqa: E731
        wrap1 = lambda f: None if f is None else (lambda x: f(unravel(x)))                       # noqa: E731
        res = _newton_cg(wrap1(fun), flat0, miniter=miniter, maxiter=maxiter, energy_reduction_factor=energy_reduction_factor,
                         old_fval=old_fval, absdelta=absdelta, norm_ord=norm_ord, xtol=xtol,
                         jac=None if jac is None else (lambda x: fl(jac(unravel(x)))),
                         fun_and_grad=None if fun_and_grad is None else (lambda x: (lambda v, g: (v, fl(g)))(*fun_and_grad(unravel(x)))),
                         hessp=None if hessp is None else (lambda x, t: fl(hessp(unravel(x), unravel(t)))),
                         cg=cg, name=name, time_threshold=time_threshold, cg_kwargs=cg_kwargs,
                         custom_gradnorm=None if custom_gradnorm is None else (lambda g: custom_gradnorm(unravel(g))),
                         hessp_at=None if hessp_at is None else (lambda x: (lambda op: (lambda t: fl(op(unravel(t)))))(hessp_at(unravel(x)))),
                         vdot=vdot, vnorm=vnorm, _size=_size)
        return res._replace(x=unravel(res.x), jac=None if res.jac is None else unravel(res.jac))
    norm_ord = 1 if norm_ord is None else norm_ord
    _dot = (lambda a, b: float(torch.dot(a, b))) if vdot is None else vdot
    _nrm = _norm if vnorm is None else vnorm
    miniter = 0 if miniter is None else miniter
    maxiter = 200 if maxiter is None else maxiter
    maxiter = np.iinfo(np.int64).max if np.isinf(maxiter) else int(maxiter)
    pos = x0.clone()
    xtol = xtol * (pos.numel() if _size is None else _size)      # _size: number of non-frozen entries (point estimates)
    cg_kwargs = {} if cg_kwargs is None else dict(cg_kwargs)
    cg_name = cg_kwargs.pop("name", name + "CG" if name is not None else None)      # optimize.py:302
    gradnorm = (lambda v: _nrm(v, norm_ord)) if custom_gradnorm is None else custom_gradnorm
    if hessp_at is None or fun_and_grad is None:
        fun, fun_and_grad, hessp = _prepare_fun_vag_hessp(fun, jac, hessp, fun_and_grad)
    _fg = fun_and_grad

    def fun_and_grad(x):                      # user functions may return 0-d tensors: the control flow below works on floats
        v, gr = _fg(x)
        return float(v), gr
    energy, g = fun_and_grad(pos)
    nfev, njev, nhev = 1, 1, 0
    if np.isnan(energy):
        raise ValueError("energy is Nan")
    status, i = -1, 0
    for i in range(1, maxiter + 1):
        if old_fval and energy_reduction_factor:
            cg_absdelta = energy_reduction_factor * (old_fval - energy)
        else:
            cg_absdelta = None if absdelta is None else absdelta / 100.0
        mag_g = _nrm(g, cg_kwargs.get("norm_ord", 1))
        cg_resnorm = min(0.5, np.sqrt(mag_g)) * mag_g
        kw = dict(absdelta=cg_absdelta, resnorm=cg_resnorm, norm_ord=1, name=cg_name, _raise_nonposdef=False)
        if vdot is not None:
            kw.update(vdot=vdot, vnorm=vnorm)
        if time_threshold is not None:
            kw["time_threshold"] = time_threshold
        kw.update(cg_kwargs)
        op = hessp_at(pos) if hessp_at is not None else (lambda v, _p=pos: hessp(_p, v))
        res = cg(op, g, **kw)
        nat_g, info = res.x, res.info
        nhev += res.nfev
        if info is not None and info < 0:
            raise ValueError("conjugate gradient failed")
        dd, grad_scaling, accepted = nat_g, 1.0, False
        ls_it = 0
        for ls_it in range(9):
            new_pos = pos - grad_scaling * dd
            new_energy, new_g = fun_and_grad(new_pos)
            nfev, njev = nfev + 1, njev + 1
            if new_energy <= energy:
                accepted = True
                break
            grad_scaling /= 2
            if ls_it == 5:
                gam = _dot(g, g)
                hv = hessp_at(pos)(g) if hessp_at is not None else hessp(pos, g)   # re-linearise at pos
                curv = _dot(g, hv)
                nhev += 1
                grad_scaling = 1.0
                dd = gam / curv * g
        if not accepted:
            status = -1
            break
        energy_diff = energy - new_energy
        old_fval = energy
        energy, pos, g = new_energy, new_pos, new_g
        descent_norm = grad_scaling * gradnorm(dd)
        if np.isnan(new_energy):
            raise ValueError("energy is NaN")
        min_cond = ls_it < 2 and i > miniter
        if absdelta is not None and 0.0 <= energy_diff < absdelta and min_cond:
            status = 0
            break
        if descent_norm <= xtol and i > miniter:
            status = 0
            break
        if time_threshold is not None and datetime.now() > time_threshold:      # optimize.py:394-396
            status = i
            break
    else:
        status = i
    return OptimizeResults(pos, True, status, float(energy), g, i, nfev, njev, nhev)


def newton_cg(*args, **kwargs):
    """``jft.newton_cg``: returns only the optimal position."""
    return _newton_cg(*args, **kwargs).x


static_newton_cg = newton_cg


def minimize(fun, x0, args=(), *, method: str, tol=None, options=None) -> OptimizeResults:
    """``jft.minimize`` (optimize.py:863-892).  Newton-CG only: the trust-region minimiser is an alternative outside the path."""
    options = {} if options is None else dict(options)
    if not isinstance(args, tuple):
        raise TypeError(f"args argument must be a tuple, got {type(args)!r}")
    fun_with_args = (lambda x: fun(x, *args)) if args else fun
    if tol is not None:
        raise ValueError("use solver-specific options")
    if method.lower() in ("newton-cg", "newtoncg", "ncg"):
        return _newton_cg(fun_with_args, x0, **options)
    if method.lower() in ("trust-ncg", "trustncg"):
        raise NotImplementedError("trust-ncg is outside the B200 path (SURVEY.md section 2); use method='newton-cg'")
    raise ValueError(f"method {method} not recognized")
