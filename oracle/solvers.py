"""Oracle (TEST INFRASTRUCTURE ONLY): conjugate gradient and Newton-CG on flat vectors.

Restates the control flow of

* ``_cg``          nifty/re/conjugate_gradient.py:77-214
* ``_newton_cg``   nifty/re/optimize.py:271-411

for plain ``numpy`` vectors (the reference runs the same recurrences on
pytrees; ``vdot``/``norm`` over a pytree equal the flat-vector ones,
nifty/re/tree_math/vector_math.py:173-223).
"""

from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional

import numpy as np

N_RESET = 20  # nifty/re/conjugate_gradient.py:17


@dataclasses.dataclass
class CGResult:
    x: Any
    nit: int
    nfev: int
    info: int
    success: bool
    energies: list  # per-iteration energy trace (oracle extra, used by parity tests)


def _vnorm(v, ord):
    if ord == 1:
        return float(np.sum(np.abs(v)))
    if ord == 2:
        return float(np.sqrt(np.vdot(v, v).real))
    if ord == np.inf:
        return float(np.max(np.abs(v)))
    return float(np.linalg.norm(v, ord=ord))


def cg(mat: Callable, j, x0=None, *, absdelta=None, resnorm=None, norm_ord=None, tol=1e-5,
       atol=0.0, miniter=None, maxiter=None, name=None, _raise_nonposdef=True) -> CGResult:
    """Solve ``mat(x) = j``; sign conventions ``r = mat(x) - j``, ``x <- x - alpha d``."""
    j = np.asarray(j)
    norm_ord = 2 if norm_ord is None else norm_ord
    maxiter_fallback = 20 * j.size
    if miniter is None:
        miniter = min(6, maxiter if maxiter is not None else maxiter_fallback)
    if maxiter is None:
        maxiter = max(min(200, maxiter_fallback), miniter)
    if absdelta is None and resnorm is None:
        resnorm = max(tol * _vnorm(j, norm_ord), atol)
    fi = np.finfo(j.dtype if np.issubdtype(j.dtype, np.floating) else np.float64)
    eps, tiny = 6.0 * fi.eps, 6.0 * fi.tiny

    trace = []
    if x0 is None:
        pos = np.zeros_like(j)
        r = -j
        d = r
        energy = 0.0
        nfev = 0
    else:
        pos = np.array(x0, copy=True)
        r = mat(pos) - j
        d = r
        energy = float(np.vdot((r - j) / 2, pos).real)
        nfev = 1
    previous_gamma = float(np.vdot(r, r).real)
    trace.append(energy)
    if previous_gamma == 0:
        return CGResult(pos, 0, nfev, 0, True, trace)

    info = -1
    i = 0
    for i in range(1, maxiter + 1):
        q = mat(d)
        nfev += 1
        curv = float(np.vdot(d, q).real)
        nm = "CG" if name is None else name
        if curv == 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: zero curvature")
            info = 0
            break
        if curv < 0.0:
            if _raise_nonposdef:
                raise ValueError(f"{nm}: negative curvature")
            if i == 1:
                pos = previous_gamma / (-curv) * (-j)
            info = 0
            break
        alpha = previous_gamma / curv
        pos = pos - alpha * d
        if i % N_RESET == 0:
            r = mat(pos) - j
            nfev += 1
        else:
            r = r - q * alpha
        gamma = float(np.vdot(r, r).real)
        if 0.0 <= gamma <= tiny:
            info = 0
            break
        if resnorm is not None:
            nrm = _vnorm(r, norm_ord)
            if nrm < resnorm and i >= miniter:
                info = 0
                break
        new_energy = float(np.vdot((r - j) / 2, pos).real)
        energy_diff = energy - new_energy
        if energy_diff < -eps * abs(new_energy):
            if _raise_nonposdef:
                raise ValueError(f"{nm}: WARNING: energy increased")
            info = i
            break
        if absdelta is not None and energy_diff < absdelta and i >= miniter:
            info = 0
            break
        energy = new_energy
        trace.append(energy)
        d = d * max(0.0, gamma / previous_gamma) + r
        previous_gamma = gamma
    info = i if info == -1 else info
    return CGResult(pos, i, nfev, info, info == 0, trace)


@dataclasses.dataclass
class NewtonResult:
    x: Any
    success: bool
    status: int
    fun: float
    jac: Any
    nit: int
    nfev: int
    njev: int
    nhev: int
    energies: list


def newton_cg(x0, fun_and_grad: Callable, hessp: Callable, *, miniter=None, maxiter=None,
              energy_reduction_factor=0.1, old_fval=None, absdelta=None, norm_ord=None,
              xtol=1e-5, name=None, cg_kwargs=None, custom_gradnorm: Optional[Callable] = None
              ) -> NewtonResult:
    """Newton-CG with successive-halving line search; ``hessp(pos, v)``."""
    norm_ord = 1 if norm_ord is None else norm_ord
    miniter = 0 if miniter is None else miniter
    maxiter = 200 if maxiter is None else maxiter
    pos = np.array(x0, copy=True)
    xtol = xtol * pos.size
    cg_kwargs = {} if cg_kwargs is None else dict(cg_kwargs)
    cg_kwargs.pop("name", None)
    gradnorm = (lambda v: _vnorm(v, norm_ord)) if custom_gradnorm is None else custom_gradnorm

    energy, g = fun_and_grad(pos)
    nfev, njev, nhev = 1, 1, 0
    if np.isnan(energy):
        raise ValueError("energy is Nan")
    trace = [float(energy)]
    status = -1
    i = 0
    for i in range(1, maxiter + 1):
        if old_fval and energy_reduction_factor:
            cg_absdelta = energy_reduction_factor * (old_fval - energy)
        else:
            cg_absdelta = None if absdelta is None else absdelta / 100.0
        mag_g = _vnorm(g, cg_kwargs.get("norm_ord", 1))
        cg_resnorm = min(0.5, np.sqrt(mag_g)) * mag_g
        kw = dict(absdelta=cg_absdelta, resnorm=cg_resnorm, norm_ord=1, _raise_nonposdef=False)
        kw.update(cg_kwargs)
        res = cg(lambda v, _p=pos: hessp(_p, v), g, **kw)
        nat_g, info = res.x, res.info
        nhev += res.nfev
        if info is not None and info < 0:
            raise ValueError("conjugate gradient failed")

        dd = nat_g
        grad_scaling = 1.0
        accepted = False
        for ls_it in range(9):
            new_pos = pos - grad_scaling * dd
            new_energy, new_g = fun_and_grad(new_pos)
            nfev, njev = nfev + 1, njev + 1
            if new_energy <= energy:
                accepted = True
                break
            grad_scaling /= 2
            if ls_it == 5:
                gam = float(np.vdot(g, g))
                curv = float(np.vdot(g, hessp(pos, g)))
                nhev += 1
                grad_scaling = 1.0
                dd = gam / curv * g
        if not accepted:
            status = -1
            break
        energy_diff = energy - new_energy
        old_fval = energy
        energy = new_energy
        pos = new_pos
        g = new_g
        trace.append(float(energy))
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
    else:
        status = i
    return NewtonResult(pos, True, status, float(energy), g, i, nfev, njev, nhev, trace)
