"""TEST INFRASTRUCTURE: run the reference's UNMODIFIED `_cg` / `_newton_cg` without JAX.

`nifty.re` needs jax, which is not installable here.  The two solver routines, however, are plain Python loops over
`jnp` array arithmetic (`/root/reference/nifty/re/conjugate_gradient.py:77-214`, `optimize.py:271-411`); this module
loads those two source files *as they are* from the reference tree into a throw-away package whose `jax` is a
NumPy-backed stand-in and whose helper modules (`.logger`, `.misc`, `.tree_math`) provide the handful of functions the
solvers use, restated for flat float64 vectors (`tree_math/vector_math.py:128-223`).  Nothing is copied: the reference
files are read and executed in place.  Used by `make_solver_golden.py` (fixture generation, this container only) and by
`tests/test_oracle_solvers.py` (live re-check whenever `/root/reference` is present).
"""
import functools
import importlib.util
import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference/nifty/re"
PKG = "_nifty_re_ref"


def available():
    return os.path.isfile(os.path.join(REF, "conjugate_gradient.py")) and os.path.isfile(os.path.join(REF, "optimize.py"))


def _numpy_as_jnp():
    jnp = types.ModuleType("jax.numpy")
    for k in dir(np):
        if not k.startswith("__"):
            setattr(jnp, k, getattr(np, k))
    jnp.ndarray = np.ndarray
    jnp.linalg = np.linalg
    return jnp


def _norm(tree, ord=2):
    x = np.asarray(tree)
    if x.ndim == 0:
        return np.abs(x)
    return np.linalg.norm(np.array([np.linalg.norm(x.ravel(), ord=ord)]), ord=ord)


def load():
    """returns (conjugate_gradient module, optimize module) executed from the reference sources"""
    if PKG + ".optimize" in sys.modules:
        return sys.modules[PKG + ".conjugate_gradient"], sys.modules[PKG + ".optimize"]
    saved = {k: sys.modules.get(k) for k in ("jax", "jax.numpy", "jax.lax", "jax.tree_util")}
    jax = types.ModuleType("jax")
    jnp = _numpy_as_jnp()
    lax = types.ModuleType("jax.lax")
    tu = types.ModuleType("jax.tree_util")
    tu.Partial = functools.partial
    jax.numpy, jax.lax, jax.tree_util = jnp, lax, tu
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.tree_util": tu})
    try:
        pkg = types.ModuleType(PKG)
        pkg.__path__ = [REF]
        sys.modules[PKG] = pkg
        lg = types.ModuleType(PKG + ".logger")
        lg.logger = logging.getLogger("nifty_re_ref")
        lg.logger.addHandler(logging.NullHandler())
        lg.logger.propagate = False
        misc = types.ModuleType(PKG + ".misc")
        misc.doc_from = lambda original: (lambda target: target)
        misc.safeguard_arguments_against_accidental_calls_into_jax = lambda f: f
        misc.conditional_call = lambda condition, fn, *a, **k: fn(*a, **k) if condition else None

        def conditional_raise(condition, exception):
            if condition:
                raise exception
        misc.conditional_raise = conditional_raise
        tm = types.ModuleType(PKG + ".tree_math")
        tm.assert_arithmetics = lambda *a, **k: None
        tm.result_type = lambda *trees: np.result_type(*[np.asarray(t).dtype for t in trees])
        tm.size = lambda a, axis=None: int(np.size(a))
        tm.vdot = lambda a, b, precision=None: np.add(np.vdot(a, b), 0.0)
        tm.where = np.where
        tm.zeros_like = np.zeros_like
        tm.norm = _norm
        tm.PyTreeString = str
        tm.hide_strings = lambda x: x
        for m in (lg, misc, tm):
            sys.modules[m.__name__] = m
        mods = []
        for name in ("conjugate_gradient", "optimize"):
            spec = importlib.util.spec_from_file_location(f"{PKG}.{name}", os.path.join(REF, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            spec.loader.exec_module(mod)
            mods.append(mod)
        setattr(pkg, "conjugate_gradient", mods[0])
        return tuple(mods)
    finally:
        for k, v in saved.items():      # the stand-in `jax` must not leak into other tests
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
