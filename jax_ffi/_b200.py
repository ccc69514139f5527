"""JAX side of the binding: registers the XLA FFI targets of ``jax_ffi/nifty_b200_jax.cc`` and builds the
JAX-transformable correlated-field function a ``nifty.re`` maintainer drops in at the two seams SURVEY.md section 8b
names (``hartley`` bound in ``CorrelatedFieldMaker.finalize``, nifty/re/correlated_field.py:865, or the closure
``correlated_field(p)`` :909-912).

Activates when JAX is importable (it is not in this repository's image: ``import jax`` fails and :func:`available`
returns False); nothing in ``nifty_b200`` imports this module.

What JAX asks of the op and how it is met (nifty/re/likelihood.py:618-619 ``jax.linearize`` + ``jax.linear_transpose``;
optimize_kl.py:106,135 ``vmap`` over samples; test_empirical_power_spectrum.py:39 ``VModel(cf, in_axes="xi")``):

* ``jit``      -- ``jax.ffi.ffi_call`` custom calls on the CUDA platform, enqueued on XLA's stream.
* ``jvp`` / ``linearize`` -- ``jax.custom_jvp``: ``cf(a, x) = offset + L(a, x)`` with ``L`` bilinear, so the tangent is
  ``L(da, x) + L(a, dx)`` -- two calls of the SAME FFI target, each wrapped in ``jax.custom_derivatives.linear_call``,
  i.e. declared linear in its last argument with an explicit transpose.  A ``custom_vjp`` would not do: it cannot be
  forward-differentiated, and ``LikelihoodWithModel.metric`` linearises first.
* ``linear_transpose`` / ``vjp`` -- the transposes registered with ``linear_call`` are the two halves of
  ``nb200_cf_apply_adjoint``: ``c -> a[pd] * g`` and ``c -> segment_sum(x * g)``.
* ``vmap``     -- ``vmap_method="broadcast_all"``: batch axes are prepended to every operand, the handlers read the batch
  extent from the buffer shapes and call the ``*_batch`` entry points (``amp`` of batch extent 1 = shared table).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libnifty_b200_jax.so")
_registered = False


def available() -> bool:
    """True if JAX is importable and the handler library was built against the XLA FFI headers."""
    try:
        import jax  # noqa: F401
        import jax.ffi  # noqa: F401
    except Exception:
        return False
    if not os.path.exists(_LIB):
        return False
    return bool(ctypes.CDLL(_LIB).nb200_jax_ffi_available())


def register():
    """``jax.ffi.register_ffi_target`` for the handlers (idempotent)."""
    global _registered
    if _registered:
        return
    import jax
    lib = ctypes.CDLL(_LIB)
    for name in ("nb200_jax_cf_apply", "nb200_jax_cf_adjoint", "nb200_jax_cf_adjoint_xi", "nb200_jax_hartley", "nb200_jax_hartley_chirpz"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def make_correlated_field(plan_handle: int, grid_shape, n_bins: int, offset: float, dtype=np.float64):
    """``cf(amp, xi) = offset + (1/V) hartley(amp[power_distributor] * xi)`` as a JAX function that supports jit, vmap,
    jvp / linearize, vjp and linear_transpose.  ``plan_handle``: address of an ``nb200_plan`` (``nifty_b200.Plan(...)._h``)
    that the caller keeps alive; ``amp``: the K-entry table ``azm * normalized_amplitude`` with ``amp[0] = zeromode * V``
    (nifty/re/correlated_field.py:889-912)."""
    import jax
    import jax.numpy as jnp
    from jax.custom_derivatives import linear_call

    register()
    grid_shape = tuple(int(s) for s in grid_shape)
    rank = len(grid_shape)
    attrs = dict(plan=np.int64(plan_handle), grid_rank=np.int64(rank))

    def _L(amp, xi, off=0.0):
        out = jax.ShapeDtypeStruct(xi.shape, xi.dtype)
        return jax.ffi.ffi_call("nb200_jax_cf_apply", out, vmap_method="broadcast_all")(amp, xi, offset=np.float64(off), **attrs)

    def _LT(amp, xi, cot):
        outs = (jax.ShapeDtypeStruct(cot.shape, cot.dtype), jax.ShapeDtypeStruct(cot.shape[:cot.ndim - rank] + (n_bins,), cot.dtype))
        return jax.ffi.ffi_call("nb200_jax_cf_adjoint", outs, vmap_method="broadcast_all")(amp, xi, cot, **attrs)

    def _LT_xi(amp, cot):
        out = jax.ShapeDtypeStruct(cot.shape, cot.dtype)
        return jax.ffi.ffi_call("nb200_jax_cf_adjoint_xi", out, vmap_method="broadcast_all")(amp, cot, **attrs)

    # L is linear in each argument: declare it so, with the transposes the library provides
    def lin_in_xi(amp, dxi):
        return linear_call(lambda a, x: _L(a, x), lambda a, c: _LT_xi(a, c), amp, dxi)

    def lin_in_amp(xi, damp):
        return linear_call(lambda x, a: _L(a, x), lambda x, c: _LT(jnp.ones((n_bins,), c.dtype), x, c)[1], xi, damp)

    @jax.custom_jvp
    def cf(amp, xi):
        return _L(amp, xi, offset)

    @cf.defjvp
    def _cf_jvp(primals, tangents):
        amp, xi = primals
        damp, dxi = tangents
        return cf(amp, xi), lin_in_amp(xi, damp) + lin_in_xi(amp, dxi)

    return cf


def make_hartley(plan_handle: int, grid_shape, dtype=np.float64, *, chirpz_tables=None, padded_shape=None):
    """The ``hartley(p, axes=all)`` seam (nifty/re/correlated_field.py:24-30, bound with ``partial`` at :865) as a JAX function:
    assign the result to ``nifty.re.correlated_field.hartley`` (or pass it where ``finalize`` binds it) before ``finalize()``.

    The transform is linear and self-adjoint, so one ``linear_call`` whose transpose is the function itself gives jit, vmap,
    jvp / linearize and linear_transpose.  ``chirpz_tables`` (device array of the per-axis chirp / filter tables,
    ``nifty_b200.BluesteinHartley(...)._tab``) and ``padded_shape`` select the chirp-z form for extents that are not powers of
    two; ``plan_handle`` is then the plan of the PADDED grid."""
    import jax
    from jax.custom_derivatives import linear_call

    register()
    grid_shape = tuple(int(s) for s in grid_shape)
    rank = len(grid_shape)
    item_bytes = int(np.prod(grid_shape)) * np.dtype(dtype).itemsize
    attrs = dict(plan=np.int64(plan_handle), grid_rank=np.int64(rank), item_bytes=np.int64(item_bytes))

    def _raw(x):
        out = jax.ShapeDtypeStruct(x.shape, x.dtype)
        if chirpz_tables is None:
            return jax.ffi.ffi_call("nb200_jax_hartley", out, vmap_method="broadcast_all")(x, **attrs)
        n = (1,) * (3 - rank) + grid_shape
        work = jax.ShapeDtypeStruct((2 * int(np.prod(padded_shape)),), x.dtype)
        res, _ = jax.ffi.ffi_call("nb200_jax_hartley_chirpz", (out, work), vmap_method="broadcast_all")(
            chirpz_tables, x, n0=np.int64(n[0]), n1=np.int64(n[1]), n2=np.int64(n[2]), **attrs)
        return res

    def hartley(p, axes=None):
        if axes is not None and tuple(sorted(a % p.ndim for a in axes)) != tuple(range(p.ndim - rank, p.ndim)):
            raise NotImplementedError("the device transform acts on all axes of the grid")
        return linear_call(lambda _, x: _raw(x), lambda _, c: _raw(c), (), p)

    return hartley
