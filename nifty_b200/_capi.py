"""ctypes binding of the C ABI declared in ``include/nifty_b200.h``.

The product only ever loads ``nifty_b200/lib/libniftyb200.so`` (hand-written sm_100a kernels); if
that library is missing or no CUDA device is present every entry point raises -- there is no CPU
fallback.  ``CApi(path)`` accepts an explicit path only so that the test-suite can bind the
sequential host emulation of the same kernel sources (``tests/emu``) for logic checks.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libniftyb200.so")

vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double


class ModelDesc(C.Structure):
    """``nb200_model_desc``"""
    _fields_ = [
        ("kind_power", i32), ("has_fluctuations", i32), ("has_deviations", i32), ("has_asperity", i32),
        ("has_scaling", i32), ("reserved", i32),
        ("offset_mean", f64),
        ("zeromode_a", f64), ("zeromode_b", f64),
        ("fluct_a", f64), ("fluct_b", f64),
        ("slope_a", f64), ("slope_b", f64),
        ("flex_a", f64), ("flex_b", f64),
        ("asp_a", f64), ("asp_b", f64),
        ("scaling_a", f64), ("scaling_b", f64),
        ("off_xi", i64), ("off_zeromode", i64), ("off_fluct", i64), ("off_slope", i64), ("off_flex", i64),
        ("off_asp", i64), ("off_spectrum", i64), ("off_scaling", i64),
        ("latent_size", i64),
        ("amplitude_type", i32), ("renormalize_amplitude", i32), ("cutoff_a", f64), ("cutoff_b", f64), ("off_cutoff", i64),
    ]


class CgOpts(C.Structure):
    """``nb200_cg_opts``"""
    _fields_ = [("absdelta", f64), ("resnorm", f64), ("tol", f64), ("atol", f64), ("norm_ord", i32),
                ("miniter", i32), ("maxiter", i32), ("raise_nonposdef", i32), ("check_every", i32),
                ("x0_is_zero", i32), ("n_frozen", i32), ("reserved", i32), ("frozen", C.POINTER(i64))]


class CgResult(C.Structure):
    """``nb200_cg_result``"""
    _fields_ = [("info", i32), ("nit", i32), ("nfev", i32), ("error", i32), ("energy", f64), ("gamma", f64)]


REDUCE_HOOK = C.CFUNCTYPE(None, vp, vp, i64, vp)      # nb200_reduce_hook

# name -> (restype, argtypes); kept in sync with include/nifty_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "nb200_last_error": (C.c_char_p, []),
    "nb200_version": (C.c_int, []),
    "nb200_launch_count": (C.c_ulonglong, []),
    "nb200_timing_begin": (C.c_int, []),
    "nb200_timing_end": (C.c_int, [C.c_char_p, i64]),
    "nb200_plan_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(i64), C.POINTER(f64), C.c_int, C.c_int]),
    "nb200_plan_destroy": (None, [vp]),
    "nb200_plan_num_modes": (i64, [vp]),
    "nb200_plan_size": (i64, [vp]),
    "nb200_plan_total_volume": (f64, [vp]),
    "nb200_plan_mode_lengths": (C.c_int, [vp, vp]),
    "nb200_plan_mode_multiplicity": (C.c_int, [vp, vp]),
    "nb200_plan_relative_log_mode_lengths": (C.c_int, [vp, vp]),
    "nb200_plan_log_volume": (C.c_int, [vp, vp]),
    "nb200_plan_power_distributor": (C.c_int, [vp, vp]),
    "nb200_plan_create_dist": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(i64), C.POINTER(f64), C.c_int, C.c_int, C.c_int, C.c_int]),
    "nb200_plan_dist_info": (C.c_int, [vp, C.POINTER(i64), i64]),
    "nb200_plan_local_map": (C.c_int, [vp, C.c_int, vp]),
    "nb200_plan_set_scratch": (C.c_int, [vp, vp, vp, vp]),
    "nb200_plan_set_chunks": (C.c_int, [vp, C.c_int]),
    "nb200_plan_set_reduce_chunks": (C.c_int, [vp, C.c_int]),
    "nb200_dist_phase": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int]),
    "nb200_hartley": (C.c_int, [vp, vp, vp, vp]),
    "nb200_hartley_chirpz": (C.c_int, [vp, vp, C.POINTER(i64), vp, vp, vp, vp]),
    "nb200_cf_apply": (C.c_int, [vp, vp, vp, vp, f64, vp]),
    "nb200_cf_apply_adjoint": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "nb200_cf_apply_batch": (C.c_int, [vp, vp, vp, i64, vp, f64, vp, i64]),
    "nb200_cf_apply_adjoint_batch": (C.c_int, [vp, vp, vp, i64, vp, vp, vp, vp, i64, i64]),
    "nb200_model_create": (C.c_int, [C.POINTER(vp), vp, C.POINTER(ModelDesc)]),
    "nb200_model_destroy": (None, [vp]),
    "nb200_model_set_likelihood": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, f64, vp]),
    "nb200_lin_create": (C.c_int, [C.POINTER(vp), vp]),
    "nb200_lin_destroy": (None, [vp]),
    "nb200_lin_update": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "nb200_lin_set_pointwise": (C.c_int, [vp, vp, vp, vp]),
    "nb200_lin_energy": (C.c_int, [vp, vp, C.POINTER(f64)]),
    "nb200_lin_amplitude": (C.c_int, [vp, vp, vp]),
    "nb200_lin_signal": (C.c_int, [vp, vp, vp]),
    "nb200_cf_forward": (C.c_int, [vp, vp, vp, vp]),
    "nb200_cf_amplitude": (C.c_int, [vp, vp, vp, vp]),
    "nb200_metric": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "nb200_metric_pair": (C.c_int, [vp, vp, vp, vp, vp, C.c_int]),
    "nb200_rsm": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "nb200_lsm": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "nb200_transformation": (C.c_int, [vp, vp, vp]),
    "nb200_normalized_residual": (C.c_int, [vp, vp, vp]),
    "nb200_cg_default_opts": (None, [C.POINTER(CgOpts)]),
    "nb200_cg_solve": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(CgOpts), C.POINTER(CgResult)]),
    "nb200_metric_multi": (C.c_int, [C.POINTER(vp), C.c_int, f64, C.c_int, vp, vp, vp, vp, vp]),
    "nb200_cg_solve_multi": (C.c_int, [C.POINTER(vp), C.c_int, f64, C.c_int, vp, vp, vp, C.POINTER(CgOpts), C.POINTER(CgResult), vp, vp]),
    "nb200_vec_axpby": (C.c_int, [vp, vp, i64, f64, vp, f64, vp, vp]),
    "nb200_vec_dot": (C.c_int, [vp, vp, i64, vp, vp, C.POINTER(f64)]),
    "nb200_vec_stats": (C.c_int, [vp, vp, i64, vp, C.POINTER(f64)]),
}


class NB200Error(RuntimeError):
    pass


class CApi:
    """Loaded library with typed entry points; ``call(name, *args)`` raises on a non-zero status."""

    def __init__(self, path: str = DEFAULT_LIB):
        if not os.path.exists(path):
            raise NB200Error(
                f"{path} not found: build the sm_100a library first (python -c 'import __graft_entry__ as g; g.build()'). "
                "nifty_b200 has no CPU fallback.")
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)
            fn.restype = res
            fn.argtypes = args

    def last_error(self) -> str:
        return self.lib.nb200_last_error().decode()

    def call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            raise NB200Error(self.last_error())
        return rc


_default = None


def default_api() -> CApi:
    global _default
    if _default is None:
        _default = CApi(DEFAULT_LIB)
    return _default
