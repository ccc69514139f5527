"""The C-ABI library loads and exports every symbol include/nifty_b200.h declares (no compute calls:
runs without a GPU), and the ctypes table matches the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nifty_b200.h")
LIB = os.path.join(ROOT, "nifty_b200", "lib", "libniftyb200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_ctypes_table():
    from nifty_b200._capi import SIGNATURES
    assert declared_symbols() == sorted(SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.nb200_version.restype = ctypes.c_int
    assert lib.nb200_version() >= 100
    lib.nb200_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.nb200_last_error(), bytes)


def test_no_cuda_device_fails_loudly():
    """Without a GPU the product refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import nifty_b200 as nb
    with pytest.raises(nb.NB200Error, match="no CPU fallback"):
        nb.Runtime()
    # the C level refuses as well
    from nifty_b200._capi import CApi
    api = CApi(LIB)
    h = ctypes.c_void_p()
    shp = (ctypes.c_int64 * 2)(16, 16)
    dst = (ctypes.c_double * 2)(1.0, 1.0)
    rc = api.lib.nb200_plan_create(ctypes.byref(h), 0, 2, shp, dst, 1, 0)
    assert rc != 0 and b"no CUDA device" in api.lib.nb200_last_error()
