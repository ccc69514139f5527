"""The jax.ffi binding (jax_ffi/) cannot run here (no JAX, no XLA headers); what CAN be checked on CPU: the translation unit
builds without the headers and says so, its gated branch type-checks against an API-shaped stand-in of `xla/ffi/api/ffi.h`
and exports the handler symbols, and the Python side degrades to `available() == False` instead of failing."""
import ctypes
import importlib.util
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "jax_ffi", "nifty_b200_jax.cc")


def _build(tmp_path, extra, name):
    out = str(tmp_path / name)
    libdir = os.path.join(ROOT, "nifty_b200", "lib")
    cmd = ["g++", "-O0", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include")] + extra + [SRC, "-o", out]
    if extra:
        cmd += ["-L" + libdir, "-lniftyb200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return out


def test_builds_without_xla_headers(tmp_path):
    lib = ctypes.CDLL(_build(tmp_path, [], "plain.so"))
    assert lib.nb200_jax_ffi_available() == 0


def test_gated_branch_type_checks_against_api_stand_in(tmp_path):
    import __graft_entry__ as g
    g.build()
    extra = ["-I" + os.path.join(ROOT, "tests", "mock_xla"), "-I/usr/local/cuda/include"]
    lib = ctypes.CDLL(_build(tmp_path, extra, "mock.so"))
    assert lib.nb200_jax_ffi_available() == 1
    for name in ("nb200_jax_cf_apply", "nb200_jax_cf_adjoint", "nb200_jax_cf_adjoint_xi", "nb200_jax_hartley", "nb200_jax_hartley_chirpz"):
        assert hasattr(lib, name)


def test_python_side_degrades_without_jax():
    spec = importlib.util.spec_from_file_location("_b200", os.path.join(ROOT, "jax_ffi", "_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        import jax  # noqa: F401
        have_jax = True
    except Exception:
        have_jax = False
    if not have_jax:
        assert mod.available() is False
