"""Repository rules: the product never touches the oracle or the emulator; only tests/, smoke() and
bench.py's CPU-baseline legs may import ``oracle``."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(base, f)


def test_product_does_not_import_oracle_or_emulator():
    for f in _py_files(os.path.join(ROOT, "nifty_b200")):
        src = open(f).read()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
        assert "libniftyb200_emu" not in src and "build_emu" not in src, f


def test_bench_uses_oracle_only_in_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"^\s*import oracle", src, flags=re.M):
        head = src[:m.start()]
        fn = re.findall(r"^def (\w+)\(", head, flags=re.M)[-1]
        assert fn == "cpu_setup", fn


def test_oracle_headers_say_test_infrastructure():
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith(".py"):
            assert "TEST INFRASTRUCTURE" in open(os.path.join(ROOT, "oracle", f)).read(), f
