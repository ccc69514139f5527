"""TEST INFRASTRUCTURE: build the sequential host emulation of the kernel sources.

``nifty_b200/csrc/*.cuh`` write every kernel body once against an execution context; the product
compiles them with nvcc for sm_100a.  Here the same translation unit is compiled with
``g++ -DNB_EMU`` so that one host "thread" runs each block in turn.  It lets the CPU-only test tier
check index arithmetic, mirror logic, epilogues and the host-side call sequences against the
oracle.  It is never built into, shipped with or loaded by the ``nifty_b200`` package.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "nifty_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libniftyb200_emu.so")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "nifty_b200.h"))
    newest = max(os.path.getmtime(f) for f in srcs)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DNB_EMU", "-x", "c++",
           os.path.join(SRC, "nb_api.cu"), "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
