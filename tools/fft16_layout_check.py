"""Design aid for csrc/nb_fft16.cuh: checks the (thread, slot) <-> position layouts of the register-resident
radix-16 line FFT (bijections, DIF correctness against numpy.fft) and counts shared-memory bank conflicts of
every exchange access pattern (8-byte words, half-warp phases) for both thread mappings.

Run: python tools/fft16_layout_check.py
"""
import itertools
import sys

import numpy as np


def bits(v, lo, hi):
    return (v >> lo) & ((1 << (hi - lo)) - 1) if hi > lo else 0


class Layout:
    def __init__(self, lg):
        self.lg = lg
        self.n = 1 << lg
        self.m = self.n // 16
        self.lpc = max(1, 4096 // self.n)
        self.lg2 = min(4, lg - 4)
        self.lg3 = lg - 4 - self.lg2

    # fields of a position
    def F(self, p):
        lg3, lg2 = self.lg3, self.lg2
        return bits(p, 0, lg3), bits(p, lg3, lg3 + lg2), bits(p, lg3 + lg2, self.lg)

    def mk(self, f3, f2, f1):
        return f3 | (f2 << self.lg3) | (f1 << (self.lg3 + self.lg2))

    def pos1(self, t, s):
        return t | (s << (self.lg - 4))

    def pos2(self, u, s):
        lg2, lg3 = self.lg2, self.lg3
        f2 = s >> (4 - lg2)
        f1hi = s & ((1 << (4 - lg2)) - 1)          # F1[lg2:4)
        # thread bits: F3 | F1[lg3:lg2) << lg3 | F1[0:lg3) << lg2
        f3 = bits(u, 0, lg3)
        f1_mid = bits(u, lg3, lg2)                 # F1[lg3:lg2)
        f1_lo = bits(u, lg2, lg2 + lg3)            # F1[0:lg3)
        f1 = f1_lo | (f1_mid << lg3) | (f1hi << lg2)
        return self.mk(f3, f2, f1)

    def pos3(self, v, s):
        lg3 = self.lg3
        if lg3 == 0:
            return self.pos2(v, s)
        f3 = s >> (4 - lg3)
        f2hi = s & ((1 << (4 - lg3)) - 1)          # F2[lg3:4)
        f1 = v & 15
        f2lo = bits(v, 4, 4 + lg3)
        return self.mk(f3, f2lo | (f2hi << lg3), f1)

    def xaddr(self, r, p):
        f3, f2, f1 = self.F(p)
        lgl = self.lpc.bit_length() - 1
        rs = (r << max(0, 4 - lgl)) & 15 if self.lpc > 1 else 0
        return r * self.n + (p ^ ((f1 ^ rs) & 15))


def dif_check(L):
    """run the layout-driven DIF on random data and compare with numpy"""
    n, m, lg = L.n, L.m, L.lg
    rng = np.random.default_rng(lg)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    mem = x.copy()                  # logical positions
    # stage 1
    new = np.empty_like(mem)
    for t in range(m):
        a = np.array([mem[L.pos1(t, s)] for s in range(16)])
        y = np.fft.fft(a)
        tw = np.exp(-2j * np.pi * t * np.arange(16) / n)
        y = y * tw
        for s in range(16):
            new[L.pos1(t, s)] = y[s]
    mem = new
    # stage 2
    r2 = 1 << L.lg2
    new = np.empty_like(mem)
    seen = set()
    for u in range(m):
        for e in range(16 // r2):
            idx = [L.pos2(u, (q << (4 - L.lg2)) | e) for q in range(r2)]
            seen.update(idx)
            a = mem[idx]
            y = np.fft.fft(a)
            f3 = L.F(idx[0])[0]
            tw = np.exp(-2j * np.pi * f3 * np.arange(r2) / (1 << (L.lg3 + L.lg2)))
            y = y * tw
            new[idx] = y
    assert len(seen) == n
    mem = new
    if L.lg3 > 0:
        r3 = 1 << L.lg3
        new = np.empty_like(mem)
        seen = set()
        for v in range(m):
            for e in range(16 // r3):
                idx = [L.pos3(v, (q << (4 - L.lg3)) | e) for q in range(r3)]
                seen.update(idx)
                new[idx] = np.fft.fft(mem[idx])
        assert len(seen) == n
        mem = new
    # final: thread v slot s holds k = v | s << (lg-4)
    ref = np.fft.fft(x)
    out = np.empty_like(mem)
    for v in range(m):
        for s in range(16):
            out[v | (s << (lg - 4))] = mem[L.pos3(v, s)]
    err = np.max(np.abs(out - ref)) / np.max(np.abs(ref))
    assert err < 1e-12, (lg, err)
    return err


def conflicts(addrs16):
    """addrs16: 8-byte word addresses of the 16 lanes of one half-warp; returns the conflict degree"""
    banks = {}
    for a in addrs16:
        banks.setdefault(a % 16, set()).add(a)
    return max(len(v) for v in banks.values())


def bank_check(L, rfast):
    worst = {}
    m, lpc = L.m, L.lpc

    def ids(tid):
        return (tid % lpc, tid // lpc) if rfast else (tid // m, tid % m)

    stages = [("w1", L.pos1), ("r2", L.pos2), ("w2", L.pos2), ("r3", L.pos3), ("w3", L.pos3)]
    for name, fn in stages:
        w = 1
        for hw in range(16):
            for s in range(16):
                addrs = []
                for lane in range(16):
                    r, t = ids(hw * 16 + lane)
                    addrs.append(L.xaddr(r, fn(t, s)))
                w = max(w, conflicts(addrs))
        worst[name] = w
    return worst


if __name__ == "__main__":
    ok = True
    for lg in range(5, 13):
        L = Layout(lg)
        err = dif_check(L)
        for rfast in (False, True):
            w = bank_check(L, rfast)
            bad = {k: v for k, v in w.items() if v > 1}
            print(f"lg={lg:2d} n={L.n:5d} m={L.m:4d} lpc={L.lpc:3d} rfast={int(rfast)} dif_err={err:.1e} conflicts={w}")
            if bad:
                ok = False
    sys.exit(0 if ok else 1)
