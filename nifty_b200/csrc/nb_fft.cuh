// nifty_b200 -- in-place power-of-two line FFTs in shared memory (forward sign, e^{-i theta}).
//
// A CTA holds `nlines` complex lines of length n = 2^lg (pitch `pitch` elements) in shared memory.
//   fft_dif : natural order in  -> digit-reversed order out   (decimation in frequency)
//   fft_dit : digit-reversed in -> natural order out          (decimation in time)
// Both use radix(m) for the stage whose sub-transform length is m, so that the output order of
// fft_dif is exactly the input order of fft_dit; `fft_pos(k, lg)` is the LOGICAL slot of element k.
//
// Design points (each one answers a measured stall, see profiles/):
//  * logical slot i lives at physical slot swz(i) = i ^ (xor of the 3-bit digits above the lowest):
//    every access pattern in which the 8 threads of a quarter-warp differ in one octal digit --
//    all radix-8 stages, and the digit-reversed reads of the combine / store phases -- is free of
//    bank conflicts for 16-byte elements, without padding (a 4096-point f64 line stays 64 KB, so
//    three CTAs fit one SM).  swz is GF(2)-linear: swz(b + (q << s)) = swz(b) ^ swz(q << s) when the
//    bits are disjoint, i.e. one swizzle per butterfly plus one XOR per element.
//  * the first DIF stage reads its inputs straight from global memory through a loader functor
//    (8 independent coalesced loads per thread) -- one shared-memory round trip and one barrier
//    less per line, and the prologue of a pass (amplitude gather etc.) is fused into it.
//  * one twiddle load per butterfly (w_m^j, coalesced, from a COMPACT per-stage table computed on the
//    host in extended precision: stage s owns the contiguous run w_m^0 .. w_m^{m/R-1}, 9 KB in all for
//    a 4096-point line, so the tables survive in the ~25 KB of L1 that three 64 KB CTAs leave);
//    w^2..w^7 are formed by multiplication (<= 3 roundings deep).
//  * every butterfly reads and writes the same r slots, so no thread carries registers across a
//    barrier (which is also what lets tests/emu run the same code sequentially on the host).
#pragma once
#include "nb_common.cuh"
#include <vector>

namespace nb {

// log2 of the radix used for a sub-transform of length 2^lm: 8 except for the 2/4/16 leftovers.
// order = 1 puts the radix-4 leftover of lengths 2^(3m+2) FIRST (4*8*8*.. instead of 8*8*..*4): the first
// stage is the one fused with the global loads, and a radix-4 batch needs half the registers.
NB_HH NB_INLINE int fft_radix_lg(int lm, int order = 0) {
  if (order == 1 && lm > 2 && lm % 3 == 2) return 2;
  return lm == 1 ? 1 : ((lm == 2 || lm == 4) ? 2 : 3);
}

// logical slot of output element k after fft_dif of length 2^lg (== slot where fft_dit expects input k)
NB_HH NB_INLINE int fft_pos(int k, int lg, int order = 0) {
  int pos = 0, lm = lg;
  while (lm > 0) {
    int lr = fft_radix_lg(lm, order);
    int p = k & ((1 << lr) - 1);
    k >>= lr;
    lm -= lr;
    pos += p << lm;
  }
  return pos;
}

// physical slot of logical slot i (within one line)
NB_HH NB_INLINE int swz(int i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9) ^ (i >> 12)) & 7); }

template <class T> struct FftConst;
template <> struct FftConst<double> { static NB_HD NB_INLINE double rsqrt2() { return 0.70710678118654752440; } };
template <> struct FftConst<float> { static NB_HD NB_INLINE float rsqrt2() { return 0.70710678118654752440f; } };

// y_p = sum_q a_q w_R^{pq},  R = 2, 4, 8 (forward sign), in place on the register array
template <class T> NB_HD NB_INLINE void dft2(cplx<T>* a) {
  cplx<T> t = a[0] - a[1];
  a[0] = a[0] + a[1];
  a[1] = t;
}
template <class T> NB_HD NB_INLINE void dft4(cplx<T>* a) {
  cplx<T> d0 = a[0] + a[2], d1 = a[0] - a[2], d2 = a[1] + a[3], d3 = cmul_mi(a[1] - a[3]);
  a[0] = d0 + d2;
  a[2] = d0 - d2;
  a[1] = d1 + d3;
  a[3] = d1 - d3;
}
template <class T> NB_HD NB_INLINE void dft8(cplx<T>* a) {
  const T h = FftConst<T>::rsqrt2();
  cplx<T> b[4], c[4];
  b[0] = a[0] + a[4];
  b[1] = a[1] + a[5];
  b[2] = a[2] + a[6];
  b[3] = a[3] + a[7];
  cplx<T> t0 = a[0] - a[4], t1 = a[1] - a[5], t2 = a[2] - a[6], t3 = a[3] - a[7];
  c[0] = t0;
  c[1] = cmake<T>((t1.x + t1.y) * h, (t1.y - t1.x) * h);    // * (1 - i)/sqrt2
  c[2] = cmul_mi(t2);                                       // * (-i)
  c[3] = cmake<T>((t3.y - t3.x) * h, -(t3.x + t3.y) * h);   // * (-1 - i)/sqrt2
  dft4(b);
  dft4(c);
  a[0] = b[0]; a[2] = b[1]; a[4] = b[2]; a[6] = b[3];
  a[1] = c[0]; a[3] = c[1]; a[5] = c[2]; a[7] = c[3];
}
template <int LR, class T> NB_HD NB_INLINE void dftR(cplx<T>* a) {
  if (LR == 1) dft2(a);
  else if (LR == 2) dft4(a);
  else dft8(a);
}

// a[q] *= w^q, q = 1..R-1, powers of w formed by multiplication
template <int LR, class T> NB_HD NB_INLINE void twiddle_apply(cplx<T>* a, cplx<T> w) {
  a[1] = cmul(a[1], w);
  if (LR >= 2) {
    cplx<T> w2 = cmul(w, w), w3 = cmul(w2, w);
    a[2] = cmul(a[2], w2);
    a[3] = cmul(a[3], w3);
    if (LR >= 3) {
      cplx<T> w4 = cmul(w2, w2);
      a[4] = cmul(a[4], w4);
      a[5] = cmul(a[5], cmul(w4, w));
      a[6] = cmul(a[6], cmul(w3, w3));
      a[7] = cmul(a[7], cmul(w4, w3));
    }
  }
}

// Host-built description of one line FFT (passed by value inside kernel parameters, i.e. read
// through the constant bank with uniform addresses): stage lengths / radices, the swizzled element
// offsets swz(q << lmr) of every stage, and a device table pos[k] = swz(fft_pos(k)).
typedef unsigned short fft_slot_t;     // physical slot within a line (lines are <= 2^14 elements: 227 KB of shared memory)
struct FftDev {
  int lg, ns;
  int lm[6], lr[6];
  int off[6][8];
  int tw_off[6];              // first entry of stage s in the compact twiddle table
  const fft_slot_t* pos;
  const void* ctw;            // cplx<T>[tw_off[ns-1] + ...]: stage s holds w_{2^lm}^j, j < 2^(lm-lr)
};
inline FftDev make_fft_dev(int lg, const fft_slot_t* pos_table, const void* ctw, int order = 0) {
  FftDev f;
  f.lg = lg; f.ns = 0; f.pos = pos_table; f.ctw = ctw;
  for (int s = 0; s < 6; ++s) { f.lm[s] = 0; f.lr[s] = 0; f.tw_off[s] = 0; for (int q = 0; q < 8; ++q) f.off[s][q] = 0; }
  int lm = lg, toff = 0;
  while (lm > 0) {
    int lr = fft_radix_lg(lm, order);
    f.lm[f.ns] = lm; f.lr[f.ns] = lr; f.tw_off[f.ns] = toff;
    toff += 1 << (lm - lr);
    for (int q = 0; q < (1 << lr); ++q) f.off[f.ns][q] = swz(q << (lm - lr));
    ++f.ns;
    lm -= lr;
  }
  return f;
}
inline void fill_pos_table(int lg, fft_slot_t* out, int order = 0) { for (int k = 0; k < (1 << lg); ++k) out[k] = (fft_slot_t)swz(fft_pos(k, lg, order)); }
// compact twiddle table of a line FFT of length 2^lg from the full table full[k] = w_{2^lg}^k
template <class C> inline void fill_compact_twiddles(int lg, const C* full, std::vector<C>& out, int order = 0) {
  out.clear();
  int lm = lg;
  while (lm > 0) {
    int lr = fft_radix_lg(lm, order);
    for (int j = 0; j < (1 << (lm - lr)); ++j) out.push_back(full[(size_t)j << (lg - lm)]);
    lm -= lr;
  }
}

// One radix-2^LR stage over all lines, data in shared memory (stage index st of `f`).
// DIF: y_p = (sum_q x_q w_R^{pq}) w_m^{jp};  DIT: y_p = sum_q (x_q w_m^{jq}) w_R^{pq}
template <int LR, bool DIT, class T>
NB_HD NB_INLINE void fft_stage(Ctx& ctx, cplx<T>* s, const FftDev& f, int st, int nlines, int pitch,
                               const cplx<T>* tw, int lg_tw) {
  constexpr int R = 1 << LR;
  const int lm = f.lm[st];
  const int lmr = lm - LR;              // log2(m / R)
  const int lbf = f.lg - LR;            // log2(butterflies per line)
  const int total = nlines << lbf;
  const cplx<T>* ctw = reinterpret_cast<const cplx<T>*>(f.ctw) + f.tw_off[st];
  (void)tw; (void)lg_tw;
  NB_FOR(ctx, t, total) {
    int line = t >> lbf, u = t & ((1 << lbf) - 1);
    int blk = u >> lmr, j = u & ((1 << lmr) - 1);
    cplx<T> w = cmake<T>(T(1), T(0));
    if (j != 0) w = ldg(ctw + j);
    cplx<T>* base = s + line * pitch;
    const int b0 = swz((blk << lm) + j);
    cplx<T> a[R];
#pragma unroll
    for (int q = 0; q < R; ++q) a[q] = base[b0 ^ f.off[st][q]];
    if (DIT && j != 0) twiddle_apply<LR>(a, w);
    dftR<LR>(a);
    if (!DIT && j != 0) twiddle_apply<LR>(a, w);
#pragma unroll
    for (int q = 0; q < R; ++q) base[b0 ^ f.off[st][q]] = a[q];
  }
  ctx.sync();
}

// First DIF stage (m = n) with the inputs taken from a loader: ld.batch<R>(line, j, lmr, a) fills
// a[q] = x[j + (q << lmr)], q < R, issuing all of its loads before any use
template <int LR, class T, class Ld>
NB_HD NB_INLINE void fft_stage_first(Ctx& ctx, cplx<T>* s, const FftDev& f, int nlines, int pitch,
                                     const cplx<T>* tw, int lg_tw, const Ld& ld) {
  constexpr int R = 1 << LR;
  const int lmr = f.lg - LR;
  const int total = nlines << lmr;
  const cplx<T>* ctw = reinterpret_cast<const cplx<T>*>(f.ctw);      // stage 0 starts the compact table
  (void)tw; (void)lg_tw;
  NB_FOR(ctx, t, total) {
    int line = t >> lmr, j = t & ((1 << lmr) - 1);
    cplx<T> a[R];
    ld.template batch<R>(line, j, lmr, a);
    cplx<T> w = cmake<T>(T(1), T(0));
    if (j != 0) w = ldg(ctw + j);
    dftR<LR>(a);
    if (j != 0) twiddle_apply<LR>(a, w);
    cplx<T>* base = s + line * pitch;
    const int b0 = swz(j);
#pragma unroll
    for (int q = 0; q < R; ++q) base[b0 ^ f.off[0][q]] = a[q];
  }
  ctx.sync();
}

template <bool DIT, class T>
NB_HD NB_INLINE void fft_stage_any(Ctx& ctx, cplx<T>* s, const FftDev& f, int st, int nlines, int pitch,
                                   const cplx<T>* tw, int lg_tw) {
  int lr = f.lr[st];
  if (lr == 3) fft_stage<3, DIT>(ctx, s, f, st, nlines, pitch, tw, lg_tw);
  else if (lr == 2) fft_stage<2, DIT>(ctx, s, f, st, nlines, pitch, tw, lg_tw);
  else fft_stage<1, DIT>(ctx, s, f, st, nlines, pitch, tw, lg_tw);
}

// natural order from `ld` -> digit-reversed (swizzled) in shared memory; returns synchronised.
template <class T, class Ld>
NB_HD NB_INLINE void fft_dif_load(Ctx& ctx, cplx<T>* s, const FftDev& f, int nlines, int pitch, const cplx<T>* tw,
                                  int lg_tw, const Ld& ld) {
  if (f.lg == 0) {
    NB_FOR(ctx, t, nlines) { cplx<T> a[1]; ld.template batch<1>(t, 0, 0, a); s[t * pitch] = a[0]; }
    ctx.sync();
    return;
  }
  int lr = f.lr[0];
  if (lr == 3) fft_stage_first<3>(ctx, s, f, nlines, pitch, tw, lg_tw, ld);
  else if (lr == 2) fft_stage_first<2>(ctx, s, f, nlines, pitch, tw, lg_tw, ld);
  else fft_stage_first<1>(ctx, s, f, nlines, pitch, tw, lg_tw, ld);
  for (int st = 1; st < f.ns; ++st) fft_stage_any<false>(ctx, s, f, st, nlines, pitch, tw, lg_tw);
}

// digit-reversed (swizzled) -> natural (swizzled); caller synchronised after filling; returns synchronised
template <class T>
NB_HD NB_INLINE void fft_dit(Ctx& ctx, cplx<T>* s, const FftDev& f, int nlines, int pitch, const cplx<T>* tw, int lg_tw) {
  for (int st = f.ns - 1; st >= 0; --st) fft_stage_any<true>(ctx, s, f, st, nlines, pitch, tw, lg_tw);
}

}  // namespace nb
