// nifty_b200 -- in-place power-of-two line FFTs in shared memory (forward sign, e^{-i theta}).
//
// A CTA holds `nlines` complex lines of length n = 2^lg (pitch `pitch` elements) in shared memory.
//   fft_dif : natural order in  -> digit-reversed order out   (decimation in frequency)
//   fft_dit : digit-reversed in -> natural order out          (decimation in time)
// Both use radix(m) for the stage whose sub-transform length is m, so that the output order of
// fft_dif is exactly the input order of fft_dit; `fft_pos(k, lg)` is the slot of element k.
// Every butterfly reads and writes the same r slots, so no thread carries registers across a
// barrier (which is also what lets tests/emu run the same code sequentially on the host).
// Twiddles come from a table tw[j] = exp(-2 pi i j / tw_n) (tw_n a multiple of n) computed on the
// host in extended precision.
#pragma once
#include "nb_common.cuh"

namespace nb {

// log2 of the radix used for a sub-transform of length 2^lm: 8 except for the 2/4/16 leftovers.
NB_HH NB_INLINE int fft_radix_lg(int lm) { return lm == 1 ? 1 : ((lm == 2 || lm == 4) ? 2 : 3); }

// slot of output element k after fft_dif of length 2^lg (== slot where fft_dit expects input k)
NB_HH NB_INLINE int fft_pos(int k, int lg) {
  int pos = 0, lm = lg;
  while (lm > 0) {
    int lr = fft_radix_lg(lm);
    int p = k & ((1 << lr) - 1);
    k >>= lr;
    lm -= lr;
    pos += p << lm;
  }
  return pos;
}

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

// One radix-2^LR stage over all lines.  lm = log2 of the sub-transform length of this stage.
// DIF: y_p = (sum_q x_q w_R^{pq}) w_m^{jp};   DIT: y_p = sum_q (x_q w_m^{jq}) w_R^{pq}
template <int LR, bool DIT, class T>
NB_HD NB_INLINE void fft_stage(Ctx& ctx, cplx<T>* s, int lg, int lm, int nlines, int pitch,
                               const cplx<T>* tw, int tw_shift /* log2(tw_n) - lm */) {
  constexpr int R = 1 << LR;
  const int lmr = lm - LR;              // log2(m / R)
  const int lbf = lg - LR;              // log2(butterflies per line)
  const int total = nlines << lbf;
  NB_FOR(ctx, t, total) {
    int line = t >> lbf, u = t & ((1 << lbf) - 1);
    int blk = u >> lmr, j = u & ((1 << lmr) - 1);
    cplx<T>* base = s + (size_t)line * pitch + (blk << lm) + j;
    cplx<T> a[R];
#pragma unroll
    for (int q = 0; q < R; ++q) a[q] = base[q << lmr];
    if (DIT && j != 0) {
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = cmul(a[q], ldg(tw + ((size_t)(j * q) << tw_shift)));
    }
    dftR<LR>(a);
    if (!DIT && j != 0) {
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = cmul(a[q], ldg(tw + ((size_t)(j * q) << tw_shift)));
    }
#pragma unroll
    for (int q = 0; q < R; ++q) base[q << lmr] = a[q];
  }
  ctx.sync();
}

template <bool DIT, class T>
NB_HD NB_INLINE void fft_stage_any(Ctx& ctx, cplx<T>* s, int lg, int lm, int nlines, int pitch,
                                   const cplx<T>* tw, int lg_tw) {
  int lr = fft_radix_lg(lm);
  if (lr == 3) fft_stage<3, DIT>(ctx, s, lg, lm, nlines, pitch, tw, lg_tw - lm);
  else if (lr == 2) fft_stage<2, DIT>(ctx, s, lg, lm, nlines, pitch, tw, lg_tw - lm);
  else fft_stage<1, DIT>(ctx, s, lg, lm, nlines, pitch, tw, lg_tw - lm);
}

// natural -> digit-reversed.  Caller must have synchronised after filling `s`; returns synchronised.
template <class T>
NB_HD NB_INLINE void fft_dif(Ctx& ctx, cplx<T>* s, int lg, int nlines, int pitch, const cplx<T>* tw, int lg_tw) {
  int lm = lg;
  while (lm > 0) {
    fft_stage_any<false>(ctx, s, lg, lm, nlines, pitch, tw, lg_tw);
    lm -= fft_radix_lg(lm);
  }
}

// digit-reversed -> natural
template <class T>
NB_HD NB_INLINE void fft_dit(Ctx& ctx, cplx<T>* s, int lg, int nlines, int pitch, const cplx<T>* tw, int lg_tw) {
  // the stage lengths of fft_dif in reverse order
  int lms[32], ns = 0, lm = lg;
  while (lm > 0) { lms[ns++] = lm; lm -= fft_radix_lg(lm); }
  for (int i = ns - 1; i >= 0; --i) fft_stage_any<true>(ctx, s, lg, lms[i], nlines, pitch, tw, lg_tw);
}

}  // namespace nb
