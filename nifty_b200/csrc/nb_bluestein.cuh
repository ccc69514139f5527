// nifty_b200 -- Hartley transform of grids whose extents are NOT powers of two (the reference's `hartley`,
// nifty/re/correlated_field.py:24-30, on shapes like the (3, 3) of test/test_re/test_correlated_field.py:123-124 or the
// 3618^2 of its published benchmark) as a chirp convolution around the power-of-two passes of this library.
//
// With c_j = exp(i pi j^2 / n) an n-point DFT is  X_k = conj(c_k) sum_j (x_j conj(c_j)) c_{k-j}, a cyclic convolution on any
// padded power-of-two length M >= 2 n - 1; in several dimensions chirp and filter are outer products of per-axis tables.  A
// complex transform of the padded grid is two real Hartley transforms recombined with their index reflections, so one
// transform of the n-grid is
//   ChirpPad    x (n-grid)            -> Re / Im of x conj(c), zero-padded           (writes 2 M-grids)
//   hartley x 2                          in place, the fused passes of nb_passes.cuh
//   ChirpFilter reflect-combine to the complex spectrum, multiply by the filter spectrum, conjugate; mirror pairs (k, -k)
//               are handled by one thread, in place                                   (reads 2, writes 2 M-grids)
//   hartley x 2
//   ChirpCrop   reflect-combine on the n-grid corner, conj(c), Re +/- Im              (reads 4 n-grids, writes 1)
// All three bodies are streaming element-wise kernels (HBM-bound; algorithmic bytes in the lines above).
#pragma once
#include "nb_common.cuh"

namespace nb {

template <class T> struct ChirpParams {
  long n[3], m[3];          // logical / padded extents, right-aligned (missing leading axes = 1)
  int lg_m[3];
  long ntot, mtot;
  const cplx<T>* cc[3];     // conj(c_j), n[a] entries per axis
  const cplx<T>* G[3];      // spectrum of the chirp filter per axis, m[a] entries; the 1 / mtot of the inverse transform is folded into G[2]
  const T* x;               // ChirpPad input (n-grid)
  T* w;                     // work: two M-grids
  T* out;                   // ChirpCrop output (n-grid)
  T s;                      // +1: Re + Im (non_canonical_hartley), -1: Re - Im
};

template <class T> struct ChirpPadBody {
  typedef ChirpParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    const long stride = (long)ctx.nblk * ctx.nthr;
    T* re = p.w; T* im = p.w + p.mtot;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.mtot; i += stride) {
      long i2 = i & (p.m[2] - 1), r = i >> p.lg_m[2];
      long i1 = r & (p.m[1] - 1), i0 = r >> p.lg_m[1];
      T vr = 0, vi = 0;
      if (i2 < p.n[2] && i1 < p.n[1] && i0 < p.n[0]) {
        cplx<T> c = cmul(cmul(p.cc[0][i0], p.cc[1][i1]), p.cc[2][i2]);
        T v = p.x[(i0 * p.n[1] + i1) * p.n[2] + i2];
        vr = v * c.x; vi = v * c.y;
      }
      re[i] = vr; im[i] = vi;
    }
  }
};

// (H_re, H_im) -> (P, Q) = (Re, -Im) of [F(re) + i F(im)] G, where for a real array F = (H + H o f)/2 + i s (H - H o f)/2
template <class T> struct ChirpFilterBody {
  typedef ChirpParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    const long stride = (long)ctx.nblk * ctx.nthr;
    T* hr = p.w; T* hi = p.w + p.mtot;
    // every pair (k, -k) has its smaller linear index in the lower half of the first axis that exists
    const long half = p.m[0] > 1 ? (p.m[0] / 2 + 1) * p.m[1] * p.m[2] : (p.m[1] > 1 ? (p.m[1] / 2 + 1) * p.m[2] : p.m[2] / 2 + 1);
    const T hs = T(0.5) * p.s;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < half; i += stride) {
      long i2 = i & (p.m[2] - 1), r = i >> p.lg_m[2];
      long i1 = r & (p.m[1] - 1), i0 = r >> p.lg_m[1];
      long j0 = (p.m[0] - i0) & (p.m[0] - 1), j1 = (p.m[1] - i1) & (p.m[1] - 1), j2 = (p.m[2] - i2) & (p.m[2] - 1);
      long ip = (((j0 << p.lg_m[1]) + j1) << p.lg_m[2]) + j2;
      if (ip < i) continue;
      T a = hr[i], ap = hr[ip], b = hi[i], bp = hi[ip];
      T frr = T(0.5) * (a + ap), fri = hs * (a - ap), fir = T(0.5) * (b + bp), fii = hs * (b - bp);
      cplx<T> g = cmul(cmul(p.G[0][i0], p.G[1][i1]), p.G[2][i2]);       // G(-k) = G(k): the filter is even along every axis
      cplx<T> B = cmul(cmake<T>(frr - fii, fri + fir), g);
      hr[i] = B.x; hi[i] = -B.y;
      if (ip != i) {
        cplx<T> Bp = cmul(cmake<T>(frr + fii, fir - fri), g);
        hr[ip] = Bp.x; hi[ip] = -Bp.y;
      }
    }
  }
};

template <class T> struct ChirpCropBody {
  typedef ChirpParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void*) {
    const long stride = (long)ctx.nblk * ctx.nthr;
    const T* hp = p.w; const T* hq = p.w + p.mtot;
    const T hs = T(0.5) * p.s;
    for (long i = (long)ctx.bid * ctx.nthr + ctx.tid; i < p.ntot; i += stride) {
      long r = i / p.n[2], i2 = i - r * p.n[2];
      long i0 = r / p.n[1], i1 = r - i0 * p.n[1];
      long j0 = (p.m[0] - i0) & (p.m[0] - 1), j1 = (p.m[1] - i1) & (p.m[1] - 1), j2 = (p.m[2] - i2) & (p.m[2] - 1);
      long l = (((i0 << p.lg_m[1]) + i1) << p.lg_m[2]) + i2, lp = (((j0 << p.lg_m[1]) + j1) << p.lg_m[2]) + j2;
      T a = hp[l], ap = hp[lp], b = hq[l], bp = hq[lp];
      T fpr = T(0.5) * (a + ap), fpi = hs * (a - ap), fqr = T(0.5) * (b + bp), fqi = hs * (b - bp);
      cplx<T> z = cmake<T>(fpr - fqi, -(fpi + fqr));                     // conj(F(P) + i F(Q))
      cplx<T> c = cmul(cmul(p.cc[0][i0], p.cc[1][i1]), p.cc[2][i2]);
      cplx<T> X = cmul(z, c);
      p.out[i] = X.x + p.s * X.y;
    }
  }
};

}  // namespace nb
