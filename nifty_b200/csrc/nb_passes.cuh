// nifty_b200 -- axis passes of the n-D Hartley transform and of the fused metric-vector product.
//
// Reference semantics: nifty/re/correlated_field.py:24-30 (`hartley` = Re(fftn) +/- Im(fftn)),
// :882-887 (outer_harmonic_transform), :909-912 (correlated_field) and, for the fused product,
// nifty/re/likelihood.py:613-621 (LikelihoodWithModel.metric).  Nothing here follows the
// reference's implementation (jnp.fft.fftn of a complexified array); the decomposition is:
//
//   P1  real lines along the last axis  -> half spectrum (n/2+1), stored TRANSPOSED so that the
//       next axis is contiguous (a CTA owns R adjacent lines and emits R-element pieces)
//   PC  complex lines along a middle axis (3-D only), same transposed store
//   P3  lines along axis 0.  The complex FFT of a line (k_last, k_mid) yields, by Hermitian
//       symmetry, the two REAL position-space lines (k_last,k_mid) and (-k_last,-k_mid).  A
//       pointwise operator acts on them (x 1/V, exp, noise weight ...), the two real lines are
//       re-packed as one complex line and transformed again (first axis of the adjoint
//       transform), split into the two half spectra and stored transposed.  Forward-only and
//       adjoint-only variants exist for linearisation / JVP / VJP.
//   P5  complex lines along the last axis + Hermitian combine -> two natural real rows, with the
//       latent-space epilogue (amplitude gather, + t, mode-bin partial sums, dot products).
//
// Position-space arrays therefore live in REVERSED axis order ("T-layout": [x_last]..[x_0],
// x_0 contiguous); latent-space (harmonic) arrays are in the reference's natural C order.
// A 2-D product is P1 -> P3 -> P5 (10 array streams), a 3-D one P1 -> PC -> P3 -> PC -> P5 (14).
#pragma once
#include "nb_common.cuh"
#include "nb_fft.cuh"
#include "nb_fft16.cuh"

namespace nb {

// pair load: one 2*sizeof(T)-byte access when the caller guarantees alignment (compile time)
template <bool ALIGNED, class T> NB_HD NB_INLINE void load_pair(const T* p, T& a, T& b) {
  if (ALIGNED) {
    cplx<T> v = ld_stream(reinterpret_cast<const cplx<T>*>(p));
    a = v.x; b = v.y;
  } else {
    a = ld_stream(p); b = ld_stream(p + 1);
  }
}

// folded mode-bin lookup shared by the amplitude prologues / epilogues
struct FoldGeom {
  int n_o, n_r, n;     // extents of (outer, row, last) axes of the natural array
  int hr1, h1;         // n_r/2+1, n/2+1
  const int* plane_loc;   // slab-decomposed grids: folded plane of the LOCAL bin table for local row o (else null)
  NB_HD NB_INLINE long base(int o, int rr) const {
    int pl = plane_loc ? ldg(plane_loc + o) : fold_idx(o, n_o);
    return ((long)pl * hr1 + fold_idx(rr, n_r)) * h1;
  }
};

// ---------------------------------------------------------------------------------------------
// P1 prologues: `batch<NQ, ALIGNED>` produces the NQ complex inputs z = x[2e] + i x[2e+1] of one
// butterfly (elements e_q = j + (q << lmr)) of the real line (o, rr).  All loads of a batch are
// issued before anything depends on them (first the streaming loads and bin indices, then the
// table gathers), with no control flow in between -- the per-element dependent chain
// index -> amplitude used to be paid 16 times in a row per thread.
// ---------------------------------------------------------------------------------------------
template <class T> NB_HH NB_INLINE bool ptr_aligned2(const T* p) { return (reinterpret_cast<uintptr_t>(p) & (2 * sizeof(T) - 1)) == 0; }

template <class T> struct ProPlain {
  static constexpr bool kUsesBins = false;
  static constexpr bool kStaged = false;
  const T* x;
  bool aligned() const { return ptr_aligned2(x); }
  template <int NQ, bool AL> NB_HD NB_INLINE void batch(int, int, long off, int j, int lmr, cplx<T>* a) const {
#pragma unroll
    for (int q = 0; q < NQ; ++q) { T u, v; load_pair<AL>(x + off + 2 * (j + (q << lmr)), u, v); a[q] = cmake<T>(u, v); }
  }
  NB_HD NB_INLINE void prefetch(Ctx& ctx, int, int, long off, long count) const { prefetch_l2(ctx, x + off, count * sizeof(T)); }
  // mirror-quad interface of P1MBody (see there)
  struct Pre { T x0, x1, x2, x3; };
  NB_HD NB_INLINE void preload(long offA, long offB, const int*, int e, int ne, Pre& q) const {
    q.x0 = ld_stream(x + offA + e); q.x1 = ld_stream(x + offA + ne); q.x2 = ld_stream(x + offB + e); q.x3 = ld_stream(x + offB + ne);
  }
  NB_HD NB_INLINE void gather(Pre&) const {}
  NB_HD NB_INLINE void finish(const Pre& q, T* v) const { v[0] = q.x0; v[1] = q.x1; v[2] = q.x2; v[3] = q.x3; }
  NB_HD NB_INLINE T single(long off, const int*, int e) const { return x[off + e]; }
  NB_HD NB_INLINE const int* bins(int, int) const { return nullptr; }
};

// a[b(k)] * xi_k   (forward model, correlated_field.py:911 with azm folded into the table)
template <class T> struct ProAmp {
  static constexpr bool kUsesBins = true;
  static constexpr bool kStaged = false;     // linearisation only (once per solve): generic body
  const T* xi; const int* idxf; const T* amp; FoldGeom fg;
  bool aligned() const { return ptr_aligned2(xi); }
  template <int NQ, bool AL> NB_HD NB_INLINE void batch(int o, int rr, long off, int j, int lmr, cplx<T>* a) const {
    const int* ip = idxf + fg.base(o, rr);
    int b0[NQ], b1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      int e = 2 * (j + (q << lmr));
      T u, v; load_pair<AL>(xi + off + e, u, v); a[q] = cmake<T>(u, v);
      b0[q] = ldg(ip + fold_idx(e, fg.n)); b1[q] = ldg(ip + fold_idx(e + 1, fg.n));
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) { a[q].x *= ldg(amp + b0[q]); a[q].y *= ldg(amp + b1[q]); }
  }
  NB_HD NB_INLINE void prefetch(Ctx& ctx, int o, int rr, long off, long count) const {
    prefetch_l2(ctx, xi + off, count * sizeof(T));
    prefetch_l2(ctx, idxf + fg.base(o, rr), (size_t)fg.h1 * sizeof(int));
  }
  struct Pre { int b; T A, x0, x1, x2, x3; };
  NB_HD NB_INLINE void preload(long offA, long offB, const int* ip, int e, int ne, Pre& q) const {
    q.b = ldg(ip + e);
    q.x0 = ld_stream(xi + offA + e); q.x1 = ld_stream(xi + offA + ne); q.x2 = ld_stream(xi + offB + e); q.x3 = ld_stream(xi + offB + ne);
  }
  NB_HD NB_INLINE void gather(Pre& q) const { q.A = ldg(amp + q.b); }
  NB_HD NB_INLINE void finish(const Pre& q, T* v) const { v[0] = q.A * q.x0; v[1] = q.A * q.x1; v[2] = q.A * q.x2; v[3] = q.A * q.x3; }
  NB_HD NB_INLINE T single(long off, const int* ip, int e) const { return ldg(amp + ldg(ip + fold_idx(e, fg.n))) * xi[off + e]; }
  NB_HD NB_INLINE const int* bins(int o, int rr) const { return idxf + fg.base(o, rr); }
};

// JVP input: A_b t_k + dA_b xi_k with dA_b = A_b (cj + kappa du_b) (b>=1), dA_0 = da0.
// `ad` interleaves (A_b, du_b) so that one 2*sizeof(T) gather serves both.
template <class T> struct ProMetric {
  static constexpr bool kUsesBins = true;
  static constexpr bool kStaged = true;      // the tangent prologue of every metric-vector product
  const T* xi; const T* t; const int* idxf; const cplx<T>* ad; const T* scal;  // scal[0]=cj, scal[1]=da0
  T kappa; FoldGeom fg;
  bool aligned() const { return ptr_aligned2(xi) && ptr_aligned2(t); }
  NB_HD NB_INLINE T one(int b, cplx<T> g, T xv, T tv, T cj, T da0) const {
    T dA = (b == 0) ? da0 : g.x * (cj + kappa * g.y);
    return g.x * tv + dA * xv;
  }
  template <int NQ, bool AL> NB_HD NB_INLINE void batch(int o, int rr, long off, int j, int lmr, cplx<T>* a) const {
    const int* ip = idxf + fg.base(o, rr);
    const T cj = ldg(scal), da0 = ldg(scal + 1);
    int b0[NQ], b1[NQ];
    cplx<T> tv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      int e = 2 * (j + (q << lmr));
      T u, v; load_pair<AL>(xi + off + e, u, v); a[q] = cmake<T>(u, v);
      load_pair<AL>(t + off + e, u, v); tv[q] = cmake<T>(u, v);
      b0[q] = ldg(ip + fold_idx(e, fg.n)); b1[q] = ldg(ip + fold_idx(e + 1, fg.n));
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      cplx<T> g0 = ldg(ad + b0[q]), g1 = ldg(ad + b1[q]);
      a[q] = cmake<T>(one(b0[q], g0, a[q].x, tv[q].x, cj, da0), one(b1[q], g1, a[q].y, tv[q].y, cj, da0));
    }
  }
  NB_HD NB_INLINE void prefetch(Ctx& ctx, int o, int rr, long off, long count) const {
    prefetch_l2(ctx, xi + off, count * sizeof(T));
    prefetch_l2(ctx, t + off, count * sizeof(T));
    prefetch_l2(ctx, idxf + fg.base(o, rr), (size_t)fg.h1 * sizeof(int));
  }
  struct Pre { int b; cplx<T> g; T x0, x1, x2, x3, t0, t1, t2, t3; };
  NB_HD NB_INLINE void preload(long offA, long offB, const int* ip, int e, int ne, Pre& q) const {
    q.b = ldg(ip + e);
    q.x0 = ld_stream(xi + offA + e); q.x1 = ld_stream(xi + offA + ne); q.x2 = ld_stream(xi + offB + e); q.x3 = ld_stream(xi + offB + ne);
    q.t0 = ld_stream(t + offA + e); q.t1 = ld_stream(t + offA + ne); q.t2 = ld_stream(t + offB + e); q.t3 = ld_stream(t + offB + ne);
  }
  NB_HD NB_INLINE void gather(Pre& q) const { q.g = ldg(ad + q.b); }
  NB_HD NB_INLINE void finish(const Pre& q, T* v) const {
    const T cj = ldg(scal), da0 = ldg(scal + 1);
    const T dA = (q.b == 0) ? da0 : q.g.x * (cj + kappa * q.g.y);
    v[0] = q.g.x * q.t0 + dA * q.x0; v[1] = q.g.x * q.t1 + dA * q.x1;
    v[2] = q.g.x * q.t2 + dA * q.x2; v[3] = q.g.x * q.t3 + dA * q.x3;
  }
  NB_HD NB_INLINE T single(long off, const int* ip, int e) const {
    int b = ldg(ip + fold_idx(e, fg.n));
    return one(b, ldg(ad + b), xi[off + e], t[off + e], ldg(scal), ldg(scal + 1));
  }
  NB_HD NB_INLINE const int* bins(int o, int rr) const { return idxf + fg.base(o, rr); }
};

template <class T, class Pro> struct P1Params {
  int lg_n, lg_R;
  int n_r, n_o;
  long in_ostride, in_rstride, out_ostride, out_kstride;
  int pitch;
  const cplx<T>* tw; int lg_tw;   // table for the REAL length n (tw_n multiple of n)
  FftDev fft;                     // complex FFT of length n/2
  cplx<T>* out;
  int ahead;                      // CTAs resident on the device: prefetch distance (0 = off)
  Pro pro;
};

template <class T, class Pro, bool AL> struct P1Body {
  typedef P1Params<T, Pro> Params;
#ifdef NB_P1_MINB
  static constexpr int kMinBlocks = NB_P1_MINB;
#endif
  struct Loader {
    const Params& p; int o, rr0; long in0;
    template <int NQ> NB_HD NB_INLINE void batch(int r, int j, int lmr, cplx<T>* a) const {
      const int rr = rr0 + r;
      p.pro.template batch<NQ, AL>(o, rr, in0 + rr * p.in_rstride, j, lmr, a);
    }
  };
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem);
    const int R = 1 << p.lg_R, lg_h = p.lg_n - 1, h = 1 << lg_h;
    const int gpo = p.n_r >> p.lg_R;          // R divides n_r (both powers of two, checked by the host)
    const int o = ctx.bid / gpo, rr0 = (ctx.bid % gpo) << p.lg_R;
    const long in0 = o * p.in_ostride;
    if (p.ahead > 0 && ctx.bid + p.ahead < ctx.nblk) {   // inputs of the CTA that will follow on this SM
      const int b2 = ctx.bid + p.ahead, o2 = b2 / gpo, r2 = (b2 % gpo) << p.lg_R;
      if (p.in_rstride == (1 << p.lg_n)) p.pro.prefetch(ctx, o2, r2, o2 * p.in_ostride + r2 * p.in_rstride, (long)R << p.lg_n);
    }
    Loader ld{p, o, rr0, in0};
    fft_dif_load(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw, ld);   // w_h^j = w_n^{2j}: same table, larger stride
    const T half = T(0.5);
    cplx<T>* outp = p.out + o * p.out_ostride + rr0;
    const int tsh = p.lg_tw - p.lg_n;
    constexpr int U = 4;
    const int total = (h + 1) << p.lg_R;
    for (int i0 = ctx.tid; i0 < total; i0 += ctx.nthr * U) {
      int pk[U], pc[U];
      cplx<T> w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int i = i0 + u * ctx.nthr;
        i = i < total ? i : 0;
        int k = i >> p.lg_R;
        pk[u] = ldg(p.fft.pos + (k & (h - 1)));
        pc[u] = ldg(p.fft.pos + ((h - k) & (h - 1)));
        w[u] = ldg(p.tw + ((size_t)k << tsh));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int i = i0 + u * ctx.nthr;
        if (i >= total) break;
        int k = i >> p.lg_R, r = i & (R - 1);
        const cplx<T>* line = s + r * p.pitch;
        cplx<T> zk = line[pk[u]];
        cplx<T> zc = cconj(line[pc[u]]);
        cplx<T> e = zk + zc, d = zk - zc;
        cplx<T> wd = cmul_mi(cmul(w[u], d));   // -i w (zk - zc)
        outp[k * p.out_kstride + r] = cmake<T>(half * (e.x + wd.x), half * (e.y + wd.y));
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// P1M: P1 for prologues that look up the mode-bin table.  Rows rr and n_r - rr of one plane and the
// elements e and n - e of a row all fall into the SAME bin, so a CTA owns R/2 mirror pairs of rows
// (pair i >= 1: rows i and n_r - i; pair 0: the two self-mirrored rows 0 and n_r/2) and fetches ONE
// bin index and ONE table entry per quad {e, n-e} x {row, mirror row}.  ncu (profiles/r2_notes.md):
// with one dependent 16-byte gather per element P1 moved 917 MB from L2 to the SMs for 541 MB of
// payload (every gather drags a 32-byte sector) and spent 55 % of its samples waiting in the load
// phase.  The scaled inputs are written to shared memory first (the first FFT stage is therefore not
// fused with the loads here); everything after the load phase is P1Body's.
// ---------------------------------------------------------------------------------------------
// MINB = CTAs per SM the register budget is cut for: 3 pays when the lines are short (256^3: 138 vs 156 us),
// 2 when three CTAs' shared memory would squeeze L1 (4096^2: 204 vs 229 us) -- chosen by the host.
template <class T, class Pro, int MINB> struct P1MBody {
  typedef P1Params<T, Pro> Params;
  static constexpr int kMinBlocks = MINB;
  // row of line r of the CTA whose first pair is i0 (R/2 pairs: lines [0,R/2) are the rows i, lines [R/2,R) their mirrors)
  static NB_HD NB_INLINE int row_of(int r, int i0, int hR, int n_r) {
    int i = i0 + (r < hR ? r : r - hR);
    if (r < hR) return i;
    return i == 0 ? (n_r >> 1) : n_r - i;
  }
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem);
    T* sr = reinterpret_cast<T*>(smem);
    const int R = 1 << p.lg_R, hR = R >> 1, lg_hR = p.lg_R - 1;
    const int n = 1 << p.lg_n, lg_h = p.lg_n - 1, h = 1 << lg_h;
    const int gpo = p.n_r >> p.lg_R;
    const int o = ctx.bid / gpo, i0 = (ctx.bid % gpo) << lg_hR;
    const long in0 = o * p.in_ostride;
    // real element x of line L lives at sr[2 * (L * pitch + swz(x >> 1)) + (x & 1)]
#define NB_P1M_SLOT(L, x) ((((L) * p.pitch + swz((x) >> 1)) << 1) + ((x) & 1))
    {   // quads (e, n-e), 0 < e < h, of the pairs i >= 1
      constexpr int U = MINB == 2 ? 4 : 2;     // measured: 4 quads per round pay with the 128-register budget only
      const int cnt = hR << lg_h;
      struct Slot { int rp, e; bool ok; typename Pro::Pre pre; };
      auto issue = [&](int q0, Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int q = q0 + u * ctx.nthr;
          bool in = q < cnt;
          q = in ? q : 0;
          int rp = q >> lg_h, e = q & (h - 1);
          int i = i0 + rp;
          sl[u].ok = in && i != 0 && e != 0;
          sl[u].rp = rp; sl[u].e = e;
          e = e != 0 ? e : 1;
          int ra = (i != 0) ? i : 1, rb = p.n_r - ra;          // safe rows for the masked-out slots
          p.pro.preload(in0 + ra * p.in_rstride, in0 + rb * p.in_rstride, p.pro.bins(o, ra), e, n - e, sl[u].pre);
        }
      };
      auto consume = [&](int, const Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!sl[u].ok) continue;
          const int rp = sl[u].rp, e = sl[u].e;
          T v[4];
          p.pro.finish(sl[u].pre, v);
          sr[NB_P1M_SLOT(rp, e)] = v[0]; sr[NB_P1M_SLOT(rp, n - e)] = v[1];
          sr[NB_P1M_SLOT(rp + hR, e)] = v[2]; sr[NB_P1M_SLOT(rp + hR, n - e)] = v[3];
        }
      };
      auto gather = [&](Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) p.pro.gather(sl[u].pre);
      };
      batched_loop<false, U, Slot>(ctx, cnt, issue, gather, consume);
    }
    {   // self-mirrored elements e = 0 and e = h of every line
      NB_FOR(ctx, k, 2 * R) {
        int r = k >> 1, e = (k & 1) ? h : 0;
        int rr = row_of(r, i0, hR, p.n_r);
        sr[NB_P1M_SLOT(r, e)] = p.pro.single(in0 + rr * p.in_rstride, p.pro.bins(o, rr), e);
      }
    }
    if (i0 == 0) {   // pair 0 = rows 0 and n_r/2: different bins, no sharing (one CTA per plane)
      NB_FOR(ctx, k, 2 * (n - 2)) {
        int r = (k & 1) ? hR : 0, x = (k >> 1) + 1;
        if (x >= h) ++x;                                  // x in [1, n) without h
        int rr = row_of(r, 0, hR, p.n_r);
        sr[NB_P1M_SLOT(r, x)] = p.pro.single(in0 + rr * p.in_rstride, p.pro.bins(o, rr), x);
      }
    }
#undef NB_P1M_SLOT
    ctx.sync();
    for (int st = 0; st < p.fft.ns; ++st) fft_stage_any<false>(ctx, s, p.fft, st, R, p.pitch, p.tw, p.lg_tw);
    // split of the half-length complex FFT Z into the spectrum of the real line:
    //   X[k] = (e - i w^k d)/2,  X[h-k] = conj(e + i w^k d)/2   with e = Z[k] + conj Z[h-k], d = Z[k] - conj Z[h-k],
    // i.e. both outputs of the pair (k, h-k) come from one pair of shared-memory reads, one twiddle and
    // one complex product; k runs over [0, h/2] (k = h/2 pairs with itself, k = 0 with h).
    const T half = T(0.5);
    cplx<T>* outp = p.out + o * p.out_ostride;
    const int tsh = p.lg_tw - p.lg_n;
    constexpr int U = 2;
    const int hh = h >> 1;
    const int total = (hh + 1) << p.lg_R;
    struct PSlot { int pk, pc; cplx<T> w; };
    auto pissue = [&](int j0, PSlot* sl) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int i = j0 + u * ctx.nthr;
        i = i < total ? i : 0;
        int k = i >> p.lg_R;
        sl[u].pk = ldg(p.fft.pos + (k & (h - 1)));
        sl[u].pc = ldg(p.fft.pos + ((h - k) & (h - 1)));
        sl[u].w = ldg(p.tw + ((size_t)k << tsh));
      }
    };
    auto pconsume = [&](int j0, const PSlot* sl) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int i = j0 + u * ctx.nthr;
        if (i >= total) break;
        int k = i >> p.lg_R, r = i & (R - 1);
        const cplx<T>* line = s + r * p.pitch;
        cplx<T> zk = line[sl[u].pk];
        cplx<T> zc = cconj(line[sl[u].pc]);
        cplx<T> e = zk + zc, d = zk - zc;
        cplx<T> m = cmul_mi(cmul(sl[u].w, d));   // -i w (zk - zc)
        const int row = row_of(r, i0, hR, p.n_r);
        outp[k * p.out_kstride + row] = cmake<T>(half * (e.x + m.x), half * (e.y + m.y));
        if (h - k != k) outp[(h - k) * p.out_kstride + row] = cmake<T>(half * (e.x - m.x), -half * (e.y - m.y));
      }
    };
    batched_loop<false, U, PSlot>(ctx, total, pissue, [](PSlot*) {}, pconsume);
  }
};

// ---------------------------------------------------------------------------------------------
// PC: complex lines, FFT, transposed store
// ---------------------------------------------------------------------------------------------
template <class T> struct PCParams {
  int lg_n, lg_R;
  int n_r, n_o;
  long in_ostride, in_rstride, out_ostride, out_kstride;
  int pitch;
  const cplx<T>* tw; int lg_tw;
  FftDev fft;
  const cplx<T>* in;
  cplx<T>* out;
  const int* omap;   // optional: outer index of the bid-th group (chunked launches of the slab pipeline)
};

// raw complex lines (stride `rstride` between the lines of a CTA); inactive lines read as zero
template <class T> struct RawLoader {
  const cplx<T>* base; long rstride; const struct LineInfo* li;
  // slab-decomposed grids: after the all-to-all a line is a sequence of per-rank chunks; element x of
  // line l lives at chunk_base[src_off[x]] + l * src_mul[x]  (src_off includes the offset inside the chunk)
  const long* src_off; const int* src_mul; const cplx<T>* chunk_base; int l0;
  template <int NQ> NB_HD NB_INLINE void batch(int r, int j, int lmr, cplx<T>* a) const;
};

// MINB: register budget (78 registers at 3); 3 measured 11 % faster for 16 KB CTAs (256^3: 64.6 vs 72.3 us), the host keeps
// 2 where three CTAs' shared memory would squeeze L1 (long lines), as for the other passes
template <class T, int MINB = 2> struct PCBody {
  typedef PCParams<T> Params;
  static constexpr int kMinBlocks = MINB;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    cplx<T>* s = reinterpret_cast<cplx<T>*>(smem);
    const int R = 1 << p.lg_R, n = 1 << p.lg_n;
    const int gpo = p.n_r >> p.lg_R;
    const int o = p.omap ? ldg(p.omap + ctx.bid / gpo) : ctx.bid / gpo, rr0 = (ctx.bid % gpo) << p.lg_R;
    const cplx<T>* inp = p.in + o * p.in_ostride + rr0 * p.in_rstride;
    RawLoader<T> ld{inp, p.in_rstride, nullptr, nullptr, nullptr, nullptr, 0};          // R divides n_r: no bounds check
    fft_dif_load(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw, ld);
    cplx<T>* outp = p.out + o * p.out_ostride + rr0;
    NB_FOR(ctx, i, n << p.lg_R) {
      int k = i >> p.lg_R, r = i & (R - 1);
      outp[k * p.out_kstride + r] = s[r * p.pitch + ldg(p.fft.pos + k)];
    }
  }
};

// ---------------------------------------------------------------------------------------------
// mirror geometry shared by P3 and P5: lines l = a*n_mid + km with a in [0, h_a]; the partner
// real line is (-a, -km).  n_mid is a power of two (1 for 2-D).
// ---------------------------------------------------------------------------------------------
struct MirrorGeom {
  int n_a, h_a, lg_mid;
  // slab decomposition of the `a` axis (dist != 0): this rank owns the planes a = a0 .. a0+cA-1 of the
  // half range [0, h_a] ("A" planes, stored first) and their mirrors n_a - a ("B" planes, stored behind
  // them in the same order; a = 0 and a = h_a have none).  skip = 1 if a = 0 is owned (its B is missing).
  int dist, a0, cA, skip;
  NB_HH NB_INLINE int nlines() const { return (dist ? cA : (h_a + 1)) << lg_mid; }
  // returns false if the line is handled by its (stored) partner; lB = -1 for self-mirrored lines.
  // lA / lB index real lines of the (local) position / latent array: plane * n_mid + km.
  NB_HD NB_INLINE bool resolve(int l, int& lA, int& lB) const {
    int a = l >> lg_mid, km = l & ((1 << lg_mid) - 1);
    int mkm = neg_idx(km, 1 << lg_mid);
    lA = l;
    if (!dist) {
      int ma = neg_idx(a, n_a);
      lB = (ma << lg_mid) + mkm;
      if (lB == lA) { lB = -1; return true; }
      if (ma <= h_a && lB < lA) return false;
      return true;
    }
    int ga = a0 + a;
    if (ga == 0 || 2 * ga == n_a) {          // self-mirrored plane: the partner line lives in the same plane
      lB = (a << lg_mid) + mkm;
      if (lB == lA) { lB = -1; return true; }
      return lB > lA;
    }
    lB = ((cA + a - skip) << lg_mid) + mkm;
    return true;
  }
};

// per-CTA table of the lines it owns (filled once, then read by every phase)
struct LineInfo { int lA, lB, active, pad; };
constexpr int LINEINFO_BYTES = 16 * 16;   // R <= 16

template <class T> template <int NQ>
NB_HD NB_INLINE void RawLoader<T>::batch(int r, int j, int lmr, cplx<T>* a) const {
  const cplx<T>* lp = base + r * rstride + j;
  if (li && !li[r].active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) a[q] = cmake<T>(0, 0);
    return;
  }
  if (src_off) {
    const long l = l0 + r;
#pragma unroll
    for (int q = 0; q < NQ; ++q) { int x = j + (q << lmr); a[q] = ld_stream(chunk_base + ldg(src_off + x) + l * ldg(src_mul + x)); }
    return;
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) a[q] = ld_stream(lp + (q << lmr));
}

// lines already sitting in the (natural-order) line buffers: the first butterfly stage of the in-place transform reads
// position j + q (n/R) and writes the swizzled slot of the same aligned group of 8, which only the 8 neighbouring
// threads of the same warp touch in this stage -- all reads of the warp are ordered before its writes
template <class T> struct SmemLoader {
  const cplx<T>* s; int pitch; const LineInfo* li;
  template <int NQ> NB_HD NB_INLINE void batch(int r, int j, int lmr, cplx<T>* a) const {
    const cplx<T>* lp = s + r * pitch + j;
    const bool act = !li || li[r].active;
#pragma unroll
    for (int q = 0; q < NQ; ++q) a[q] = act ? lp[q << lmr] : cmake<T>(0, 0);
#ifndef NB_EMU
    __syncwarp(__activemask());
#endif
  }
};

NB_HD NB_INLINE void fill_line_info(Ctx& ctx, LineInfo* li, const MirrorGeom& mg, int l0, int R) {
  NB_FOR(ctx, r, R) {
    int l = l0 + r, lA = 0, lB = -1;
    bool act = l < mg.nlines() && mg.resolve(l, lA, lB);
    li[r].lA = lA; li[r].lB = lB; li[r].active = act ? 1 : 0; li[r].pad = 0;
  }
  ctx.sync();
}

// ---------------------------------------------------------------------------------------------
// P3 pointwise operator (position space, T-layout index ridx = line*n + x)
// ---------------------------------------------------------------------------------------------
enum PointMode {
  PM_METRIC = 0,      // r = jl_a jl_b (v/V + cshift)                       acc0 += r
  PM_LINEARIZE = 1,   // f = off + v/V; s = sc exp(f); store s, jl; r = dE/df   acc0 += E, acc1 += r
  PM_JVP_OUT = 2,     // pos_out = (v/V + cshift) * (jl_a ? jl_a : 1)        (FWD only)
  PM_FIELD_OUT = 3,   // pos_out = off + v/V                                 (FWD only)
  PM_LOAD = 4,        // r = in_pos * (jl_a ? jl_a : 1)                      acc0 += r (ADJ only)
};
enum LhKind { LH_GAUSS = 0, LH_POISSON = 1 };

template <class T> struct PointOp {
  int mode;
  int natural;          // PM_FIELD_OUT / PM_JVP_OUT / PM_LOAD address natural-layout arrays
  int lh_kind, nl_exp;  // nl_exp: 0 identity, 1 exp, 2 tabulated (nl_s / nl_ds: signal and d signal / d field per position, T-layout)
  const T* nl_s; const T* nl_ds;
  T invV, offset, cshift_scale;   // cshift = cshift_scale * (*cshift_ptr) if cshift_ptr
  const T* cshift_ptr;
  T sc;                 // multiplicative scaling of the signal (1 if absent); read from sc_ptr if set
  const T* sc_ptr;
  T w_scalar; const T* w_arr;     // Gaussian inverse noise variance
  const T* data;                  // T-layout
  const T* jl_a; const T* jl_b;
  const T* in_pos;
  T* s_out; T* jl_out; T* pos_out;
  T* partials;          // [nblk][2]
  int n_a_full, n_mid, n;         // for natural addressing: T-layout (a, km, x) <-> natural (x, km, a)

  NB_HD NB_INLINE long addr(int lfull, int x) const {
    if (!natural) return (long)lfull * n + x;
    int a = lfull / n_mid, km = lfull - a * n_mid;
    return ((long)x * n_mid + km) * n_a_full + a;
  }
  NB_HD NB_INLINE T load(int lfull, int x, T& acc0) const {
    long i = addr(lfull, x);
    T r = in_pos[i];
    if (jl_a) r *= jl_a[(long)lfull * n + x];
    acc0 += r;
    return r;
  }
  template <int MODE>
  NB_HD NB_INLINE T point(int lfull, int x, T v, T cshift, T scv, T& acc0, T& acc1) const {
    long i = (long)lfull * n + x;
    if (MODE == PM_METRIC) {
      T r = jl_a[i] * jl_b[i] * (v * invV + cshift);
      acc0 += r;
      return r;
    } else if (MODE == PM_LINEARIZE) {
      T f = offset + v * invV;
      T s, jd;
      if (nl_exp == 2) { s = nl_s[i]; jd = nl_ds[i]; }       // (evaluated by the host at this field: any pointwise map)
      else { s = nl_exp ? scv * nb_exp(f) : f; jd = nl_exp ? s : T(1); }
      T e, cot, l;
      if (lh_kind == LH_GAUSS) {
        T w = w_arr ? w_arr[i] : w_scalar;
        T d = data ? data[i] : T(0);
        T r = s - d;
        e = T(0.5) * w * r * r;
        cot = w * r * jd;
        l = nb_sqrt(w);
      } else {
        T d = data ? data[i] : T(0);
        e = s - d * nb_log(s);
        cot = (T(1) - d / s) * jd;
        l = T(1) / nb_sqrt(s);
      }
      if (s_out) s_out[i] = s;
      if (jl_out) jl_out[i] = jd * l;
      acc0 += e;
      acc1 += cot;
      return cot;
    } else if (MODE == PM_JVP_OUT) {
      T r = v * invV + cshift;
      if (jl_a) r *= jl_a[i];
      pos_out[addr(lfull, x)] = r;
      return r;
    } else {  // PM_FIELD_OUT
      T r = offset + v * invV;
      pos_out[addr(lfull, x)] = r;
      return r;
    }
  }
};

template <class T> struct P3Params {
  int lg_n, lg_R;
  MirrorGeom mg;
  int pitch;
  const cplx<T>* tw; int lg_tw;
  FftDev fft;
  T hsign;                 // +1: Re+Im (non_canonical_hartley), -1: Re-Im
  const cplx<T>* in;       // [l][n]   (or the all-to-all receive buffer when src_off != null)
  const long* src_off; const int* src_mul;
  cplx<T>* out;            // [k in 0..n/2][out_kstride]
  long out_kstride;
  int ahead;
  int line0;               // first line of this launch (chunked launches); multiple of R
  PointOp<T> op;
};

// MINB = CTAs per SM the register budget is cut for (80 registers at 3, ~106 at 2).  Three 64 KB CTAs leave
// only ~25 KB of L1 for the slot / twiddle tables every phase reads; which side wins was measured per shape
// (profiles/r2_notes.md) and is chosen by the host for the fused metric pass.
template <class T, bool FWD, bool ADJ, int MODE, int MINB = 2> struct P3Body {
  typedef P3Params<T> Params;
  static constexpr int kMinBlocks = MINB;
  // Hermitian combine of the pair (x, y = -x) of one line + pointwise operator
  static NB_HD NB_INLINE void pair(const Params& p, cplx<T>* line, const LineInfo& li, int x, int y, T cshift, T scv,
                                   T& acc0, T& acc1) {
    const T sg = p.hsign;
    int px = ldg(p.fft.pos + x), py = ldg(p.fft.pos + y);
    cplx<T> cx = line[px], cy = line[py];
    T aX = cx.x + sg * cx.y, bY = cx.x - sg * cx.y;
    T aY = cy.x + sg * cy.y, bX = cy.x - sg * cy.y;
    T a2x = p.op.template point<MODE>(li.lA, x, aX, cshift, scv, acc0, acc1), b2x = 0, a2y = 0, b2y = 0;
    if (y != x) a2y = p.op.template point<MODE>(li.lA, y, aY, cshift, scv, acc0, acc1);
    if (li.lB >= 0) {
      b2x = p.op.template point<MODE>(li.lB, x, bX, cshift, scv, acc0, acc1);
      if (y != x) b2y = p.op.template point<MODE>(li.lB, y, bY, cshift, scv, acc0, acc1);
    }
    if (ADJ) {
      line[px] = cmake<T>(a2x, b2x);
      if (y != x) line[py] = cmake<T>(a2y, b2y);
    }
  }
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    LineInfo* li = reinterpret_cast<LineInfo*>(smem);
    cplx<T>* s = reinterpret_cast<cplx<T>*>(reinterpret_cast<unsigned char*>(smem) + LINEINFO_BYTES);
    const int R = 1 << p.lg_R, n = 1 << p.lg_n, h = n >> 1, lg_h = p.lg_n - 1;
    const int l0 = p.line0 + (ctx.bid << p.lg_R);
    const int pbid = (p.line0 >> p.lg_R) + ctx.bid;   // slot of this CTA's partial sums
    fill_line_info(ctx, li, p.mg, l0, R);
    T acc0 = 0, acc1 = 0;
    T cshift = p.op.cshift_ptr ? p.op.cshift_scale * ldg(p.op.cshift_ptr) : T(0);
    T scv = p.op.sc_ptr ? ldg(p.op.sc_ptr) : p.op.sc;
    if (FWD && MODE == PM_METRIC && p.ahead > 0) {
      for (int r = 0; r < R; ++r) {
        if (!li[r].active) continue;
        prefetch_l2(ctx, p.op.jl_a + (long)li[r].lA * n, (size_t)n * sizeof(T));
        if (p.op.jl_b != p.op.jl_a) prefetch_l2(ctx, p.op.jl_b + (long)li[r].lA * n, (size_t)n * sizeof(T));
        if (li[r].lB >= 0) {
          prefetch_l2(ctx, p.op.jl_a + (long)li[r].lB * n, (size_t)n * sizeof(T));
          if (p.op.jl_b != p.op.jl_a) prefetch_l2(ctx, p.op.jl_b + (long)li[r].lB * n, (size_t)n * sizeof(T));
        }
      }
    }
    if (FWD) {
      if (p.ahead > 0 && ctx.bid + p.ahead < ctx.nblk)
        prefetch_l2(ctx, p.in + ((long)(ctx.bid + p.ahead) << (p.lg_R + p.lg_n)), sizeof(cplx<T>) << (p.lg_R + p.lg_n));
      const cplx<T>* inp = p.in + (long)l0 * n;
      RawLoader<T> ld{inp, (long)n, li, p.src_off, p.src_mul, p.in, l0};
      fft_dif_load(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw, ld);
      if (n == 1) {
        NB_FOR(ctx, r, R) if (li[r].active) pair(p, s + r * p.pitch, li[r], 0, 0, cshift, scv, acc0, acc1);
      } else if (MODE == PM_METRIC && ADJ) {
        // hot path: pairs (x, n-x), 0 < x < h, of lines that have a partner; every load of a batch of U
        // pairs is issued (from always-valid addresses) before anything depends on it
        constexpr int U = 2;
        const T* ja = p.op.jl_a; const T* jb = p.op.jl_b;
        const bool same = (ja == jb);
        const T sg = p.hsign, iv = p.op.invV;
        const int cnt = R << lg_h;
        struct Slot { int px, py, r; bool ok; T mAx, mAy, mBx, mBy; };
        auto issue = [&](int i0, Slot* sl) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            int i = i0 + u * ctx.nthr;
            bool in = i < cnt;
            i = in ? i : 0;
            int r = i >> lg_h, x = i & (h - 1);
            sl[u].ok = in && li[r].active && li[r].lB >= 0 && x != 0;
            sl[u].r = r;
            x = x != 0 ? x : 1;
            int y = n - x;
            long iA = (long)li[r].lA * n, iB = (long)(li[r].lB >= 0 ? li[r].lB : li[r].lA) * n;
            sl[u].px = ldg(p.fft.pos + x); sl[u].py = ldg(p.fft.pos + y);
            T aX = ld_stream(ja + iA + x), aY = ld_stream(ja + iA + y), bX = ld_stream(ja + iB + x), bY = ld_stream(ja + iB + y);
            if (!same) { aX *= ld_stream(jb + iA + x); aY *= ld_stream(jb + iA + y); bX *= ld_stream(jb + iB + x); bY *= ld_stream(jb + iB + y); }
            else { aX *= aX; aY *= aY; bX *= bX; bY *= bY; }
            sl[u].mAx = aX; sl[u].mAy = aY; sl[u].mBx = bX; sl[u].mBy = bY;
          }
        };
        auto consume = [&](int, const Slot* sl) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!sl[u].ok) continue;
            cplx<T>* line = s + sl[u].r * p.pitch;
            cplx<T> cx = line[sl[u].px], cy = line[sl[u].py];
            T aX = sl[u].mAx * ((cx.x + sg * cx.y) * iv + cshift), bY = sl[u].mBy * ((cx.x - sg * cx.y) * iv + cshift);
            T aY = sl[u].mAy * ((cy.x + sg * cy.y) * iv + cshift), bX = sl[u].mBx * ((cy.x - sg * cy.y) * iv + cshift);
            acc0 += (aX + aY) + (bX + bY);
            line[sl[u].px] = cmake<T>(aX, bX);
            line[sl[u].py] = cmake<T>(aY, bY);
          }
        };
        batched_loop<false, U, Slot>(ctx, cnt, issue, [](Slot*) {}, consume);
        // the self-paired columns x = 0 and x = h, and lines without a partner
        NB_FOR(ctx, i, 2 * R) {
          int r = i >> 1;
          if (li[r].active) { int x = (i & 1) ? h : 0; pair(p, s + r * p.pitch, li[r], x, x, cshift, scv, acc0, acc1); }
        }
        for (int r = 0; r < R; ++r) {
          if (!li[r].active || li[r].lB >= 0) continue;
          NB_FOR(ctx, x, h) if (x != 0) pair(p, s + r * p.pitch, li[r], x, n - x, cshift, scv, acc0, acc1);
        }
      } else {
        // pairs (x, n-x) for x in [0, h) (x = 0 is its own partner), then the self-paired x = h
        NB_FOR(ctx, i, R << lg_h) {
          int r = i >> lg_h, x = i & (h - 1);
          if (!li[r].active) continue;
          pair(p, s + r * p.pitch, li[r], x, (n - x) & (n - 1), cshift, scv, acc0, acc1);
        }
        NB_FOR(ctx, r, R) if (li[r].active) pair(p, s + r * p.pitch, li[r], h, h, cshift, scv, acc0, acc1);
      }
    } else {
      NB_FOR(ctx, i, R << p.lg_n) {
        int r = i >> p.lg_n, x = i & (n - 1);
        cplx<T> v = cmake<T>(0, 0);
        if (li[r].active) {
          v.x = p.op.load(li[r].lA, x, acc0);
          if (li[r].lB >= 0) v.y = p.op.load(li[r].lB, x, acc0);
        }
        s[r * p.pitch + ldg(p.fft.pos + x)] = v;
      }
    }
    if (ADJ) {
      ctx.sync();
      fft_dit(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw);
      const T half = T(0.5);
      NB_FOR(ctx, i, (h + 1) << p.lg_R) {
        int k = i >> p.lg_R, r = i & (R - 1);
        if (!li[r].active) continue;
        cplx<T> zk = s[r * p.pitch + swz(k)], zc = cconj(s[r * p.pitch + swz((n - k) & (n - 1))]);
        cplx<T> e = zk + zc, d = cmul_mi(zk - zc);
        cplx<T>* orow = p.out + k * p.out_kstride;
        orow[li[r].lA] = cmake<T>(half * e.x, half * e.y);
        if (li[r].lB >= 0) orow[li[r].lB] = cmake<T>(half * d.x, half * d.y);
      }
    }
    if (p.op.partials) {
      // scratch for the reduction lives behind the line buffers
      void* scratch = reinterpret_cast<void*>(s + (size_t)R * p.pitch);
      acc0 = ctx.block_sum(acc0, scratch);
      acc1 = ctx.block_sum(acc1, scratch);
      if (ctx.tid == 0) { p.op.partials[2 * pbid] = acc0; p.op.partials[2 * pbid + 1] = acc1; }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// P5 epilogues (latent space, natural layout)
// ---------------------------------------------------------------------------------------------
template <class T> struct EpiPlain {
  static constexpr bool BATCHED = false;
  struct Pre {};
  NB_HD NB_INLINE void preload(long, long, long, int, int, Pre&) const {}
  NB_HD NB_INLINE void preload_bin(long, long, int, int, int, Pre&) const {}
  NB_HD NB_INLINE void gather(Pre&) const {}
  NB_HD NB_INLINE void finish(long, long, long, int, int, T, T, T, T, const Pre&, T&) const {}
  T* out; T scale;
  T* partials;   // unused
  NB_HD NB_INLINE void prefetch(Ctx&, long, long, long, int) const {}
  NB_HD NB_INLINE void emit(long rowA, long rowB, long, long, int x, int y, T gAx, T gAy, T gBx, T gBy, T&) const {
    out[rowA + x] = gAx * scale;
    if (y != x) out[rowA + y] = gAy * scale;
    if (rowB >= 0) {
      out[rowB + x] = gBx * scale;
      if (y != x) out[rowB + y] = gBy * scale;
    }
  }
};

// out_k = A_b g_k / V (+ add_k); W[l][x] = sum over the (up to four) mirror points of xi_k g_k / V;
// acc += add_k out_k  (the xi-block of <t, M t> for conjugate gradient)
template <class T> struct EpiAdjoint {
  static constexpr bool BATCHED = true;
  T* out; const T* add; const T* xi; const int* idxf; const T* amp; T* W; T invV;
  T* partials;   // [nblk]
  // split form of emit() for a pair with distinct x != y and a partner row: preload (streaming
  // loads + bin index), gather (amplitude table), finish (arithmetic + stores)
  struct Pre { int b; T A, a0, a1, a2, a3, x0, x1, x2, x3; };
  NB_HD NB_INLINE void preload(long rowA, long rowB, long fbase, int x, int y, Pre& q) const {
    q.b = ldg(idxf + fbase + x);
    q.a0 = q.a1 = q.a2 = q.a3 = q.x0 = q.x1 = q.x2 = q.x3 = 0;
    if (add) { q.a0 = ld_stream(add + rowA + x); q.a1 = ld_stream(add + rowA + y); q.a2 = ld_stream(add + rowB + x); q.a3 = ld_stream(add + rowB + y); }
    if (xi) { q.x0 = ld_stream(xi + rowA + x); q.x1 = ld_stream(xi + rowA + y); q.x2 = ld_stream(xi + rowB + x); q.x3 = ld_stream(xi + rowB + y); }
  }
  NB_HD NB_INLINE void gather(Pre& q) const { q.A = ldg(amp + q.b); }
  // the same with the bin index already at hand (index rows staged in shared memory): the table gather goes out together
  // with the streaming loads, no dependent second round
  NB_HD NB_INLINE void preload_bin(long rowA, long rowB, int b, int x, int y, Pre& q) const {
    q.b = b; q.A = ldg(amp + b);
    q.a0 = q.a1 = q.a2 = q.a3 = q.x0 = q.x1 = q.x2 = q.x3 = 0;
    if (add) { q.a0 = ld_stream(add + rowA + x); q.a1 = ld_stream(add + rowA + y); q.a2 = ld_stream(add + rowB + x); q.a3 = ld_stream(add + rowB + y); }
    if (xi) { q.x0 = ld_stream(xi + rowA + x); q.x1 = ld_stream(xi + rowA + y); q.x2 = ld_stream(xi + rowB + x); q.x3 = ld_stream(xi + rowB + y); }
  }
  NB_HD NB_INLINE void finish(long rowA, long rowB, long wbase, int x, int y, T gAx, T gAy, T gBx, T gBy, const Pre& q,
                              T& acc) const {
    gAx *= invV; gAy *= invV; gBx *= invV; gBy *= invV;
    T o0 = q.A * gAx + q.a0, o1 = q.A * gAy + q.a1, o2 = q.A * gBx + q.a2, o3 = q.A * gBy + q.a3;
    acc += (q.a0 * o0 + q.a1 * o1) + (q.a2 * o2 + q.a3 * o3);
    out[rowA + x] = o0; out[rowA + y] = o1; out[rowB + x] = o2; out[rowB + y] = o3;
    if (W) W[wbase + x] = (q.x0 * gAx + q.x1 * gAy) + (q.x2 * gBx + q.x3 * gBy);   // (an L2 evict_last hint here did not help the segment sum)
  }
  // all loads first (the outputs may alias the inputs as far as the compiler knows), then the stores
  NB_HD NB_INLINE void emit(long rowA, long rowB, long fbase, long wbase, int x, int y, T gAx, T gAy, T gBx, T gBy,
                            T& acc) const {
    const bool two = (y != x), hasB = rowB >= 0;
    int b = ldg(idxf + fbase + x);
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0, x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    if (add) {
      a0 = add[rowA + x];
      if (two) a1 = add[rowA + y];
      if (hasB) { a2 = add[rowB + x]; if (two) a3 = add[rowB + y]; }
    }
    if (xi) {
      x0 = xi[rowA + x];
      if (two) x1 = xi[rowA + y];
      if (hasB) { x2 = xi[rowB + x]; if (two) x3 = xi[rowB + y]; }
    }
    T A = ldg(amp + b);
    gAx *= invV; gAy *= invV; gBx *= invV; gBy *= invV;
    T o0 = A * gAx + a0, o1 = A * gAy + a1, o2 = A * gBx + a2, o3 = A * gBy + a3;
    T ws = x0 * gAx;
    acc += a0 * o0;
    out[rowA + x] = o0;
    if (two) { out[rowA + y] = o1; ws += x1 * gAy; acc += a1 * o1; }
    if (hasB) {
      out[rowB + x] = o2; ws += x2 * gBx; acc += a2 * o2;
      if (two) { out[rowB + y] = o3; ws += x3 * gBy; acc += a3 * o3; }
    }
    if (W) W[wbase + x] = ws;
  }
  // rows touched by the epilogue of a line (for L2 prefetch at CTA start)
  NB_HD NB_INLINE void prefetch(Ctx& ctx, long rowA, long rowB, long fbase, int n) const {
    if (add) { prefetch_l2(ctx, add + rowA, (size_t)n * sizeof(T)); if (rowB >= 0) prefetch_l2(ctx, add + rowB, (size_t)n * sizeof(T)); }
    if (xi) { prefetch_l2(ctx, xi + rowA, (size_t)n * sizeof(T)); if (rowB >= 0) prefetch_l2(ctx, xi + rowB, (size_t)n * sizeof(T)); }
    prefetch_l2(ctx, idxf + fbase, (size_t)(n / 2 + 1) * sizeof(int));
  }
};

template <class T, class Epi> struct P5Params {
  int lg_n, lg_R;
  MirrorGeom mg;
  int hmid1;      // n_mid/2 + 1 (folded extent of the middle axis)
  int pitch;
  const cplx<T>* tw; int lg_tw;
  FftDev fft;
  T hsign;
  const cplx<T>* in;   // [l][n]   (or the all-to-all receive buffer when src_off != null)
  const long* src_off; const int* src_mul;
  int ahead;
  int line0;
  int sidx_off;   // byte offset of the staged bin-index rows in shared memory (0: the epilogue reads them from global memory)
  Epi epi;
  // staged chain (nb_passes2.cuh): the lines are columns of the row-major output of the previous pass and are
  // gathered into the line buffers by the TMA engine (gather != 0) instead of being read from `in`
  int gather;
  TmaDesc desc;
};

// MINB / PIPE (chosen by the host, profiles/r2_notes.md): long lines (64 KB CTAs) run two CTAs per SM with the pipelined
// epilogue (128 registers; at 80 it spills: 251 vs 151 us at 4096^2); short lines run three CTAs' worth of registers
// without the pipeline (80 registers, no spills: 121 vs 140 us at 256^3).
template <class T, class Epi, int MINB = 2, bool PIPE = true> struct P5Body {
  typedef P5Params<T, Epi> Params;
  static constexpr int kMinBlocks = MINB;
  static NB_HD NB_INLINE void pair(const Params& p, const cplx<T>* line, const LineInfo& li, int x, int y, T& acc) {
    const T sg = p.hsign;
    const int n = 1 << p.lg_n, h = n >> 1, nmid = 1 << p.mg.lg_mid;
    cplx<T> cx = line[ldg(p.fft.pos + x)], cy = line[ldg(p.fft.pos + y)];
    T gAx = cx.x + sg * cx.y, gBy = cx.x - sg * cx.y;
    T gAy = cy.x + sg * cy.y, gBx = cy.x - sg * cy.y;
    int a = li.lA >> p.mg.lg_mid, km = li.lA & (nmid - 1);
    long fbase = ((long)a * p.hmid1 + fold_idx(km, nmid)) * (h + 1);   // a <= h_a is already folded
    p.epi.emit((long)li.lA * n, li.lB >= 0 ? (long)li.lB * n : -1L, fbase, (long)li.lA * (h + 1), x, y, gAx, gAy, gBx, gBy, acc);
  }
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    LineInfo* li = reinterpret_cast<LineInfo*>(smem);
    cplx<T>* s = reinterpret_cast<cplx<T>*>(reinterpret_cast<unsigned char*>(smem) + LINEINFO_BYTES);
    const int R = 1 << p.lg_R, n = 1 << p.lg_n, h = n >> 1, lg_h = p.lg_n - 1;
    const int l0 = p.line0 + (ctx.bid << p.lg_R);
    const int pbid = (p.line0 >> p.lg_R) + ctx.bid;
    fill_line_info(ctx, li, p.mg, l0, R);
    T acc = 0;
    // bin-index rows of the CTA's lines -> shared memory behind the line buffers (eight loads in flight per thread, landing
    // while the lines are gathered): the epilogue then issues its amplitude gathers together with its streaming loads
    int* sidx = nullptr;
    if constexpr (Epi::BATCHED) {
      if (p.sidx_off) {
        sidx = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(smem) + p.sidx_off);
        const int nmid = 1 << p.mg.lg_mid, tot = R * (h + 1);
        constexpr int U = 8;
        for (int i0 = ctx.tid; i0 < tot; i0 += ctx.nthr * U) {
          int rv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * ctx.nthr, ii = i < tot ? i : 0, r = ii / (h + 1), k = ii - r * (h + 1);
            const int a = li[r].lA >> p.mg.lg_mid, km = li[r].lA & (nmid - 1);
            const long fbase = ((long)a * p.hmid1 + fold_idx(km, nmid)) * (h + 1);
            rv[u] = li[r].active ? ldg(p.epi.idxf + fbase + k) : 0;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { const int i = i0 + u * ctx.nthr; if (i < tot) sidx[i] = rv[u]; }
        }
      }
    }
    for (int r = 0; r < R && p.ahead > 0; ++r) {
      if (!li[r].active) continue;
      const int nmid = 1 << p.mg.lg_mid;
      int a = li[r].lA >> p.mg.lg_mid, km = li[r].lA & (nmid - 1);
      long fbase = ((long)a * p.hmid1 + fold_idx(km, nmid)) * (h + 1);
      p.epi.prefetch(ctx, (long)li[r].lA * n, li[r].lB >= 0 ? (long)li[r].lB * n : -1L, fbase, n);
    }
    if (p.ahead > 0 && ctx.bid + p.ahead < ctx.nblk)
      prefetch_l2(ctx, p.in + ((long)(ctx.bid + p.ahead) << (p.lg_R + p.lg_n)), sizeof(cplx<T>) << (p.lg_R + p.lg_n));
    if (p.gather) {
      // column l0 + r of the previous pass's row-major output -> line buffer r (natural order), box_rows rows per copy
      Mbar* bar = reinterpret_cast<Mbar*>(s + (size_t)R * p.pitch);      // reduction scratch: free until the end
      if (ctx.tid == 0) {
        mbar_init(bar, 1);
        const int br = p.desc.box_rows, nb = n / br;
        unsigned bytes = 0;
        for (int r = 0; r < R; ++r) if (li[r].active) bytes += (unsigned)(n * sizeof(cplx<T>));
        mbar_expect(bar, bytes);
        for (int r = 0; r < R; ++r) {
          if (!li[r].active) continue;
          for (int b = 0; b < nb; ++b) tma_gather(s + r * p.pitch + b * br, &p.desc, l0 + r, b * br, bar);
        }
      }
      ctx.sync();
      mbar_wait(bar, 0);
#ifdef NB_EMU
      std::vector<cplx<T>> copy(s, s + (size_t)R * p.pitch);      // the sequential emulation cannot order reads before writes
      SmemLoader<T> ld{copy.data(), p.pitch, li};
#else
      SmemLoader<T> ld{s, p.pitch, li};
#endif
      fft_dif_load(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw, ld);
    } else {
      const cplx<T>* inp = p.in + (long)l0 * n;
      RawLoader<T> ld{inp, (long)n, li, p.src_off, p.src_mul, p.in, l0};
      fft_dif_load(ctx, s, p.fft, R, p.pitch, p.tw, p.lg_tw, ld);
    }
    if (n == 1) {
      NB_FOR(ctx, r, R) if (li[r].active) pair(p, s + r * p.pitch, li[r], 0, 0, acc);
    } else if (Epi::BATCHED) {
      // hot path: pairs (x, n-x), 0 < x < h, of lines with a partner row, U pairs per batch.  The loop is
      // software-pipelined: the streaming loads and bin indices of batch k+1 are issued before batch k is
      // finished (arithmetic + stores), its table gathers right after, so that a thread always has one
      // batch of global loads in flight (ncu r2a: 52 % of this pass's samples sat in this loop, 37 % of
      // them waiting on the two dependent load rounds of the unpipelined form).
      constexpr int U = 2;
      const T sg = p.hsign;
      const int cnt = R << lg_h, nmid = 1 << p.mg.lg_mid;
      struct Slot { int px, py, r, x; bool ok; typename Epi::Pre pre; };
      auto issue = [&](int i0, Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int i = i0 + u * ctx.nthr;
          bool in = i < cnt;
          i = in ? i : 0;
          int r = i >> lg_h, x = i & (h - 1);
          sl[u].ok = in && li[r].active && li[r].lB >= 0 && x != 0;
          x = x != 0 ? x : 1;
          sl[u].r = r; sl[u].x = x;
          int a = li[r].lA >> p.mg.lg_mid, km = li[r].lA & (nmid - 1);
          long fbase = ((long)a * p.hmid1 + fold_idx(km, nmid)) * (h + 1);
          sl[u].px = ldg(p.fft.pos + x); sl[u].py = ldg(p.fft.pos + (n - x));
          if (sidx) p.epi.preload_bin((long)li[r].lA * n, (long)(li[r].lB >= 0 ? li[r].lB : li[r].lA) * n, sidx[r * (h + 1) + x], x, n - x, sl[u].pre);
          else p.epi.preload((long)li[r].lA * n, (long)(li[r].lB >= 0 ? li[r].lB : li[r].lA) * n, fbase, x, n - x, sl[u].pre);
        }
      };
      auto consume = [&](int, const Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!sl[u].ok) continue;
          const int r = sl[u].r, x = sl[u].x;
          const cplx<T>* line = s + r * p.pitch;
          cplx<T> cx = line[sl[u].px], cy = line[sl[u].py];
          p.epi.finish((long)li[r].lA * n, (long)li[r].lB * n, (long)li[r].lA * (h + 1), x, n - x, cx.x + sg * cx.y, cy.x + sg * cy.y,
                       cy.x - sg * cy.y, cx.x - sg * cx.y, sl[u].pre, acc);
        }
      };
      auto gather = [&](Slot* sl) {
        if (sidx) return;
#pragma unroll
        for (int u = 0; u < U; ++u) p.epi.gather(sl[u].pre);
      };
      batched_loop<PIPE, U, Slot>(ctx, cnt, issue, gather, consume);
      NB_FOR(ctx, i, 2 * R) {
        int r = i >> 1;
        if (li[r].active) { int x = (i & 1) ? h : 0; pair(p, s + r * p.pitch, li[r], x, x, acc); }
      }
      for (int r = 0; r < R; ++r) {
        if (!li[r].active || li[r].lB >= 0) continue;
        NB_FOR(ctx, x, h) if (x != 0) pair(p, s + r * p.pitch, li[r], x, n - x, acc);
      }
    } else {
      NB_FOR(ctx, i, R << lg_h) {
        int r = i >> lg_h, x = i & (h - 1);
        if (!li[r].active) continue;
        pair(p, s + r * p.pitch, li[r], x, (n - x) & (n - 1), acc);
      }
      NB_FOR(ctx, r, R) if (li[r].active) pair(p, s + r * p.pitch, li[r], h, h, acc);
    }
    if (p.epi.partials) {
      void* scratch = reinterpret_cast<void*>(s + (size_t)R * p.pitch);
      acc = ctx.block_sum(acc, scratch);
      if (ctx.tid == 0) p.epi.partials[pbid] = acc;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// axis-reversing copy (natural <-> T-layout), used at the API boundary for data / noise arrays and
// for user-visible position-space outputs.  dims (d0,d1,d2) natural -> out[x2][x1][x0].
// ---------------------------------------------------------------------------------------------
template <class T> struct RevParams { const T* in; T* out; int d0, d1, d2; };
template <class T> struct RevBody {
  typedef RevParams<T> Params;
  // one block per (x1, 32x32 tile of (x0,x2)); smem 32*33 T
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    T* tile = reinterpret_cast<T*>(smem);
    int t0 = (p.d0 + 31) / 32, t2 = (p.d2 + 31) / 32;
    int b = ctx.bid;
    int i2 = b % t2; b /= t2;
    int i0 = b % t0; b /= t0;
    int x1 = b;
    NB_FOR(ctx, i, 1024) {
      int r = i >> 5, c = i & 31;        // r: x0 offset, c: x2 offset (contiguous in the input)
      int x0 = i0 * 32 + r, x2 = i2 * 32 + c;
      if (x0 < p.d0 && x2 < p.d2) tile[r * 33 + c] = p.in[((long)x0 * p.d1 + x1) * p.d2 + x2];
    }
    ctx.sync();
    NB_FOR(ctx, i, 1024) {
      int r = i >> 5, c = i & 31;        // r: x2 offset, c: x0 offset (contiguous in the output)
      int x0 = i0 * 32 + c, x2 = i2 * 32 + r;
      if (x0 < p.d0 && x2 < p.d2) p.out[((long)x2 * p.d1 + x1) * p.d0 + x0] = tile[c * 33 + r];
    }
  }
};

}  // namespace nb
