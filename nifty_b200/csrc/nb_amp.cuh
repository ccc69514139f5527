// nifty_b200 -- the O(K) amplitude model on the device: forward, tangent and cotangent chains.
//
// Restates (not ports) nifty/re/correlated_field.py:481-516 (NonParametricAmplitude.__call__),
// :297-299 (_remove_slope), :807-821 (get_normalized_amplitudes), nifty/re/gauss_markov.py:102-114
// (integrated_wiener_process) and nifty/re/num/stats_distributions.py:42-98 (normal/lognormal
// priors); the tangent / cotangent chains are hand-derived (the reference gets them from JAX AD).
//
// The two chained cumulative sums of the integrated Wiener process
//     y_{j+1} = y_j + r1_j ,   x_{j+1} = x_j + dt_j y_j + r0_j
// are ONE scan over affine maps (a,b,c): (x,y) -> (x + a y + b, y + c), composed associatively as
// (a1+a2, b1+b2+a2 c1, c1+c2).  The cotangent chain is the same recurrence run backwards.
// Scans are two launches (chunk aggregates + their ordered scan in the last block to finish, apply) with
// fixed chunking, hence bit-reproducible; global sums use per-block partials finished by the last block in fixed order.
#pragma once
#include "nb_common.cuh"
#include <cstdlib>

namespace nb {

// indices into the per-linearisation scalar block (device array of T)
enum AmpScalIdx {
  SC_FLU = 0, SC_SLOPE, SC_SIG, SC_ASP, SC_Z, SC_S, SC_CWL, SC_TWLAST,
  SC_CJ = 8, SC_DA0 = 9,        // tangent side, consumed by ProMetric (must stay adjacent)
  SC_DSLOPE = 10, SC_DTWLAST = 11,
  SC_SG = 12, SC_SGL = 13, SC_ABAR0 = 14, SC_UBL = 15,   // cotangent side
  SC_SCALING = 16,              // value of the multiplicative scaling (1 if absent)
  SC_ENERGY = 17,               // likelihood energy of the last linearisation
  SC_SUMCOT = 18,               // sum of dE/df (scaling gradient)
  SC_DOT = 19,                  // <add, out> of the last adjoint application
  SC_CTF = 20, SC_CWC = 21, SC_SGC = 22, SC_UBC = 23,   // Matern: cutoff value, sum wS c, sum g c, sum ubar c
  SC_COUNT = 32
};

template <class T> struct AmpModel {
  int K;                 // number of mode bins
  int kind_power;        // 1: "power", 0: "amplitude"   (correlated_field.py:506-513)
  int has_dev, has_flu, has_asp, has_scaling;
  int matern, renorm;    // Matern amplitude (correlated_field.py:302-395): u_b = slope/4 log1p((k_b/cutoff)^2), optional renormalisation
  const T* modes;        // [K] mode lengths k_b (Matern)
  T ctf_a, ctf_b; long off_ctf;
  const T* ell;          // [K] relative log mode lengths
  const T* mult;         // [K] multiplicities
  const T* dt;           // [K-2] log volumes
  T V;                   // total volume
  T flu_a, flu_b, slp_a, slp_b, flx_a, flx_b, asp_a, asp_b, zm_a, zm_b, scl_a, scl_b;
  long off_flu, off_slp, off_flx, off_asp, off_spec, off_zm, off_xi, off_scl;
  long L;
};

template <class T> struct Aff { T a, b, c; };
template <class T> NB_HD NB_INLINE Aff<T> aff_id() { Aff<T> e; e.a = 0; e.b = 0; e.c = 0; return e; }
template <class T> NB_HD NB_INLINE Aff<T> aff_compose(Aff<T> f, Aff<T> s) {
  Aff<T> r; r.a = f.a + s.a; r.b = f.b + s.b + s.a * f.c; r.c = f.c + s.c; return r;
}

// this thread's share of sum_i partials[i*stride], i < n (i = tid, tid + nthr, ...; fixed order -> deterministic); the
// loads of four terms are issued before the first is added: a plain loop pays one global round trip PER TERM
template <class T> NB_HD NB_INLINE T strided_partial_sum(Ctx& ctx, const T* partials, int n, int stride) {
  T s = 0;
  for (int i0 = ctx.tid; i0 < n; i0 += ctx.nthr * 4) {
    T r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * ctx.nthr; r[u] = partials[(size_t)(i < n ? i : i0) * stride]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * ctx.nthr; if (i < n) s += r[u]; }
  }
  return s;
}
// sum of partials[i*stride], i < n, by all threads of the calling block (fixed order -> deterministic)
template <class T> NB_HD NB_INLINE T block_total(Ctx& ctx, const T* partials, int n, int stride, void* scratch) {
  return ctx.block_sum(strided_partial_sum(ctx, partials, n, stride), scratch);
}

// ---- ordered block scan of affine maps -----------------------------------------------------------
// Every thread contributes the composition `v` of its own (contiguous) elements.  Returns the
// composition of all EARLIER threads (exclusive prefix) and, in `total`, of all threads.  Fixed
// shuffle tree -> bit-reproducible.  `scratch` holds >= 2*32+2 Aff<T>.
#ifdef NB_EMU
template <class T> inline Aff<T> block_scan_aff(Ctx&, Aff<T> v, Aff<T>& total, void*) { total = v; return aff_id<T>(); }
#else
template <class T> __device__ NB_INLINE Aff<T> shfl_up_aff(Aff<T> v, int o) {
  Aff<T> r; r.a = __shfl_up_sync(0xffffffffu, v.a, o); r.b = __shfl_up_sync(0xffffffffu, v.b, o); r.c = __shfl_up_sync(0xffffffffu, v.c, o);
  return r;
}
template <class T> __device__ NB_INLINE Aff<T> block_scan_aff(Ctx& ctx, Aff<T> v, Aff<T>& total, void* scratch) {
  Aff<T>* wagg = reinterpret_cast<Aff<T>*>(scratch);       // [32] inclusive warp totals, then their scan
  const int lane = ctx.tid & 31, warp = ctx.tid >> 5, nw = (ctx.nthr + 31) >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Aff<T> n = shfl_up_aff(v, o);
    if (lane >= o) v = aff_compose(n, v);
  }
  Aff<T> excl = shfl_up_aff(v, 1);
  if (lane == 0) excl = aff_id<T>();
  __syncthreads();
  if (lane == 31) wagg[warp] = v;
  __syncthreads();
  if (warp == 0) {
    Aff<T> w = lane < nw ? wagg[lane] : aff_id<T>();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      Aff<T> n = shfl_up_aff(w, o);
      if (lane >= o) w = aff_compose(n, w);
    }
    Aff<T> we = shfl_up_aff(w, 1);
    if (lane == 0) we = aff_id<T>();
    wagg[32 + lane] = we;                 // exclusive prefix of each warp
    if (lane == 31) wagg[64] = w;         // block total
  }
  __syncthreads();
  Aff<T> res = aff_compose(wagg[32 + warp], excl);
  total = wagg[64];
  __syncthreads();
  return res;
}
#endif

constexpr int SCAN_NT = 256;
// A chunk is SCAN_NT * E elements (E = 4 or 8 per thread, chosen by the host: E = 4 keeps more blocks
// resident and measured 2-4 % faster per product on large tables; E = 8 lets a table of <= 2048 bins run
// as a single chunk, i.e. without the aggregate kernel).
// Staging slot of chunk element i: one pad element per thread run of E, so that the thread-strided
// accesses el[tid*(E+1) + e] (stride 54 or 30 words: an odd multiple of 2) are free of bank conflicts;
// without the pad the stride is 48 / 24 words and 8-byte accesses collide 16 / 8 ways.
template <int E> NB_HH NB_INLINE int sidx(int i) { return i + i / E; }
template <class T> inline size_t scan_smem_bytes(int E) { return (size_t)(SCAN_NT * E + SCAN_NT + 80) * sizeof(Aff<T>); }
inline int scan_pick_e(long n) {
  if (const char* e = std::getenv("NB200_SCAN_E")) { if (e[0] == '4') return 4; if (e[0] == '8') return 8; }   // test / tuning knob
  return n <= SCAN_NT * 8 ? 8 : 4;
}

// ---- generic scan: (1) per-chunk aggregates [skipped for a single chunk], (2) apply -----------------
template <class T, class Elem> struct ScanAggParams { long n; Elem elem; Aff<T>* agg; Aff<T>* pre; unsigned* counter; };
#ifndef NB_AGG_MINB
#define NB_AGG_MINB 4      // 64 registers: these kernels are latency bound, 4 resident CTAs measured 20 % faster than 2
#endif
#ifndef NB_APPLY_MINB
#define NB_APPLY_MINB 3    // 80 registers (10 % faster than the unconstrained 84-90)
#endif
template <class T, class Elem, int E> struct ScanAggBody {
  typedef ScanAggParams<T, Elem> Params;
  static constexpr int kMinBlocks = NB_AGG_MINB;
  static constexpr int CH = SCAN_NT * E, STAGE = CH + SCAN_NT;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    Aff<T>* el = reinterpret_cast<Aff<T>*>(smem);
    long p0 = (long)ctx.bid * CH;
    NB_FOR(ctx, i, CH) el[sidx<E>(i)] = (p0 + i < p.n) ? p.elem.get(p0 + i) : aff_id<T>();
    ctx.sync();
#ifdef NB_EMU
    Aff<T> tot = aff_id<T>();
    for (int i = 0; i < CH; ++i) tot = aff_compose(tot, el[sidx<E>(i)]);
#else
    const Aff<T>* mine_el = el + ctx.tid * (E + 1);
    Aff<T> a = mine_el[0];
#pragma unroll
    for (int e = 1; e < E; ++e) a = aff_compose(a, mine_el[e]);
    Aff<T> tot;
    block_scan_aff(ctx, a, tot, reinterpret_cast<void*>(el + STAGE));
#endif
    if (ctx.tid == 0) p.agg[ctx.bid] = tot;
    // the last block to finish turns the chunk aggregates into exclusive prefixes (pre[c], state
    // before chunk c) and the grand total pre[nblk] -- one ordered block scan instead of two per
    // apply block
    if (ctx.last_block(p.counter)) {
      void* scratch = reinterpret_cast<void*>(el + STAGE);
      const int n = ctx.nblk, per = (n + ctx.nthr - 1) / ctx.nthr;
      const int lo = ctx.tid * per, hi = (lo + per < n) ? lo + per : n;
      Aff<T> v = aff_id<T>();
      for (int i = lo; i < hi; ++i) v = aff_compose(v, p.agg[i]);
      Aff<T> total;
      Aff<T> pre = block_scan_aff(ctx, v, total, scratch);
      for (int i = lo; i < hi; ++i) { p.pre[i] = pre; pre = aff_compose(pre, p.agg[i]); }
      if (ctx.tid == 0) p.pre[n] = total;
    }
  }
};
// Out::put(pos, pre, x0, y0, x1, y1, total_x, acc[4]) is called once per element with the state before /
// after; Out::finish(ctx, acc, total_x, scratch) runs in every block afterwards.
template <class T, class Elem, class Out> struct ScanApplyParams { long n; Elem elem; Out out; const Aff<T>* pre; int nchunks; };
template <class T, class Elem, class Out, int E> struct ScanApplyBody {
  typedef ScanApplyParams<T, Elem, Out> Params;
  static constexpr int kMinBlocks = NB_APPLY_MINB;
  static constexpr int CH = SCAN_NT * E, STAGE = CH + SCAN_NT;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    Aff<T>* el = reinterpret_cast<Aff<T>*>(smem);
    void* scratch = reinterpret_cast<void*>(el + STAGE);
    long p0 = (long)ctx.bid * CH;
    NB_FOR(ctx, i, CH) el[sidx<E>(i)] = (p0 + i < p.n) ? p.elem.get(p0 + i) : aff_id<T>();
    ctx.sync();
    // carry-in of this chunk and the grand total from the chunk aggregates (nchunks > 1)
    Aff<T> carry = aff_id<T>(), grand = aff_id<T>();
    if (p.nchunks > 1) { carry = p.pre[ctx.bid]; grand = p.pre[p.nchunks]; }
#ifdef NB_EMU
    {
      // one host "thread": walk the whole chunk sequentially
      Aff<T> blk = aff_id<T>();
      for (int i = 0; i < CH; ++i) blk = aff_compose(blk, el[sidx<E>(i)]);
      if (p.nchunks <= 1) grand = blk;
      T acc[4] = {0, 0, 0, 0};
      T x = carry.b, y = carry.c;
      for (int i = 0; i < CH; ++i) {
        long pos = p0 + i;
        if (pos >= p.n) break;
        Aff<T> a = el[sidx<E>(i)];
        T x1 = x + a.a * y + a.b, y1 = y + a.c;
        p.out.put(pos, p.out.load(pos), x, y, x1, y1, grand.b, acc);
        x = x1; y = y1;
      }
      p.out.finish(ctx, acc, grand.b, scratch);
    }
#else
    const Aff<T>* mine_el = el + ctx.tid * (E + 1);
    Aff<T> mine = mine_el[0];
#pragma unroll
    for (int e = 1; e < E; ++e) mine = aff_compose(mine, mine_el[e]);
    Aff<T> blk;
    Aff<T> pre = block_scan_aff(ctx, mine, blk, scratch);
    if (p.nchunks <= 1) grand = blk;
    pre = aff_compose(carry, pre);
    T acc[4] = {0, 0, 0, 0};
    T x = pre.b, y = pre.c;          // state = prefix applied to the zero state
    // all global loads of the E outputs of this thread are issued before the sequential walk
    typename Out::Pre pl[E];
    if (p.n > 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        long pos = p0 + ctx.tid * E + e;
        pl[e] = p.out.load(pos < p.n ? pos : p.n - 1);
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      long pos = p0 + ctx.tid * E + e;
      if (pos < p.n) {
        Aff<T> a = mine_el[e];
        T x1 = x + a.a * y + a.b, y1 = y + a.c;
        p.out.put(pos, pl[e], x, y, x1, y1, grand.b, acc);
        x = x1; y = y1;
      }
    }
    p.out.finish(ctx, acc, grand.b, scratch);
#endif
  }
};

// ---- helpers ---------------------------------------------------------------------------------
template <class T> NB_HD NB_INLINE T prior_ln(T a, T b, T xi) { return nb_exp(a + b * xi); }

// scalars of the amplitude model at `pos` (every thread may call it; a handful of exps)
template <class T> struct AmpPoint { T flu, slope, sig, asp, z, scl, ctf; };
template <class T> NB_HD NB_INLINE AmpPoint<T> amp_point(const AmpModel<T>& m, const T* pos) {
  AmpPoint<T> a;
  a.flu = m.has_flu ? prior_ln(m.flu_a, m.flu_b, pos[m.off_flu]) : T(1);
  a.slope = m.slp_a + m.slp_b * pos[m.off_slp];
  a.sig = m.has_dev ? prior_ln(m.flx_a, m.flx_b, pos[m.off_flx]) : T(0);
  a.asp = (m.has_dev && m.has_asp) ? prior_ln(m.asp_a, m.asp_b, pos[m.off_asp]) : T(0);
  a.z = prior_ln(m.zm_a, m.zm_b, pos[m.off_zm]);
  a.scl = m.has_scaling ? prior_ln(m.scl_a, m.scl_b, pos[m.off_scl]) : T(1);
  a.ctf = m.matern ? prior_ln(m.ctf_a, m.ctf_b, pos[m.off_ctf]) : T(1);
  return a;
}

// ---- forward chain ---------------------------------------------------------------------------
// scan position p = bin b; elements are the IWP increments for b >= 2 (j = b-2), identity below
template <class T> struct FwdElem {
  AmpModel<T> m; const T* pos;
  NB_HD NB_INLINE Aff<T> get(long b) const {
    if (!m.has_dev) return aff_id<T>();
    long j = b >= 2 ? b - 2 : 0;
    AmpPoint<T> ap = amp_point(m, pos);
    T dt = m.dt[j], sd = ap.sig * nb_sqrt(dt), q = nb_sqrt(dt * dt / T(12) + ap.asp);
    T x0 = pos[m.off_spec + 2 * j], x1 = pos[m.off_spec + 2 * j + 1];
    T r1 = sd * x1, r0 = sd * x0 * q + T(0.5) * dt * r1;
    Aff<T> e; e.a = dt; e.b = r0; e.c = r1;
    if (b < 2) e = aff_id<T>();
    return e;
  }
};
// P_b = exp(slope l_b + tw_b - tw_last l_b / l_last); partial S
template <class T> struct FwdOut {
  AmpModel<T> m; const T* pos;
  T* P; T* partials; unsigned* counter; T* scal;
  T* ellv; T* cv;     // Matern: per-linearisation tables a_b = du_b/dslope and c_b = du_b/dcutoff
  struct Pre { T l, mu; };
  NB_HD NB_INLINE Pre load(long b) const { Pre q; q.l = m.matern ? m.modes[b] : m.ell[b]; q.mu = m.mult[b]; return q; }
  NB_HD NB_INLINE void put(long b, const Pre& q, T, T, T x1, T, T total, T* acc) const {
    AmpPoint<T> ap = amp_point(m, pos);
    T u;
    if (m.matern) {
      T kk = q.l / ap.ctf, lg = log1p(kk * kk);
      u = T(0.25) * ap.slope * lg;
      ellv[b] = T(0.25) * lg;
      cv[b] = T(0.25) * ap.slope * (T(-2) * q.l * q.l / (ap.ctf * ap.ctf * ap.ctf)) / (T(1) + kk * kk);
    } else {
      T l = q.l, llast = m.ell[m.K - 1];
      u = ap.slope * l;
      if (m.has_dev) u += x1 - total * (l / llast);
    }
    T Pb = nb_exp(u);
    P[b] = Pb;
    if (b >= 1) acc[0] += q.mu * (m.kind_power ? Pb : Pb * Pb);
  }
  NB_HD NB_INLINE void finish(Ctx& ctx, T* acc, T total, void* scratch) const {
    T s = ctx.block_sum(acc[0], scratch);
    if (ctx.tid == 0) partials[ctx.bid] = s;
    if (ctx.last_block(counter)) {
      T S = block_total(ctx, partials, ctx.nblk, 1, scratch);
      if (ctx.tid == 0) {
        AmpPoint<T> ap = amp_point(m, pos);
        if (m.matern && !m.renorm) S = m.V;       // norm = 1: amp = scale sqrt(V) shape (see AmpTabBody)
        scal[SC_CTF] = ap.ctf;
        scal[SC_S] = S; scal[SC_FLU] = ap.flu; scal[SC_SLOPE] = ap.slope; scal[SC_SIG] = ap.sig;
        scal[SC_ASP] = ap.asp; scal[SC_Z] = ap.z; scal[SC_SCALING] = ap.scl;
        scal[SC_TWLAST] = m.has_dev ? total : T(0);
      }
    }
  }
};
// A_b (with A_0 = z V), wS_b, sum wS_b l_b
template <class T> struct AmpTabParams { int ch; AmpModel<T> m; const T* P; T* amp; T* wS; T* partials; unsigned* counter; T* scal; const T* ellv; const T* cv; };
template <class T> struct AmpTabBody {
  typedef AmpTabParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    const AmpModel<T>& m = p.m;
    T S = p.scal[SC_S], flu = p.scal[SC_FLU], z = p.scal[SC_Z];
    T acc = 0, acc2 = 0;
    const bool norm = !m.matern || m.renorm;
    long b0 = (long)ctx.bid * p.ch;
    NB_FOR(ctx, i, p.ch) {
      long b = b0 + i;
      if (b >= m.K) break;
      T Pb = p.P[b], A, w;
      if (m.kind_power) { A = flu * m.V * nb_sqrt(Pb / S); w = m.mult[b] * Pb / S; }
      else { A = flu * m.V * Pb / nb_sqrt(S); w = T(2) * m.mult[b] * Pb * Pb / S; }
      if (!norm) w = 0;
      if (b == 0) { A = z * m.V; w = 0; }
      p.amp[b] = A; p.wS[b] = w;
      acc += w * p.ellv[b];
      if (p.cv) acc2 += w * p.cv[b];
    }
    { T v2[2] = {acc, acc2}; ctx.template block_sum_n<2>(v2, smem); acc = v2[0]; acc2 = v2[1]; }
    if (ctx.tid == 0) { p.partials[2 * ctx.bid] = acc; p.partials[2 * ctx.bid + 1] = acc2; }
    if (ctx.last_block(p.counter)) {
      T c = block_total(ctx, p.partials, ctx.nblk, 2, smem);
      T c2 = block_total(ctx, p.partials + 1, ctx.nblk, 2, smem);
      if (ctx.tid == 0) { p.scal[SC_CWL] = c; p.scal[SC_CWC] = c2; }
    }
  }
};

// ---- tangent chain ---------------------------------------------------------------------------
template <class T> struct JvpElem {
  AmpModel<T> m; const T* pos; const T* t; const T* scal;
  NB_HD NB_INLINE Aff<T> get(long b) const {
    if (!m.has_dev) return aff_id<T>();
    long j = b >= 2 ? b - 2 : 0;
    AmpPoint<T> ap; ap.sig = scal[SC_SIG]; ap.asp = scal[SC_ASP];
    T dt = m.dt[j], sd = ap.sig * nb_sqrt(dt), q = nb_sqrt(dt * dt / T(12) + ap.asp);
    T x0 = pos[m.off_spec + 2 * j], x1 = pos[m.off_spec + 2 * j + 1];
    T d0 = t[m.off_spec + 2 * j], d1 = t[m.off_spec + 2 * j + 1];
    T dsig_rel = m.flx_b * t[m.off_flx];
    T dasp = m.has_asp ? ap.asp * m.asp_b * t[m.off_asp] : T(0);
    T dr1 = sd * (dsig_rel * x1 + d1);
    T dr0 = sd * q * (dsig_rel * x0 + d0) + sd * x0 * dasp / (T(2) * q) + T(0.5) * dt * dr1;
    Aff<T> e; e.a = dt; e.b = dr0; e.c = dr1;
    if (b < 2) e = aff_id<T>();
    return e;
  }
};
template <class T> struct JvpOut {
  AmpModel<T> m; const T* pos; const T* t; const T* wS; const T* amp; const T* ellv; const T* cv;
  cplx<T>* ad; T* partials; unsigned* counter; T* scal;   // ad[b] = (A_b, du_b)
  struct Pre { T l, w, A, c; };
  NB_HD NB_INLINE Pre load(long b) const { Pre q; q.l = ellv[b]; q.w = wS[b]; q.A = amp[b]; q.c = cv ? cv[b] : T(0); return q; }
  NB_HD NB_INLINE void put(long b, const Pre& q, T, T, T x1, T, T total, T* acc) const {
    T l = q.l;
    T d = m.slp_b * t[m.off_slp] * l;
    if (m.matern) d += q.c * (scal[SC_CTF] * m.ctf_b * t[m.off_ctf]);
    if (m.has_dev) d += x1 - total * (l / m.ell[m.K - 1]);
    ad[b] = cmake<T>(q.A, d);
    acc[0] += q.w * d;
  }
  NB_HD NB_INLINE void finish(Ctx& ctx, T* acc, T, void* scratch) const {
    T s = ctx.block_sum(acc[0], scratch);
    if (ctx.tid == 0) partials[ctx.bid] = s;
    T tflu = 0, tzm = 0, z = 0;      // fetched ahead of the grid-wide handshake (latency off the serial tail)
    if (ctx.tid == 0) { if (m.has_flu) tflu = t[m.off_flu]; tzm = t[m.off_zm]; z = scal[SC_Z]; }
    if (ctx.last_block(counter)) {
      T sw = block_total(ctx, partials, ctx.nblk, 1, scratch);
      if (ctx.tid == 0) {
        T dflu_rel = m.has_flu ? m.flu_b * tflu : T(0);
        scal[SC_CJ] = dflu_rel - T(0.5) * sw;
        scal[SC_DA0] = m.V * z * m.zm_b * tzm;
      }
    }
  }
};

// ---- cotangent chain -------------------------------------------------------------------------
// segment sum over the folded W array through a CSR of W positions per bin (fixed order), a power-of-two
// group of lanes per bin:  g_b = A_b * sum_{p in bin b} W[p]  (g_0 = 0).  (Storing W in bin order instead,
// i.e. scattering in the last adjoint pass, was measured: the segment sum gained 12 us, that pass lost 32.)
template <class T> struct SegSumParams {
  AmpModel<T> m; const T* W; const int* order; const int* offs; const T* amp;
  T* g; T* abar /* optional raw bin sums */; T* partials /* [nblk][2] */; unsigned* counter; T* scal;
  const T* abar_in;   // if set: skip the gather, take the (all-reduced) bin sums from here
  const T* ellv; const T* cv;
  int lg_lpb;   // log2(lanes per bin); 256 threads -> 256 >> lg_lpb bins per block
};
template <class T> struct SegSumBody {
  typedef SegSumParams<T> Params;
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    T* sm = reinterpret_cast<T*>(smem);          // [256] lane partials, then reduction scratch
    const AmpModel<T>& m = p.m;
    const int lpb = 1 << p.lg_lpb, bpb = 256 >> p.lg_lpb;
    long b0 = (long)ctx.bid * bpb;
    NB_FOR(ctx, i, 256) {
      long b = b0 + (i >> p.lg_lpb);
      int lane = i & (lpb - 1);
      T s = 0;
      if (p.abar_in) {
        if (b < m.K && lane == 0) s = p.abar_in[b];
      } else if (b < m.K) {
        // measured alternatives (profiles/r2_notes.md): branch-free rounds of 4 or 8 entries were 8-16 % slower
        // than this loop (bins hold 3.5 entries on average in 2-D), storing W in bin order cost the producer 32 us
        int beg = p.offs[b], end = p.offs[b + 1];
        int q = beg + lane;
        for (; q + 3 * lpb < end; q += 4 * lpb) {     // four independent index -> value chains in flight
          int o0 = p.order[q], o1 = p.order[q + lpb], o2 = p.order[q + 2 * lpb], o3 = p.order[q + 3 * lpb];
          T w0 = p.W[o0], w1 = p.W[o1], w2 = p.W[o2], w3 = p.W[o3];
          s += w0; s += w1; s += w2; s += w3;
        }
        for (; q < end; q += lpb) s += p.W[p.order[q]];
      }
      sm[i] = s;
    }
    ctx.sync();
    T a0 = 0, a1 = 0, a2 = 0;
    NB_FOR(ctx, i, bpb) {
      long b = b0 + i;
      if (b < m.K) {
        T abar = 0;
        for (int l = 0; l < lpb; ++l) abar += sm[(i << p.lg_lpb) + l];   // fixed order
        if (p.abar) p.abar[b] = abar;
        if (p.g) {
          T gb = (b == 0) ? T(0) : abar * p.amp[b];
          p.g[b] = gb;
          if (b == 0) p.scal[SC_ABAR0] = abar;
          a0 += gb; a1 += gb * p.ellv[b];
          if (p.cv) a2 += gb * p.cv[b];
        }
      }
    }
    if (!p.g) return;
    void* scratch = reinterpret_cast<void*>(sm + 256);
    { T v3[3] = {a0, a1, a2}; ctx.template block_sum_n<3>(v3, scratch); a0 = v3[0]; a1 = v3[1]; a2 = v3[2]; }
    if (ctx.tid == 0) { p.partials[3 * ctx.bid] = a0; p.partials[3 * ctx.bid + 1] = a1; p.partials[3 * ctx.bid + 2] = a2; }
    T cwl = 0, cwc = 0;
    if (ctx.tid == 0) { cwl = p.scal[SC_CWL]; cwc = p.scal[SC_CWC]; }
    if (ctx.last_block(p.counter)) {
      T v[3] = {0, 0, 0};
      NB_FOR(ctx, i, ctx.nblk) { v[0] += p.partials[3 * (size_t)i]; v[1] += p.partials[3 * (size_t)i + 1]; v[2] += p.partials[3 * (size_t)i + 2]; }
      ctx.template block_sum_n<3>(v, scratch);
      T sg = v[0], sgl = v[1], sgc = v[2];
      if (ctx.tid == 0) {
        T kappa = m.kind_power ? T(0.5) : T(1);
        p.scal[SC_SG] = sg; p.scal[SC_SGL] = sgl; p.scal[SC_SGC] = sgc;
        p.scal[SC_UBL] = kappa * sgl - T(0.5) * sg * cwl;   // sum_b ubar_b l_b
        p.scal[SC_UBC] = kappa * sgc - T(0.5) * sg * cwc;   // sum_b ubar_b c_b (Matern cutoff)
      }
    }
  }
};

// reverse scan over j' = (K-3) - j ; element c_j = twbar_{j+2}
template <class T> NB_HD NB_INLINE T twbar_at(const AmpModel<T>& m, const T* g, const T* wS, const T* scal, long b) {
  T kappa = m.kind_power ? T(0.5) : T(1);
  T ub = kappa * g[b] - T(0.5) * scal[SC_SG] * wS[b];
  if (b == m.K - 1) ub -= scal[SC_UBL] / m.ell[m.K - 1];
  return ub;
}
template <class T> struct VjpElem {
  AmpModel<T> m; const T* g; const T* wS; const T* scal;
  NB_HD NB_INLINE Aff<T> get(long p) const {
    long j = (m.K - 3) - p;
    T dt = m.dt[j], c = twbar_at(m, g, wS, scal, j + 2);
    Aff<T> e; e.a = dt; e.b = dt * c; e.c = c; return e;
  }
};

// everything that finishes an adjoint application: spectrum cotangents from the reverse scan,
// the scalar leaves, "+ add" and the dot product <add, out>
template <class T> struct VjpOut {
  AmpModel<T> m; const T* pos; const T* g; const T* wS; const T* scal_in;
  T* out; const T* add;
  T* partials /* [nblk][3] */; unsigned* counter; T* scal;
  const T* p3_partials; int n_p3;       // [n_p3][2]: acc0 = sum of position-space cotangent
  const T* p5_partials; int n_p5;       // [n_p5]: xi-block of <add, out>
  T scl_factor;                         // factor applied to the p3 sum for the scaling leaf
  struct Pre { T dt, xi0, xi1, a0, a1; };
  NB_HD NB_INLINE Pre load(long p) const {
    long j = (m.K - 3) - p;
    Pre q; q.dt = m.dt[j]; q.xi0 = pos[m.off_spec + 2 * j]; q.xi1 = pos[m.off_spec + 2 * j + 1]; q.a0 = 0; q.a1 = 0;
    if (add) { q.a0 = add[m.off_spec + 2 * j]; q.a1 = add[m.off_spec + 2 * j + 1]; }
    return q;
  }
  NB_HD NB_INLINE void put(long p, const Pre& pq, T x0, T, T, T y1, T, T* acc) const {
    long j = (m.K - 3) - p;
    T sig = scal_in[SC_SIG], asp = scal_in[SC_ASP];
    T dt = pq.dt, sq = nb_sqrt(dt), sd = sig * sq, q = nb_sqrt(dt * dt / T(12) + asp);
    T r0bar = y1, r1bar = x0 + T(0.5) * dt * r0bar;
    T xi0 = pq.xi0, xi1 = pq.xi1;
    T o0 = r0bar * sd * q, o1 = r1bar * sd;
    acc[0] += (r0bar * xi0 * q + r1bar * xi1) * sq;           // sigbar
    acc[1] += r0bar * sd * xi0 / (T(2) * q);                  // aspbar
    if (add) {
      T a0 = pq.a0, a1 = pq.a1;
      o0 += a0; o1 += a1;
      acc[2] += a0 * o0 + a1 * o1;
    }
    out[m.off_spec + 2 * j] = o0; out[m.off_spec + 2 * j + 1] = o1;
  }
  NB_HD NB_INLINE void leaf(long off, T v, T a, T& dot) const {
    if (add) { v += a; dot += a * v; }
    out[off] = v;
  }
  NB_HD NB_INLINE void finish(Ctx& ctx, T* acc, T, void* scratch) const {
    ctx.template block_sum_n<3>(acc, scratch);
    if (ctx.tid == 0) { partials[3 * ctx.bid] = acc[0]; partials[3 * ctx.bid + 1] = acc[1]; partials[3 * ctx.bid + 2] = acc[2]; }
    // inputs of the scalar leaves: fetched before the grid-wide handshake so that their latency is
    // not paid serially (seven dependent global round trips) by the last block's thread 0
    T la[7] = {0, 0, 0, 0, 0, 0, 0}, sg = 0, ubl = 0, ubc = 0, ctf = 0, sig = 0, asp = 0, ab0 = 0, z = 0;
    if (ctx.tid == 0) {
      if (add) {
        if (m.has_flu) la[0] = add[m.off_flu];
        la[1] = add[m.off_slp];
        if (m.matern) la[2] = add[m.off_ctf];
        if (m.has_dev) { la[3] = add[m.off_flx]; if (m.has_asp) la[4] = add[m.off_asp]; }
        la[5] = add[m.off_zm];
        if (m.has_scaling) la[6] = add[m.off_scl];
      }
      sg = scal_in[SC_SG]; ubl = scal_in[SC_UBL]; ubc = scal_in[SC_UBC]; ctf = scal_in[SC_CTF];
      sig = scal_in[SC_SIG]; asp = scal_in[SC_ASP]; ab0 = scal_in[SC_ABAR0]; z = scal_in[SC_Z];
    }
    if (ctx.last_block(counter)) {
      // all five totals in ONE reduction (each used to cost its own pair of barriers)
      T v[5] = {0, 0, 0, 0, 0};
      v[0] = strided_partial_sum(ctx, partials, ctx.nblk, 3); v[1] = strided_partial_sum(ctx, partials + 1, ctx.nblk, 3);
      v[2] = strided_partial_sum(ctx, partials + 2, ctx.nblk, 3);
      v[3] = strided_partial_sum(ctx, p5_partials, n_p5, 1);
      if (m.has_scaling) v[4] = strided_partial_sum(ctx, p3_partials, n_p3, 2);
      ctx.template block_sum_n<5>(v, scratch);
      if (ctx.tid == 0) {
        T sigbar = v[0], aspbar = v[1], dot = v[2] + v[3], sp = v[4];
        if (m.has_flu) leaf(m.off_flu, sg * m.flu_b, la[0], dot);
        leaf(m.off_slp, ubl * m.slp_b, la[1], dot);
        if (m.matern) leaf(m.off_ctf, ubc * ctf * m.ctf_b, la[2], dot);
        if (m.has_dev) {
          leaf(m.off_flx, sigbar * sig * m.flx_b, la[3], dot);
          if (m.has_asp) leaf(m.off_asp, aspbar * asp * m.asp_b, la[4], dot);
        }
        leaf(m.off_zm, ab0 * m.V * z * m.zm_b, la[5], dot);
        if (m.has_scaling) leaf(m.off_scl, sp * scl_factor, la[6], dot);
        scal[SC_DOT] = dot;
      }
    }
  }
};

// element functor for K <= 2 or no spectrum: nothing to scan
template <class T> struct NoElem { NB_HD NB_INLINE Aff<T> get(long) const { return aff_id<T>(); } };

}  // namespace nb
