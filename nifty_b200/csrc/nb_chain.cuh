// nifty_b200 -- the O(K) amplitude chains of one metric-vector product as TWO persistent kernels.
//
// The tangent chain (spectrum tangent -> du table + scalars for the first pass) and the cotangent chain (mode-bin
// sums of the last pass -> spectrum / scalar cotangents) of nb_amp.cuh used to be five launches (segment sum, two
// chunk-aggregate kernels, two apply kernels), each a few resident waves of short, latency-bound blocks, together
// 27 % of a 4096^2 product.  Here every chain is ONE cooperative launch of resident CTAs with grid-wide barriers
// between its phases:
//
//   tangent   phase 0  ordered scan of the affine maps (nb_amp.cuh) inside each CTA's contiguous range; the maps
//                      relative to the start of the range are parked in the output table itself; CTA aggregates
//             phase 1  carry-in from the aggregates of the CTAs before (fixed order), streaming output
//   cotangent phase 0  segment sum N -> K: the W positions of a window of bins are gathered through ONE flat, coalesced,
//                      deeply unrolled loop into shared memory (two dependent global round trips instead of the three
//                      per bin of SegSumBody), then summed per bin in fixed order; g_b and partial sums
//             phase 1  global sums (every CTA, fixed order), ordered scan inside each CTA's range (reverse bin order),
//                      thread prefixes kept in shared memory; CTA aggregates
//             phase 2  carry-in, outputs (spectrum cotangent, scalar leaves, <add, out>)
//
// Reference semantics: nifty/re/correlated_field.py:481-516 and nifty/re/gauss_markov.py:102-114 through the element /
// output functors of nb_amp.cuh (JvpElem / JvpOut / VjpOut), which this file reuses unchanged.  Everything is
// deterministic: fixed CTA ranges, fixed composition order, no floating-point atomics.
//
// tests/emu runs the phases block after block with one shared-memory image per block (launch_coop in nb_backend.cuh).
#pragma once
#include "nb_amp.cuh"
#include "nb_fft16.cuh"

#ifndef NB_CHAIN_MINB
#define NB_CHAIN_MINB 3      // resident CTAs per SM the register budget allows (80 registers)
#endif

namespace nb {

// developer aid (NB200_COOP_DBG=1): per-CTA timestamps at the phase boundaries of the cotangent kernel
#ifdef NB_EMU
inline void chain_stamp(Ctx&, long long*, int) {}
#else
__device__ NB_INLINE void chain_stamp(Ctx& ctx, long long* dbg, int slot) {
  if (dbg && ctx.tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[(size_t)ctx.bid * 12 + slot] = t; }
}
#endif

// ---- ordered scan of 256 affine maps held in shared memory --------------------------------------
// in: tg[0..256) ; out: tg[v] = composition of tg[0..v) (exclusive), tg[256] = composition of all.  Called by every
// thread between two barriers (the first warp does the work).
#ifdef NB_EMU
template <class T> inline void scan256_aff(Ctx&, Aff<T>* tg) {
  Aff<T> run = aff_id<T>();
  for (int v = 0; v < SCAN_NT; ++v) { Aff<T> e = tg[v]; tg[v] = run; run = aff_compose(run, e); }
  tg[SCAN_NT] = run;
}
#else
template <class T> __device__ NB_INLINE void scan256_aff(Ctx& ctx, Aff<T>* tg) {
  if (ctx.tid >= 32) return;
  const int lane = ctx.tid;
  Aff<T> v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = tg[8 * lane + k];
  Aff<T> inc = v[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) inc = aff_compose(inc, v[k]);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Aff<T> n = shfl_up_aff(inc, o);
    if (lane >= o) inc = aff_compose(n, inc);
  }
  Aff<T> run = shfl_up_aff(inc, 1);
  if (lane == 0) run = aff_id<T>();
#pragma unroll
  for (int k = 0; k < 8; ++k) { tg[8 * lane + k] = run; run = aff_compose(run, v[k]); }
  if (lane == 31) tg[SCAN_NT] = inc;
}
#endif

// Ordered scan of the elements [lo, hi) of a CTA in rounds of CH = 256 * E.  Per round: every thread FETCHES the raw
// inputs of its E (coalesced) elements before any of them is turned into a map and staged (one global round trip per
// round instead of one per element: the compiler does not move loads across the shared-memory stores), composes its E
// consecutive staged maps, the 256 thread aggregates are scanned, and the thread rewrites its staged maps as
// store(prefix before the element, prefix after the element) relative to lo; `flush(position, staged value)` then runs
// coalesced over the round.  Returns the composition of the whole range (in every thread).
template <class T, int E, class Elem, class Store, class Flush>
NB_HD NB_INLINE Aff<T> cta_scan(Ctx& ctx, Aff<T>* el, Aff<T>* tg, long lo, long hi, const Elem& elem, const Store& store, const Flush& flush) {
  constexpr int CH = SCAN_NT * E;
  Aff<T> run = aff_id<T>();
  for (long p0 = lo; p0 < hi; p0 += CH) {
    for (int i0 = ctx.tid; i0 < CH; i0 += ctx.nthr * E) {
      typename Elem::Raw raw[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { const long q = p0 + i0 + e * ctx.nthr; raw[e] = elem.fetch(q < hi ? q : lo); }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = i0 + e * ctx.nthr;
        el[sidx<E>(i)] = (p0 + i < hi) ? elem.make(p0 + i, raw[e]) : aff_id<T>();
      }
    }
    ctx.sync();
    NB_FOR(ctx, vt, SCAN_NT) {
      const Aff<T>* mine = el + vt * (E + 1);
      Aff<T> a = mine[0];
#pragma unroll
      for (int e = 1; e < E; ++e) a = aff_compose(a, mine[e]);
      tg[vt] = a;
    }
    ctx.sync();
    scan256_aff(ctx, tg);
    ctx.sync();
    NB_FOR(ctx, vt, SCAN_NT) {
      Aff<T>* mine = el + vt * (E + 1);
      Aff<T> cur = aff_compose(run, tg[vt]);
#pragma unroll
      for (int e = 0; e < E; ++e) { const Aff<T> nxt = aff_compose(cur, mine[e]); mine[e] = store(cur, nxt); cur = nxt; }
    }
    run = aff_compose(run, tg[SCAN_NT]);
    ctx.sync();
    NB_FOR(ctx, i, CH) { if (p0 + i < hi) flush(p0 + i, el[sidx<E>(i)]); }
    ctx.sync();
  }
  return run;
}

// carry-in of CTA `bid` (composition of agg[0..bid)) and the composition of all nblk aggregates, fixed order.  One global
// round trip: every thread loads its (<= CARRY_PER) aggregates before composing them; the thread that owns `bid` also
// leaves the composition of its aggregates before `bid` in tg[257].
constexpr int CARRY_PER = 8;        // 256 * 8 CTAs at most (a cooperative grid has <= 148 * 8)
template <class T>
NB_HD NB_INLINE void cta_carry(Ctx& ctx, Aff<T>* tg, const Aff<T>* agg, int nblk, int bid, Aff<T>& carry, Aff<T>& grand) {
  const int per = (nblk + SCAN_NT - 1) / SCAN_NT;
  const int owner = bid / per;
  NB_FOR(ctx, vt, SCAN_NT) {
    Aff<T> a[CARRY_PER];
#pragma unroll
    for (int k = 0; k < CARRY_PER; ++k) { const int i = vt * per + k; a[k] = (k < per && i < nblk) ? agg[i] : aff_id<T>(); }
    Aff<T> v = aff_id<T>();
#pragma unroll
    for (int k = 0; k < CARRY_PER; ++k) {
      if (vt == owner && vt * per + k == bid) tg[SCAN_NT + 1] = v;
      v = aff_compose(v, a[k]);
    }
    tg[vt] = v;
  }
  ctx.sync();
  scan256_aff(ctx, tg);
  ctx.sync();
  carry = aff_compose(tg[owner], tg[SCAN_NT + 1]);
  grand = tg[SCAN_NT];
  ctx.sync();
}

template <class T, int E> struct ChainSmem {
  // staging array (CH + 256 padded slots) + the 256 thread aggregates and their total
  static constexpr int CH = SCAN_NT * E, STAGE = CH + SCAN_NT;
  static constexpr size_t STAGE_BYTES = (size_t)STAGE * sizeof(Aff<T>);
  static constexpr size_t TG_BYTES = (size_t)(SCAN_NT + 8) * sizeof(Aff<T>);
  static constexpr size_t BYTES = STAGE_BYTES + TG_BYTES;
};

// ---- tangent chain -----------------------------------------------------------------------------------
// JvpElem (nb_amp.cuh) with the inputs of an element fetched separately from the arithmetic, the scalars read once
template <class T> struct JvpElemB {
  const T* dt; const T* xs; const T* ts;       // log volumes, spectrum excitations, spectrum tangent
  T sig, asp, dsig_rel, dasp; bool dev;
  NB_HD NB_INLINE JvpElemB(const JvpElem<T>& e) {
    const AmpModel<T>& m = e.m;
    dev = m.has_dev != 0; dt = m.dt; xs = e.pos + m.off_spec; ts = e.t + m.off_spec;
    sig = asp = dsig_rel = dasp = 0;
    if (dev) {
      sig = e.scal[SC_SIG]; asp = e.scal[SC_ASP];
      dsig_rel = m.flx_b * e.t[m.off_flx];
      dasp = m.has_asp ? asp * m.asp_b * e.t[m.off_asp] : T(0);
    }
  }
  struct Raw { T dt, x0, x1, d0, d1; };
  NB_HD NB_INLINE Raw fetch(long b) const {
    const long j = b >= 2 ? b - 2 : 0;
    Raw r; r.dt = dt[j]; r.x0 = xs[2 * j]; r.x1 = xs[2 * j + 1]; r.d0 = ts[2 * j]; r.d1 = ts[2 * j + 1];
    return r;
  }
  NB_HD NB_INLINE Aff<T> make(long b, const Raw& r) const {
    if (b < 2) return aff_id<T>();
    const T sd = sig * nb_sqrt(r.dt), q = nb_sqrt(r.dt * r.dt / T(12) + asp);
    const T dr1 = sd * (dsig_rel * r.x1 + r.d1);
    const T dr0 = sd * q * (dsig_rel * r.x0 + r.d0) + sd * r.x0 * dasp / (T(2) * q) + T(0.5) * r.dt * dr1;
    Aff<T> e; e.a = r.dt; e.b = dr0; e.c = dr1;
    return e;
  }
};

template <class T> struct TanChainParams {
  JvpElem<T> elem; JvpOut<T> out;
  long n;                  // K
  long per;                // scan positions per CTA (multiple of 256 E)
  Aff<T>* agg;             // [nblk]
  unsigned* bar;           // grid barrier words {arrivals, generation}
  int nph;                 // phases to run (kPhases; fewer only as a timing aid)
};
template <class T, int E> struct TanChainBody {
  typedef TanChainParams<T> Params;
  static constexpr int kMinBlocks = NB_CHAIN_MINB;
  static constexpr int kPhases = 2;
  typedef ChainSmem<T, E> SM;
  static size_t smem_bytes() { return SM::BYTES; }
  static NB_HD void phase(int ph, Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    Aff<T>* el = reinterpret_cast<Aff<T>*>(sm);
    Aff<T>* tg = reinterpret_cast<Aff<T>*>(sm + SM::STAGE_BYTES);
    const bool dev = p.elem.m.has_dev != 0;
    const long lo = (long)ctx.bid * p.per, hi = lo + p.per < p.n ? lo + p.per : p.n;
    cplx<T>* ad = p.out.ad;
    if (ph == 0) {
      if (!dev) return;
      // the x-row of the map up to and including every element, relative to lo, parked in the output table
      const JvpElemB<T> elem(p.elem);
      auto store = [](const Aff<T>&, const Aff<T>& after) { return after; };
      auto flush = [&](long q, const Aff<T>& a) { ad[q] = cmake<T>(a.a, a.b); };
      Aff<T> tot = cta_scan<T, E>(ctx, el, tg, lo, hi, elem, store, flush);
      if (ctx.tid == 0) p.agg[ctx.bid] = tot;
    } else {
      Aff<T> carry = aff_id<T>(), grand = aff_id<T>();
      if (dev) cta_carry(ctx, tg, p.agg, ctx.nblk, ctx.bid, carry, grand);
      const T X = carry.b, Y = carry.c;
      T acc[4] = {0, 0, 0, 0};
      for (long q0 = lo + ctx.tid; q0 < hi; q0 += (long)ctx.nthr * 4) {       // four positions per thread in flight
        typename JvpOut<T>::Pre pl[4]; cplx<T> ab[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long q = q0 + (long)u * ctx.nthr, qq = q < hi ? q : lo;
          pl[u] = p.out.load(qq);
          ab[u] = dev ? ad[qq] : cmake<T>(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long q = q0 + (long)u * ctx.nthr;
          if (q < hi) p.out.put(q, pl[u], T(0), T(0), X + ab[u].x * Y + ab[u].y, T(0), grand.b, acc);
        }
      }
      p.out.finish(ctx, acc, grand.b, reinterpret_cast<void*>(el));
    }
  }
};

// ---- cotangent chain ---------------------------------------------------------------------------------
// element of the reverse scan (VjpElem of nb_amp.cuh) with the global sums passed by value: they are produced inside
// the same kernel
template <class T> struct VjpElemB {
  const T* dt; const T* g; const T* wS; long K; T sg, ubl_last, kappa;
  struct Raw { T dt, g, w; };
  NB_HD NB_INLINE Raw fetch(long p) const {
    const long j = (K - 3) - p;
    Raw r; r.dt = dt[j]; r.g = g[j + 2]; r.w = wS[j + 2];
    return r;
  }
  NB_HD NB_INLINE Aff<T> make(long p, const Raw& r) const {
    T c = kappa * r.g - T(0.5) * sg * r.w;
    if (p == 0) c -= ubl_last;             // bin K - 1
    Aff<T> e; e.a = r.dt; e.b = r.dt * c; e.c = c; return e;
  }
};

template <class T> struct CotChainParams {
  VjpOut<T> out;
  long nj;                 // K - 2 scan positions (0: no spectrum)
  long per;
  Aff<T>* agg;
  unsigned* bar;
  int nph;
  Aff<T>* rel;             // [nj] (x before, y after) of every element relative to the start of its CTA's range
  // segment sum
  const T* W; const int* order; const int* offs; const T* amp; const T* ellv; const T* cv;
  T* g;                    // [K]
  T* segpart;              // [nblk][3]
  T* scal;
  const int* seg_b0;       // [nblk + 1] first bin of every CTA in the segment sum
  const int* seg_lg;       // [nblk] log2(lanes per bin) of every CTA
  long long* dbg;          // developer aid: [nblk][12] timestamps (or null)
  int dbg_twice;
};
template <class T, int E> struct CotChainBody {
  typedef CotChainParams<T> Params;
  static constexpr int kMinBlocks = NB_CHAIN_MINB;
  static constexpr int kPhases = 3;
  typedef ChainSmem<T, E> SM;
  // Phase 0: segment sum N -> K.  A CTA owns a contiguous range of bins (cut by the host so that every CTA has about the
  // same number of W positions) and a lane count L = 2^lg per bin chosen from the mean population of its range.  Item
  // (bin, lane) adds the values at the positions beg + lane, beg + lane + L, ... of the bin in that order -- four position
  // -> value chains of two items in flight per thread, no barrier inside a round of 1024 items -- and the lanes of a bin
  // are combined in lane order: a fixed order, hence bit-reproducible.  g_b = A_b * sum; partial sums of g, g l, g c.
  // (Measured alternatives, profiles/r3_notes.md: staging the values of a window of bins in shared memory through one flat
  // gather loop -- with or without a bin-ordered copy of W in global memory -- was slower: 50 - 64 us against 41 us.)
  static NB_HD void seg_phase(Ctx& ctx, const Params& p, unsigned char* sm, int slot0 = 0) {
    T* ssum = reinterpret_cast<T*>(sm);          // [1024] lane partials
    const AmpModel<T>& m = p.out.m;
    const long B0 = p.seg_b0[ctx.bid], B1 = p.seg_b0[ctx.bid + 1];
    const int lg = p.seg_lg[ctx.bid], L = 1 << lg;
    T a0 = 0, a1 = 0, a2 = 0;
    chain_stamp(ctx, p.dbg, slot0);
    const long nitem = (B1 - B0) << lg;
    for (long base = 0; base < nitem; base += 1024) {
      const long lim = base + 1024 < nitem ? base + 1024 : nitem;
      constexpr int U = 2, J = 4;
      for (long i0 = base + ctx.tid; i0 < lim; i0 += ctx.nthr * U) {
        int beg[U], end[U]; T s[U];
        int maxc = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long i = i0 + (long)u * ctx.nthr;
          beg[u] = end[u] = 0; s[u] = 0;
          if (i < lim) { const long b = B0 + (i >> lg); beg[u] = p.offs[b] + (int)(i & (L - 1)); end[u] = p.offs[b + 1]; }
          const int c = end[u] > beg[u] ? (end[u] - beg[u] + L - 1) >> lg : 0;
          maxc = c > maxc ? c : maxc;
        }
        for (int k = 0; k < maxc; k += J) {
          int o[U][J]; T v[U][J];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < J; ++j) { const int q = beg[u] + ((k + j) << lg); o[u][j] = p.order[q < end[u] ? q : 0]; }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < J; ++j) v[u][j] = p.W[o[u][j]];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < J; ++j) { const int q = beg[u] + ((k + j) << lg); if (q < end[u]) s[u] += v[u][j]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { const long i = i0 + (long)u * ctx.nthr; if (i < lim) ssum[i - base] = s[u]; }
      }
      ctx.sync();
      const long nb = (lim - base) >> lg, bb0 = B0 + (base >> lg);
      constexpr int FU = 4;                      // (the table values of FU bins are loaded before any is used)
      for (long i0 = ctx.tid; i0 < nb; i0 += ctx.nthr * FU) {
        T ra[FU], rl[FU], rc[FU];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const long bi = i0 + (long)u * ctx.nthr, b = bb0 + (bi < nb ? bi : 0);
          ra[u] = p.amp[b]; rl[u] = p.ellv[b]; rc[u] = p.cv ? p.cv[b] : T(0);
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const long bi = i0 + (long)u * ctx.nthr, b = bb0 + bi;
          if (bi >= nb) continue;
          T abar = 0;
          for (int l = 0; l < L; ++l) abar += ssum[(bi << lg) + l];
          const T gb = (b == 0) ? T(0) : abar * ra[u];
          p.g[b] = gb;
          if (b == 0) p.scal[SC_ABAR0] = abar;
          a0 += gb; a1 += gb * rl[u]; a2 += gb * rc[u];
        }
      }
      ctx.sync();
    }
    chain_stamp(ctx, p.dbg, slot0 + 1);
    T v3[3] = {a0, a1, a2};
    ctx.template block_sum_n<3>(v3, reinterpret_cast<void*>(ssum));
    if (ctx.tid == 0) { p.segpart[3 * ctx.bid] = v3[0]; p.segpart[3 * ctx.bid + 1] = v3[1]; p.segpart[3 * ctx.bid + 2] = v3[2]; }
    chain_stamp(ctx, p.dbg, slot0 + 2);
  }
  static size_t smem_bytes() { return SM::BYTES; }

  static NB_HD void phase(int ph, Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    Aff<T>* el = reinterpret_cast<Aff<T>*>(sm);
    Aff<T>* tg = reinterpret_cast<Aff<T>*>(sm + SM::STAGE_BYTES);
    const AmpModel<T>& m = p.out.m;
    if (ph == 0) {
      seg_phase(ctx, p, sm);
      if (p.dbg_twice) { ctx.sync(); seg_phase(ctx, p, sm, 8); }     // developer aid: the same pass again with its inputs hot in L2
      return;
    }
    const long lo = (long)ctx.bid * p.per, hi = lo + p.per < p.nj ? lo + p.per : p.nj;
    if (ph == 1) {
      chain_stamp(ctx, p.dbg, 3);
      // the global sums, by every CTA in the same fixed order
      const T cwl = p.scal[SC_CWL], cwc = p.scal[SC_CWC], llast = m.ell[m.K - 1];      // (in flight with the partial sums)
      T v[3];
      v[0] = strided_partial_sum(ctx, p.segpart, ctx.nblk, 3); v[1] = strided_partial_sum(ctx, p.segpart + 1, ctx.nblk, 3);
      v[2] = strided_partial_sum(ctx, p.segpart + 2, ctx.nblk, 3);
      ctx.template block_sum_n<3>(v, reinterpret_cast<void*>(tg));
      const T kappa = m.kind_power ? T(0.5) : T(1);
      const T sg = v[0], sgl = v[1], sgc = v[2];
      const T ubl = kappa * sgl - T(0.5) * sg * cwl, ubc = kappa * sgc - T(0.5) * sg * cwc;
      if (ctx.bid == 0 && ctx.tid == 0) {
        p.scal[SC_SG] = sg; p.scal[SC_SGL] = sgl; p.scal[SC_SGC] = sgc; p.scal[SC_UBL] = ubl; p.scal[SC_UBC] = ubc;
      }
      if (p.nj <= 0) return;
      VjpElemB<T> elem; elem.dt = m.dt; elem.g = p.g; elem.wS = p.out.wS; elem.K = m.K; elem.sg = sg; elem.kappa = kappa;
      elem.ubl_last = ubl / llast;
      // what the output of an element needs: x BEFORE and y AFTER it (VjpOut::put), with the dt sum before it
      auto store = [](const Aff<T>& before, const Aff<T>& after) { Aff<T> r; r.a = before.a; r.b = before.b; r.c = after.c; return r; };
      auto flush = [&](long q, const Aff<T>& a) { p.rel[q] = a; };
      Aff<T> tot = cta_scan<T, E>(ctx, el, tg, lo, hi, elem, store, flush);
      if (ctx.tid == 0) p.agg[ctx.bid] = tot;
      chain_stamp(ctx, p.dbg, 4);
      return;
    }
    // phase 2: carry-in, outputs (a flat, coalesced loop)
    chain_stamp(ctx, p.dbg, 5);
    T acc[4] = {0, 0, 0, 0};
    Aff<T> carry = aff_id<T>(), grand = aff_id<T>();
    if (p.nj > 0) {
      cta_carry(ctx, tg, p.agg, ctx.nblk, ctx.bid, carry, grand);
      const T X = carry.b, Y = carry.c;
      for (long q0 = lo + ctx.tid; q0 < hi; q0 += (long)ctx.nthr * 4) {
        typename VjpOut<T>::Pre pl[4]; Aff<T> r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long q = q0 + (long)u * ctx.nthr, qq = q < hi ? q : lo;
          pl[u] = p.out.load(qq);
          r[u] = p.rel[qq];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long q = q0 + (long)u * ctx.nthr;
          if (q < hi) p.out.put(q, pl[u], X + r[u].a * Y + r[u].b, T(0), T(0), Y + r[u].c, grand.b, acc);
        }
      }
    }
    p.out.finish(ctx, acc, grand.b, reinterpret_cast<void*>(el));
    chain_stamp(ctx, p.dbg, 6);
  }
};

}  // namespace nb
