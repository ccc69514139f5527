// nifty_b200 -- staged axis passes: the hot chain of the fused metric-vector product (P1 -> [PC] -> P3 -> [PC] -> P5,
// nb_passes.cuh documents the decomposition and the reference semantics) rebuilt around the register-resident
// radix-16 line FFT of nb_fft16.cuh for line lengths 32 .. 4096, float64.
//
// Execution model (same skeleton for every pass):
//   * persistent CTAs (2 per SM, 256 threads, 128 registers), tile = 4096 complex elements = LPC lines;
//     CTA b owns tiles b, b + grid, ... (fixed assignment: partial sums stay bit-reproducible);
//   * TRANSPOSES ARE LOADS.  Every pass WRITES its lines contiguously (line-major, fully coalesced stores from
//     registers); the next pass, which needs lines along another axis, GATHERS them through a 2-D tensor map
//     (cp.async.bulk.tensor, box = 1 element x 256 rows, completion on an mbarrier) into a 64 KB staging buffer
//     while the previous tile is being transformed.  The TMA engine walks the strided rows (measured 4.9 TB/s
//     for 16-byte elements at 64 KB stride, tools/microbench/tma_gather_bench.cu) -- no scattered LSU traffic
//     (the transposed 16 / 32-byte stores of the generic bodies cost ~2 clocks per element);
//   * the later-phase streams of the next tile (Jacobian weights, tangent, excitations) are pulled into L2 with
//     bulk prefetches when its gather is issued;
//   * a thread reads its 16 elements of the staged tile ONCE and keeps them in registers through all butterfly
//     stages (two 32 KB exchanges per transform), the pointwise operator and, in P3, the second transform.
//
// Intermediate layouts of the staged chain (complex, row-major):
//   P1 out  [j0 (,j1)][k_last <= n_last/2]            PCa out [j0][k_last][k_mid]
//   P3 out  [x_last (,x_mid)][k_0 <= n_0/2]           PCb out [x_last][k_0][k_mid]
// i.e. P3 / P5 read line l = a * n_mid + km as the column l of a [n_line][(h+1) n_mid] matrix.
//
// Every body is written against Team<> (nb_fft16.cuh), so tests/emu runs the same source with 256 virtual
// threads per block on the host.  Shapes / modes these bodies do not cover run the generic bodies of
// nb_passes.cuh (the host decides per operator, Plan::staged_chain).
#pragma once
#include "nb_passes.cuh"
#include "nb_fft16.cuh"

#ifndef NB_P5F_CHK
#define NB_P5F_CHK 8      // amplitude gathers in flight per thread in the staged epilogue of the last pass
#endif

namespace nb {

// shared-memory carve-up of a staged pass
template <class T, int LG> struct StageLayout {
  typedef F16Geom<LG> G;
  typedef Fft16<T, LG, false> F;
  static constexpr size_t STAGE_BYTES = (size_t)F16_TILE * sizeof(cplx<T>);
  static constexpr size_t XB_OFF = STAGE_BYTES;
  static constexpr size_t INFO_OFF = XB_OFF + F::XBYTES;
  static constexpr size_t INFO_BYTES = (size_t)G::LPC * 48;
  static constexpr size_t BAR_OFF = INFO_OFF + INFO_BYTES;
  static constexpr size_t SCRATCH_OFF = BAR_OFF + 64;
  static constexpr size_t BYTES = SCRATCH_OFF;           // launch() appends the 512-byte reduction scratch
};

// block-wide sum of a per-thread value (fixed order)
#ifdef NB_EMU
template <class TM, class Get> inline auto team_sum(TM& tm, void*, const Get& get) -> decltype(get(tm.ts[0])) {
  decltype(get(tm.ts[0])) s = 0;
  for (int vt = 0; vt < F16_NT; ++vt) s += get(tm.ts[vt]);
  return s;
}
#else
template <class TM, class Get> __device__ NB_INLINE auto team_sum(TM& tm, void* scratch, const Get& get) -> decltype(get(tm.ts)) {
  return tm.ctx.block_sum(get(tm.ts), scratch);
}
#endif

// Gather of the lines of one tile (first warp, all lanes): line i of the tile is column col(i) of the tensor,
// fetched in pieces of box_rows rows; every lane announces the bytes of its own pieces, lane 0 arrives last.
// `col(i)` returns < 0 for lines that are not loaded.  Contiguous sources (lines stored one after the other by a
// generic body) use plain bulk copies instead.
template <class T, int LG, class ColFn>
NB_HD NB_INLINE void issue_lines(int lane, cplx<T>* stage, Mbar* bar, const TmaDesc* desc, const cplx<T>* contig, const ColFn& col, int row0) {
  typedef F16Geom<LG> G;
  constexpr int N = G::N, LPC = G::LPC;
  if (contig) {
    for (int i = lane; i < LPC; i += 32) {
      const long c = col(i);
      if (c < 0) continue;
      constexpr unsigned LINE = (unsigned)(N * sizeof(cplx<T>));
      constexpr unsigned PIECE = LINE > 32768u ? 32768u : LINE;
      mbar_expect_tx(bar, LINE);
      for (unsigned o = 0; o < LINE; o += PIECE)
        bulk_g2s(reinterpret_cast<char*>(stage + (size_t)i * N) + o, reinterpret_cast<const char*>(contig + c * N) + o, PIECE, bar);
    }
  } else {
    const int br = desc->box_rows, nb = N / br, items = LPC * nb;
    for (int it = lane; it < items; it += 32) {
      const int i = it / nb, b = it - i * nb;
      const long c = col(i);
      if (c < 0) continue;
      mbar_expect_tx(bar, (unsigned)(br * sizeof(cplx<T>)));
      tma_gather(stage + (size_t)i * N + (size_t)b * br, desc, (int)c, row0 + b * br, bar);
    }
  }
  warp_sync();
  if (lane == 0) mbar_arrive(bar);
}

// ---------------------------------------------------------------------------------------------
// Hermitian partner exchange shared by P3 and P1: after a transform thread (r, t) holds Z[k], k = t + M s.
// The split of the transform of two packed real sequences needs the pairs (Z[k], Z[n - k]) for k < n/2: the
// thread keeps k_i = t + M i, i < 8 (its slots 0..7) and receives Z[n - k_i] from thread (M - t) (its slots
// 15 - i) into a[8 + i]; thread 0 owns all its partners itself.  One pass: real and imaginary parts use the
// two halves of the exchange buffer.
// ---------------------------------------------------------------------------------------------
template <class T, int LG, class TM>
NB_HD NB_INLINE void partner_exchange(TM& tm, typename XWord<T>::type* xb) {
  typedef F16Geom<LG> G;
  constexpr int M = G::M;
  T* w = reinterpret_cast<T*>(xb);                 // 16 x 256 scalars
  // thread t sends its slots 8..15 (Z[t + M (8 + j)]) as words j; thread 0 is its own partner and needs
  // Z[0], Z[15 M], .., Z[9 M] instead: it sends slots (0, 15, .., 9) as words (7, 6, .., 0)
  tm.all([&](int vt, typename TM::State& S) {
    const bool t0 = S.th.t == 0;
    S.zh = S.a[8];                                  // Z[n/2] (meaningful for t == 0)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const cplx<T> v = t0 ? S.a[j == 7 ? 0 : 9 + j] : S.a[8 + j];
      w[j * F16_NT + vt] = v.x; w[(8 + j) * F16_NT + vt] = v.y;
    }
  });
  tm.sync();
  tm.all([&](int vt, typename TM::State& S) {
    const int t = S.th.t;
    const int pv = vt - t + ((M - t) & (M - 1));     // partner thread of the same line
#pragma unroll
    for (int i = 0; i < 8; ++i) S.a[8 + i] = cmake<T>(w[(7 - i) * F16_NT + pv], w[(15 - i) * F16_NT + pv]);
  });
}

// ---------------------------------------------------------------------------------------------
// P5 (staged): complex lines along the last axis -> two natural real rows + latent-space epilogue
// ---------------------------------------------------------------------------------------------
template <class T> struct P5FParams {
  TmaDesc desc;            // gather source (staged chain); unused when `contig` is set
  const cplx<T>* contig;   // lines stored contiguously [l][n] by a generic body (else null)
  MirrorGeom mg;
  int hmid1;
  const cplx<T>* tw;       // w_n^j, n entries
  T hsign;
  int line0, nlines;       // line range of this launch
  int ntiles;
  int prefetch;            // L2 bulk prefetch of the epilogue rows of the next tile
  int staged_epi;          // epilogue inputs through shared memory (run_staged)
  EpiAdjoint<T> epi;
};

struct L5Info { long rowA, rowB, fbase, wbase; int active, pad; };

template <class T, int LG, bool SE = false> struct P5FBody {
  typedef P5FParams<T> Params;
  static constexpr int kMinBlocks = 2;
  typedef Fft16<T, LG, false> F;
  typedef F16Geom<LG> G;
  typedef StageLayout<T, LG> SL;
  struct TS { cplx<T> a[16]; typename F::Th th; T acc; };
  static constexpr int N = G::N, H = N / 2, LPC = G::LPC, M = G::M;

  static NB_HD NB_INLINE bool line_info(const Params& p, int l, L5Info& q) {
    int lA = 0, lB = -1;
    const bool act = l < p.line0 + p.nlines && p.mg.resolve(l, lA, lB);
    const int nmid = 1 << p.mg.lg_mid;
    const int a = lA >> p.mg.lg_mid, km = lA & (nmid - 1);
    q.rowA = (long)lA * N; q.rowB = lB >= 0 ? (long)lB * N : -1L;
    q.fbase = ((long)a * p.hmid1 + fold_idx(km, nmid)) * (H + 1);
    q.wbase = (long)lA * (H + 1);
    q.active = act ? 1 : 0; q.pad = 0;
    return act;
  }
  static NB_HD NB_INLINE void issue(const Params& p, int lane, int tile, cplx<T>* stage, Mbar* bar) {
    const int l0 = p.line0 + tile * LPC;
    auto col = [&](int i) -> long { L5Info q; return line_info(p, l0 + i, q) ? (long)(l0 + i) : -1L; };
    issue_lines<T, LG>(lane, stage, bar, &p.desc, p.contig, col, 0);
    if (p.prefetch) {
      const unsigned rb = (unsigned)(N * sizeof(T));
      for (int i = lane; i < LPC; i += 32) {
        L5Info q;
        if (!line_info(p, l0 + i, q)) continue;
        if (p.epi.add) { bulk_prefetch_l2(p.epi.add + q.rowA, rb); if (q.rowB >= 0) bulk_prefetch_l2(p.epi.add + q.rowB, rb); }
        if (p.epi.xi) { bulk_prefetch_l2(p.epi.xi + q.rowA, rb); if (q.rowB >= 0) bulk_prefetch_l2(p.epi.xi + q.rowB, rb); }
      }
    }
  }

  // ---- epilogue with its inputs in shared memory ----------------------------------------------------------
  // The register-resident epilogue below pays two dependent global round trips per chunk of four elements (bin index ->
  // amplitude, beside the streaming rows): eight per tile, the largest part of the tile time at 16 warps per SM.  Here
  //   * the bin-index rows of the tile are copied to shared memory while the gather of the tile lands,
  //   * the `add` rows (A and B of every line: exactly the 64 KB of the staging buffer) arrive by bulk copies during the
  //     transform, the `xi` A rows (32 KB: the exchange buffer) by bulk copies issued right after it,
  //   * so the epilogue is two rounds of eight amplitude gathers (+ the xi B row from global memory) per thread.
  // `add` may alias `out`: the rows of a tile are copied before any of its outputs is stored.
  static constexpr size_t SIDX_OFF = SL::BYTES + 512;
  static constexpr size_t SIDX_BYTES = (size_t)(F16_TILE / 2 + LPC + 8) * sizeof(int);
  static constexpr size_t BYTES_STAGED = SL::BYTES + SIDX_BYTES;
  static NB_HD NB_INLINE void issue_rows(const Params& p, int lane, int tile, const T* src, T* dst, bool both, Mbar* bar) {
    const int l0 = p.line0 + tile * LPC;
    constexpr unsigned ROW = (unsigned)(N * sizeof(T));
    constexpr unsigned PIECE = ROW > 32768u ? 32768u : ROW;
    for (int i = lane; i < LPC; i += 32) {
      L5Info q;
      if (!line_info(p, l0 + i, q)) continue;
      const bool hasB = both && q.rowB >= 0;
      mbar_expect_tx(bar, hasB ? 2 * ROW : ROW);
      T* d = dst + (size_t)i * (both ? 2 * N : N);
      for (unsigned o = 0; o < ROW; o += PIECE) {
        bulk_g2s(reinterpret_cast<char*>(d) + o, reinterpret_cast<const char*>(src + q.rowA) + o, PIECE, bar);
        if (hasB) bulk_g2s(reinterpret_cast<char*>(d + N) + o, reinterpret_cast<const char*>(src + q.rowB) + o, PIECE, bar);
      }
    }
    warp_sync();
    if (lane == 0) mbar_arrive(bar);
  }
  static NB_HD void run_staged(Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
    T* sadd = reinterpret_cast<T*>(sm);
    typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + SL::XB_OFF);
    T* sxi = reinterpret_cast<T*>(sm + SL::XB_OFF);
    L5Info* li = reinterpret_cast<L5Info*>(sm + SL::INFO_OFF);
    Mbar* bar = reinterpret_cast<Mbar*>(sm + SL::BAR_OFF);
    Mbar* bar2 = bar + 1; Mbar* bar3 = bar + 2;
    void* scratch = reinterpret_cast<void*>(sm + SL::SCRATCH_OFF);
    int* sidx = reinterpret_cast<int*>(sm + SIDX_OFF);
    Team<TS> tm(ctx);
    tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); S.acc = 0; });
    tm.one([&]() { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); });
    tm.sync();
    int tile = ctx.bid;
    unsigned phase = 0;
    if (tile < p.ntiles) tm.warp0([&](int lane) { issue(p, lane, tile, stage, bar); });
    const T sg = p.hsign, iv = p.epi.invV;
    const EpiAdjoint<T>& E = p.epi;
    for (; tile < p.ntiles; tile += ctx.nblk) {
      const int l0 = p.line0 + tile * LPC;
      tm.all([&](int vt, TS&) { if (vt < LPC) { L5Info q; line_info(p, l0 + vt, q); li[vt] = q; } });
      tm.sync();
      tm.coop([&](Ctx& c) {              // bin-index rows -> shared memory (eight loads in flight per thread)
        constexpr int U = 8, TOT = LPC * (H + 1);
        for (int i0 = c.tid; i0 < TOT; i0 += c.nthr * U) {
          int r[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * c.nthr, ii = i < TOT ? i : 0, ln = ii / (H + 1), k = ii - ln * (H + 1);
            r[u] = li[ln].active ? ldg(E.idxf + li[ln].fbase + k) : 0;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { const int i = i0 + u * c.nthr; if (i < TOT) sidx[i] = r[u]; }
        }
      });
      tm.all([&](int, TS& S) {
        mbar_wait(bar, phase);
        if (li[S.th.r].active != 0) {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = stage[S.th.r * N + F::elem(S.th, s)];
        } else {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = cmake<T>(0, 0);
        }
      });
      tm.sync();
      if (E.add) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue_rows(p, lane, tile, E.add, sadd, true, bar2); });
      tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); });
      F::run(tm, xb);
      tm.sync();
      if (E.xi && E.W) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue_rows(p, lane, tile, E.xi, sxi, false, bar3); });
      tm.all([&](int, TS& S) {
        if (E.add) mbar_wait(bar2, phase);
        const L5Info q = li[S.th.r];
        if (!q.active) return;
        const bool hasB = q.rowB >= 0;
        const T* aArow = sadd + (size_t)S.th.r * 2 * N;
        const T* aBrow = aArow + N;
        const int* irow = sidx + S.th.r * (H + 1);
        constexpr int CHK = NB_P5F_CHK;
#pragma unroll
        for (int c = 0; c < 16; c += CHK) {
          T A[CHK], xB[CHK];
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const int x = F::elem(S.th, c + u), y = (N - x) & (N - 1);
            A[u] = ldg(E.amp + irow[fold_idx(x, N)]);
            xB[u] = (E.xi && E.W && hasB) ? ld_ro(E.xi + q.rowB + y) : T(0);
          }
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const int x = F::elem(S.th, c + u), y = (N - x) & (N - 1);
            const cplx<T> z = S.a[c + u];
            const T gA = (z.x + sg * z.y) * iv, gB = hasB ? (z.x - sg * z.y) * iv : T(0);
            const T aA = E.add ? aArow[x] : T(0);
            const T oA = A[u] * gA + aA;
            E.out[q.rowA + x] = oA;
            T dot = aA * oA;
            if (hasB) {
              const T aB = E.add ? aBrow[y] : T(0);
              const T oB = A[u] * gB + aB;
              E.out[q.rowB + y] = oB;
              dot += aB * oB;
            }
            S.acc += dot;
            S.a[c + u] = cmake<T>(gA, xB[u] * gB);     // kept for the mode-bin sums below
          }
        }
        if (E.W) {
          if (E.xi) mbar_wait(bar3, phase);
          const T* xrow = sxi + (size_t)S.th.r * N;
#pragma unroll
          for (int s = 0; s < 16; ++s) {
            const int x = F::elem(S.th, s);
            S.a[s].x = (E.xi ? xrow[x] : T(0)) * S.a[s].x + S.a[s].y;      // this thread's two mirror points
          }
        }
      });
      phase ^= 1u;
      tm.sync();          // the staged rows have been consumed: the staging buffer may take the next tile, the exchange buffer the sums
      if (tile + ctx.nblk < p.ntiles) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue(p, lane, tile + ctx.nblk, stage, bar); });
      if (E.W) {
        tm.all([&](int, TS& S) {
          if (!li[S.th.r].active) return;
          T* wv = reinterpret_cast<T*>(xb);
#pragma unroll
          for (int s = 0; s < 16; ++s) wv[S.th.r * N + F::elem(S.th, s)] = S.a[s].x;
        });
        tm.sync();
        tm.all([&](int, TS& S) {
          const L5Info q = li[S.th.r];
          if (!q.active) return;
          const T* xl = reinterpret_cast<const T*>(xb) + S.th.r * N;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int xf = S.th.t + M * j;
            T w = xl[xf];
            if (xf != 0) w += xl[N - xf];
            E.W[q.wbase + xf] = w;
          }
          if (S.th.t == 0) E.W[q.wbase + H] = xl[H];
        });
      }
      tm.sync();
    }
    if (E.partials) {
      T tot = team_sum(tm, scratch, [](const TS& S) { return S.acc; });
      tm.one([&]() { E.partials[ctx.bid] = tot; });
    }
  }

  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    if constexpr (SE) { run_staged(ctx, p, smem); return; }
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
    typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + SL::XB_OFF);
    L5Info* li = reinterpret_cast<L5Info*>(sm + SL::INFO_OFF);
    Mbar* bar = reinterpret_cast<Mbar*>(sm + SL::BAR_OFF);
    void* scratch = reinterpret_cast<void*>(sm + SL::SCRATCH_OFF);
    Team<TS> tm(ctx);
    tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); S.acc = 0; });
    tm.one([&]() { mbar_init(bar, 1); });
    tm.sync();
    int tile = ctx.bid;
    unsigned phase = 0;
    if (tile < p.ntiles) tm.warp0([&](int lane) { issue(p, lane, tile, stage, bar); });
    const T sg = p.hsign, iv = p.epi.invV;
    const EpiAdjoint<T>& E = p.epi;
    for (; tile < p.ntiles; tile += ctx.nblk) {
      const int l0 = p.line0 + tile * LPC;
      tm.all([&](int vt, TS&) { if (vt < LPC) { L5Info q; line_info(p, l0 + vt, q); li[vt] = q; } });
      tm.sync();
      tm.all([&](int, TS& S) {
        mbar_wait(bar, phase);
        if (li[S.th.r].active != 0) {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = stage[S.th.r * N + F::elem(S.th, s)];
        } else {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = cmake<T>(0, 0);
        }
      });
      phase ^= 1u;
      tm.sync();
      if (tile + ctx.nblk < p.ntiles) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue(p, lane, tile + ctx.nblk, stage, bar); });
      tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); });      // recomputed per tile: shorter live ranges, no spills
      F::run(tm, xb);
      tm.sync();
      // epilogue: thread holds Z[x], x = t + M s: gA[x] = Re + sg Im, gB[n-x] = Re - sg Im
      tm.all([&](int, TS& S) {
        const L5Info q = li[S.th.r];
        if (!q.active) return;
        const bool hasB = q.rowB >= 0;
        const long rB = hasB ? q.rowB : q.rowA;
        constexpr int CHK = 4;
#pragma unroll
        for (int c = 0; c < 16; c += CHK) {
          int b[CHK]; T aA[CHK], aB[CHK], xA[CHK], xB[CHK], A[CHK];
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const int x = F::elem(S.th, c + u), y = (N - x) & (N - 1);
            b[u] = ldg(E.idxf + q.fbase + fold_idx(x, N));
            aA[u] = E.add ? ld_ro(E.add + q.rowA + x) : T(0);
            aB[u] = E.add ? ld_ro(E.add + rB + y) : T(0);
            xA[u] = E.xi ? ld_ro(E.xi + q.rowA + x) : T(0);
            xB[u] = E.xi ? ld_ro(E.xi + rB + y) : T(0);
          }
#pragma unroll
          for (int u = 0; u < CHK; ++u) A[u] = ldg(E.amp + b[u]);
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const int x = F::elem(S.th, c + u), y = (N - x) & (N - 1);
            const cplx<T> z = S.a[c + u];
            const T gA = (z.x + sg * z.y) * iv, gB = hasB ? (z.x - sg * z.y) * iv : T(0);
            const T oA = A[u] * gA + aA[u];
            E.out[q.rowA + x] = oA;
            T dot = aA[u] * oA, ws = xA[u] * gA;
            if (hasB) {
              const T oB = A[u] * gB + aB[u];
              E.out[q.rowB + y] = oB;
              dot += aB[u] * oB; ws += xB[u] * gB;
            }
            S.acc += dot;
            S.a[c + u].x = ws;          // partial of the mode-bin sum W (this thread's two mirror points)
          }
        }
        if (E.W) {
          T* wv = reinterpret_cast<T*>(xb);
#pragma unroll
          for (int s = 0; s < 16; ++s) wv[S.th.r * N + F::elem(S.th, s)] = S.a[s].x;
        }
      });
      if (E.W) {
        tm.sync();
        tm.all([&](int, TS& S) {
          const L5Info q = li[S.th.r];
          if (!q.active) return;
          const T* xl = reinterpret_cast<const T*>(xb) + S.th.r * N;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int xf = S.th.t + M * j;
            T w = xl[xf];
            if (xf != 0) w += xl[N - xf];
            E.W[q.wbase + xf] = w;
          }
          if (S.th.t == 0) E.W[q.wbase + H] = xl[H];
        });
      }
      tm.sync();
    }
    if (E.partials) {
      T tot = team_sum(tm, scratch, [](const TS& S) { return S.acc; });
      tm.one([&]() { E.partials[ctx.bid] = tot; });
    }
  }
};

// ---------------------------------------------------------------------------------------------
// P3 (staged), fused metric pass: lines along axis 0 -> two real position-space lines -> x jl_a jl_b / V
// -> one complex line (A2[x] + i B2[n-x]) -> transform -> split into the two half spectra, stored line-major
// ---------------------------------------------------------------------------------------------
template <class T> struct P3FParams {
  TmaDesc desc;            // gather source: column l of [n][(h_a+1) n_mid]
  MirrorGeom mg;
  const cplx<T>* tw;
  T hsign;
  cplx<T>* out;            // [position line][k in 0..n/2]
  int line0, nlines, ntiles;
  int prefetch;
  int stage_jl;            // Jacobian weights through the staging buffer (bulk copies) instead of per-thread global loads
  PointOp<T> op;
};

template <class T, int LG> struct P3FBody {
  typedef P3FParams<T> Params;
  static constexpr int kMinBlocks = 2;
  typedef Fft16<T, LG, false> F;
  typedef F16Geom<LG> G;
  typedef StageLayout<T, LG> SL;
  struct TS { cplx<T> a[16]; typename F::Th th; T acc; cplx<T> zh; };
  static constexpr int N = G::N, H = N / 2, LPC = G::LPC, M = G::M;

  static NB_HD NB_INLINE bool line_info(const Params& p, int l, LineInfo& q) {
    int lA = 0, lB = -1;
    const bool act = l < p.line0 + p.nlines && p.mg.resolve(l, lA, lB);
    q.lA = lA; q.lB = lB; q.active = act ? 1 : 0; q.pad = 0;
    return act;
  }
  static NB_HD NB_INLINE void issue(const Params& p, int lane, int tile, cplx<T>* stage, Mbar* bar) {
    const int l0 = p.line0 + tile * LPC;
    auto col = [&](int i) -> long { LineInfo q; return line_info(p, l0 + i, q) ? (long)(l0 + i) : -1L; };
    issue_lines<T, LG>(lane, stage, bar, &p.desc, (const cplx<T>*)nullptr, col, 0);
    if (p.prefetch) {
      const unsigned rb = (unsigned)(N * sizeof(T));
      for (int i = lane; i < LPC; i += 32) {
        LineInfo q;
        if (!line_info(p, l0 + i, q)) continue;
        bulk_prefetch_l2(p.op.jl_a + (long)q.lA * N, rb);
        if (q.lB >= 0) bulk_prefetch_l2(p.op.jl_a + (long)q.lB * N, rb);
        if (p.op.jl_b != p.op.jl_a) {
          bulk_prefetch_l2(p.op.jl_b + (long)q.lA * N, rb);
          if (q.lB >= 0) bulk_prefetch_l2(p.op.jl_b + (long)q.lB * N, rb);
        }
      }
    }
  }

  // the Jacobian weights of the lines of a tile -> the staging buffer (jl of line i: A row at doubles [2 N i, 2 N i + N),
  // B row behind it), bulk copies completing on `bar`
  static NB_HD NB_INLINE void issue_jl(const Params& p, int lane, int tile, T* sj, Mbar* bar) {
    const int l0 = p.line0 + tile * LPC;
    constexpr unsigned ROW = (unsigned)(N * sizeof(T));
    constexpr unsigned PIECE = ROW > 32768u ? 32768u : ROW;
    for (int i = lane; i < LPC; i += 32) {
      LineInfo q;
      if (!line_info(p, l0 + i, q)) continue;
      mbar_expect_tx(bar, q.lB >= 0 ? 2 * ROW : ROW);
      for (unsigned o = 0; o < ROW; o += PIECE) {
        bulk_g2s(reinterpret_cast<char*>(sj + (size_t)i * 2 * N) + o, reinterpret_cast<const char*>(p.op.jl_a + (long)q.lA * N) + o, PIECE, bar);
        if (q.lB >= 0)
          bulk_g2s(reinterpret_cast<char*>(sj + (size_t)i * 2 * N + N) + o, reinterpret_cast<const char*>(p.op.jl_a + (long)q.lB * N) + o, PIECE, bar);
      }
    }
    warp_sync();
    if (lane == 0) mbar_arrive(bar);
  }

  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
    T* sj = reinterpret_cast<T*>(sm);
    typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + SL::XB_OFF);
    LineInfo* li = reinterpret_cast<LineInfo*>(sm + SL::INFO_OFF);
    Mbar* bar = reinterpret_cast<Mbar*>(sm + SL::BAR_OFF);
    Mbar* bar2 = bar + 1;
    void* scratch = reinterpret_cast<void*>(sm + SL::SCRATCH_OFF);
    Team<TS> tm(ctx);
    tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); S.acc = 0; S.zh = cmake<T>(0, 0); });
    tm.one([&]() { mbar_init(bar, 1); mbar_init(bar2, 1); });
    tm.sync();
    int tile = ctx.bid;
    unsigned phase = 0;
    if (tile < p.ntiles) tm.warp0([&](int lane) { issue(p, lane, tile, stage, bar); });
    const T sg = p.hsign, iv = p.op.invV;
    const T cshift = p.op.cshift_ptr ? p.op.cshift_scale * ldg(p.op.cshift_ptr) : T(0);
    const T* ja = p.op.jl_a; const T* jb = p.op.jl_b;
    const bool same = (ja == jb);
    // With one set of weights (the metric of ONE linearisation) the weights of the tile travel through the staging buffer
    // too: bulk copies issued when the spectrum has been read into registers, landing during the first transform -- the
    // pointwise loop then reads shared memory instead of paying four dependent L2 round trips per tile.
    const bool sjl = same && p.stage_jl != 0;
    for (; tile < p.ntiles; tile += ctx.nblk) {
      const int l0 = p.line0 + tile * LPC;
      tm.all([&](int vt, TS&) { if (vt < LPC) { LineInfo q; line_info(p, l0 + vt, q); li[vt] = q; } });
      tm.sync();
      tm.all([&](int, TS& S) {
        mbar_wait(bar, phase);
        if (li[S.th.r].active != 0) {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = stage[S.th.r * N + F::elem(S.th, s)];
        } else {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = cmake<T>(0, 0);
        }
      });
      tm.sync();
      if (sjl) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue_jl(p, lane, tile, sj, bar2); });
      else if (tile + ctx.nblk < p.ntiles) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue(p, lane, tile + ctx.nblk, stage, bar); });
      F::run(tm, xb);
      // pointwise operator on Z[x] -> A[x] = Re + sg Im (line lA), B[n-x] = Re - sg Im (line lB)
      tm.all([&](int, TS& S) {
        const LineInfo q = li[S.th.r];
        if (sjl) mbar_wait(bar2, phase);
        if (!q.active) return;
        const bool hasB = q.lB >= 0;
        const long iA = (long)q.lA * N, iB = (long)(hasB ? q.lB : q.lA) * N;
        const T* sA = sj + (size_t)S.th.r * 2 * N;
        const T* sB = hasB ? sA + N : sA;
        constexpr int CHK = 4;
#pragma unroll
        for (int c = 0; c < 16; c += CHK) {
          T mA[CHK], mB[CHK];
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const int x = F::elem(S.th, c + u), y = (N - x) & (N - 1);
            if (sjl) { mA[u] = sA[x]; mB[u] = sB[y]; mA[u] *= mA[u]; mB[u] *= mB[u]; }
            else {
              mA[u] = ld_ro(ja + iA + x); mB[u] = ld_ro(ja + iB + y);
              if (!same) { mA[u] *= ld_ro(jb + iA + x); mB[u] *= ld_ro(jb + iB + y); }
              else { mA[u] *= mA[u]; mB[u] *= mB[u]; }
            }
          }
#pragma unroll
          for (int u = 0; u < CHK; ++u) {
            const cplx<T> z = S.a[c + u];
            const T aX = mA[u] * ((z.x + sg * z.y) * iv + cshift);
            const T bY = hasB ? mB[u] * ((z.x - sg * z.y) * iv + cshift) : T(0);
            S.acc += aX + bY;
            S.a[c + u] = cmake<T>(aX, bY);
          }
        }
      });
      phase ^= 1u;
      tm.sync();
      if (sjl && tile + ctx.nblk < p.ntiles) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue(p, lane, tile + ctx.nblk, stage, bar); });
      // (recomputing the per-thread transform constants here instead of carrying them across the tile keeps the
      // kernel inside its 128-register budget: 692 bytes of spills per thread otherwise)
      tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); });
      F::run(tm, xb);
      tm.sync();
      partner_exchange<T, LG>(tm, xb);
      // FA[k] = (Z[k] + conj Z[n-k]) / 2 -> line lA,  FB[k] = (Z[n-k] - conj Z[k]) / 2i -> line lB
      tm.all([&](int, TS& S) {
        const LineInfo q = li[S.th.r];
        if (!q.active) return;
        const T half = T(0.5);
        cplx<T>* oA = p.out + (long)q.lA * (H + 1);
        cplx<T>* oB = p.out + (long)(q.lB >= 0 ? q.lB : q.lA) * (H + 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = S.th.t + M * i;
          const cplx<T> zk = S.a[i], zc = S.a[8 + i];
          oA[k] = cmake<T>(half * (zk.x + zc.x), half * (zk.y - zc.y));
          if (q.lB >= 0) oB[k] = cmake<T>(half * (zk.y + zc.y), half * (zk.x - zc.x));
        }
        if (S.th.t == 0) {
          oA[H] = cmake<T>(S.zh.x, T(0));
          if (q.lB >= 0) oB[H] = cmake<T>(S.zh.y, T(0));
        }
      });
      tm.sync();
    }
    if (p.op.partials) {
      T tot = team_sum(tm, scratch, [](const TS& S) { return S.acc; });
      tm.one([&]() { p.op.partials[2 * ctx.bid] = tot; p.op.partials[2 * ctx.bid + 1] = T(0); });
    }
  }
};

// ---------------------------------------------------------------------------------------------
// PC (staged): complex lines along the middle axis.  Line c = o * ncols + col is the column `col` of the rows
// o * n .. o * n + n - 1 of the source matrix; the result is stored as line c of out[c][k].
// ---------------------------------------------------------------------------------------------
template <class T> struct PCFParams {
  TmaDesc desc;
  int ncols;               // lines per outer index
  long nlines;             // total lines
  const int* omap;         // optional: outer index of the i-th group (chunked launches)
  const cplx<T>* tw;
  cplx<T>* out;
  int ntiles;
};

template <class T, int LG> struct PCFBody {
  typedef PCFParams<T> Params;
  static constexpr int kMinBlocks = 2;
  typedef Fft16<T, LG, false> F;
  typedef F16Geom<LG> G;
  typedef StageLayout<T, LG> SL;
  struct TS { cplx<T> a[16]; typename F::Th th; };
  static constexpr int N = G::N, LPC = G::LPC, M = G::M;

  // global line index of line i of the tile (through the outer map for chunked launches), -1 past the end
  static NB_HD NB_INLINE long line_of(const Params& p, int tile, int i) {
    const long c = (long)tile * LPC + i;
    if (c >= p.nlines) return -1;
    if (!p.omap) return c;
    const long oi = c / p.ncols;
    return (long)ldg(p.omap + oi) * p.ncols + (c - oi * p.ncols);
  }
  static NB_HD NB_INLINE void issue(const Params& p, int lane, int tile, cplx<T>* stage, Mbar* bar) {
    // every line has its own row offset
    const int br = p.desc.box_rows, nb = N / br, items = LPC * nb;
    for (int it = lane; it < items; it += 32) {
      const int i = it / nb, b = it - i * nb;
      const long c = line_of(p, tile, i);
      if (c < 0) continue;
      const long o = c / p.ncols;
      mbar_expect_tx(bar, (unsigned)(br * sizeof(cplx<T>)));
      tma_gather(stage + (size_t)i * N + (size_t)b * br, &p.desc, (int)(c - o * p.ncols), (int)(o * N) + b * br, bar);
    }
    warp_sync();
    if (lane == 0) mbar_arrive(bar);
  }
  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
    typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + SL::XB_OFF);
    Mbar* bar = reinterpret_cast<Mbar*>(sm + SL::BAR_OFF);
    Team<TS> tm(ctx);
    tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); });
    tm.one([&]() { mbar_init(bar, 1); });
    tm.sync();
    int tile = ctx.bid;
    unsigned phase = 0;
    if (tile < p.ntiles) tm.warp0([&](int lane) { issue(p, lane, tile, stage, bar); });
    for (; tile < p.ntiles; tile += ctx.nblk) {
      tm.all([&](int, TS& S) {
        mbar_wait(bar, phase);
        if (line_of(p, tile, S.th.r) >= 0) {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = stage[S.th.r * N + F::elem(S.th, s)];
        } else {
#pragma unroll
          for (int s = 0; s < 16; ++s) S.a[s] = cmake<T>(0, 0);
        }
      });
      phase ^= 1u;
      tm.sync();
      if (tile + ctx.nblk < p.ntiles) tm.warp0([&](int lane) { if (lane == 0) fence_async_smem(); warp_sync(); issue(p, lane, tile + ctx.nblk, stage, bar); });
      tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 1); });      // recomputed per tile: shorter live ranges, no spills
      F::run(tm, xb);
      tm.all([&](int, TS& S) {
        const long c = line_of(p, tile, S.th.r);
        if (c < 0) return;
        cplx<T>* o = p.out + c * N;
#pragma unroll
        for (int s = 0; s < 16; ++s) o[F::elem(S.th, s)] = S.a[s];
      });
      tm.sync();
    }
  }
};

// ---------------------------------------------------------------------------------------------
// P1 (staged) for the prologues that read the mode-bin table (mirror quads, see P1MBody): a tile is LPC/2
// mirror pairs of real rows; the scaled inputs are written to the staging buffer by the quad loop, the
// half-length complex transform runs in registers, the real-FFT split follows the partner exchange and the
// half spectrum of every row is stored contiguously (row-major).
// ---------------------------------------------------------------------------------------------
template <class T, class Pro> struct P1FParams {
  int lg_n;                // log2 of the REAL line length (complex length N = 2^(lg_n-1))
  int n_r, n_o;
  long in_ostride, in_rstride;
  const cplx<T>* tw; int lg_tw;     // table for the real length (2^lg_tw entries, lg_tw >= lg_n)
  cplx<T>* out;            // [o * n_r + row][k in 0..N]
  int ntiles;
  Pro pro;
};

template <class T, class Pro, int LG> struct P1FBody {
  typedef P1FParams<T, Pro> Params;
  static constexpr int kMinBlocks = 2;
  typedef Fft16<T, LG, false> F;
  typedef F16Geom<LG> G;
  typedef StageLayout<T, LG> SL;
  struct TS { cplx<T> a[16]; typename F::Th th; cplx<T> zh; };
  static constexpr int N = G::N, LPC = G::LPC, M = G::M, HR = LPC / 2;
  static_assert(LPC >= 2, "the mirror-pair tile needs at least two lines");

  static NB_HD NB_INLINE int row_of(int r, int i0, int n_r) {
    const int i = i0 + (r < HR ? r : r - HR);
    if (r < HR) return i;
    return i == 0 ? (n_r >> 1) : n_r - i;
  }
  // the quad loop of P1MBody, writing real element x of line L to sr[L * 2 N + x]
  static NB_HD NB_INLINE void prologue(Ctx& ctx, const Params& p, T* sr, int o, int i0) {
    const int n = 2 * N, h = N, lg_h = LG;
    const long in0 = (long)o * p.in_ostride;
#define NB_P1F_SLOT(L, x) ((L) * (2 * N) + (x))
    {
      constexpr int U = 4;
      const int cnt = HR << lg_h;
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
          int ra = (i != 0) ? i : 1, rb = p.n_r - ra;
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
          sr[NB_P1F_SLOT(rp, e)] = v[0]; sr[NB_P1F_SLOT(rp, n - e)] = v[1];
          sr[NB_P1F_SLOT(rp + HR, e)] = v[2]; sr[NB_P1F_SLOT(rp + HR, n - e)] = v[3];
        }
      };
      auto gather = [&](Slot* sl) {
#pragma unroll
        for (int u = 0; u < U; ++u) p.pro.gather(sl[u].pre);
      };
      batched_loop<false, U, Slot>(ctx, cnt, issue, gather, consume);
    }
    NB_FOR(ctx, k, 2 * LPC) {        // self-mirrored elements e = 0 and e = h of every line
      int r = k >> 1, e = (k & 1) ? h : 0;
      int rr = row_of(r, i0, p.n_r);
      sr[NB_P1F_SLOT(r, e)] = p.pro.single(in0 + rr * p.in_rstride, p.pro.bins(o, rr), e);
    }
    if (i0 == 0) {                   // pair 0 = rows 0 and n_r/2: different bins, no sharing
      NB_FOR(ctx, k, 2 * (n - 2)) {
        int r = (k & 1) ? HR : 0, x = (k >> 1) + 1;
        if (x >= h) ++x;
        int rr = row_of(r, 0, p.n_r);
        sr[NB_P1F_SLOT(r, x)] = p.pro.single(in0 + rr * p.in_rstride, p.pro.bins(o, rr), x);
      }
    }
#undef NB_P1F_SLOT
  }

  static NB_HD void run(Ctx& ctx, const Params& p, void* smem) {
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem);
    cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
    typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + SL::XB_OFF);
    Team<TS> tm(ctx);
    const int tsh = p.lg_tw - p.lg_n;
    tm.all([&](int vt, TS& S) { F::init(S.th, vt, p.tw, 2 << tsh); S.zh = cmake<T>(0, 0); });
    const int gpo = p.n_r / LPC;
    const T half = T(0.5);
    for (int tile = ctx.bid; tile < p.ntiles; tile += ctx.nblk) {
      const int o = tile / gpo, i0 = (tile % gpo) * HR;
      tm.coop([&](Ctx& c) { prologue(c, p, reinterpret_cast<T*>(stage), o, i0); });
      tm.sync();
      tm.all([&](int, TS& S) {
#pragma unroll
        for (int s = 0; s < 16; ++s) S.a[s] = stage[S.th.r * N + F::elem(S.th, s)];
      });
      F::run(tm, xb);
      tm.sync();
      partner_exchange<T, LG>(tm, xb);
      // X[k] = (e - i w^k d)/2, X[N-k] = conj(e + i w^k d)/2, e = Z[k] + conj Z[N-k], d = Z[k] - conj Z[N-k]
      tm.all([&](int, TS& S) {
        const int row = row_of(S.th.r, i0, p.n_r);
        cplx<T>* orow = p.out + ((long)o * p.n_r + row) * (N + 1);
        cplx<T> w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = ldg(p.tw + ((size_t)(S.th.t + M * i) << tsh));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = S.th.t + M * i;
          const cplx<T> zk = S.a[i], zc = cconj(S.a[8 + i]);
          const cplx<T> e = zk + zc, d = zk - zc;
          const cplx<T> m = cmul_mi(cmul(w[i], d));
          orow[k] = cmake<T>(half * (e.x + m.x), half * (e.y + m.y));
          orow[N - k] = cmake<T>(half * (e.x - m.x), -half * (e.y - m.y));
        }
        if (S.th.t == 0) orow[N / 2] = cmake<T>(S.zh.x, -S.zh.y);      // k = N/2 pairs with itself: X[N/2] = conj(Z[N/2])
      });
      tm.sync();
    }
  }
};

}  // namespace nb
