// nifty_b200 -- register-resident radix-16 line FFTs (forward sign, e^{-i theta}) for the staged pass kernels.
//
// A CTA of 256 threads owns a TILE of 4096 complex elements = LPC = 4096 / n lines of length n = 2^LG
// (2^5 <= n <= 2^12).  A line is transformed by its team of n/16 threads; every thread keeps 16 elements in
// registers for the whole transform:
//
//   stage 1   radix-16 butterfly over the top 4 bits of the index (thread t holds x[t + (n/16) q], q < 16),
//             post-twiddle w_n^{t k1}
//   exchange  through shared memory (each thread writes its 16 values and reads 16 others)
//   stage 2   radix-16 (or the remaining 2^(LG-4) < 16) over the next bits, post-twiddle
//   exchange
//   stage 3   radix 2^(LG-8) over the last bits (LG > 8)
//
// and ends with thread t holding X[t + (n/16) q], q < 16 -- the SAME distribution the input had, so that a
// second transform (the adjoint half of pass P3) needs no reordering.  Compared with the in-place radix-8
// shared-memory FFT of nb_fft.cuh a 4096-point line makes 2 shared-memory round trips instead of 4, issues no
// barrier-separated butterfly reads, and exposes 16 independent elements per thread to the FP64 pipe.
//
// Index bookkeeping.  The transform is the in-place DIF on LOGICAL positions p (LG bits) split into the digit
// fields F3 = p[0:LG3), F2 = p[LG3:LG3+LG2), F1 = p[LG-4:LG) (LG2 = min(4, LG-4), LG3 = LG-4-LG2); after the
// transform F1, F2, F3 hold the output digits k1, k2, k3 with k = k1 | k2 << 4 | k3 << (4+LG2).  Which position a
// thread's register slot s holds in stage j is "thread bits | slot bits" (disjoint), and the exchange buffer
// address of a position is p ^ ((F1 ^ rs) & 15) + line * n (rs: a per-line constant) -- an XOR swizzle of the
// low nibble that makes every write and read pattern of both thread mappings free of bank conflicts for 8-byte
// words (tools/fft16_layout_check.py enumerates them).  All slot parts are compile-time constants, so an access
// costs one XOR.
//
// The exchange buffer holds 4096 8-byte words (32 KB): double data goes through it as two passes (real parts,
// then imaginary parts), float data in one pass of (re, im) pairs.
//
// Thread mappings: t-fast (consecutive threads = consecutive elements of one line; natural-layout global
// traffic coalesces) and r-fast (consecutive threads = the same element of consecutive lines; TRANSPOSED
// global traffic coalesces).  Moving a tile from line-major global memory into the r-fast register
// distribution is what the bulk-copy staging buffer does (the "staged transpose" of the passes).
#pragma once
#include "nb_common.cuh"
#include "nb_fft.cuh"
#include <cstring>
#include <vector>
#ifndef NB_EMU
#include <cuda.h>
#endif

namespace nb {

constexpr int F16_NT = 256;      // threads per CTA
constexpr int F16_TILE = 4096;   // complex elements per tile

template <int LG> struct F16Geom {
  static_assert(LG >= 5 && LG <= 12, "line length out of range for the register-resident FFT");
  static constexpr int N = 1 << LG;
  static constexpr int M = N / 16;              // threads per line
  static constexpr int LPC = F16_TILE / N;      // lines per tile
  static constexpr int LGL = 12 - LG;
  static constexpr int LG2 = (LG - 4) < 4 ? (LG - 4) : 4;
  static constexpr int LG3 = LG - 4 - LG2;
  static constexpr int R2 = 1 << LG2, R3 = 1 << LG3;
  // slot parts of the logical position in the three stage layouts
  static NB_HH constexpr int sb1(int s) { return s << (LG - 4); }
  static NB_HH constexpr int sb2(int s) { return ((s >> (4 - LG2)) << LG3) | ((s & ((1 << (4 - LG2)) - 1)) << (LG3 + 2 * LG2)); }
  static NB_HH constexpr int sb3(int s) { return LG3 == 0 ? sb2(s) : ((s >> (4 - LG3)) | ((s & ((1 << (4 - LG3)) - 1)) << (2 * LG3))); }
  // slot part of F1 (swizzle nibble) in the stage-2 layout: F1[LG2:4)
  static NB_HH constexpr int f1s2(int s) { return (s & ((1 << (4 - LG2)) - 1)) << LG2; }
  // exchange address constants: address = thread base ^ C(s)
  static NB_HH constexpr int c1(int s) { return sb1(s) ^ s; }
  static NB_HH constexpr int c2(int s) { return sb2(s) ^ f1s2(s); }
  static NB_HH constexpr int c3(int s) { return LG3 == 0 ? c2(s) : sb3(s); }
  // thread parts
  static NB_HH constexpr int tb2(int u) {
    const int f3 = u & ((1 << LG3) - 1);
    const int f1_mid = (u >> LG3) & ((1 << (LG2 - LG3)) - 1);    // F1[LG3:LG2)
    const int f1_lo = (u >> LG2) & ((1 << LG3) - 1);             // F1[0:LG3)
    return f3 | (f1_lo << (LG3 + LG2)) | (f1_mid << (LG3 + LG2 + LG3));
  }
  static NB_HH constexpr int f1t2(int u) {       // thread part of F1 in the stage-2 layout
    const int f1_mid = (u >> LG3) & ((1 << (LG2 - LG3)) - 1);
    const int f1_lo = (u >> LG2) & ((1 << LG3) - 1);
    return f1_lo | (f1_mid << LG3);
  }
  static NB_HH constexpr int tb3(int v) { return LG3 == 0 ? tb2(v) : (((v >> 4) << LG3) | ((v & 15) << (LG3 + 4))); }
  static NB_HH constexpr int f1t3(int v) { return LG3 == 0 ? f1t2(v) : (v & 15); }
  static NB_HH constexpr int rs(int r) { return LPC > 1 ? ((r << (LGL < 4 ? 4 - LGL : 0)) & 15) : 0; }
  // pitch (complex elements) of a line in the staging buffer: padded so that r-fast reads are conflict free
  template <class T> static NB_HH constexpr int stage_pitch(bool rfast) {
    const int ph = 128 / (int)sizeof(cplx<T>);          // lanes per shared-memory phase
    return N + ((rfast && LPC > 1) ? (ph / LPC > 1 ? ph / LPC : 1) : 0);
  }
};

template <class T> struct F16Const;
template <> struct F16Const<double> {
  static NB_HD NB_INLINE double c1() { return 0.92387953251128675613; }
  static NB_HD NB_INLINE double s1() { return 0.38268343236508977173; }
};
template <> struct F16Const<float> {
  static NB_HD NB_INLINE float c1() { return 0.92387953251128675613f; }
  static NB_HD NB_INLINE float s1() { return 0.38268343236508977173f; }
};

// 16-point DFT in registers, natural order in and out (two radix-4 passes)
template <class T> NB_HD NB_INLINE void dft16(cplx<T>* a) {
  const T h = FftConst<T>::rsqrt2(), c1 = F16Const<T>::c1(), s1 = F16Const<T>::s1();
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) {
    cplx<T> b[4] = {a[q0], a[q0 + 4], a[q0 + 8], a[q0 + 12]};
    dft4(b);
    a[q0] = b[0]; a[q0 + 4] = b[1]; a[q0 + 8] = b[2]; a[q0 + 12] = b[3];     // a[q0 + 4 k0] = B[q0][k0]
  }
  // twiddle a[q0 + 4 k0] by w16^(q0 k0)
  a[5] = cmul(a[5], cmake<T>(c1, -s1));                                       // w^1
  { cplx<T> v = a[6]; a[6] = cmake<T>((v.x + v.y) * h, (v.y - v.x) * h); }    // w^2
  a[7] = cmul(a[7], cmake<T>(s1, -c1));                                       // w^3
  { cplx<T> v = a[9]; a[9] = cmake<T>((v.x + v.y) * h, (v.y - v.x) * h); }    // w^2
  a[10] = cmul_mi(a[10]);                                                     // w^4 = -i
  { cplx<T> v = a[11]; a[11] = cmake<T>((v.y - v.x) * h, -(v.x + v.y) * h); } // w^6
  a[13] = cmul(a[13], cmake<T>(s1, -c1));                                     // w^3
  { cplx<T> v = a[14]; a[14] = cmake<T>((v.y - v.x) * h, -(v.x + v.y) * h); } // w^6
  a[15] = cmul(a[15], cmake<T>(-c1, s1));                                     // w^9
  cplx<T> o[16];
#pragma unroll
  for (int k0 = 0; k0 < 4; ++k0) {
    cplx<T> b[4] = {a[4 * k0], a[4 * k0 + 1], a[4 * k0 + 2], a[4 * k0 + 3]};
    dft4(b);
    o[k0] = b[0]; o[k0 + 4] = b[1]; o[k0 + 8] = b[2]; o[k0 + 12] = b[3];      // X[k0 + 4 k1]
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = o[k];
}

// 16 >> LR butterflies of radix 2^LR over the slots e + (q << (4 - LR))
template <int LR, class T> NB_HD NB_INLINE void dft_slots(cplx<T>* a) {
  if (LR == 0) return;
  if (LR == 4) { dft16(a); return; }
  constexpr int R = 1 << LR, G = 16 >> LR;
#pragma unroll
  for (int e = 0; e < G; ++e) {
    cplx<T> b[R];
#pragma unroll
    for (int q = 0; q < R; ++q) b[q] = a[e + q * G];
    dftR<LR>(b);
#pragma unroll
    for (int q = 0; q < R; ++q) a[e + q * G] = b[q];
  }
}

// a[e + q * G] *= w^q, q = 1 .. R-1 (every butterfly group e), powers formed by multiplication
template <int LR, class T> NB_HD NB_INLINE void twiddle_slots(cplx<T>* a, cplx<T> w) {
  constexpr int R = 1 << LR, G = 16 >> LR;
  cplx<T> wp = w;
#pragma unroll
  for (int q = 1; q < R; ++q) {
#pragma unroll
    for (int e = 0; e < G; ++e) a[e + q * G] = cmul(a[e + q * G], wp);
    if (q + 1 < R) wp = cmul(wp, w);
  }
}

// 8-byte exchange word
template <class T> struct XWord;
template <> struct XWord<double> {
  typedef double type;
  static constexpr int PASSES = 2;
  static NB_HD NB_INLINE double get(const cplx<double>& v, int part) { return part ? v.y : v.x; }
  static NB_HD NB_INLINE void set(cplx<double>& v, int part, double w) { if (part) v.y = w; else v.x = w; }
};
template <> struct XWord<float> {
  typedef cplx<float> type;
  static constexpr int PASSES = 1;
  static NB_HD NB_INLINE cplx<float> get(const cplx<float>& v, int) { return v; }
  static NB_HD NB_INLINE void set(cplx<float>& v, int, cplx<float> w) { v = w; }
};

template <class T, int LG, bool RFAST> struct Fft16 {
  typedef F16Geom<LG> G;
  typedef typename XWord<T>::type word_t;
  static constexpr int PASSES = XWord<T>::PASSES;
  static constexpr size_t XBYTES = (size_t)F16_TILE * sizeof(word_t);

  struct Th {
    int r, t;            // line of the tile, thread of the line team
    int b1, b2, b3;      // exchange base addresses of the three stage layouts
    cplx<T> w1, w2;      // w_n^t and w_(n/16)^(F3)
  };
  // tw: table of w_(n * stride)^j (stride 1: the table of the line length itself)
  static NB_HD NB_INLINE void init(Th& th, int tid, const cplx<T>* tw, int stride) {
    th.r = RFAST ? (tid & (G::LPC - 1)) : (tid >> (LG - 4));
    th.t = RFAST ? (tid >> G::LGL) : (tid & (G::M - 1));
    const int t = th.t, rsw = G::rs(th.r), rn = th.r * G::N;
    th.b1 = rn ^ t ^ rsw;
    th.b2 = rn ^ G::tb2(t) ^ G::f1t2(t) ^ rsw;
    th.b3 = rn ^ G::tb3(t) ^ G::f1t3(t) ^ rsw;
    th.w1 = ldg(tw + (size_t)t * stride);
    th.w2 = ldg(tw + (size_t)16 * (t & (G::R3 - 1)) * stride);
  }
  // index held by register slot s before the transform (x) and after it (k)
  static NB_HD NB_INLINE int elem(const Th& th, int s) { return th.t + s * G::M; }

  static NB_HD NB_INLINE void stage1(cplx<T>* a, const Th& th) {
    dft16(a);
    if (LG > 4) twiddle_slots<4>(a, th.w1);
  }
  static NB_HD NB_INLINE void stage2(cplx<T>* a, const Th& th) {
    dft_slots<G::LG2>(a);
    if (G::LG3 > 0) twiddle_slots<G::LG2>(a, th.w2);
  }
  static NB_HD NB_INLINE void stage3(cplx<T>* a, const Th&) { dft_slots<G::LG3>(a); }

  template <int J> static NB_HD NB_INLINE int xaddr(const Th& th, int s) {
    return J == 1 ? (th.b1 ^ G::c1(s)) : (J == 2 ? (th.b2 ^ G::c2(s)) : (th.b3 ^ G::c3(s)));
  }
  template <int J> static NB_HD NB_INLINE void xwrite(const cplx<T>* a, const Th& th, word_t* xb, int part) {
#pragma unroll
    for (int s = 0; s < 16; ++s) xb[xaddr<J>(th, s)] = XWord<T>::get(a[s], part);
  }
  template <int J> static NB_HD NB_INLINE void xread(cplx<T>* a, const Th& th, const word_t* xb, int part) {
#pragma unroll
    for (int s = 0; s < 16; ++s) XWord<T>::set(a[s], part, xb[xaddr<J>(th, s)]);
  }

  // The whole transform for a team object (see nb_passes2.cuh): TS must expose `a` and `th`.  The exchange
  // buffer must be free on entry; it is free again on return (ends with a barrier only if `sync_after`).
  // The exchanges only connect the threads of one line team (t-fast mapping: M consecutive threads): teams of
  // <= 32 threads synchronise with __syncwarp, larger ones on their own named barrier -- the lines of a tile
  // drift apart instead of meeting at 8 block-wide barriers per transform.
  template <class Team> static NB_HD NB_INLINE void tsync(Team& tm) {
    if (RFAST) tm.sync(); else tm.template team_sync<G::M>();
  }
  template <class Team> static NB_HD NB_INLINE void run(Team& tm, word_t* xb) {
    tm.all([&](int, typename Team::State& S) { stage1(S.a, S.th); xwrite<1>(S.a, S.th, xb, 0); });
    tsync(tm);
    if (PASSES == 2) {
      tm.all([&](int, typename Team::State& S) { xread<2>(S.a, S.th, xb, 0); });
      tsync(tm);
      tm.all([&](int, typename Team::State& S) { xwrite<1>(S.a, S.th, xb, 1); });
      tsync(tm);
      tm.all([&](int, typename Team::State& S) { xread<2>(S.a, S.th, xb, 1); });
    } else {
      tm.all([&](int, typename Team::State& S) { xread<2>(S.a, S.th, xb, 0); });
    }
    if (G::LG3 == 0) {
      tm.all([&](int, typename Team::State& S) { stage2(S.a, S.th); });
      return;
    }
    tsync(tm);
    tm.all([&](int, typename Team::State& S) { stage2(S.a, S.th); xwrite<2>(S.a, S.th, xb, 0); });
    tsync(tm);
    if (PASSES == 2) {
      tm.all([&](int, typename Team::State& S) { xread<3>(S.a, S.th, xb, 0); });
      tsync(tm);
      tm.all([&](int, typename Team::State& S) { xwrite<2>(S.a, S.th, xb, 1); });
      tsync(tm);
      tm.all([&](int, typename Team::State& S) { xread<3>(S.a, S.th, xb, 1); });
    } else {
      tm.all([&](int, typename Team::State& S) { xread<3>(S.a, S.th, xb, 0); });
    }
    tm.all([&](int, typename Team::State& S) { stage3(S.a, S.th); });
  }
};

// ---------------------------------------------------------------------------------------------
// Team: how a body addresses "every thread of the CTA".  CUDA: the calling thread with its state in
// registers.  tests/emu: 256 virtual threads run phase by phase on the host (barriers are phase ends).
// ---------------------------------------------------------------------------------------------
#ifdef NB_EMU
template <class TS> struct Team {
  typedef TS State;
  Ctx& ctx;
  std::vector<TS> ts;
  explicit Team(Ctx& c) : ctx(c), ts(F16_NT) {}
  template <class F> void all(const F& f) { for (int vt = 0; vt < F16_NT; ++vt) f(vt, ts[vt]); }
  template <class F> void one(const F& f) { f(); }
  template <class F> void coop(const F& f) { f(ctx); }      // old-style cooperative loop (NB_FOR / batched_loop over ctx)
  template <class F> void warp0(const F& f) { for (int lane = 0; lane < 32; ++lane) f(lane); }
  void sync() {}
  template <int TEAM> void team_sync() {}
};
#else
template <class TS> struct Team {
  typedef TS State;
  Ctx& ctx;
  TS ts;
  __device__ NB_INLINE explicit Team(Ctx& c) : ctx(c) {}
  template <class F> __device__ NB_INLINE void all(const F& f) { f(ctx.tid, ts); }
  template <class F> __device__ NB_INLINE void one(const F& f) { if (ctx.tid == 0) f(); }
  template <class F> __device__ NB_INLINE void coop(const F& f) { f(ctx); }
  template <class F> __device__ NB_INLINE void warp0(const F& f) { if (ctx.tid < 32) f(ctx.tid); }   // first warp, all lanes
  __device__ NB_INLINE void sync() { __syncthreads(); }
  // barrier among the TEAM consecutive threads that contain the caller (TEAM a power of two <= 256)
  template <int TEAM> __device__ NB_INLINE void team_sync() {
    if (TEAM >= F16_NT) __syncthreads();
    else if (TEAM <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + ctx.tid / TEAM), "r"(TEAM) : "memory");
  }
};
#endif

// ---------------------------------------------------------------------------------------------
// asynchronous staging: 1-D bulk copies (cp.async.bulk, the TMA engine) completing on an mbarrier
// ---------------------------------------------------------------------------------------------
#ifdef NB_EMU
struct Mbar { int dummy; };
inline void mbar_init(Mbar*, int) {}
inline void mbar_expect(Mbar*, unsigned) {}
inline void mbar_expect_tx(Mbar*, unsigned) {}
inline void mbar_arrive(Mbar*) {}
inline void warp_sync() {}
inline void mbar_wait(Mbar*, unsigned) {}
// 2-D tensor of 16-byte elements [nrows][row_len]; a gather fetches the column `col` of `box_rows` consecutive rows
struct TmaDesc { const unsigned char* base; long row_stride_bytes; int box_rows, pad; };
inline void tma_gather(void* dst, const TmaDesc* d, int col, int row, Mbar*) {
  for (int i = 0; i < d->box_rows; ++i)
    std::memcpy(reinterpret_cast<unsigned char*>(dst) + (size_t)i * 16, d->base + (size_t)(row + i) * d->row_stride_bytes + (size_t)col * 16, 16);
}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, Mbar*) { std::memcpy(dst, src, bytes); }
inline void bulk_prefetch_l2(const void*, unsigned) {}
inline void fence_async_smem() {}
#else
typedef unsigned long long Mbar;
__device__ NB_INLINE unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ NB_INLINE void mbar_init(Mbar* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ NB_INLINE void mbar_expect(Mbar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// transaction bytes without an arrival (every issuing lane announces its own copy; one lane arrives afterwards)
__device__ NB_INLINE void mbar_expect_tx(Mbar* b, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ NB_INLINE void mbar_arrive(Mbar* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ NB_INLINE void warp_sync() { __syncwarp(); }
// 2-D tensor map over 16-byte elements viewed as pairs of doubles: dims [2 * row_len][nrows], box {2, box_rows}.
// A gather brings the column `col` of `box_rows` consecutive rows into contiguous shared memory -- the
// transposing load of the staged passes (the TMA engine walks the strided rows; no LSU instructions).
struct TmaDesc { CUtensorMap map; int box_rows, pad; };
__device__ NB_INLINE void tma_gather(void* dst, const TmaDesc* d, int col, int row, Mbar* b) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(&d->map), "r"(2 * col), "r"(row), "r"(smem_u32(b)) : "memory");
}
__device__ NB_INLINE void mbar_wait(Mbar* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes a multiple of 16, both addresses 16-byte aligned
__device__ NB_INLINE void bulk_g2s(void* dst, const void* src, unsigned bytes, Mbar* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ NB_INLINE void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// orders earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes
__device__ NB_INLINE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

}  // namespace nb
