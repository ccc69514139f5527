// nifty_b200 -- common definitions shared by every kernel body.
//
// Kernel *bodies* are written once as templates over an execution context `Ctx` that supplies
// (tid, nthr, bid, nblk), a block barrier and block reductions.  The CUDA build instantiates them
// with the device `Ctx` inside `__global__` wrappers (this is the only thing the product library
// libniftyb200.so contains).  tests/emu compiles the very same bodies with `-DNB_EMU` for a
// sequential host context (one "thread" per block, blocks run one after the other) so that index
// arithmetic, mirror logic and epilogues can be checked against the oracle in a container without
// a GPU.  The emulator is test infrastructure: it is never built into or loaded by the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>

#ifdef NB_EMU
#define NB_HD
#define NB_HH
#define NB_DEV
#define NB_INLINE inline
#else
#include <cuda_runtime.h>
#define NB_HD __device__
#define NB_HH __host__ __device__   // pure helpers also usable from host code
#define NB_DEV __device__
#define NB_INLINE __forceinline__
#endif

namespace nb {

template <class T>
struct alignas(2 * sizeof(T)) cplx {
  T x, y;
};

template <class T> NB_HH NB_INLINE cplx<T> cmake(T x, T y) { cplx<T> c; c.x = x; c.y = y; return c; }
template <class T> NB_HH NB_INLINE cplx<T> operator+(cplx<T> a, cplx<T> b) { return cmake<T>(a.x + b.x, a.y + b.y); }
template <class T> NB_HH NB_INLINE cplx<T> operator-(cplx<T> a, cplx<T> b) { return cmake<T>(a.x - b.x, a.y - b.y); }
template <class T> NB_HH NB_INLINE cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return cmake<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class T> NB_HH NB_INLINE cplx<T> cconj(cplx<T> a) { return cmake<T>(a.x, -a.y); }
// multiply by -i
template <class T> NB_HH NB_INLINE cplx<T> cmul_mi(cplx<T> a) { return cmake<T>(a.y, -a.x); }

// ---------------------------------------------------------------------------------------------
// execution contexts
// ---------------------------------------------------------------------------------------------
#ifdef NB_EMU
struct Ctx {
  int tid, nthr, bid, nblk;
  void sync() {}
  template <class T> T block_sum(T v, void*) { return v; }
  template <int N, class T> void block_sum_n(T*, void*) {}
  template <class T> T block_max(T v, void*) { return v; }
  // true for exactly one block: the one that finishes last (blocks run in order in the emulator)
  bool last_block(unsigned* counter) {
    unsigned old = *counter;
    *counter = (old >= (unsigned)(nblk - 1)) ? 0u : old + 1u;
    return old == (unsigned)(nblk - 1);
  }
};
template <class T> inline T ldg(const T* p) { return *p; }
#else
struct Ctx {
  int tid, nthr, bid, nblk;
  __device__ NB_INLINE void sync() { __syncthreads(); }
  // Deterministic block reduction (fixed shuffle tree, fixed warp order).  `scratch` must hold
  // >= 32 values of T and may be reused after the call returns (the call ends with a barrier).
  template <class T> __device__ NB_INLINE T block_sum(T v, void* scratch) {
    T* s = reinterpret_cast<T*>(scratch);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = tid >> 5, l = tid & 31, nw = (nthr + 31) >> 5;
    __syncthreads();
    if (l == 0) s[w] = v;
    __syncthreads();
    T r = 0;
    if (w == 0) {
      r = (l < nw) ? s[l] : T(0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
      if (l == 0) s[0] = r;
    }
    __syncthreads();
    r = s[0];
    __syncthreads();
    return r;
  }
  // N sums at once (same barrier count as one): v[0..N) reduced in place, result in every thread.
  // `scratch` must hold >= 32*N values of T.
  template <int N, class T> __device__ NB_INLINE void block_sum_n(T* v, void* scratch) {
    T* s = reinterpret_cast<T*>(scratch);
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    }
    int w = tid >> 5, l = tid & 31, nw = (nthr + 31) >> 5;
    __syncthreads();
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) s[k * 32 + w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        T r = (l < nw) ? s[k * 32 + l] : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
        if (l == 0) s[k * 32] = r;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = s[k * 32];
    __syncthreads();
  }
  template <class T> __device__ NB_INLINE T block_max(T v, void* scratch) {
    T* s = reinterpret_cast<T*>(scratch);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T u = __shfl_down_sync(0xffffffffu, v, o); v = u > v ? u : v; }
    int w = tid >> 5, l = tid & 31, nw = (nthr + 31) >> 5;
    __syncthreads();
    if (l == 0) s[w] = v;
    __syncthreads();
    T r = 0;
    if (w == 0) {
      r = (l < nw) ? s[l] : s[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { T u = __shfl_down_sync(0xffffffffu, r, o); r = u > r ? u : r; }
      if (l == 0) s[0] = r;
    }
    __syncthreads();
    r = s[0];
    __syncthreads();
    return r;
  }
  // "last block done" with a self-resetting counter; all global writes of the other blocks that
  // were issued before their call are visible to the block for which this returns true.
  __device__ NB_INLINE bool last_block(unsigned* counter) {
    __shared__ unsigned s_old;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_old = atomicInc(counter, (unsigned)(nblk - 1));
    __syncthreads();
    bool last = (s_old == (unsigned)(nblk - 1));
    if (last) __threadfence();
    return last;
  }
};
template <class T> __device__ NB_INLINE T ldg(const T* p) { return __ldg(p); }
template <> __device__ NB_INLINE cplx<double> ldg(const cplx<double>* p) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cmake<double>(v.x, v.y);
}
template <> __device__ NB_INLINE cplx<float> ldg(const cplx<float>* p) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return cmake<float>(v.x, v.y);
}
#endif

// Streaming load: data that is read once per pass bypasses L1 (ld.global.cg).  The passes keep 3 x 64 KB of
// shared memory per SM, which leaves ~25 KB of L1 for the twiddle / slot tables every butterfly reads; with
// the line data allocating in L1 as well those tables kept missing (ncu: 19 % L1 hit rate in P3).
#if defined(NB_EMU)
template <class T> inline T ld_stream(const T* p) { return *p; }
#elif defined(NB_NO_STREAM_LD)
template <class T> __device__ NB_INLINE T ld_stream(const T* p) { return *p; }
#else
template <class T> __device__ NB_INLINE T ld_stream(const T* p) { return __ldcg(p); }
template <> __device__ NB_INLINE cplx<double> ld_stream(const cplx<double>* p) {
  double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cmake<double>(v.x, v.y);
}
template <> __device__ NB_INLINE cplx<float> ld_stream(const cplx<float>* p) {
  float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  return cmake<float>(v.x, v.y);
}
#endif

// Read-only streaming load (ld.global.nc, no L1 allocation) for data the kernel never writes: unlike ld_stream the
// compiler may move it across stores, i.e. hoist the loads of a whole epilogue ahead of its first store.
#if defined(NB_EMU)
template <class T> inline T ld_ro(const T* p) { return *p; }
#else
__device__ NB_INLINE double ld_ro(const double* p) { double v; asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ NB_INLINE float ld_ro(const float* p) { float v; asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
#endif

// cooperative L2 prefetch of [p, p+bytes): turns the DRAM latency of a later phase into an L2 hit
#ifdef NB_EMU
inline void prefetch_l2(Ctx&, const void*, size_t) {}
#else
__device__ NB_INLINE void prefetch_l2(Ctx& ctx, const void* p, size_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  for (size_t o = (size_t)ctx.tid * 128; o < bytes; o += (size_t)ctx.nthr * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
}
#endif

// Batched strided loop over [0, cnt): a thread handles the items i0 + u * nthr, u < U, per round.
//   issue(i0, slot)    every independent global load of the round (from always-valid addresses)
//   gather(slot)       loads that depend on those (table lookups)
//   consume(i0, slot)  arithmetic, shared memory, stores
// PIPE = true issues round k+1 before round k is consumed (one round of loads always in flight; costs a
// second set of slot registers).  Measured per loop (profiles/r2_notes.md): it pays in the P5 epilogue only.
template <bool PIPE, int U, class Slot, class Issue, class Gather, class Consume>
NB_HD NB_INLINE void batched_loop(Ctx& ctx, int cnt, const Issue& issue, const Gather& gather, const Consume& consume) {
  const int step = ctx.nthr * U;
  Slot cur[U];
  int i0 = ctx.tid;
  if (i0 < cnt) { issue(i0, cur); gather(cur); }
  for (; i0 < cnt; i0 += step) {
    const bool more = i0 + step < cnt;
    if (PIPE) {
      Slot nxt[U];
      if (more) issue(i0 + step, nxt);
      consume(i0, cur);
      if (more) {
        gather(nxt);
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
      }
    } else {
      consume(i0, cur);
      if (more) { issue(i0 + step, cur); gather(cur); }
    }
  }
}

#define NB_FOR(ctx, i, count) for (int i = (ctx).tid; i < (int)(count); i += (ctx).nthr)

NB_HH NB_INLINE int fold_idx(int x, int n) { return x <= n - x ? x : n - x; }
NB_HH NB_INLINE int neg_idx(int x, int n) { return x == 0 ? 0 : n - x; }

template <class T> NB_HD NB_INLINE T nb_exp(T x);
template <> NB_HD NB_INLINE double nb_exp(double x) { return exp(x); }
template <> NB_HD NB_INLINE float nb_exp(float x) { return expf(x); }
template <class T> NB_HD NB_INLINE T nb_log(T x);
template <> NB_HD NB_INLINE double nb_log(double x) { return log(x); }
template <> NB_HD NB_INLINE float nb_log(float x) { return logf(x); }
template <class T> NB_HD NB_INLINE T nb_sqrt(T x);
template <> NB_HD NB_INLINE double nb_sqrt(double x) { return sqrt(x); }
template <> NB_HD NB_INLINE float nb_sqrt(float x) { return sqrtf(x); }

}  // namespace nb
