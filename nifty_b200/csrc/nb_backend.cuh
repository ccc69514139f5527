// nifty_b200 -- launch / memory shims.  CUDA build: real kernels, cudaMalloc, streams.
// tests/emu build (-DNB_EMU): the same bodies run block after block on the host (see nb_common.cuh).
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <typeinfo>
#include <vector>
#include "nb_common.cuh"
#include "nb_fft16.cuh"

namespace nb {

struct Error { std::string msg; };
inline std::string& last_error_ref() { static thread_local std::string s; return s; }
inline int fail(const std::string& m) { last_error_ref() = m; return 1; }

#ifdef NB_EMU
typedef void* stream_t;
inline int dev_set(int) { return 0; }
inline void* dev_alloc(size_t bytes) { void* p = std::calloc(bytes ? bytes : 1, 1); if (!p) throw Error{"out of host memory (emulator)"}; return p; }
inline void dev_free(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n, stream_t) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, stream_t) { std::memcpy(h, d, n); }
inline void d2d(void* d, const void* s, size_t n, stream_t) { std::memmove(d, s, n); }
inline void dev_zero(void* d, size_t n, stream_t) { std::memset(d, 0, n); }
inline void stream_sync(stream_t) {}
inline size_t max_smem_per_block() { return 227 * 1024; }
inline int sm_count() { return 148; }
inline void make_tma_desc(TmaDesc& d, const void* base, uint64_t row_len, uint64_t nrows, int box_rows) {
  (void)nrows;
  d.base = reinterpret_cast<const unsigned char*>(base); d.row_stride_bytes = (long)(row_len * 16); d.box_rows = box_rows; d.pad = 0;
}

template <class Body>
inline void launch(int grid, int block, size_t smem, stream_t, const typename Body::Params& p) {
  (void)block;
  static const bool trace = std::getenv("NB200_EMU_TRACE") != nullptr;    // which bodies ran (test aid)
  if (trace) std::fprintf(stderr, "[emu] %s grid=%d\n", typeid(Body).name(), grid);
  std::vector<char> sm(smem + 4096);
  for (int b = 0; b < grid; ++b) {
    Ctx ctx{0, 1, b, grid};
    Body::run(ctx, p, sm.data());
  }
}

// cooperative (multi-phase) bodies: `Body::phase(ph, ctx, params, smem)` for ph < Body::kPhases with a grid-wide barrier
// between phases; shared memory persists across the phases of a block.  Emulation: one shared-memory image per block.
inline int coop_max_blocks_per_sm(const void*, int, size_t smem) { return smem <= 100 * 1024 ? 2 : 1; }
template <class Body> inline int coop_blocks_per_sm(size_t smem) { return coop_max_blocks_per_sm(nullptr, 256, smem); }
template <class Body>
inline void launch_coop(int grid, int block, size_t smem, stream_t, const typename Body::Params& p) {
  (void)block;
  static const bool trace = std::getenv("NB200_EMU_TRACE") != nullptr;
  if (trace) std::fprintf(stderr, "[emu] coop %s grid=%d\n", typeid(Body).name(), grid);
  std::vector<std::vector<char>> sm((size_t)grid, std::vector<char>(smem + 4096));
  for (int ph = 0; ph < Body::kPhases && ph < p.nph; ++ph)
    for (int b = 0; b < grid; ++b) {
      Ctx ctx{0, 1, b, grid};
      Body::phase(ph, ctx, p, sm[(size_t)b].data());
    }
}
#else
typedef cudaStream_t stream_t;
#define NB_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) throw ::nb::Error{std::string(#expr) + ": " + cudaGetErrorString(e_)}; \
  } while (0)

inline int dev_set(int d) { NB_CUDA_CHECK(cudaSetDevice(d)); return 0; }
inline void* dev_alloc(size_t bytes) {
  void* p = nullptr;
  NB_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
  NB_CUDA_CHECK(cudaMemset(p, 0, bytes ? bytes : 1));
  return p;
}
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline void h2d(void* d, const void* h, size_t n, stream_t s) { NB_CUDA_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void* h, const void* d, size_t n, stream_t s) { NB_CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void* d, const void* s_, size_t n, stream_t s) { NB_CUDA_CHECK(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void dev_zero(void* d, size_t n, stream_t s) { NB_CUDA_CHECK(cudaMemsetAsync(d, 0, n, s)); }
inline void stream_sync(stream_t s) { NB_CUDA_CHECK(cudaStreamSynchronize(s)); }
inline size_t max_smem_per_block() {
  int dev = 0, v = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  NB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return (size_t)v;
}
inline int sm_count() {
  int dev = 0, v = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  NB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  return v;
}

// 2-D tensor map over a row-major matrix of 16-byte elements [nrows][row_len] (viewed as pairs of doubles), box =
// one element x box_rows rows: the transposing gather of the staged passes (nb_passes2.cuh)
inline void make_tma_desc(TmaDesc& d, const void* base, uint64_t row_len, uint64_t nrows, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult qr;
    void* fn = nullptr;
    NB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) throw Error{"nb200: cuTensorMapEncodeTiled is not available in this driver"};
    enc = reinterpret_cast<EncodeFn>(fn);
  }
  cuuint64_t dims[2] = {2 * row_len, nrows};
  cuuint64_t strides[1] = {row_len * 16};
  cuuint32_t box[2] = {2, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(&d.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error{"nb200: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"};
  d.box_rows = box_rows; d.pad = 0;
}

// bodies may define `static constexpr int kMinBlocks` (register budget hint for ptxas)
template <class B, class = void> struct MinBlocks { static constexpr int value = 1; };
template <class B> struct MinBlocks<B, decltype((void)B::kMinBlocks)> { static constexpr int value = B::kMinBlocks; };

template <class Body>
__global__ void __launch_bounds__(256, MinBlocks<Body>::value) nb_kernel(const __grid_constant__ typename Body::Params p) {
  extern __shared__ __align__(16) unsigned char nb_smem[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
  Body::run(ctx, p, nb_smem);
}

// number of kernel launches issued by this library (bench.py reports it as gpu_launches)
// (atomic: XLA / several host threads may launch concurrently, one thread per device)
inline std::atomic<unsigned long long>& launch_counter() { static std::atomic<unsigned long long> c{0}; return c; }

// optional per-kernel CUDA-event timing (instrumentation for bench.py's roofline section)
struct KernelTimer {
  bool on = false;
  struct Rec { const char* name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
};
inline KernelTimer& kernel_timer() { static thread_local KernelTimer t; return t; }      // (per host thread)

template <class Body>
inline void launch(int grid, int block, size_t smem, stream_t s, const typename Body::Params& p) {
  static std::atomic<size_t> configured[64];          // per device: largest dynamic shared-memory size set so far
  int dev = 0;
  NB_CUDA_CHECK(cudaGetDevice(&dev));
  smem += 512;   // reduction scratch behind the line buffers
  if (dev >= 0 && dev < 64 && smem > configured[dev].load(std::memory_order_acquire)) {
    // (racing threads may both set the attribute: harmless, the attribute only grows)
    NB_CUDA_CHECK(cudaFuncSetAttribute(nb_kernel<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t cur = configured[dev].load(std::memory_order_relaxed);
    while (cur < smem && !configured[dev].compare_exchange_weak(cur, smem, std::memory_order_release)) {}
  }
  static const bool trace = std::getenv("NB200_TRACE") != nullptr;     // developer aid: every launch on stderr
  if (trace) std::fprintf(stderr, "[nb200] launch %s grid=%d block=%d smem=%zu\n", typeid(Body).name(), grid, block, smem);
  KernelTimer& kt = kernel_timer();
  if (kt.on) {
    KernelTimer::Rec r; r.name = typeid(Body).name();
    NB_CUDA_CHECK(cudaEventCreate(&r.a)); NB_CUDA_CHECK(cudaEventCreate(&r.b));
    NB_CUDA_CHECK(cudaEventRecord(r.a, s));
    nb_kernel<Body><<<grid, block, smem, s>>>(p);
    NB_CUDA_CHECK(cudaEventRecord(r.b, s));
    kt.recs.push_back(r);
  } else {
    nb_kernel<Body><<<grid, block, smem, s>>>(p);
  }
  NB_CUDA_CHECK(cudaGetLastError());
  if (trace) { cudaError_t e = cudaStreamSynchronize(s); std::fprintf(stderr, "[nb200]   -> %s\n", cudaGetErrorString(e)); }
  ++launch_counter();
}
// ---- cooperative (multi-phase) kernels --------------------------------------------------------------------------------
// Grid-wide barrier for kernels launched with cudaLaunchCooperativeKernel (all CTAs resident): bar[0] counts arrivals
// (self-resetting), bar[1] is the generation the waiters spin on.  The fences make the global writes of every CTA before
// the barrier visible to every CTA after it.
__device__ NB_INLINE void grid_barrier(Ctx& ctx, unsigned* bar) {
  __syncthreads();
  if (ctx.tid == 0) {
    volatile unsigned* gen = bar + 1;
    const unsigned g = *gen;                 // read before arriving: the generation cannot advance until this CTA has arrived
    __threadfence();
    if (atomicInc(bar, (unsigned)(ctx.nblk - 1)) == (unsigned)(ctx.nblk - 1)) {
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}

template <class Body>
__global__ void __launch_bounds__(256, MinBlocks<Body>::value) nb_kernel_coop(const __grid_constant__ typename Body::Params p) {
  extern __shared__ __align__(16) unsigned char nb_smem[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
#pragma unroll 1
  for (int ph = 0; ph < Body::kPhases && ph < p.nph; ++ph) {     // (nph < kPhases: developer timing aid)
    if (ph) grid_barrier(ctx, p.bar);
    Body::phase(ph, ctx, p, nb_smem);
  }
}

// resident CTAs per SM of a cooperative body at this dynamic shared-memory size (0: does not fit)
template <class Body> inline int coop_blocks_per_sm(size_t smem) {
  smem += 512;
  NB_CUDA_CHECK(cudaFuncSetAttribute(nb_kernel_coop<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int nb = 0;
  NB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, nb_kernel_coop<Body>, 256, smem));
  return nb;
}

template <class Body>
inline void launch_coop(int grid, int block, size_t smem, stream_t s, const typename Body::Params& p) {
  smem += 512;
  static const bool trace = std::getenv("NB200_TRACE") != nullptr;
  if (trace) std::fprintf(stderr, "[nb200] coop launch %s grid=%d block=%d smem=%zu\n", typeid(Body).name(), grid, block, smem);
  void* args[1] = {const_cast<void*>(reinterpret_cast<const void*>(&p))};
  KernelTimer& kt = kernel_timer();
  KernelTimer::Rec r; r.name = typeid(Body).name();
  if (kt.on) {
    NB_CUDA_CHECK(cudaEventCreate(&r.a)); NB_CUDA_CHECK(cudaEventCreate(&r.b));
    NB_CUDA_CHECK(cudaEventRecord(r.a, s));
  }
  NB_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(nb_kernel_coop<Body>), dim3(grid), dim3(block), args, smem, s));
  if (kt.on) { NB_CUDA_CHECK(cudaEventRecord(r.b, s)); kt.recs.push_back(r); }
  if (trace) { cudaError_t e = cudaStreamSynchronize(s); std::fprintf(stderr, "[nb200]   -> %s\n", cudaGetErrorString(e)); }
  ++launch_counter();
}
#endif

template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { dev_free(p); }
  void alloc(size_t count) { dev_free(p); p = nullptr; p = reinterpret_cast<T*>(dev_alloc(count * sizeof(T))); n = count; }
  void upload(const std::vector<T>& h, stream_t s = 0) { alloc(h.size()); if (!h.empty()) { h2d(p, h.data(), h.size() * sizeof(T), s); stream_sync(s); } }
};

}  // namespace nb
