// Microbenchmark: how fast can the TMA engine gather a TRANSPOSED line of 16-byte elements?
//   source layouts for a logical array X[l][k] (l: lines, k: elements; we gather the column k = const):
//     strided  : addr = (l * K + k) * 16                         (line-major, element stride K * 16 B)
//     blocked  : addr = ((k / 8) * L + l) * 128 + (k % 8) * 16   (8 consecutive k share one 128-byte line)
//   a "tile" is one column of 4096 elements fetched by 16 tensor loads of box {16 B x 256}; the CTA then writes
//   the 64 KB tile to a line-major output with plain coalesced stores.  Also times the LSU alternative
//   (scattered LDG.128) and a plain contiguous bulk copy as the reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/microbench/tma_gather_bench.cu -o tools/microbench/tma_gather_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: strided tensor gather, 1: blocked tensor gather, 2: LSU scattered loads (strided), 3: contiguous bulk copy, 4: LSU scattered (blocked)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_gather(const __grid_constant__ CUtensorMap tmap, const double2* src, double2* out, int K, int L) {
  extern __shared__ __align__(128) unsigned char sm[];
  double2* st0 = reinterpret_cast<double2*>(sm);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 65536);
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  unsigned phase = 0;
  for (int k = blockIdx.x; k < K; k += gridDim.x) {
    if (MODE == 0 || MODE == 1 || MODE == 3) {
      if (tid == 0) {
        mbar_expect(bar, 65536);
        if (MODE == 0) { for (int c = 0; c < 16; ++c) tma_2d(st0 + c * 256, &tmap, 2 * k, c * 256, bar); }
        else if (MODE == 1) { for (int c = 0; c < 16; ++c) tma_3d(st0 + c * 256, &tmap, 2 * (k & 7), c * 256, k >> 3, bar); }
        else { bulk_g2s(st0, src + (size_t)k * 4096, 32768, bar); bulk_g2s(st0 + 2048, src + (size_t)k * 4096 + 2048, 32768, bar); }
      }
      mbar_wait(bar, phase); phase ^= 1;
#pragma unroll
      for (int s = 0; s < 16; ++s) out[(size_t)k * 4096 + tid + 256 * s] = st0[tid + 256 * s];
      __syncthreads();
    } else {
      double2 v[16];
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int l = tid + 256 * s;
        v[s] = MODE == 2 ? src[(size_t)l * K + k] : src[((size_t)(k >> 3) * L + l) * 8 + (k & 7)];
      }
#pragma unroll
      for (int s = 0; s < 16; ++s) out[(size_t)k * 4096 + tid + 256 * s] = v[s];
    }
  }
}

template <int MODE> void run(const char* name, const CUtensorMap& tm, const double2* src, double2* out, int K, int L) {
  const size_t smem = 65536 + 64;
  CK(cudaFuncSetAttribute(k_gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<MODE>, 256, smem));
  const int grid = std::min(K, 148 * occ);
  for (int i = 0; i < 2; ++i) k_gather<MODE><<<grid, 256, smem>>>(tm, src, out, K, L);
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int reps = 10;
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) k_gather<MODE><<<grid, 256, smem>>>(tm, src, out, K, L);
  CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
  // verify a few entries
  std::vector<double2> h(4096);
  int bad = 0;
  for (int k : {0, 5, K / 2 + 3, K - 1}) {
    CK(cudaMemcpy(h.data(), out + (size_t)k * 4096, 4096 * sizeof(double2), cudaMemcpyDeviceToHost));
    for (int l = 0; l < 4096; l += 97) {
      double want = MODE == 3 ? (double)(k * 4096 + l) : (double)l * 10000.0 + k;
      if (h[l].x != want) ++bad;
    }
  }
  const double gb = 2.0 * (double)K * 4096 * 16 / 1e9;
  printf("%-34s occ=%d  %8.3f ms  %7.1f GB/s (read+write)  %s\n", name, occ, ms, gb / (ms * 1e-3), bad ? "MISMATCH" : "ok");
}

int main() {
  const int K = 4096, L = 4096;      // 256 MiB array
  const size_t n = (size_t)K * L;
  std::vector<double2> hs(n), hb(n), hc(n);
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < K; ++k) {
      double2 v = make_double2((double)l * 10000.0 + k, 0.5);
      hs[(size_t)l * K + k] = v;
      hb[((size_t)(k >> 3) * L + l) * 8 + (k & 7)] = v;
    }
  for (size_t i = 0; i < n; ++i) hc[i] = make_double2((double)i, 0.0);
  double2 *ds, *db, *dc, *dout;
  CK(cudaMalloc(&ds, n * 16)); CK(cudaMalloc(&db, n * 16)); CK(cudaMalloc(&dc, n * 16)); CK(cudaMalloc(&dout, n * 16));
  CK(cudaMemcpy(ds, hs.data(), n * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), n * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dc, hc.data(), n * 16, cudaMemcpyHostToDevice));
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  CUtensorMap tms, tmb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)2 * K, (cuuint64_t)L};
    cuuint64_t strides[1] = {(cuuint64_t)K * 16};
    cuuint32_t box[2] = {2, 256}, es[2] = {1, 1};
    CUresult r = enc(&tms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ds, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode strided failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[3] = {16, (cuuint64_t)L, (cuuint64_t)K / 8};
    cuuint64_t strides[2] = {128, (cuuint64_t)L * 128};
    cuuint32_t box[3] = {2, 256, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tmb, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode blocked failed %d\n", (int)r); return 1; }
  }
  run<3>("contiguous bulk copy (reference)", tms, dc, dout, K, L);
  run<0>("TMA gather, strided [l][k]", tms, ds, dout, K, L);
  run<1>("TMA gather, blocked [k/8][l][8]", tmb, db, dout, K, L);
  run<2>("LSU gather, strided", tms, ds, dout, K, L);
  run<4>("LSU gather, blocked", tms, db, dout, K, L);
  return 0;
}
