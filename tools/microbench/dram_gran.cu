// Microbenchmark (developer tool, not part of the product path):
//  * DRAM efficiency of strided chunk gathers/scatters as a function of chunk size
//  * FP64 FMA throughput, DSMEM read bandwidth inside an 8-CTA cluster
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_gran dram_gran.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void copy_kernel(const int4* __restrict__ in, int4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    int4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}

// Array = rows x rowbytes. Tile = all rows x chunk bytes. gather: read strided tile, write contiguous.
// chunk16 = chunk size in int4 units; rowlen16 = row length in int4 units.
template <bool GATHER>
__global__ void tile_kernel(const int4* __restrict__ in, int4* __restrict__ out, int rows, int rowlen16, int chunk16) {
  int ntiles = rowlen16 / chunk16;
  size_t per_tile = (size_t)rows * chunk16;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    for (size_t e = threadIdx.x; e < per_tile; e += 4 * blockDim.x) {
      int4 v[4];
      size_t idx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        size_t ee = e + (size_t)u * blockDim.x;
        if (ee < per_tile) {
          size_t r = ee / chunk16, c = ee % chunk16;
          size_t strided = r * rowlen16 + (size_t)t * chunk16 + c;
          size_t contig = (size_t)t * per_tile + ee;
          idx[u] = GATHER ? contig : strided;
          v[u] = in[GATHER ? strided : contig];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        size_t ee = e + (size_t)u * blockDim.x;
        if (ee < per_tile) out[idx[u]] = v[u];
      }
    }
  }
}

__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// DSMEM: each CTA of an 8-cluster reads the smem of CTA (rank+1)%8 repeatedly.
__global__ void __cluster_dims__(8, 1, 1) dsmem_kernel(double* out, int iters, int nwords) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) sm[i] = i;
  cluster.sync();
  unsigned r = cluster.block_rank();
  const double* remote = cluster.map_shared_rank(sm, (r + 1) % 8);
  double acc = 0;
  for (int it = 0; it < iters; ++it)
    for (int i = threadIdx.x * 2; i < nwords; i += blockDim.x * 2) {
      double2 v = *reinterpret_cast<const double2*>(remote + i);
      acc += v.x + v.y;
    }
  cluster.sync();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void smem_local_kernel(double* out, int iters, int nwords) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  for (int it = 0; it < iters; ++it)
    for (int i = threadIdx.x * 2; i < nwords; i += blockDim.x * 2) {
      double2 v = *reinterpret_cast<const double2*>(sm + i);
      acc += v.x + v.y;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> float time_it(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d smem/block optin %zu L2 %d MB\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
  const int rows = 4096, rowbytes = 65536;  // 4096 x 4096 complex f64
  size_t bytes = (size_t)rows * rowbytes;    // 256 MiB
  int4 *in, *out; CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&out, bytes));
  CK(cudaMemset(in, 1, bytes)); CK(cudaMemset(out, 0, bytes));
  size_t n16 = bytes / 16;
  for (int g : {148 * 4, 148 * 8, 148 * 16}) {
    float ms = time_it([&] { copy_kernel<<<g, 512>>>(in, out, n16); });
    printf("copy grid %d: %.3f ms  %.1f GB/s\n", g, ms, 2.0 * bytes / ms * 1e-6);
  }
  for (int chunk : {16, 32, 64, 128, 256, 512, 1024}) {
    for (int threads : {512}) {
      int c16 = chunk / 16; int ntiles = (rowbytes / 16) / c16;
      int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;
      float msg = time_it([&] { tile_kernel<true><<<grid, threads>>>(in, out, rows, rowbytes / 16, c16); });
      float mss = time_it([&] { tile_kernel<false><<<grid, threads>>>(in, out, rows, rowbytes / 16, c16); });
      printf("chunk %4d B: gather %.3f ms %.1f GB/s | scatter %.3f ms %.1f GB/s (grid %d)\n", chunk, msg, 2.0 * bytes / msg * 1e-6, mss, 2.0 * bytes / mss * 1e-6, grid);
    }
  }
  {
    double* o; CK(cudaMalloc(&o, 148 * 8 * 256 * 8));
    int iters = 200000;
    float ms = time_it([&] { dfma_kernel<<<148 * 8, 256>>>(o, iters); }, 3);
    double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
    printf("DFMA: %.3f ms  %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  {
    double* o; CK(cudaMalloc(&o, 148 * 8 * 1024 * 8));
    int nwords = 8192; int iters = 2000; size_t smem = nwords * 8;
    int grid = 144;  // multiple of 8
    float ms = time_it([&] { dsmem_kernel<<<grid, 512, smem>>>(o, iters, nwords); }, 3);
    CK(cudaGetLastError());
    double b = (double)grid * iters * nwords * 8;
    printf("DSMEM remote read: %.3f ms  %.1f GB/s total, %.1f B/clk/SM @1.9GHz\n", ms, b / ms * 1e-6, b / ms * 1e-6 / grid / 1.9);
    float ms2 = time_it([&] { smem_local_kernel<<<grid, 512, smem>>>(o, iters, nwords); }, 3);
    printf("SMEM local read:   %.3f ms  %.1f GB/s total, %.1f B/clk/SM @1.9GHz\n", ms2, b / ms2 * 1e-6, b / ms2 * 1e-6 / grid / 1.9);
  }
  return 0;
}
