// Microbenchmark of the register-resident radix-16 line FFT (csrc/nb_fft16.cuh) inside the staging skeleton the
// pass kernels use: persistent CTAs, bulk-copy (cp.async.bulk + mbarrier) of the next tile into a staging
// buffer while the current tile is transformed, direct stores from registers.
//   t-fast: out[line][k]            (natural store)
//   r-fast: out[k][line] transposed (the staged transpose)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I nifty_b200/csrc tools/microbench/fft16_bench.cu -o tools/microbench/fft16_bench
// Run on a B200: tools/microbench/fft16_bench   (prints GB/s per variant and the error against a host DFT)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <complex>
#include <vector>
#include "nb_fft16.cuh"

using namespace nb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <class T, int LG, bool RF>
__global__ void __launch_bounds__(256, 2) k_fft(const cplx<T>* in, cplx<T>* out, const cplx<T>* tw, int ntiles, long nlines) {
  typedef Fft16<T, LG, RF> F;
  typedef F16Geom<LG> G;
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int PITCH = G::template stage_pitch<T>(RF);
  cplx<T>* stage = reinterpret_cast<cplx<T>*>(sm);
  typename F::word_t* xb = reinterpret_cast<typename F::word_t*>(sm + (size_t)G::LPC * PITCH * sizeof(cplx<T>));
  Mbar* bar = reinterpret_cast<Mbar*>(sm + (size_t)G::LPC * PITCH * sizeof(cplx<T>) + F::XBYTES);
  Ctx ctx{(int)threadIdx.x, 256, (int)blockIdx.x, (int)gridDim.x};
  struct TS { cplx<T> a[16]; typename F::Th th; };
  Team<TS> tm(ctx);
  F::init(tm.ts.th, ctx.tid, tw, 1);
  if (ctx.tid == 0) mbar_init(bar, 1);
  __syncthreads();
  auto issue = [&](int tile) {
    if (ctx.tid == 0) mbar_expect(bar, (unsigned)(F16_TILE * sizeof(cplx<T>)));
    __syncwarp();
    if (ctx.tid < G::LPC) bulk_g2s(stage + ctx.tid * PITCH, in + ((long)tile * G::LPC + ctx.tid) * G::N, G::N * sizeof(cplx<T>), bar);
  };
  int tile = ctx.bid;
  unsigned phase = 0;
  if (tile < ntiles && ctx.tid < 32) issue(tile);
  for (; tile < ntiles; tile += ctx.nblk) {
    mbar_wait(bar, phase); phase ^= 1;
    const TS& S0 = tm.ts;
#pragma unroll
    for (int s = 0; s < 16; ++s) tm.ts.a[s] = stage[S0.th.r * PITCH + F::elem(S0.th, s)];
    __syncthreads();
    if (tile + ctx.nblk < ntiles && ctx.tid < 32) { fence_async_smem(); issue(tile + ctx.nblk); }
    F::run(tm, xb);
    const long line = (long)tile * G::LPC + tm.ts.th.r;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int k = F::elem(tm.ts.th, s);
      if (RF) out[(long)k * nlines + line] = tm.ts.a[s];
      else out[line * G::N + k] = tm.ts.a[s];
    }
    __syncthreads();
  }
}

template <class T, int LG, bool RF> void run(const char* name, long total_elems) {
  typedef Fft16<T, LG, RF> F;
  typedef F16Geom<LG> G;
  const long nlines = total_elems >> LG;
  const int ntiles = (int)(nlines / G::LPC);
  std::vector<cplx<T>> h(total_elems), tw(G::N);
  srand(1);
  for (auto& v : h) v = cmake<T>((T)(rand() / (double)RAND_MAX - 0.5), (T)(rand() / (double)RAND_MAX - 0.5));
  for (int j = 0; j < G::N; ++j) { long double a = -6.283185307179586476925286766559L * j / G::N; tw[j] = cmake<T>((T)cosl(a), (T)sinl(a)); }
  cplx<T>*din, *dout, *dtw;
  CK(cudaMalloc(&din, total_elems * sizeof(cplx<T>))); CK(cudaMalloc(&dout, total_elems * sizeof(cplx<T>))); CK(cudaMalloc(&dtw, G::N * sizeof(cplx<T>)));
  CK(cudaMemcpy(din, h.data(), total_elems * sizeof(cplx<T>), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dtw, tw.data(), G::N * sizeof(cplx<T>), cudaMemcpyHostToDevice));
  constexpr int PITCH = G::template stage_pitch<T>(RF);
  const size_t smem = (size_t)G::LPC * PITCH * sizeof(cplx<T>) + F::XBYTES + 64;
  CK(cudaFuncSetAttribute(k_fft<T, LG, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fft<T, LG, RF>, 256, smem));
  const int grid = std::min(ntiles, 148 * occ);
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) k_fft<T, LG, RF><<<grid, 256, smem>>>(din, dout, dtw, ntiles, nlines);
  CK(cudaDeviceSynchronize());
  const int reps = 20;
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) k_fft<T, LG, RF><<<grid, 256, smem>>>(din, dout, dtw, ntiles, nlines);
  CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
  std::vector<cplx<T>> o(total_elems);
  CK(cudaMemcpy(o.data(), dout, total_elems * sizeof(cplx<T>), cudaMemcpyDeviceToHost));
  double err = 0, nrm = 0;
  for (long line : {0L, nlines / 3, nlines - 1}) {
    for (int k = 0; k < G::N; k += 37) {
      std::complex<long double> s = 0;
      for (int j = 0; j < G::N; ++j) { long double ang = -6.283185307179586476925286766559L * ((long)j * k % G::N) / G::N; s += std::complex<long double>(h[line * G::N + j].x, h[line * G::N + j].y) * std::complex<long double>(cosl(ang), sinl(ang)); }
      cplx<T> g = RF ? o[(long)k * nlines + line] : o[line * G::N + k];
      err = std::max(err, (double)std::abs(std::complex<long double>(g.x, g.y) - s)); nrm = std::max(nrm, (double)std::abs(s));
    }
  }
  const double gb = 2.0 * total_elems * sizeof(cplx<T>) / 1e9;
  printf("%-28s n=%5d occ=%d grid=%4d smem=%6zu  %8.3f ms  %7.1f GB/s  rel err %.1e\n", name, G::N, occ, grid, smem, ms, gb / (ms * 1e-3), err / nrm);
  cudaFree(din); cudaFree(dout); cudaFree(dtw);
}

int main() {
  const long E = 1L << 25;    // 32 Mi complex elements: 512 MiB in + 512 MiB out (f64)
  run<double, 12, false>("f64 4096 t-fast", E);
  run<double, 11, false>("f64 2048 t-fast", E);
  run<double, 11, true>("f64 2048 r-fast", E);
  run<double, 10, true>("f64 1024 r-fast", E);
  run<double, 8, false>("f64 256 t-fast", E);
  run<double, 8, true>("f64 256 r-fast", E);
  run<double, 7, true>("f64 128 r-fast", E);
  run<float, 12, false>("f32 4096 t-fast", E);
  run<float, 8, true>("f32 256 r-fast", E);
  return 0;
}
