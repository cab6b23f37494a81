// Follow-up microbenchmark: is the random-gather ceiling per-SM (outstanding L1 misses) or chip-wide (L2/HBM)?
// 8-byte probes with 16 in flight per thread: vary the number of SMs used, the cache operator and the footprint.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ uint64_t ld_nc(const uint64_t *p) { return __ldg(p); }
__device__ __forceinline__ uint64_t ld_cg(const uint64_t *p) { return __ldcg(p); }
__device__ __forceinline__ uint64_t ld_na(const uint64_t *p) {
  uint64_t v; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}

template <int MODE, int MLP>
__global__ void gather(const uint64_t *__restrict__ a, uint64_t n_words, int iters, uint64_t *out, uint64_t seed) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint64_t v[MLP];
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      const uint64_t p = mix(seed + tid * 1315423911ULL + (uint64_t)(it * MLP + m) * 2654435761ULL) % n_words;
      v[m] = MODE == 0 ? ld_nc(a + p) : MODE == 1 ? ld_cg(a + p) : ld_na(a + p);
    }
#pragma unroll
    for (int m = 0; m < MLP; ++m) acc += v[m];
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int MODE>
void run(const char *name, const uint64_t *a, uint64_t n_words, uint64_t *out, int blocks, int threads) {
  const int iters = 64; constexpr int MLP = 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<MODE, MLP><<<blocks, threads>>>(a, n_words, 4, out, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  gather<MODE, MLP><<<blocks, threads>>>(a, n_words, iters, out, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double acc = (double)blocks * threads * iters * MLP;
  printf("  %-10s blocks %5d x %4d thr  footprint %6.2f GB: %7.2f G probes/s (%.2f ms)\n", name, blocks, threads,
         n_words * 8 / 1e9, acc / ms / 1e6, ms);
}

int main() {
  const uint64_t max_words = (uint64_t)8e9 / 8;
  uint64_t *a, *out;
  cudaMalloc(&a, max_words * 8); cudaMalloc(&out, 8);
  cudaMemset(a, 1, max_words * 8);
  for (double gb : {0.05, 0.5, 2.0, 8.0}) {
    const uint64_t n = (uint64_t)(gb * 1e9) / 8;
    run<0>("ld.nc", a, n, out, 148 * 2, 1024);
  }
  const uint64_t n = (uint64_t)2e9 / 8;
  for (int b : {37, 74, 148, 296}) run<0>("ld.nc", a, n, out, b, 1024);
  for (int t : {128, 256, 512, 1024}) run<0>("ld.nc", a, n, out, 148, t);
  run<1>("ld.cg", a, n, out, 296, 1024);
  run<2>("no_alloc", a, n, out, 296, 1024);
  return 0;
}
