// Random-gather microbenchmark: what the B200 memory system sustains for the access shapes of the
// mapper (8-byte counter probes, 24/40/48-byte packed-genome windows at 8-byte alignment), and whether
// cudaLimitMaxL2FetchGranularity changes it.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

template <int WORDS, int MLP>
__global__ void gather(const uint64_t *__restrict__ a, uint64_t n_words, int iters, uint64_t *out, uint64_t seed) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint64_t v[MLP][WORDS];
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      const uint64_t p = mix(seed + tid * 1315423911ULL + (uint64_t)(it * MLP + m) * 2654435761ULL) % (n_words - WORDS);
#pragma unroll
      for (int w = 0; w < WORDS; ++w) v[m][w] = __ldg(a + p + w);
    }
#pragma unroll
    for (int m = 0; m < MLP; ++m)
#pragma unroll
      for (int w = 0; w < WORDS; ++w) acc += v[m][w];
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int WORDS, int MLP>
void run(const char *name, const uint64_t *a, uint64_t n_words, uint64_t *out) {
  const int threads = 256, blocks = 148 * 8, iters = 64;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<WORDS, MLP><<<blocks, threads>>>(a, n_words, 4, out, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  gather<WORDS, MLP><<<blocks, threads>>>(a, n_words, iters, out, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double acc = (double)blocks * threads * iters * MLP;
  printf("  %-28s %8.2f G gathers/s  useful %7.1f GB/s  (%.2f ms)\n", name, acc / ms / 1e6, acc * WORDS * 8 / ms / 1e6, ms);
}

int main(int argc, char **argv) {
  const uint64_t n_words = (uint64_t)(argc > 1 ? atof(argv[1]) : 2e9) / 8;
  uint64_t *a, *out;
  cudaMalloc(&a, n_words * 8); cudaMalloc(&out, 8);
  cudaMemset(a, 1, n_words * 8);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("default cudaLimitMaxL2FetchGranularity = %zu, array %.2f GB\n", lim, n_words * 8 / 1e9);
  for (int g : {0, 32, 64, 128}) {
    if (g) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
      cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
      printf("set %d -> %s, now %zu\n", g, cudaGetErrorString(e), lim);
    }
    run<1, 8>("8B probe, 8 in flight", a, n_words, out);
    run<1, 16>("8B probe, 16 in flight", a, n_words, out);
    run<3, 4>("24B window, 4 in flight", a, n_words, out);
    run<5, 4>("40B window, 4 in flight", a, n_words, out);
    run<6, 4>("48B window, 4 in flight", a, n_words, out);
    run<8, 2>("64B window, 2 in flight", a, n_words, out);
  }
  return 0;
}
