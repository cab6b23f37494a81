// Third random-gather microbenchmark: which unit does the B200 memory system charge for a random access --
// the 32-byte sector, the 64-byte DRAM burst pair or the 128-byte line?  Aligned blocks of 32/64/128 bytes
// are read at random addresses (16-byte vector loads, all issued before any use) over a 3 GB array, next to
// the mapper's actual shapes (8-byte probe, 40-byte window at 8-byte alignment).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench3 gather_bench3.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// BYTES per access, aligned to ALIGN bytes; MLP accesses in flight per thread
template <int BYTES, int ALIGN, int MLP>
__global__ void gather(const unsigned char *__restrict__ a, uint64_t n_bytes, int iters, uint64_t *out, uint64_t seed) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n_slots = (n_bytes - 256) / ALIGN;
  uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (BYTES >= 16) {
      uint4 v[MLP][BYTES / 16 > 0 ? BYTES / 16 : 1];
#pragma unroll
      for (int m = 0; m < MLP; ++m) {
        const uint64_t p = (mix(seed + tid * 1315423911ULL + (uint64_t)(it * MLP + m) * 2654435761ULL) % n_slots) * ALIGN;
#pragma unroll
        for (int w = 0; w < BYTES / 16; ++w) v[m][w] = __ldg(reinterpret_cast<const uint4 *>(a + p) + w);
      }
#pragma unroll
      for (int m = 0; m < MLP; ++m)
#pragma unroll
        for (int w = 0; w < BYTES / 16; ++w) acc += v[m][w].x + v[m][w].y + v[m][w].z + v[m][w].w;
    }
    else {
      uint64_t v[MLP][BYTES / 8 > 0 ? BYTES / 8 : 1];
#pragma unroll
      for (int m = 0; m < MLP; ++m) {
        const uint64_t p = (mix(seed + tid * 1315423911ULL + (uint64_t)(it * MLP + m) * 2654435761ULL) % n_slots) * ALIGN;
#pragma unroll
        for (int w = 0; w < BYTES / 8; ++w) v[m][w] = __ldg(reinterpret_cast<const uint64_t *>(a + p) + w);
      }
#pragma unroll
      for (int m = 0; m < MLP; ++m)
#pragma unroll
        for (int w = 0; w < BYTES / 8; ++w) acc += v[m][w];
    }
  }
  if (acc == 0x1234567) out[0] = acc;
}

// 40-byte window at 8-byte alignment (the 2-bit genome compare), as five 8-byte loads
template <int MLP>
__global__ void window40(const uint64_t *__restrict__ a, uint64_t n_words, int iters, uint64_t *out, uint64_t seed) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint64_t v[MLP][5];
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      const uint64_t p = mix(seed + tid * 1315423911ULL + (uint64_t)(it * MLP + m) * 2654435761ULL) % (n_words - 8);
#pragma unroll
      for (int w = 0; w < 5; ++w) v[m][w] = __ldg(a + p + w);
    }
#pragma unroll
    for (int m = 0; m < MLP; ++m)
#pragma unroll
      for (int w = 0; w < 5; ++w) acc += v[m][w];
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <class K>
void timeit(const char *name, K launch, double accesses, double bytes_each) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(4, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch(64, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("  %-40s %8.2f G accesses/s  useful %7.1f GB/s  (%.2f ms)\n", name, accesses / ms / 1e6, accesses * bytes_each / ms / 1e6, ms);
}

int main(int argc, char **argv) {
  const uint64_t n_bytes = (uint64_t)(argc > 1 ? atof(argv[1]) : 3e9);
  unsigned char *a; uint64_t *out;
  cudaMalloc(&a, n_bytes); cudaMalloc(&out, 8);
  cudaMemset(a, 1, n_bytes);
  const int threads = 256, blocks = 148 * 8;
  const double per_iter = (double)blocks * threads * 64;
  printf("array %.2f GB, %d x %d threads\n", n_bytes / 1e9, blocks, threads);
#define RUN(B, A, M) timeit(#B "B aligned " #A ", " #M " in flight", [&](int it, uint64_t s) { gather<B, A, M><<<blocks, threads>>>(a, n_bytes, it, out, s); }, per_iter * M, B)
  RUN(8, 8, 8);
  RUN(8, 8, 16);
  RUN(16, 16, 8);
  RUN(32, 32, 8);
  RUN(32, 32, 16);
  RUN(64, 64, 4);
  RUN(64, 64, 8);
  RUN(128, 128, 2);
  RUN(128, 128, 4);
  RUN(32, 8, 8);
  RUN(64, 32, 4);
  RUN(160, 32, 2);  // a bucket of five contiguous seed-context records
  RUN(160, 32, 4);
  timeit("40B window aligned 8, 4 in flight", [&](int it, uint64_t s) { window40<4><<<blocks, threads>>>((const uint64_t *)a, n_bytes / 8, it, out, s); }, per_iter * 4, 40);
  timeit("40B window aligned 8, 8 in flight", [&](int it, uint64_t s) { window40<8><<<blocks, threads>>>((const uint64_t *)a, n_bytes / 8, it, out, s); }, per_iter * 8, 40);
  return 0;
}
