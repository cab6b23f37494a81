// Fourth microbenchmark: is a BINNED prefilter worth building?  Models the two new passes of a seed pipeline
// that bins the batch's (strand, offset) seeds by the address range of their seed-context records:
//   scatter : N tuples (16-byte header read in stream order) appended to one of NB bins through one atomicAdd
//             per tuple, 16-byte header + 32-byte payload written at the bin's cursor;
//   filter  : tuples consumed in bin order, every tuple compares CNT contiguous 32-byte records at a random
//             place inside the bin's slice of a 20 GB record array (one lane per record, popcount work as in
//             the real prefilter).  Slice sizes from 16 MB to the whole array (= today's unbinned gather).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench4 gather_bench4.cu
// Usage: gather_bench4 [n_tuples = 2^29] [record_GB = 20]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// headers in stream order: x = record slot (random over the whole array), y = strand|offset, z = count, w = flags
__global__ void make_headers(uint4 *h, uint64_t n, uint64_t n_rec, uint32_t cnt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = mix(i * 0x9e3779b97f4a7c15ULL + 12345) % (n_rec - 64);
    h[i] = make_uint4((uint32_t)r, (uint32_t)(r >> 32), cnt, (uint32_t)i);
  }
}

__global__ void histogram(const uint4 *__restrict__ h, uint64_t n, uint32_t shift, uint32_t *hist) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 x = h[i];
    const uint64_t r = x.x | ((uint64_t)x.y << 32);
    atomicAdd(hist + (r >> shift), 1u);
  }
}

// one atomicAdd per tuple on the bin cursor; 16 + 32 bytes written at the returned position
__global__ void scatter(const uint4 *__restrict__ h, uint64_t n, uint32_t shift, unsigned long long *cursor,
                        uint4 *__restrict__ hb, uint4 *__restrict__ pb) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 x = h[i];
    const uint64_t r = x.x | ((uint64_t)x.y << 32);
    const unsigned long long at = atomicAdd(cursor + (r >> shift), 1ull);
    hb[at] = x;
    const uint32_t s = x.w * 2654435761u;  // payload = "read planes" (computed, not loaded)
    pb[2 * at] = make_uint4(s, s ^ 0x55555555u, s * 3u, s * 5u);
    pb[2 * at + 1] = make_uint4(s * 7u, s * 11u, s * 13u, s * 17u);
  }
}

// HINT: 0 = plain __ldg; 1 = records evict_last, tuple stream evict_first (L2 cache-policy hints)
template <int HINT>
__device__ __forceinline__ uint4 ld_rec(const uint4 *p, uint64_t pol) {
  if (HINT == 0) return __ldg(p);
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}

// warp = 32 tuples; rounds of 32 records: lane l of round r handles record (r*32+l) % cnt of tuple (r*32+l)/cnt
template <int HINT>
__global__ void filter(const uint4 *__restrict__ hb, const uint4 *__restrict__ pb, uint64_t n, const uint4 *__restrict__ rec,
                       uint32_t cnt, unsigned long long *work, unsigned long long *n_pass) {
  uint64_t pol_last = 0, pol_first = 0;
  if (HINT) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  }
  const int lane = threadIdx.x & 31;
  unsigned long long pass = 0;
  for (;;) {
    unsigned long long w0 = 0;
    if (lane == 0) w0 = atomicAdd(work, 32ull);
    w0 = __shfl_sync(0xffffffffu, w0, 0);
    if (w0 >= n) break;
    const uint32_t total = 32u * cnt;
    for (uint32_t c = lane; c < total; c += 32) {
      const uint64_t t = w0 + c / cnt;
      if (t >= n) break;
      const uint4 x = ld_rec<HINT>(hb + t, pol_first);
      const uint64_t r = (x.x | ((uint64_t)x.y << 32)) + c % cnt;
      const uint4 p0 = ld_rec<HINT>(pb + 2 * t, pol_first), p1 = ld_rec<HINT>(pb + 2 * t + 1, pol_first);
      const uint4 a = ld_rec<HINT>(rec + 2 * r, pol_last), b = ld_rec<HINT>(rec + 2 * r + 1, pol_last);
      // four 32-base chunks of bit-sliced compare
      int lb = 0;
      lb += __popc(((a.x & p0.x) | (~a.x & p0.y)) & ((a.y & p0.z) | (~a.y & p0.w)));
      lb += __popc(((a.z & p0.y) | (~a.z & p0.z)) & ((a.w & p0.w) | (~a.w & p0.x)));
      lb += __popc(((b.x & p1.x) | (~b.x & p1.y)) & ((b.y & p1.z) | (~b.y & p1.w)));
      lb += __popc(((b.z & p1.y) | (~b.z & p1.z)) & ((b.w & p1.w) | (~b.w & p1.x)));
      pass += lb > 100;
    }
  }
  if (pass) atomicAdd(n_pass, pass);
}

int main(int argc, char **argv) {
  const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 0) : (1ull << 29);
  const double rec_gb = argc > 2 ? std::atof(argv[2]) : 20.0;
  const uint64_t n_rec = (uint64_t)(rec_gb * 1e9 / 32);
  const uint32_t cnt = 8;
  uint4 *rec, *h, *hb, *pb;
  CK(cudaMalloc(&rec, n_rec * 32));
  CK(cudaMemset(rec, 0x5a, n_rec * 32));
  CK(cudaMalloc(&h, n * 16));
  CK(cudaMalloc(&hb, n * 16));
  CK(cudaMalloc(&pb, n * 32));
  uint32_t *hist;
  unsigned long long *cursor, *work;
  const uint32_t max_bins = 1u << 16;
  CK(cudaMalloc(&hist, max_bins * 4));
  CK(cudaMalloc(&cursor, max_bins * 8));
  CK(cudaMalloc(&work, 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  make_headers<<<sms * 8, 256>>>(h, n, n_rec, cnt);
  CK(cudaDeviceSynchronize());
  std::printf("tuples %llu, records %.1f GB, %u records per tuple (%.2f G record reads), %d SMs\n", (unsigned long long)n, rec_gb, cnt,
              n * (double)cnt / 1e9, sms);
  // slice sizes in records: 2^19 (16 MB) .. whole array
  for (uint32_t shift : {19u, 20u, 21u, 22u, 40u}) {
    const uint32_t n_bins = (uint32_t)((n_rec >> shift) + 1);
    CK(cudaMemset(hist, 0, max_bins * 4));
    float ms_h = 0, ms_s = 0;
    CK(cudaEventRecord(e0));
    histogram<<<sms * 8, 256>>>(h, n, shift, hist);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_h, e0, e1));
    std::vector<uint32_t> hh;
    uint32_t *hp = (uint32_t *)std::malloc(n_bins * 4);
    CK(cudaMemcpy(hp, hist, n_bins * 4, cudaMemcpyDeviceToHost));
    unsigned long long *cp = (unsigned long long *)std::malloc(n_bins * 8);
    unsigned long long acc = 0;
    for (uint32_t b = 0; b < n_bins; ++b) { cp[b] = acc; acc += hp[b]; }
    CK(cudaMemcpy(cursor, cp, n_bins * 8, cudaMemcpyHostToDevice));
    CK(cudaEventRecord(e0));
    scatter<<<sms * 8, 256>>>(h, n, shift, cursor, hb, pb);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_s, e0, e1));
    for (int hint = 0; hint < 2; ++hint) {
      float best = 1e9f;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(work, 0, 16));
        float ms = 0;
        CK(cudaEventRecord(e0));
        if (hint) filter<1><<<sms * 8, 256>>>(hb, pb, n, rec, cnt, work, work + 1);
        else filter<0><<<sms * 8, 256>>>(hb, pb, n, rec, cnt, work, work + 1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
      }
      std::printf("slice %8.1f MB (%5u bins): histogram %6.2f ms, scatter %6.2f ms, filter%s %7.2f ms = %.1f G records/s\n",
                  32.0 * (double)(1ull << (shift > 35 ? 35 : shift)) / 1e6, n_bins, ms_h, ms_s, hint ? " (L2 hints)" : "           ",
                  best, n * (double)cnt / best / 1e6);
    }
    std::free(hp);
    std::free(cp);
  }
  return 0;
}
