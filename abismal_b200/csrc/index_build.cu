// index_build.cu -- AbismalIndex construction on the GPU.
//
// Reproduces, array for array, what the reference's AbismalIndex::create_index
// computes (src/AbismalIndex.cpp:281-331): bucket sizes (:333-410), two- vs
// three-letter selection (:471-543), the windowed DP that picks which
// positions to keep (:726-855), masked bucket sizes, bucket fill (:545-641) and
// the per-bucket sort by the following bases (:857-978).  The FASTA side
// (padding, N-run handling, 4-bit encoding) stays on the host.
//
// Decomposition: position-parallel kernels for everything that is a pure
// function of the genome (hashing, counting with atomics, selection, fill);
// the DP is sequential inside a block of <= 1M positions exactly as in the
// reference, so it runs one THREAD per block (thousands of independent blocks
// on a 3.1 Gbp genome); buckets are sorted one thread (small) or one CTA
// (large) per bucket with the reference's comparator, ties broken by
// descending position (the reference fills buckets in descending order and
// uses a stable sort).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "abismal_b200.h"
#include "abismal_b200_index.h"

namespace {

thread_local std::string g_ierr;

constexpr uint32_t kKeyW = 25, kKeyW3 = 16, kSortPos = 256;  // seed::window_size (20, or 12) is a run-time argument
constexpr uint32_t kMask25 = (1u << 25) - 1u;
constexpr uint32_t kPow3 = 43046721u;
constexpr uint64_t kBlockSize = 1000000ull;
constexpr int kRun = 64;  // consecutive positions hashed by one thread

__device__ __forceinline__ uint32_t nib(const uint64_t *g, uint64_t pos) {
  return (uint32_t)(g[pos >> 4] >> ((pos & 15u) << 2)) & 15u;
}
__device__ __forceinline__ uint32_t bit2(uint32_t nt) { return (nt & 5u) == 0u; }
__device__ __forceinline__ uint32_t num_t(uint32_t nt) { return (((nt & 4u) != 0u) << 1) | ((nt & 1u) != 0u); }
__device__ __forceinline__ uint32_t num_a(uint32_t nt) { return (((nt & 8u) != 0u) << 1) | ((nt & 2u) != 0u); }

// Counting / fill passes of the reference walk the excluded intervals with
//   if (i < nidx->first && ...) count;  if (nidx->second <= i) ++nidx;
// (src/AbismalIndex.cpp:355-364, :585-594): the interval pointer advances only
// AFTER position `second` was tested against the old interval, so position
// `second` itself is skipped too.  i is skipped iff first <= i <= second.
__device__ __forceinline__ bool skipped_by_walk(const uint64_t *ex, uint32_t n_ex, uint64_t i) {
  uint32_t lo = 0, hi = n_ex;  // first interval with second >= i
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (ex[2 * mid + 1] < i) lo = mid + 1;
    else hi = mid;
  }
  return lo < n_ex && ex[2 * lo] <= i;
}
// block-based passes (selection, DP) cover exactly the positions outside [first, second)
__device__ __forceinline__ bool excluded(const uint64_t *ex, uint32_t n_ex, uint64_t i) {
  uint32_t lo = 0, hi = n_ex;  // first interval with second > i
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (ex[2 * mid + 1] <= i) lo = mid + 1;
    else hi = mid;
  }
  return lo < n_ex && ex[2 * lo] <= i;
}

struct Hashes {
  uint32_t two, t, a;
};

// true 25-mer / 16-mer hashes of position p (get_1bit_hash / get_base_3_hash)
__device__ __forceinline__ Hashes hash_at(const uint64_t *g, uint64_t p) {
  Hashes h{0, 0, 0};
  for (uint32_t j = 0; j < kKeyW; ++j) h.two = (h.two << 1) | bit2(nib(g, p + j));
  for (uint32_t j = 0; j < kKeyW3; ++j) {
    const uint32_t x = nib(g, p + j);
    h.t = h.t * 3u + num_t(x);
    h.a = h.a * 3u + num_a(x);
  }
  // IUPAC nibbles give "digits" of 3: the reference's rolling key is the window sum mod 3^16
  h.t %= kPow3;
  h.a %= kPow3;
  return h;
}

// mode 0: count all hashable positions (initialize_bucket_sizes<false>)
// mode 1: is_two_letter selection (select_two_letter_positions)
// mode 2: masked counts (initialize_bucket_sizes<true>)
// mode 3: fill buckets (hash_genome), cursors = running offsets
__global__ void pos_pass_kernel(int mode, const uint64_t *g, uint64_t lim2, uint64_t lim3, const uint64_t *ex,
                                uint32_t n_ex, uint32_t *c2, uint32_t *ct, uint32_t *ca, uint8_t *is_two,
                                const uint8_t *keep, uint32_t *idx2, uint32_t *idxt, uint32_t *idxa) {
  const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kRun;
  if (start >= lim3) return;
  const uint64_t end = min(start + (uint64_t)kRun, lim3);
  Hashes h = hash_at(g, start);
  for (uint64_t i = start; i < end; ++i) {
    if (i != start) {
      h.two = ((h.two << 1) | bit2(nib(g, i + kKeyW - 1))) & kMask25;
      const uint32_t x = nib(g, i + kKeyW3 - 1);
      h.t = (h.t * 3u + num_t(x)) % kPow3;
      h.a = (h.a * 3u + num_a(x)) % kPow3;
    }
    if (mode == 1 ? excluded(ex, n_ex, i) : skipped_by_walk(ex, n_ex, i)) continue;
    const bool in2 = i < lim2;
    if (mode == 0) {
      if (in2) atomicAdd(c2 + h.two, 1u);
      atomicAdd(ct + h.t, 1u);
      atomicAdd(ca + h.a, 1u);
    }
    else if (mode == 1) {
      if (in2) is_two[i] = c2[h.two] <= ((ct[h.t] + ca[h.a]) >> 1);
    }
    else if (mode == 2) {
      if (keep[i]) {
        if (is_two[i]) {
          if (in2) atomicAdd(c2 + h.two, 1u);
        }
        else {
          atomicAdd(ct + h.t, 1u);
          atomicAdd(ca + h.a, 1u);
        }
      }
    }
    else {
      if (in2 && keep[i]) {
        if (is_two[i]) idx2[atomicAdd(c2 + h.two, 1u)] = (uint32_t)i;
        else {
          idxt[atomicAdd(ct + h.t, 1u)] = (uint32_t)i;
          idxa[atomicAdd(ca + h.a, 1u)] = (uint32_t)i;
        }
      }
    }
  }
}

// compress_dp (src/AbismalIndex.cpp:726-855): one thread per block.
__global__ void dp_kernel(const uint64_t *g, const uint64_t *blocks, uint32_t n_blocks, const uint32_t *c2,
                          const uint32_t *ct, const uint32_t *ca, const uint8_t *is_two, uint32_t *prev_arr,
                          uint8_t *keep, uint32_t kWindow) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const uint64_t block_start = blocks[2 * b];
  const uint32_t size = (uint32_t)(blocks[2 * b + 1] - block_start);
  if (size < kWindow) return;
  uint32_t *prev = prev_arr + block_start;  // prev[i] for i in [0, size]; one slot beyond is spare

  // spool the hashes exactly as the reference does (short blocks spool fewer bases)
  uint32_t h2 = 0, ht = 0, ha = 0;
  uint64_t p2 = block_start, p3 = block_start;
  const uint32_t spool2 = min(size, kKeyW - 1);
  for (uint32_t j = 0; j < spool2; ++j) h2 = ((h2 << 1) | bit2(nib(g, p2++))) & kMask25;
  for (uint32_t j = 0; j < kKeyW3 - 1; ++j) {
    const uint32_t x = nib(g, p3++);
    ht = (ht * 3u + num_t(x)) % kPow3;
    ha = (ha * 3u + num_a(x)) % kPow3;
  }
  // monotone deque over the last `window` solutions (fixed_ring_buffer, qsz = 32)
  unsigned long long qcost[32];
  uint32_t qprev[32];
  uint32_t f = 0, bk = 0;
  unsigned long long cost_i = 0;
  for (uint32_t i = 0; i < size; ++i) {
    h2 = ((h2 << 1) | bit2(nib(g, p2++))) & kMask25;
    const uint32_t x = nib(g, p3++);
    ht = (ht * 3u + num_t(x)) % kPow3;
    ha = (ha * 3u + num_a(x)) % kPow3;
    const unsigned long long c =
      is_two[block_start + i] ? (unsigned long long)c2[h2] : (unsigned long long)((ct[ht] + ca[ha]) >> 1);
    if (i < kWindow) {
      cost_i = c;
      prev[i] = 0xffffffffu;
    }
    else {
      cost_i = qcost[f] + c;
      prev[i] = qprev[f];
    }
    // add_sol(helper, i, cost_i)
    while (f != bk && qcost[(bk - 1) & 31u] > cost_i) bk = (bk - 1) & 31u;
    qcost[bk] = cost_i;
    qprev[bk] = i;
    bk = (bk + 1) & 31u;
    while (qprev[f] + kWindow <= i) f = (f + 1) & 31u;
  }
  // start of the traceback: the best of the last `window` solutions.  The deque
  // holds exactly those (indices > size-1-window), in increasing index order
  // with non-decreasing cost; the reference scans from the end with a strict
  // '<', i.e. it takes the LARGEST index among the minima.
  unsigned long long best = ~0ull;
  uint32_t last = 0xffffffffu;
  // costs of the last window are not all in the deque (dominated ones were
  // popped), but a popped entry is never a minimum with a larger index than a
  // surviving one of equal cost, except equal-cost entries which are kept; so
  // the answer is the last deque entry whose cost equals the front's cost.
  {
    const unsigned long long mn = qcost[f];
    uint32_t k = f;
    while (k != bk) {
      if (qcost[k] == mn) {
        best = mn;
        last = qprev[k];
      }
      k = (k + 1) & 31u;
    }
  }
  (void)best;
  uint32_t cur = last;
  while (cur != 0xffffffffu) {
    keep[block_start + cur] = 1;
    cur = prev[cur];
  }
}

// BucketLess / BucketLessThree (src/AbismalIndex.cpp:857-903) with the tie
// rule implied by descending fill + stable sort: larger position first.
template <int KIND>  // 0 two-letter, 1 three-letter c_to_t, 2 three-letter g_to_a
__device__ __forceinline__ bool bucket_less(const uint64_t *g, uint32_t a, uint32_t b) {
  const uint32_t first = KIND == 0 ? kKeyW : kKeyW3;
  for (uint32_t j = first; j < kSortPos; ++j) {
    const uint32_t xa = nib(g, (uint64_t)a + j), xb = nib(g, (uint64_t)b + j);
    const uint32_t ca = KIND == 0 ? bit2(xa) : (KIND == 1 ? (xa & 5u) : (xa & 10u));
    const uint32_t cb = KIND == 0 ? bit2(xb) : (KIND == 1 ? (xb & 5u) : (xb & 10u));
    if (ca != cb) return ca < cb;
  }
  return a > b;
}

constexpr uint32_t kSmallBucket = 48;

template <int KIND>
__global__ void sort_small_kernel(const uint64_t *g, const uint32_t *counter, uint64_t n_buckets, uint32_t *idx,
                                  uint32_t *big_list, uint32_t *n_big) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_buckets) return;
  const uint32_t s = counter[k], e = counter[k + 1];
  const uint32_t n = e - s;
  if (n < 2) return;
  if (n > kSmallBucket) {
    big_list[atomicAdd(n_big, 1u)] = (uint32_t)k;
    return;
  }
  uint32_t *v = idx + s;
  for (uint32_t i = 1; i < n; ++i) {  // insertion sort
    const uint32_t x = v[i];
    uint32_t j = i;
    while (j > 0 && bucket_less<KIND>(g, x, v[j - 1])) {
      v[j] = v[j - 1];
      --j;
    }
    v[j] = x;
  }
}

// one CTA per large bucket: normalized bitonic sort (all comparisons ascending,
// virtual +inf padding past the end)
template <int KIND>
__global__ void sort_big_kernel(const uint64_t *g, const uint32_t *counter, const uint32_t *big_list, uint32_t *idx) {
  const uint32_t k = big_list[blockIdx.x];
  const uint32_t s = counter[k], n = counter[k + 1] - s;
  uint32_t *v = idx + s;
  uint32_t n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    // first stage of the merge: partner is the mirror position inside the block
    for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
      const uint32_t blk = t / (size / 2), off = t % (size / 2);
      const uint32_t i = blk * size + off, l = blk * size + size - 1 - off;
      if (l < n) {
        const uint32_t a = v[i], b = v[l];
        if (bucket_less<KIND>(g, b, a)) {
          v[i] = b;
          v[l] = a;
        }
      }
    }
    __syncthreads();
    for (uint32_t stride = size / 4; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
        const uint32_t i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const uint32_t l = i + stride;
        if (l < n) {
          const uint32_t a = v[i], b = v[l];
          if (bucket_less<KIND>(g, b, a)) {
            v[i] = b;
            v[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// exclusive scan of n+1 counters (counts in [0,n), slot n ignored) -> start offsets, total in [n]
__global__ void scan_block_sums(const uint32_t *in, uint64_t n, uint32_t *block_sums) {
  __shared__ uint32_t sh[256];
  const uint64_t per = 4096;
  const uint64_t base = (uint64_t)blockIdx.x * per;
  uint32_t s = 0;
  for (uint64_t i = base + threadIdx.x; i < min(base + per, n); i += blockDim.x) s += in[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = sh[0];
}
__global__ void scan_apply(uint32_t *data, uint64_t n, const uint32_t *block_offsets) {
  // one warp-serial pass per 4096-chunk: thread 0 of each block walks its chunk
  // in 16 sub-chunks handled by 16 threads with a small shared scan
  __shared__ uint32_t sub[16];
  const uint64_t per = 4096, base = (uint64_t)blockIdx.x * per;
  const uint32_t t = threadIdx.x;  // 16 threads, 256 elements each
  const uint64_t lo = base + (uint64_t)t * 256, hi = min(lo + 256, min(base + per, n));
  uint32_t s = 0;
  for (uint64_t i = lo; i < hi; ++i) s += data[i];
  sub[t] = s;
  __syncthreads();
  uint32_t off = block_offsets[blockIdx.x];
  for (uint32_t k = 0; k < t; ++k) off += sub[k];
  for (uint64_t i = lo; i < hi; ++i) {
    const uint32_t x = data[i];
    data[i] = off;
    off += x;
  }
}

int ifail(int code, const std::string &m) {
  g_ierr = m;
  return code;
}

#define IB_CUDA(call)                                                                   \
  do {                                                                                  \
    const cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                            \
      free_all();                                                                       \
      return ifail(ABG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));   \
    }                                                                                   \
  } while (0)

// counts -> start offsets (exclusive scan); returns total via *total
int exclusive_scan(uint32_t *d, uint64_t n, uint32_t *total) {
  const uint64_t per = 4096;
  const uint32_t nb = (uint32_t)((n + per - 1) / per);
  uint32_t *d_sums = nullptr;
  if (cudaMalloc(&d_sums, (size_t)nb * 4) != cudaSuccess) return ABG_ERR_CUDA;
  scan_block_sums<<<nb, 256>>>(d, n, d_sums);
  std::vector<uint32_t> sums(nb);
  if (cudaMemcpy(sums.data(), d_sums, (size_t)nb * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaFree(d_sums);
    return ABG_ERR_CUDA;
  }
  uint32_t run = 0;
  for (uint32_t i = 0; i < nb; ++i) {
    const uint32_t x = sums[i];
    sums[i] = run;
    run += x;
  }
  cudaMemcpy(d_sums, sums.data(), (size_t)nb * 4, cudaMemcpyHostToDevice);
  scan_apply<<<nb, 16>>>(d, n, d_sums);
  cudaMemcpy(d + n, &run, 4, cudaMemcpyHostToDevice);
  cudaFree(d_sums);
  *total = run;
  return cudaGetLastError() == cudaSuccess ? ABG_OK : ABG_ERR_CUDA;
}

// get_block_bounds (src/AbismalIndex.cpp:438-469)
std::vector<uint64_t> block_bounds(uint64_t start_pos, uint64_t step, uint64_t end_pos, const uint64_t *ex,
                                   uint32_t n_ex) {
  std::vector<uint64_t> blocks;
  uint64_t block_start = start_pos;
  uint32_t i = 0;
  while (block_start < end_pos && i != n_ex) {
    if (block_start < ex[2 * i]) {
      const uint64_t block_end = std::min({ex[2 * i], block_start + step, end_pos});
      blocks.push_back(block_start);
      blocks.push_back(block_end);
      block_start += step;
      if (block_start >= ex[2 * i + 1]) block_start = ex[2 * i++ + 1];
    }
    else block_start = ex[2 * i++ + 1];
  }
  while (block_start < end_pos) {
    const uint64_t block_end = std::min(block_start + step, end_pos);
    blocks.push_back(block_start);
    blocks.push_back(block_end);
    block_start += step;
  }
  return blocks;
}

template <int KIND>
int sort_buckets(const uint64_t *d_g, const uint32_t *d_counter, uint64_t n_buckets, uint32_t *d_idx) {
  uint32_t *d_big = nullptr, *d_nbig = nullptr;
  const uint32_t max_big = 1u << 22;
  if (cudaMalloc(&d_big, (size_t)max_big * 4) != cudaSuccess || cudaMalloc(&d_nbig, 4) != cudaSuccess) return ABG_ERR_CUDA;
  cudaMemset(d_nbig, 0, 4);
  const uint32_t grid = (uint32_t)((n_buckets + 255) / 256);
  sort_small_kernel<KIND><<<grid, 256>>>(d_g, d_counter, n_buckets, d_idx, d_big, d_nbig);
  uint32_t n_big = 0;
  cudaMemcpy(&n_big, d_nbig, 4, cudaMemcpyDeviceToHost);
  int rc = ABG_OK;
  if (n_big > max_big) rc = ABG_ERR_INVALID;
  else if (n_big > 0) sort_big_kernel<KIND><<<n_big, 256>>>(d_g, d_counter, d_big, d_idx);
  if (cudaDeviceSynchronize() != cudaSuccess) rc = ABG_ERR_CUDA;
  cudaFree(d_big);
  cudaFree(d_nbig);
  return rc;
}

}  // namespace

extern "C" {

const char *abg_index_build_last_error(void) { return g_ierr.c_str(); }

void abg_built_index_free(abg_built_index *b) {
  if (!b) return;
  std::free(b->counter);
  std::free(b->counter_t);
  std::free(b->counter_a);
  std::free(b->index);
  std::free(b->index_t);
  std::free(b->index_a);
  std::memset(b, 0, sizeof *b);
}

int abg_build_index(const uint64_t *genome, uint64_t genome_size, const uint64_t *exclude, uint32_t n_exclude,
                    int device, abg_built_index *out) {
  return abg_build_index_w(genome, genome_size, exclude, n_exclude, 20u, device, out);
}

int abg_build_index_w(const uint64_t *genome, uint64_t genome_size, const uint64_t *exclude, uint32_t n_exclude,
                      uint32_t window_size, int device, abg_built_index *out) {
  if (!genome || !exclude || !out || n_exclude == 0 || genome_size < 2 * 32767ull)
    return ifail(ABG_ERR_INVALID, "abg_build_index: bad argument");
  if (window_size != 12u && window_size != 20u)
    return ifail(ABG_ERR_INVALID, "abg_build_index: window_size must be 20, or 12 (--enable-short)");
  if (genome_size >= (1ull << 32)) return ifail(ABG_ERR_INVALID, "abg_build_index: genome must be < 2^32 bases");
  std::memset(out, 0, sizeof *out);
  if (cudaSetDevice(device) != cudaSuccess) return ifail(ABG_ERR_CUDA, "abg_build_index: cannot select device");

  const uint64_t n_words = (genome_size + 15) / 16;
  const uint64_t lim2 = genome_size - kKeyW + 1, lim3 = genome_size - kKeyW3 + 1;
  const uint64_t n2 = 1ull << 25, n3 = kPow3;

  uint64_t *d_g = nullptr, *d_ex = nullptr, *d_blocks = nullptr;
  uint32_t *d_c2 = nullptr, *d_ct = nullptr, *d_ca = nullptr, *d_prev = nullptr;
  uint32_t *d_i2 = nullptr, *d_it = nullptr, *d_ia = nullptr;
  uint32_t *d_s2 = nullptr, *d_st = nullptr, *d_sa = nullptr;  // start offsets (final counters)
  uint8_t *d_is_two = nullptr, *d_keep = nullptr;
  const auto free_all = [&]() {
    cudaFree(d_g); cudaFree(d_ex); cudaFree(d_blocks); cudaFree(d_c2); cudaFree(d_ct); cudaFree(d_ca);
    cudaFree(d_prev); cudaFree(d_i2); cudaFree(d_it); cudaFree(d_ia); cudaFree(d_s2); cudaFree(d_st);
    cudaFree(d_sa); cudaFree(d_is_two); cudaFree(d_keep);
  };

  IB_CUDA(cudaMalloc(&d_g, (n_words + 32) * 8));
  IB_CUDA(cudaMemset(d_g, 0, (n_words + 32) * 8));
  IB_CUDA(cudaMemcpy(d_g, genome, n_words * 8, cudaMemcpyHostToDevice));
  IB_CUDA(cudaMalloc(&d_ex, (size_t)n_exclude * 16));
  IB_CUDA(cudaMemcpy(d_ex, exclude, (size_t)n_exclude * 16, cudaMemcpyHostToDevice));
  IB_CUDA(cudaMalloc(&d_c2, (n2 + 1) * 4));
  IB_CUDA(cudaMalloc(&d_ct, (n3 + 1) * 4));
  IB_CUDA(cudaMalloc(&d_ca, (n3 + 1) * 4));
  IB_CUDA(cudaMemset(d_c2, 0, (n2 + 1) * 4));
  IB_CUDA(cudaMemset(d_ct, 0, (n3 + 1) * 4));
  IB_CUDA(cudaMemset(d_ca, 0, (n3 + 1) * 4));
  IB_CUDA(cudaMalloc(&d_is_two, genome_size));
  IB_CUDA(cudaMemset(d_is_two, 0, genome_size));
  IB_CUDA(cudaMalloc(&d_keep, genome_size));
  IB_CUDA(cudaMemset(d_keep, 0, genome_size));

  const uint64_t n_threads = (lim3 + kRun - 1) / kRun;
  const uint32_t grid = (uint32_t)((n_threads + 255) / 256);

  // 1. bucket sizes over all hashable positions
  pos_pass_kernel<<<grid, 256>>>(0, d_g, lim2, lim3, d_ex, n_exclude, d_c2, d_ct, d_ca, nullptr, nullptr, nullptr,
                                 nullptr, nullptr);
  // 2. two- vs three-letter selection (only inside blocks == hashable positions < lim2)
  pos_pass_kernel<<<grid, 256>>>(1, d_g, lim2, lim3, d_ex, n_exclude, d_c2, d_ct, d_ca, d_is_two, nullptr, nullptr,
                                 nullptr, nullptr);
  IB_CUDA(cudaGetLastError());

  // 3. DP over blocks
  const std::vector<uint64_t> blocks = block_bounds(0, kBlockSize, lim2, exclude, n_exclude);
  const uint32_t n_blocks = (uint32_t)(blocks.size() / 2);
  if (n_blocks > 0) {
    IB_CUDA(cudaMalloc(&d_blocks, blocks.size() * 8));
    IB_CUDA(cudaMemcpy(d_blocks, blocks.data(), blocks.size() * 8, cudaMemcpyHostToDevice));
    IB_CUDA(cudaMalloc(&d_prev, (genome_size + 1) * 4));
    dp_kernel<<<(n_blocks + 31) / 32, 32>>>(d_g, d_blocks, n_blocks, d_c2, d_ct, d_ca, d_is_two, d_prev, d_keep, window_size);
    IB_CUDA(cudaDeviceSynchronize());
    cudaFree(d_prev);
    d_prev = nullptr;
  }

  // 4. masked bucket sizes
  IB_CUDA(cudaMemset(d_c2, 0, (n2 + 1) * 4));
  IB_CUDA(cudaMemset(d_ct, 0, (n3 + 1) * 4));
  IB_CUDA(cudaMemset(d_ca, 0, (n3 + 1) * 4));
  pos_pass_kernel<<<grid, 256>>>(2, d_g, lim2, lim3, d_ex, n_exclude, d_c2, d_ct, d_ca, d_is_two, d_keep, nullptr,
                                 nullptr, nullptr);
  IB_CUDA(cudaDeviceSynchronize());

  // 5. start offsets + fill.  NB the reference's masked three-letter counts run
  // to lim3 while the fill stops at lim2 (hash_genome :576); positions in
  // [lim2, lim3) lie in the end padding (excluded), so both agree.
  uint32_t tot2 = 0, tott = 0, tota = 0;
  if (exclusive_scan(d_c2, n2, &tot2) || exclusive_scan(d_ct, n3, &tott) || exclusive_scan(d_ca, n3, &tota)) {
    free_all();
    return ifail(ABG_ERR_CUDA, "abg_build_index: scan failed");
  }
  IB_CUDA(cudaMalloc(&d_s2, (n2 + 1) * 4));
  IB_CUDA(cudaMalloc(&d_st, (n3 + 1) * 4));
  IB_CUDA(cudaMalloc(&d_sa, (n3 + 1) * 4));
  IB_CUDA(cudaMemcpy(d_s2, d_c2, (n2 + 1) * 4, cudaMemcpyDeviceToDevice));
  IB_CUDA(cudaMemcpy(d_st, d_ct, (n3 + 1) * 4, cudaMemcpyDeviceToDevice));
  IB_CUDA(cudaMemcpy(d_sa, d_ca, (n3 + 1) * 4, cudaMemcpyDeviceToDevice));
  IB_CUDA(cudaMalloc(&d_i2, std::max<size_t>(tot2, 1) * 4));
  IB_CUDA(cudaMalloc(&d_it, std::max<size_t>(tott, 1) * 4));
  IB_CUDA(cudaMalloc(&d_ia, std::max<size_t>(tota, 1) * 4));
  pos_pass_kernel<<<grid, 256>>>(3, d_g, lim2, lim3, d_ex, n_exclude, d_c2, d_ct, d_ca, d_is_two, d_keep, d_i2, d_it,
                                 d_ia);
  IB_CUDA(cudaDeviceSynchronize());

  // 6. sort buckets
  if (sort_buckets<0>(d_g, d_s2, n2, d_i2) || sort_buckets<1>(d_g, d_st, n3, d_it) ||
      sort_buckets<2>(d_g, d_sa, n3, d_ia)) {
    free_all();
    return ifail(ABG_ERR_CUDA, "abg_build_index: bucket sort failed");
  }

  // 7. download
  out->counter = (uint32_t *)std::malloc((n2 + 1) * 4);
  out->counter_t = (uint32_t *)std::malloc((n3 + 1) * 4);
  out->counter_a = (uint32_t *)std::malloc((n3 + 1) * 4);
  out->index = (uint32_t *)std::malloc(std::max<size_t>(tot2, 1) * 4);
  out->index_t = (uint32_t *)std::malloc(std::max<size_t>(tott, 1) * 4);
  out->index_a = (uint32_t *)std::malloc(std::max<size_t>(tota, 1) * 4);
  if (!out->counter || !out->counter_t || !out->counter_a || !out->index || !out->index_t || !out->index_a) {
    free_all();
    abg_built_index_free(out);
    return ifail(ABG_ERR_INVALID, "abg_build_index: out of host memory");
  }
  IB_CUDA(cudaMemcpy(out->counter, d_s2, (n2 + 1) * 4, cudaMemcpyDeviceToHost));
  IB_CUDA(cudaMemcpy(out->counter_t, d_st, (n3 + 1) * 4, cudaMemcpyDeviceToHost));
  IB_CUDA(cudaMemcpy(out->counter_a, d_sa, (n3 + 1) * 4, cudaMemcpyDeviceToHost));
  IB_CUDA(cudaMemcpy(out->index, d_i2, (size_t)tot2 * 4, cudaMemcpyDeviceToHost));
  IB_CUDA(cudaMemcpy(out->index_t, d_it, (size_t)tott * 4, cudaMemcpyDeviceToHost));
  IB_CUDA(cudaMemcpy(out->index_a, d_ia, (size_t)tota * 4, cudaMemcpyDeviceToHost));
  out->counter_size = n2;
  out->counter_size_three = n3;
  out->index_size = tot2;
  out->index_size_three = tott;
  out->max_candidates = 100;  // compress_dp :852
  free_all();
  if (tott != tota) {
    abg_built_index_free(out);
    return ifail(ABG_ERR_INVALID, "abg_build_index: three-letter index sizes disagree");
  }
  return ABG_OK;
}

}  // extern "C"
