// mapper.cu -- C ABI (include/abismal_b200.h) over the sm_100a kernels in
// mapper_kernels.cuh.  Owns device memory, the stream and pinned staging;
// callers own every host buffer they pass in.  No CPU fallback exists: every
// entry point fails with ABG_ERR_CUDA when the device is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "abismal_b200.h"
#include "mapper_kernels.cuh"

namespace {

constexpr int kDefaultMinB = 2;

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define ABG_CUDA(call)                                                                      \
  do {                                                                                      \
    const cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ABG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

// Emptiness bitmap of one counter table (bit k = bucket k non-empty).  Kept only when the table is sparse
// enough for the L2-resident filter to save HBM probes; *out stays null otherwise.
int make_bitmap(const uint32_t *d_counter, uint64_t n_buckets, uint32_t **out) {
  *out = nullptr;
  uint32_t *bits = nullptr;
  unsigned long long *d_n = nullptr, n_set = 0;
  const uint64_t words = (n_buckets + 31) / 32;
  ABG_CUDA(cudaMalloc(reinterpret_cast<void **>(&bits), words * 4));
  ABG_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_n), 8));
  ABG_CUDA(cudaMemset(d_n, 0, 8));
  ab2dev::bucket_bitmap_kernel<<<148 * 8, 256>>>(d_counter, n_buckets, bits, d_n);
  ABG_CUDA(cudaGetLastError());
  ABG_CUDA(cudaMemcpy(&n_set, d_n, 8, cudaMemcpyDeviceToHost));
  cudaFree(d_n);
  if (n_set * 2 > n_buckets) {  // dense table: almost every probe would need the counters anyway
    cudaFree(bits);
    return ABG_OK;
  }
  *out = bits;
  return ABG_OK;
}

template <class T>
int upload(const T *host, uint64_t n, uint64_t n_alloc, T **dev) {
  *dev = nullptr;
  ABG_CUDA(cudaMalloc(reinterpret_cast<void **>(dev), std::max<uint64_t>(n_alloc, 1) * sizeof(T)));
  if (n_alloc > n) ABG_CUDA(cudaMemset(*dev + n, 0, (n_alloc - n) * sizeof(T)));
  if (n) ABG_CUDA(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return ABG_OK;
}

}  // namespace

struct abg_index {
  int device = 0;
  ab2dev::IndexDev dev{};
  uint64_t bytes = 0;
  uint64_t *genome = nullptr;
  uint32_t *counter = nullptr, *counter_t = nullptr, *counter_a = nullptr;
  uint32_t *index = nullptr, *index_t = nullptr, *index_a = nullptr;
  uint32_t *bits = nullptr, *bits_t = nullptr, *bits_a = nullptr;
};

struct abg_mapper {
  abg_index *idx = nullptr;
  abg_params params{};
  uint32_t max_batch = 0, max_read_len = 0, ml = 0;
  bool paired = false, count_work = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  // launch shape
  int grid = 0, minb = 0;
  size_t smem = 0;
  const void *kernel = nullptr;
  // device batch
  char *d_seq[2] = {nullptr, nullptr};
  uint32_t *d_off[2] = {nullptr, nullptr};
  size_t seq_cap = 0;
  // device results
  abg_hit *d_pe_r1 = nullptr, *d_pe_r2 = nullptr, *d_se[2] = {nullptr, nullptr};
  uint32_t *d_cigar[2] = {nullptr, nullptr}, *d_ncigar[2] = {nullptr, nullptr};
  // scratch
  uint64_t *d_pe_overflow = nullptr;
  int16_t *d_mem_scr = nullptr;
  uint64_t *d_tb = nullptr;
  uint32_t tb_words = 0;
  unsigned int *d_work = nullptr;   // [0] work counter, [1] error flag
  unsigned long long *d_counters = nullptr;
  // pinned staging
  char *h_seq[2] = {nullptr, nullptr};
  uint32_t *h_off[2] = {nullptr, nullptr};
  abg_hit *h_pe_r1 = nullptr, *h_pe_r2 = nullptr, *h_se[2] = {nullptr, nullptr};
  uint32_t *h_cigar[2] = {nullptr, nullptr}, *h_ncigar[2] = {nullptr, nullptr};
  unsigned int *h_flags = nullptr;
  abg_work_counters counters{};
  uint32_t cur_n = 0;
  bool timed = false;  // ev0/ev1 have been recorded
};

extern "C" {

const char *abg_last_error(void) { return g_err.c_str(); }

int abg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int abg_index_create(const abg_index_view *v, int device, abg_index **out) {
  if (!v || !out || !v->genome || !v->counter || !v->counter_t || !v->counter_a)
    return fail(ABG_ERR_INVALID, "abg_index_create: null argument");
  if (v->counter_size != (1ull << 25) || v->counter_size_three != 43046721ull)
    return fail(ABG_ERR_INVALID, "abg_index_create: unexpected counter sizes");
  int n_dev = 0;
  ABG_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return fail(ABG_ERR_CUDA, "abg_index_create: no such CUDA device");
  ABG_CUDA(cudaSetDevice(device));
  // The path is random 8..88-byte gathers: ask L2 to fetch single 32-byte sectors from HBM instead of
  // promoting every miss to 64/128 bytes (a hint; ignored where unsupported).
  if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32) != cudaSuccess) (void)cudaGetLastError();
  abg_index *ix = new (std::nothrow) abg_index();
  if (!ix) return fail(ABG_ERR_INVALID, "out of host memory");
  ix->device = device;
  int rc;
  // two spare zero words: the compare's look-ahead word and 16-byte loads
  if ((rc = upload(v->genome, v->genome_words, v->genome_words + 4, &ix->genome)) ||
      (rc = upload(v->counter, v->counter_size + 1, v->counter_size + 1, &ix->counter)) ||
      (rc = upload(v->counter_t, v->counter_size_three + 1, v->counter_size_three + 1, &ix->counter_t)) ||
      (rc = upload(v->counter_a, v->counter_size_three + 1, v->counter_size_three + 1, &ix->counter_a)) ||
      (rc = upload(v->index, v->index_size, v->index_size, &ix->index)) ||
      (rc = upload(v->index_t, v->index_size_three, v->index_size_three, &ix->index_t)) ||
      (rc = upload(v->index_a, v->index_size_three, v->index_size_three, &ix->index_a))) {
    abg_index_destroy(ix);
    return rc;
  }
  if ((rc = make_bitmap(ix->counter, v->counter_size, &ix->bits)) ||
      (rc = make_bitmap(ix->counter_t, v->counter_size_three, &ix->bits_t)) ||
      (rc = make_bitmap(ix->counter_a, v->counter_size_three, &ix->bits_a))) {
    abg_index_destroy(ix);
    return rc;
  }
  ix->dev.bits = ix->bits;
  ix->dev.bits_t = ix->bits_t;
  ix->dev.bits_a = ix->bits_a;
  ix->bytes = (v->genome_words + 4) * 8 + (v->counter_size + 1) * 4 + 2 * (v->counter_size_three + 1) * 4 +
              v->index_size * 4 + 2 * v->index_size_three * 4;
  ix->dev.genome = ix->genome;
  ix->dev.counter = ix->counter;
  ix->dev.counter_t = ix->counter_t;
  ix->dev.counter_a = ix->counter_a;
  ix->dev.index = ix->index;
  ix->dev.index_t = ix->index_t;
  ix->dev.index_a = ix->index_a;
  ix->dev.max_candidates = v->max_candidates;
  *out = ix;
  return ABG_OK;
}

void abg_index_destroy(abg_index *ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  cudaFree(ix->genome);
  cudaFree(ix->counter);
  cudaFree(ix->counter_t);
  cudaFree(ix->counter_a);
  cudaFree(ix->index);
  cudaFree(ix->index_t);
  cudaFree(ix->index_a);
  cudaFree(ix->bits);
  cudaFree(ix->bits_t);
  cudaFree(ix->bits_a);
  delete ix;
}

uint64_t abg_index_device_bytes(const abg_index *ix) { return ix ? ix->bytes : 0; }

int abg_mapper_create(abg_index *ix, const abg_params *p, uint32_t max_batch, uint32_t max_read_len,
                      int count_work, abg_mapper **out) {
  if (!ix || !p || !out || max_batch == 0) return fail(ABG_ERR_INVALID, "abg_mapper_create: bad argument");
  if (p->cigar_stride < 4) return fail(ABG_ERR_INVALID, "abg_mapper_create: cigar_stride must be >= 4");
  if (max_read_len < 44) max_read_len = 44;
  if (max_read_len > 4096) return fail(ABG_ERR_TOO_LONG, "abg_mapper_create: reads longer than 4096 are not supported");
  ABG_CUDA(cudaSetDevice(ix->device));
  abg_mapper *m = new (std::nothrow) abg_mapper();
  if (!m) return fail(ABG_ERR_INVALID, "out of host memory");
  m->idx = ix;
  m->params = *p;
  m->max_batch = max_batch;
  m->max_read_len = max_read_len;
  m->ml = (max_read_len + 31u) & ~31u;
  m->paired = p->mode & ABG_MODE_PAIRED;
  m->count_work = count_work != 0;
  const int n_ends = m->paired ? 2 : 1;
  const uint32_t stride = p->cigar_stride;

#define ABG_M(call)                                                                         \
  do {                                                                                      \
    const cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                                \
      abg_mapper_destroy(m);                                                                \
      return fail(ABG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    }                                                                                       \
  } while (0)

  ABG_M(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  ABG_M(cudaEventCreate(&m->ev0));
  ABG_M(cudaEventCreate(&m->ev1));

  // launch shape: persistent grid, as many CTAs per SM as shared memory/registers allow
  m->smem = ab2dev::block_smem_bytes(m->ml, m->paired);
  {
    // register-allocation variant (CTAs per SM the kernel is bounded for); ABISMAL_B200_MINB overrides for tuning
    const char *e = std::getenv("ABISMAL_B200_MINB");
    const int v = e ? std::atoi(e) : 0;
    m->minb = (v >= 2 && v <= 4) ? v : kDefaultMinB;
  }
  m->kernel = m->minb == 2 ? (const void *)ab2dev::map_reads_kernel<2>
            : m->minb == 4 ? (const void *)ab2dev::map_reads_kernel<4>
                           : (const void *)ab2dev::map_reads_kernel<3>;
  ABG_M(cudaFuncSetAttribute(m->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem));
  int n_sm = 0, per_sm = 0;
  ABG_M(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ix->device));
  ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->kernel, ab2dev::kThreadsPerBlock, m->smem));
  if (per_sm < 1) {
    abg_mapper_destroy(m);
    return fail(ABG_ERR_CUDA, "abg_mapper_create: kernel does not fit on an SM");
  }
  m->grid = n_sm * per_sm;
  const size_t slots = (size_t)m->grid * ab2dev::kWarpsPerBlock;

  m->seq_cap = (size_t)max_batch * max_read_len;
  for (int e = 0; e < n_ends; ++e) {
    ABG_M(cudaMalloc(&m->d_seq[e], m->seq_cap + 16));
    ABG_M(cudaMalloc(&m->d_off[e], ((size_t)max_batch + 1) * 4));
    ABG_M(cudaMalloc(&m->d_se[e], (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_cigar[e], (size_t)max_batch * stride * 4));
    ABG_M(cudaMalloc(&m->d_ncigar[e], (size_t)max_batch * 4));
    ABG_M(cudaMallocHost(&m->h_seq[e], m->seq_cap + 16));
    ABG_M(cudaMallocHost(&m->h_off[e], ((size_t)max_batch + 1) * 4));
    ABG_M(cudaMallocHost(&m->h_se[e], (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_cigar[e], (size_t)max_batch * stride * 4));
    ABG_M(cudaMallocHost(&m->h_ncigar[e], (size_t)max_batch * 4));
  }
  if (m->paired) {
    ABG_M(cudaMalloc(&m->d_pe_r1, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_pe_r2, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_pe_r1, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_pe_r2, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_pe_overflow, slots * 2 * ab2dev::kPeLarge * sizeof(uint64_t)));
    ABG_M(cudaMalloc(&m->d_mem_scr, slots * ab2dev::kPeLarge * sizeof(int16_t)));
  }
  m->tb_words = ab2dev::tb_sm_words(m->ml);
  ABG_M(cudaMalloc(&m->d_tb, slots * 2 * m->tb_words * 32 * sizeof(uint64_t)));
  ABG_M(cudaMalloc(&m->d_work, 2 * sizeof(unsigned int)));
  ABG_M(cudaMallocHost(&m->h_flags, 2 * sizeof(unsigned int)));
  if (m->count_work) ABG_M(cudaMalloc(&m->d_counters, 6 * sizeof(unsigned long long)));
#undef ABG_M
  *out = m;
  return ABG_OK;
}

void abg_mapper_destroy(abg_mapper *m) {
  if (!m) return;
  cudaSetDevice(m->idx->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (int e = 0; e < 2; ++e) {
    cudaFree(m->d_seq[e]);
    cudaFree(m->d_off[e]);
    cudaFree(m->d_se[e]);
    cudaFree(m->d_cigar[e]);
    cudaFree(m->d_ncigar[e]);
    cudaFreeHost(m->h_seq[e]);
    cudaFreeHost(m->h_off[e]);
    cudaFreeHost(m->h_se[e]);
    cudaFreeHost(m->h_cigar[e]);
    cudaFreeHost(m->h_ncigar[e]);
  }
  cudaFree(m->d_pe_r1);
  cudaFree(m->d_pe_r2);
  cudaFreeHost(m->h_pe_r1);
  cudaFreeHost(m->h_pe_r2);
  cudaFree(m->d_pe_overflow);
  cudaFree(m->d_mem_scr);
  cudaFree(m->d_tb);
  cudaFree(m->d_work);
  cudaFree(m->d_counters);
  cudaFreeHost(m->h_flags);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int abg_mapper_upload(abg_mapper *m, const abg_batch *b) {
  if (!m || !b || !b->seq1 || !b->off1) return fail(ABG_ERR_INVALID, "abg_mapper_upload: null argument");
  if (b->n > m->max_batch) return fail(ABG_ERR_INVALID, "abg_mapper_upload: batch larger than max_batch");
  if (m->paired && (!b->seq2 || !b->off2)) return fail(ABG_ERR_INVALID, "abg_mapper_upload: paired mode needs two ends");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  const int n_ends = m->paired ? 2 : 1;
  const char *seqs[2] = {b->seq1, b->seq2};
  const uint32_t *offs[2] = {b->off1, b->off2};
  for (int e = 0; e < n_ends; ++e) {
    const uint32_t *off = offs[e];
    const size_t bytes = off[b->n] - off[0];
    if (bytes > m->seq_cap) return fail(ABG_ERR_TOO_LONG, "abg_mapper_upload: batch sequence bytes exceed capacity");
    for (uint32_t i = 0; i < b->n; ++i) {
      const uint32_t len = off[i + 1] - off[i];
      if (len > m->max_read_len) return fail(ABG_ERR_TOO_LONG, "abg_mapper_upload: read longer than max_read_len");
      if (len != 0 && len < 44)
        return fail(ABG_ERR_INVALID, "abg_mapper_upload: reads shorter than 44 bases must be passed as empty");
      m->h_off[e][i] = off[i] - off[0];
    }
    m->h_off[e][b->n] = off[b->n] - off[0];
    std::memcpy(m->h_seq[e], seqs[e] + off[0], bytes);
    ABG_CUDA(cudaMemcpyAsync(m->d_seq[e], m->h_seq[e], bytes, cudaMemcpyHostToDevice, m->stream));
    ABG_CUDA(cudaMemcpyAsync(m->d_off[e], m->h_off[e], ((size_t)b->n + 1) * 4, cudaMemcpyHostToDevice, m->stream));
  }
  m->cur_n = b->n;
  return ABG_OK;
}

int abg_mapper_run(abg_mapper *m) {
  if (!m) return fail(ABG_ERR_INVALID, "abg_mapper_run: null mapper");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  ab2dev::KernelParams P;
  std::memset(&P, 0, sizeof P);
  P.ix = m->idx->dev;
  P.n = m->cur_n;
  for (int e = 0; e < 2; ++e) {
    P.seq[e] = m->d_seq[e];
    P.off[e] = m->d_off[e];
    P.se[e] = m->d_se[e];
    P.cigar[e] = m->d_cigar[e];
    P.n_cigar[e] = m->d_ncigar[e];
  }
  P.pe_r1 = m->d_pe_r1;
  P.pe_r2 = m->d_pe_r2;
  P.cigar_stride = m->params.cigar_stride;
  P.mode = m->params.mode;
  P.allow_ambig = m->params.allow_ambig;
  P.min_dist = m->params.min_dist;
  P.max_dist = m->params.max_dist;
  P.max_candidates = m->params.max_candidates ? m->params.max_candidates : m->idx->dev.max_candidates;
  P.valid_frac = m->params.valid_frac;
  P.ml = m->ml;
  P.pe_overflow = m->d_pe_overflow;
  P.mem_scr = m->d_mem_scr;
  P.tb = m->d_tb;
  P.tb_words = m->tb_words;
  P.work_counter = m->d_work;
  P.error_flag = m->d_work + 1;
  P.counters = m->d_counters;
  ABG_CUDA(cudaMemsetAsync(m->d_work, 0, 2 * sizeof(unsigned int), m->stream));
  if (m->d_counters) ABG_CUDA(cudaMemsetAsync(m->d_counters, 0, 6 * sizeof(unsigned long long), m->stream));
  ABG_CUDA(cudaEventRecord(m->ev0, m->stream));
  if (P.n > 0) {
    const int grid = (int)std::min<uint64_t>((uint64_t)m->grid, ((uint64_t)P.n + ab2dev::kWarpsPerBlock - 1) / ab2dev::kWarpsPerBlock);
    void *args[] = {&P};
    ABG_CUDA(cudaLaunchKernel(m->kernel, dim3(grid), dim3(ab2dev::kThreadsPerBlock), args, m->smem, m->stream));
  }
  ABG_CUDA(cudaEventRecord(m->ev1, m->stream));
  m->timed = true;
  return ABG_OK;
}

int abg_mapper_sync(abg_mapper *m) {
  if (!m) return fail(ABG_ERR_INVALID, "abg_mapper_sync: null mapper");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  ABG_CUDA(cudaStreamSynchronize(m->stream));
  if (!m->timed || cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1) != cudaSuccess) {
    m->last_ms = 0.f;
    (void)cudaGetLastError();  // do not leave a stale error behind
  }
  return ABG_OK;
}

int abg_mapper_download(abg_mapper *m, abg_results *r) {
  if (!m || !r || !r->se1) return fail(ABG_ERR_INVALID, "abg_mapper_download: null argument");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  const uint32_t n = m->cur_n;
  const uint32_t stride = m->params.cigar_stride;
  const int n_ends = m->paired ? 2 : 1;
  for (int e = 0; e < n_ends; ++e) {
    ABG_CUDA(cudaMemcpyAsync(m->h_se[e], m->d_se[e], (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, m->stream));
    ABG_CUDA(cudaMemcpyAsync(m->h_cigar[e], m->d_cigar[e], (size_t)n * stride * 4, cudaMemcpyDeviceToHost, m->stream));
    ABG_CUDA(cudaMemcpyAsync(m->h_ncigar[e], m->d_ncigar[e], (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  }
  if (m->paired) {
    ABG_CUDA(cudaMemcpyAsync(m->h_pe_r1, m->d_pe_r1, (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, m->stream));
    ABG_CUDA(cudaMemcpyAsync(m->h_pe_r2, m->d_pe_r2, (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, m->stream));
  }
  ABG_CUDA(cudaMemcpyAsync(m->h_flags, m->d_work, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, m->stream));
  if (m->d_counters)
    ABG_CUDA(cudaMemcpyAsync(&m->counters, m->d_counters, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             m->stream));
  ABG_CUDA(cudaStreamSynchronize(m->stream));
  if (!m->timed || cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1) != cudaSuccess) {
    m->last_ms = 0.f;
    (void)cudaGetLastError();  // do not leave a stale error behind
  }
  if (m->h_flags[1] != 0u)
    return fail(ABG_ERR_CIGAR_OVERFLOW, "abg_map_batch: a CIGAR needed more than cigar_stride operations");
  std::memcpy(r->se1, m->h_se[0], (size_t)n * sizeof(abg_hit));
  if (r->cigar1) std::memcpy(r->cigar1, m->h_cigar[0], (size_t)n * stride * 4);
  if (r->n_cigar1) std::memcpy(r->n_cigar1, m->h_ncigar[0], (size_t)n * 4);
  if (m->paired) {
    if (!r->pe_r1 || !r->pe_r2 || !r->se2) return fail(ABG_ERR_INVALID, "abg_mapper_download: paired results need pe_r1/pe_r2/se2");
    std::memcpy(r->pe_r1, m->h_pe_r1, (size_t)n * sizeof(abg_hit));
    std::memcpy(r->pe_r2, m->h_pe_r2, (size_t)n * sizeof(abg_hit));
    std::memcpy(r->se2, m->h_se[1], (size_t)n * sizeof(abg_hit));
    if (r->cigar2) std::memcpy(r->cigar2, m->h_cigar[1], (size_t)n * stride * 4);
    if (r->n_cigar2) std::memcpy(r->n_cigar2, m->h_ncigar[1], (size_t)n * 4);
  }
  return ABG_OK;
}

int abg_map_batch(abg_mapper *m, const abg_batch *b, abg_results *r) {
  int rc;
  if ((rc = abg_mapper_upload(m, b)) != ABG_OK) return rc;
  if ((rc = abg_mapper_run(m)) != ABG_OK) return rc;
  return abg_mapper_download(m, r);
}

float abg_mapper_last_kernel_ms(const abg_mapper *m) { return m ? m->last_ms : 0.f; }
uint32_t abg_mapper_launches_per_run(const abg_mapper *m) { return (m && m->cur_n) ? 1u : 0u; }

int abg_mapper_get_counters(const abg_mapper *m, abg_work_counters *out) {
  if (!m || !out) return fail(ABG_ERR_INVALID, "abg_mapper_get_counters: null argument");
  if (!m->count_work) return fail(ABG_ERR_INVALID, "abg_mapper_get_counters: mapper created with count_work == 0");
  *out = m->counters;
  return ABG_OK;
}

}  // extern "C"
