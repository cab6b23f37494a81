// mapper.cu -- C ABI (include/abismal_b200.h) over the sm_100a kernels in
// mapper_kernels.cuh.  Owns device memory, the stream and pinned staging;
// callers own every host buffer they pass in.  No CPU fallback exists: every
// entry point fails with ABG_ERR_CUDA when the device is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <chrono>
#include <thread>
#include <atomic>
#include <vector>

#include "abismal_b200.h"
#include "mapper_kernels.cuh"
#include "align_tasks.cuh"
#include "seed_bins.cuh"

namespace {

constexpr int kDefaultMinB = 4;       // measured: 4 CTAs/SM (<= 64 registers) beats 3 (<= 80) on both kernels
constexpr int kDefaultMinBSeed = 4;

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

// Grid of the one-off index preparation kernels: `per_sm` CTAs for every SM of the current device.
int prep_grid(int per_sm) {
  int dev = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n_sm < 1) {
    (void)cudaGetLastError();
    n_sm = 148;
  }
  return n_sm * per_sm;
}

// Emptiness bitmap of one counter table (bit k = bucket k non-empty).  Kept only when the table is sparse
// enough for the L2-resident filter to save HBM probes; *out stays null otherwise.
int make_bitmap(const uint32_t *d_counter, uint64_t n_buckets, uint32_t **out) {
  *out = nullptr;
  uint32_t *bits = nullptr;
  unsigned long long *d_n = nullptr, n_set = 0;
  const uint64_t words = (n_buckets + 31) / 32;
  ABG_CUDA(cudaMalloc(reinterpret_cast<void **>(&bits), words * 4));
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_n), 8);
  if (e == cudaSuccess) e = cudaMemset(d_n, 0, 8);
  if (e == cudaSuccess) {
    ab2dev::bucket_bitmap_kernel<<<prep_grid(8), 256>>>(d_counter, n_buckets, bits, d_n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&n_set, d_n, 8, cudaMemcpyDeviceToHost);
  cudaFree(d_n);
  if (e != cudaSuccess) {
    cudaFree(bits);
    return fail(ABG_ERR_CUDA, std::string("bucket_bitmap_kernel: ") + cudaGetErrorString(e));
  }
  if (n_set * 2 > n_buckets) {  // dense table: almost every probe would need the counters anyway
    cudaFree(bits);
    return ABG_OK;
  }
  *out = bits;
  return ABG_OK;
}

// Host -> device copy of a large array that is usually a read-only file mapping (pageable, not yet touched):
// several threads copy 16 MB pieces into their own page-locked slots and send them on their own streams, so
// the page faults, the host copies and the DMA of different pieces overlap.
int copy_to_device(void *dst, const void *src, size_t bytes) {
  constexpr size_t kPiece = 16u << 20;
  cudaPointerAttributes at;
  const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
  (void)cudaGetLastError();
  // Off by default: measured next to the front end's reader threads on a 16-core box it was slower than the
  // single cudaMemcpy (1.8-2.3 s against 1.5 s for 2.7 GB).  ABISMAL_B200_UPLOAD_THREADS=N turns it on.
  unsigned n_thr = 1;
  if (const char *e = std::getenv("ABISMAL_B200_UPLOAD_THREADS")) n_thr = (unsigned)std::max(1, std::min(16, std::atoi(e)));
  if (pinned || bytes < 4 * kPiece || n_thr < 2) {
    ABG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return ABG_OK;
  }
  int device = 0;
  ABG_CUDA(cudaGetDevice(&device));
  const size_t n_pieces = (bytes + kPiece - 1) / kPiece;
  std::atomic<int> err{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < n_thr; ++t)
    th.emplace_back([&, t] {
      void *slot = nullptr;
      cudaStream_t st = nullptr;
      if (cudaSetDevice(device) != cudaSuccess || cudaMallocHost(&slot, kPiece) != cudaSuccess ||
          cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        err = 1;
      }
      else {
        for (size_t c = t; c < n_pieces && !err; c += n_thr) {
          const size_t o = c * kPiece, len = std::min(kPiece, bytes - o);
          if (cudaStreamSynchronize(st) != cudaSuccess) err = 1;  // the slot's previous piece has left
          std::memcpy(slot, static_cast<const char *>(src) + o, len);
          if (cudaMemcpyAsync(static_cast<char *>(dst) + o, slot, len, cudaMemcpyHostToDevice, st) != cudaSuccess) err = 1;
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) err = 1;
      }
      if (st) cudaStreamDestroy(st);
      if (slot) cudaFreeHost(slot);
    });
  for (std::thread &x : th) x.join();
  if (err) {
    (void)cudaGetLastError();
    ABG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));  // plain copy; reports its own error if any
  }
  return ABG_OK;
}

template <class T>
int upload(const T *host, uint64_t n, uint64_t n_alloc, T **dev) {
  *dev = nullptr;
  ABG_CUDA(cudaMalloc(reinterpret_cast<void **>(dev), std::max<uint64_t>(n_alloc, 1) * sizeof(T)));
  if (n_alloc > n) ABG_CUDA(cudaMemset(*dev + n, 0, (n_alloc - n) * sizeof(T)));
  if (n) {
    int rc = copy_to_device(*dev, host, n * sizeof(T));
    if (rc != ABG_OK) return rc;
  }
  return ABG_OK;
}

}  // namespace

struct abg_index {
  int device = 0;
  ab2dev::IndexDev dev{};
  uint64_t bytes = 0, bytes_extra = 0;
  uint64_t *genome = nullptr;
  uint32_t *counter = nullptr, *counter_t = nullptr, *counter_a = nullptr;
  uint32_t *index = nullptr, *index_t = nullptr, *index_a = nullptr;
  uint32_t *bits = nullptr, *bits_t = nullptr, *bits_a = nullptr;
  uint64_t *g2 = nullptr;
  uint32_t *gx = nullptr;
  bool has_iupac = false;  // the genome holds multi-bit codes: records near them are sentinels (IndexDev::ctx)
  uint4 *ctx = nullptr, *ctx_t = nullptr, *ctx_a = nullptr;  // seed-context records (IndexDev::ctx)
  uint4 *cc = nullptr;                                        // compact two-letter counters (IndexDev::cc)
  uint64_t cc_bytes = 0;
};

constexpr uint32_t kInlineOps = 16;   // CIGAR ops per read copied back with the batch; longer ones are fetched afterwards
constexpr uint32_t kMaxChunks = 256;  // sub-batches one abg_map_batch call is pipelined over
// device words per sub-batch: work counters of map / seed / align, redo count, work counter of enum_kernel,
// task slots handed out per class [3] + traceback units, dp_kernel's cursors [3]
constexpr uint32_t kChunkWords = 14;  // ... [12] hash_kernel's work counter (binned seeding), [13] spare
constexpr uint32_t kWorkWords = 2 + kChunkWords * kMaxChunks;  // [0] error flag, [1] cursor of the set overflow arena
constexpr uint32_t kOvfPerItem = 256;  // overflow arena entries per pair of max_batch (sets beyond set_slots entries):
// 3 GB per 2^20 pairs; a repeat-rich 100 Mbp test genome needs ~100 per pair (tools/repeat_perf.py)
// task-parallel alignment: slots per read / pair in the three task lists (bands <= 16 / <= 32 / <= 61 columns) and
// traceback allocations per read / pair (in tasks of the first class); beyond these the alignment runs in the warp
constexpr uint32_t kTaskCapPe[3] = {40, 32, 8}, kTaskCapSe[3] = {24, 8, 8};  // 2 GB per 2^20 pairs
constexpr uint32_t kTbTasksPe = 4, kTbTasksSe = 2;
constexpr uint32_t kMaxFilterCursors = 64, kBinWorkWords = 2 + kMaxFilterCursors;
constexpr int kScatterCtasPerSmMax = 2;  // count_kernel / scatter_kernel: at most this many CTAs per SM

struct abg_mapper {
  abg_index *idx = nullptr;
  abg_params params{};
  uint32_t max_batch = 0, max_read_len = 0, ml = 0, chunk = 0, chunk2 = 0;
  bool paired = false, count_work = false;
  cudaStream_t stream = nullptr;                       // upload / run / download (split API)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;       // pipelined abg_map_batch
  cudaStream_t s_run[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_start = nullptr;
  cudaEvent_t ev_ph[2] = {nullptr, nullptr};           // abg_mapper_run: after seed_kernel, after align_kernel
  float phase_ms[3] = {0.f, 0.f, 0.f};                 // seed, align, redo (two-phase) or {whole, 0, 0}
  cudaEvent_t ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {}, ev_out[kMaxChunks] = {};
  float last_ms = 0.f;
  // launch shape
  int grid = 0, minb = 0, grid_scratch = 0;
  size_t smem = 0, smem_s = 0, smem_a = 0;   // dynamic shared memory: full layout, seeding kernel, alignment kernel
  const void *kernel = nullptr;
  // two-phase launch: seed_kernel -> align_kernel -> map_reads_kernel over the redo list
  bool split = false;
  const void *kernel_s = nullptr, *kernel_a = nullptr;
  int grid_s = 0, grid_a = 0;
  // overlapped launch: align_kernel (grid_a_aux) next to seed_kernel (grid_s) on a second stream, then the rest
  // of the alignment (grid_a) behind the seeding; see launch()
  bool overlap = false;
  int grid_a_aux = 0;
  uint32_t wait_ns = 0;
  unsigned int *d_ready = nullptr;
  cudaStream_t s_aux[3] = {nullptr, nullptr, nullptr};          // partners of stream, s_run[0], s_run[1]
  cudaEvent_t ev_go[3] = {nullptr, nullptr, nullptr}, ev_aux[3] = {nullptr, nullptr, nullptr};
  uint32_t n_pass = 1, set_slots = 0;
  // task-parallel alignment (align_tasks.cuh): enum_kernel -> dp_kernel between seed_kernel and align_kernel
  bool use_tasks = false;
  const void *kernel_e = nullptr;
  int grid_e = 0, grid_d = 0;
  size_t smem_d = 0;
  uint32_t task_cap_item[3] = {0, 0, 0}, task_slack = 0, tb_cap_item = 0, tb_slack = 0, n_chunks_max = 1;
  uint64_t task_class_off[3] = {0, 0, 0};
  ab2dev::AlignTask *d_tasks = nullptr;
  ab2dev::TaskResult *d_task_res = nullptr;
  uint64_t *d_task_tb = nullptr;
  uint32_t *d_task_of = nullptr;
  uint64_t *d_set_ovf = nullptr;   // entries of stored paired-end sets beyond set_slots
  uint32_t *d_task_ovf = nullptr;  // their task ids
  uint32_t ovf_cap = 0;
  cudaEvent_t ev_t[2] = {nullptr, nullptr};  // abg_mapper_run: after enum_kernel, after dp_kernel
  float task_ms[2] = {0.f, 0.f};
  // binned seeding (seed_bins.cuh): hash_kernel -> scatter_kernel -> filter_kernel in front of seed_kernel
  bool use_bins = false;
  uint32_t spi = 1, bin_shift = 0, n_bins = 0, tup_cap = 0, pw = 0, surv_cap = 0, acc_cap = 32;
  const void *kernel_h = nullptr;
  int grid_h = 0, grid_sc = 0, grid_f = 0;
  bool scatter_sorted = true;  // scatter_sorted_kernel: tuples leave in runs per bin (shared-memory tile sort); measured 7.4 vs 12.6 ms
  uint32_t filter_cursors = 1;  // interleaved work cursors of the filter
  uint32_t filter_cache = 1;  // bit 0 planes loaded evict-first (measured 24.2 vs 25.3 ms), bit 1 records loaded with the L2 evict-last hint (no gain)
  bool filter_minb8 = false;  // tuning: filter_kernel bounded for 8 CTAs per SM (32 registers)
  bool filter_pipe = false;  // filter_kernel<true>: two record gathers per lane in flight
  uint32_t filter_grab = 256;  // tuples per work-cursor atomic (measured at 2^20 pairs: 32 -> 44 ms, 64 -> 30 ms, 128 -> 21 ms, 256 -> 20 ms)
  ab2dev::SeedTuple *d_tup = nullptr, *d_tup_b = nullptr;
  uint32_t *d_planes = nullptr, *d_bin_hist = nullptr, *d_surv_count = nullptr;
  uint8_t *d_strand_flag = nullptr;
  uint2 *d_surv = nullptr;
  unsigned int *d_bin_work = nullptr;  // [0] tuple slots handed out, [1] tuples binned, [2 ..] filter_kernel's work cursors (kBinWorkWords in all)
  cudaEvent_t ev_b[3] = {nullptr, nullptr, nullptr};  // abg_mapper_run: after hash_kernel, scatter_kernel, filter_kernel
  cudaEvent_t ev_bins = nullptr;                      // abg_map_batch: the batch's tuples are filtered
  float bin_ms[3] = {0.f, 0.f, 0.f};
  uint64_t *d_sets = nullptr;
  unsigned int *d_redo_flag = nullptr;
  uint32_t *d_redo_list = nullptr;
  // device batch
  char *d_seq[2] = {nullptr, nullptr};
  uint32_t *d_off[2] = {nullptr, nullptr};
  size_t seq_cap = 0;
  // device results
  abg_hit *d_pe_r1 = nullptr, *d_pe_r2 = nullptr, *d_se[2] = {nullptr, nullptr};
  uint32_t *d_cigar[2] = {nullptr, nullptr}, *d_ncigar[2] = {nullptr, nullptr};
  uint32_t *d_cigar_inline[2] = {nullptr, nullptr};  // [max_batch][inline_ops]
  // scratch
  uint64_t *d_pe_overflow = nullptr;
  int16_t *d_mem_scr = nullptr, *d_mem_scr2 = nullptr;
  uint32_t heavy_min = 64;
  uint64_t *d_tb = nullptr;
  uint32_t tb_words = 0;
  unsigned int *d_work = nullptr;   // [0] error flag, [1] overflow arena cursor, [2 + kChunkWords j ..] sub-batch j (see kChunkWords)
  unsigned long long *d_counters = nullptr;
  // pinned staging (used when the caller's buffers are pageable)
  char *h_seq[2] = {nullptr, nullptr};
  uint32_t *h_off[2] = {nullptr, nullptr};
  abg_hit *h_pe_r1 = nullptr, *h_pe_r2 = nullptr, *h_se[2] = {nullptr, nullptr};
  uint32_t *h_cigar[2] = {nullptr, nullptr}, *h_ncigar[2] = {nullptr, nullptr};  // h_cigar: kInlineOps per read
  unsigned int *h_flags = nullptr;
  abg_work_counters counters{};
  uint32_t cur_n = 0;
  bool timed = false;  // ev0/ev1 have been recorded
  bool run_pending = false;  // abg_mapper_run's error flag has not been looked at yet
};

namespace {

bool is_pinned(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Offsets of reads [c0, c1] rebased to the start of the batch, with the ReadLoader contract checked.
int stage_offsets(abg_mapper *m, const uint32_t *off, uint32_t c0, uint32_t c1, uint32_t *dst) {
  const uint32_t base = off[0];
  for (uint32_t i = c0; i < c1; ++i) {
    const uint32_t len = off[i + 1] - off[i];
    if (len > m->max_read_len) return fail(ABG_ERR_TOO_LONG, "abg_map_batch: read longer than max_read_len");
    if (len != 0 && len < m->idx->dev.window_size + 24u)  // ReadLoader::min_read_length, abismal.cpp:212-213
      return fail(ABG_ERR_INVALID, "abg_map_batch: reads shorter than 44 bases (36 with window 12) must be passed as empty");
    dst[i] = off[i] - base;
  }
  dst[c1] = off[c1] - base;
  return ABG_OK;
}

void fill_params(const abg_mapper *m, ab2dev::KernelParams &P, uint32_t c0, uint32_t n, uint32_t chunk_idx) {
  std::memset(&P, 0, sizeof P);
  unsigned int *work = m->d_work + 2 + kChunkWords * chunk_idx;
  P.set_ovf = m->d_set_ovf;
  P.task_ovf = m->d_task_ovf;
  P.ovf_count = m->d_work + 1;
  P.ovf_cap = m->ovf_cap;
  if (m->use_tasks) {
    // the sub-batch's own regions of the task lists, results, traceback arena: items [c0, c0 + n) + its slack
    P.tasks = m->d_tasks;
    P.task_res = m->d_task_res;
    P.task_of = m->d_task_of + (size_t)c0 * m->n_pass * m->set_slots;
    for (int c = 0; c < 3; ++c) {
      P.task_base[c] = (uint32_t)(m->task_class_off[c] + (uint64_t)c0 * m->task_cap_item[c] + (uint64_t)chunk_idx * m->task_slack);
      P.task_cap[c] = n * m->task_cap_item[c] + m->task_slack;
    }
    P.task_tb = m->d_task_tb + ((uint64_t)c0 * m->tb_cap_item + (uint64_t)chunk_idx * m->tb_slack) * 8u;
    P.task_tb_cap = n * m->tb_cap_item + m->tb_slack;
    P.task_count = work + 5;
    P.task_cursor = work + 9;
  }
  if (m->split) {
    P.sets = m->d_sets + (size_t)c0 * m->n_pass * (ab2dev::kSetStateWords + m->set_slots);
    P.set_slots = m->set_slots;
    P.n_pass = m->n_pass;
    P.redo_flag = m->d_redo_flag + c0;
    P.redo_list = m->d_redo_list + c0;
    P.redo_count = work + 3;
  }
  if (m->use_bins) {
    ab2dev::BinParams &B = P.bp;
    B.tup = m->d_tup;
    B.tup_count = m->d_bin_work;
    B.tup_cap = m->tup_cap;
    B.pw = m->pw;
    B.planes = m->d_planes;
    B.strand_flag = m->d_strand_flag;
    B.bin_hist = m->d_bin_hist;
    B.bin_shift = m->bin_shift;
    B.n_bins = m->n_bins;
    B.rec_base[0] = 0;
    B.rec_base[1] = (uint64_t)ab2dev::kCtxArrays * m->idx->dev.n_ctx;
    B.rec_base[2] = B.rec_base[1] + (uint64_t)ab2dev::kCtxArrays * m->idx->dev.n_ctx3;
    B.surv_count = m->d_surv_count;
    B.surv = m->d_surv;
    B.surv_cap = m->surv_cap;
    B.acc_cap = m->acc_cap;
    B.spi = m->spi;
    B.sid_base = c0 * m->spi;
  }
  P.ix = m->idx->dev;
  P.n = n;
  const uint32_t stride = m->params.cigar_stride;
  for (int e = 0; e < 2; ++e) {
    P.seq[e] = m->d_seq[e];
    P.off[e] = m->d_off[e] ? m->d_off[e] + c0 : nullptr;
    P.se[e] = m->d_se[e] ? m->d_se[e] + c0 : nullptr;
    P.cigar[e] = m->d_cigar[e] ? m->d_cigar[e] + (size_t)c0 * stride : nullptr;
    P.n_cigar[e] = m->d_ncigar[e] ? m->d_ncigar[e] + c0 : nullptr;
    P.cigar_inline[e] = m->d_cigar_inline[e] ? m->d_cigar_inline[e] + (size_t)c0 * std::min(kInlineOps, stride) : nullptr;
  }
  P.inline_ops = std::min(kInlineOps, stride);
  P.pe_r1 = m->d_pe_r1 ? m->d_pe_r1 + c0 : nullptr;
  P.pe_r2 = m->d_pe_r2 ? m->d_pe_r2 + c0 : nullptr;
  P.cigar_stride = stride;
  P.mode = m->params.mode;
  P.allow_ambig = m->params.allow_ambig;
  P.min_dist = m->params.min_dist;
  P.max_dist = m->params.max_dist;
  P.max_candidates = m->params.max_candidates ? m->params.max_candidates : m->idx->dev.max_candidates;
  P.valid_frac = m->params.valid_frac;
  P.window_size = m->idx->dev.window_size;
  P.ml = m->ml;
  P.pe_overflow = m->d_pe_overflow;
  P.mem_scr = m->d_mem_scr;
  P.mem_scr2 = m->d_mem_scr2;
  P.heavy_min = m->heavy_min;
  P.tb = m->d_tb;
  P.tb_words = m->tb_words;
  P.work_counter = work;
  P.error_flag = m->d_work;
  P.counters = m->d_counters;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (kernel, device) pair, not to a mapper: mappers
// created with different max_read_len on one device (the front end runs two workers per GPU and recreates
// them as read lengths grow) must never lower it under another mapper's launches.  Only ever raise it.
std::mutex g_smem_mu;
std::map<std::pair<int, const void *>, size_t> g_smem_cap;

cudaError_t raise_smem_cap(int device, const void *kernel, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_smem_mu);
  size_t &cap = g_smem_cap[std::make_pair(device, kernel)];
  if (bytes <= cap) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cap = bytes;
  return e;
}

int launch_one(const void *kernel, int grid, size_t smem, ab2dev::KernelParams &P, cudaStream_t st) {
  void *args[] = {&P};
  ABG_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(ab2dev::kThreadsPerBlock), args, smem, st));
  return ABG_OK;
}

// ---- binned seeding (seed_bins.cuh) ----
ab2dev::FilterParams filter_params(const abg_mapper *m) {
  ab2dev::FilterParams F;
  std::memset(&F, 0, sizeof F);
  const ab2dev::IndexDev &ix = m->idx->dev;
  F.tup = m->d_tup;
  F.tup_count = m->d_bin_work;
  F.tup_cap = m->tup_cap;
  F.planes = m->d_planes;
  F.pw = m->pw;
  F.bin_shift = m->bin_shift;
  F.n_bins = m->n_bins;
  F.rec_base[0] = 0;
  F.rec_base[1] = (uint64_t)ab2dev::kCtxArrays * ix.n_ctx;
  F.rec_base[2] = F.rec_base[1] + (uint64_t)ab2dev::kCtxArrays * ix.n_ctx3;
  F.bin_hist = m->d_bin_hist;
  F.n_binned = m->d_bin_work + 1;
  F.tup_b = m->d_tup_b;
  F.ctx[0] = ix.ctx;
  F.ctx[1] = ix.ctx_t;
  F.ctx[2] = ix.ctx_a;
  F.n_tab[0] = ix.n_ctx;
  F.n_tab[1] = F.n_tab[2] = ix.n_ctx3;
  F.surv_count = m->d_surv_count;
  F.surv = m->d_surv;
  F.surv_cap = m->surv_cap;
  F.work = m->d_bin_work + 2;
  F.grab = m->filter_grab;
  F.cache = m->filter_cache;
  F.n_cursors = m->filter_cursors;
  return F;
}

// clears the batch-wide state of the binned kernels (n = reads / pairs of the whole batch)
int reset_bins(const abg_mapper *m, uint32_t n, cudaStream_t st) {
  ABG_CUDA(cudaMemsetAsync(m->d_bin_work, 0, kBinWorkWords * sizeof(unsigned int), st));
  ABG_CUDA(cudaMemsetAsync(m->d_surv_count, 0, (size_t)n * m->spi * 4, st));
  return ABG_OK;
}

// hash_kernel over the strands of sub-batch P: appends to the batch's tuple buffer and histogram
int launch_hash(const abg_mapper *m, const ab2dev::KernelParams &P, cudaStream_t st) {
  if (P.n == 0) return ABG_OK;
  ab2dev::KernelParams Q = P;
  Q.work_counter = P.work_counter + 12;
  Q.layout_kind = ab2dev::kLayoutSeed;
  const uint64_t wpb = ab2dev::kWarpsPerBlock;
  const uint64_t n_work = P.n;  // one warp per read / pair: all its strands
  void *args[] = {&Q};
  ABG_CUDA(cudaLaunchKernel(m->kernel_h, dim3((unsigned)std::min<uint64_t>((uint64_t)m->grid_h, (n_work + wpb - 1) / wpb)),
                            dim3(ab2dev::kThreadsPerBlock), args, m->smem_s, st));
  return ABG_OK;
}

// bins of everything hashed since reset_bins: write cursors, scatter, prefilter -> survivor lists
int launch_bins(const abg_mapper *m, cudaStream_t st, const cudaEvent_t *ev_b = nullptr) {
  ab2dev::FilterParams F = filter_params(m);
  const size_t sm = (size_t)m->n_bins * 4;
  ab2dev::count_kernel<<<m->grid_sc, ab2dev::kScatterThreads, sm, st>>>(F);
  ab2dev::bin_prefix_kernel<<<1, 1024, 0, st>>>(F, (uint32_t)m->grid_sc);
  if (m->scatter_sorted) ab2dev::scatter_sorted_kernel<<<m->grid_sc, ab2dev::kScatterThreads, ab2dev::sort_scatter_smem(m->n_bins), st>>>(F);
  else ab2dev::scatter_kernel<<<m->grid_sc, ab2dev::kScatterThreads, sm, st>>>(F);
  if (ev_b) ABG_CUDA(cudaEventRecord(ev_b[1], st));
  if (m->filter_pipe) ab2dev::filter_kernel<true, false><<<m->grid_f, 256, 0, st>>>(F);
  else if (m->filter_minb8) ab2dev::filter_kernel<false, false, 8><<<m->grid_f, 256, 0, st>>>(F);
  else if (m->filter_cache & 2u) ab2dev::filter_kernel<false, true><<<m->grid_f, 256, 0, st>>>(F);
  else ab2dev::filter_kernel<false, false><<<m->grid_f, 256, 0, st>>>(F);
  if (ev_b) ABG_CUDA(cudaEventRecord(ev_b[2], st));
  ABG_CUDA(cudaGetLastError());
  return ABG_OK;
}

// One sub-batch on stream st.  Two-phase mode: seeding (one warp per read strand), then alignment/mating (one
// warp per pair) from the stored candidate sets, then the single-kernel path for the few pairs whose sets
// outgrew the stored form.  P.work_counter points at this chunk's {map, seed, align} counters + redo count.
int launch(const abg_mapper *m, ab2dev::KernelParams &P, cudaStream_t st, const cudaEvent_t *ev_ph = nullptr,
           bool bins_done = false) {
  if (P.n == 0) return ABG_OK;
  const uint64_t wpb = ab2dev::kWarpsPerBlock;
  int rc;
  if (m->use_bins && !bins_done) {  // the sub-batch is the whole batch: hash, bin and filter it here
    if ((rc = reset_bins(m, P.n, st))) return rc;
    if ((rc = launch_hash(m, P, st))) return rc;
    if (ev_ph) ABG_CUDA(cudaEventRecord(m->ev_b[0], st));
    if ((rc = launch_bins(m, st, ev_ph ? m->ev_b : nullptr))) return rc;
  }
  if (!m->split) {
    const int grid = (int)std::min<uint64_t>((uint64_t)m->grid, ((uint64_t)P.n + wpb - 1) / wpb);
    P.layout_kind = ab2dev::kLayoutFull;
    return launch_one(m->kernel, grid, m->smem, P, st);
  }
  unsigned int *work = P.work_counter;
  ABG_CUDA(cudaMemsetAsync(P.redo_flag, 0, (size_t)P.n * sizeof(unsigned int), st));
  ab2dev::KernelParams Q = P;
  const uint64_t n_work = m->paired ? (uint64_t)P.n * m->n_pass : P.n;
  const int grid_s = (int)std::min<uint64_t>((uint64_t)m->grid_s, (n_work + wpb - 1) / wpb);
  const int s_idx = st == m->stream ? 0 : (st == m->s_run[0] ? 1 : 2);
  const bool overlap = m->overlap;
  if (overlap) {
    // Seeding is bound by DRAM latency and leaves most issue slots idle; alignment is bound by issue slots
    // and needs no DRAM.  Run them on the same SMs at the same time: seed_kernel with grid_s CTAs, next to it
    // (second stream) align_kernel with grid_a_aux CTAs that consumes pairs as their sets are published,
    // and behind the seeding a second align_kernel launch that takes the SM share the seeding frees.
    Q.ready = m->d_ready + (P.redo_flag - m->d_redo_flag);
    Q.ready_need = m->paired ? m->n_pass : 1u;
    Q.wait_ns = m->wait_ns;
    ABG_CUDA(cudaMemsetAsync(Q.ready, 0, (size_t)P.n * sizeof(unsigned int), st));
    ABG_CUDA(cudaEventRecord(m->ev_go[s_idx], st));
    ABG_CUDA(cudaStreamWaitEvent(m->s_aux[s_idx], m->ev_go[s_idx], 0));
  }
  Q.work_counter = work + 1;
  Q.layout_kind = ab2dev::kLayoutSeed;
  Q.slot_base = P.slot_base;
  if ((rc = launch_one(m->kernel_s, grid_s, m->smem_s, Q, st))) return rc;
  if (ev_ph) ABG_CUDA(cudaEventRecord(ev_ph[0], st));
  Q.layout_kind = ab2dev::kLayoutAlign;
  const uint64_t a_blocks = ((uint64_t)P.n + wpb - 1) / wpb;
  if (m->use_tasks) {
    // enumerate the alignments (sorted sets + task lists), run them task-parallel; align_kernel then looks them up
    Q.work_counter = work + 4;
    if ((rc = launch_one(m->kernel_e, (int)std::min<uint64_t>((uint64_t)m->grid_e, a_blocks), m->smem_a, Q, st))) return rc;
    if (ev_ph) ABG_CUDA(cudaEventRecord(m->ev_t[0], st));
    if ((rc = launch_one((const void *)ab2dev::dp_kernel, m->grid_d, m->smem_d, Q, st))) return rc;
    if (ev_ph) ABG_CUDA(cudaEventRecord(m->ev_t[1], st));
  }
  Q.work_counter = work + 2;
  if (overlap) {
    ab2dev::KernelParams A = Q;
    A.slot_base = P.slot_base + (uint32_t)m->grid_s * (uint32_t)wpb;  // disjoint from the seeding kernel's slots
    if ((rc = launch_one(m->kernel_a, (int)std::min<uint64_t>((uint64_t)m->grid_a_aux, a_blocks), m->smem_a, A, m->s_aux[s_idx])))
      return rc;
    ABG_CUDA(cudaEventRecord(m->ev_aux[s_idx], m->s_aux[s_idx]));
  }
  // behind the seeding (stream order): every set is stored, nothing to wait for; takes the seeding's slots
  Q.ready = nullptr;
  const int grid_a_main = overlap ? m->grid_s : m->grid_a;  // overlapped: exactly the slots the seeding kernel had
  if ((rc = launch_one(m->kernel_a, (int)std::min<uint64_t>((uint64_t)grid_a_main, a_blocks), m->smem_a, Q, st))) return rc;
  if (overlap) ABG_CUDA(cudaStreamWaitEvent(st, m->ev_aux[s_idx], 0));
  if (ev_ph) ABG_CUDA(cudaEventRecord(ev_ph[1], st));
  Q.work_counter = work;
  Q.item_list = P.redo_list;
  Q.n_items_ptr = P.redo_count;
  Q.layout_kind = ab2dev::kLayoutFull;
  Q.slot_base = P.slot_base;
  Q.tasks = nullptr;  // the redo kernel maps its pairs from scratch, alignments in the warp
  return launch_one(m->kernel, m->grid, m->smem, Q, st);
}

struct ResultDst {  // where hit records and CIGAR lengths land: the caller's buffers when pinned, else staging
  abg_hit *pe_r1, *pe_r2, *se[2];
  uint32_t *ncig[2];
  bool direct;
};

ResultDst pick_result_dst(abg_mapper *m, const abg_results *r) {
  ResultDst d{};
  bool direct = is_pinned(r->se1) && is_pinned(r->n_cigar1);
  if (m->paired) direct = direct && is_pinned(r->pe_r1) && is_pinned(r->pe_r2) && is_pinned(r->se2) && is_pinned(r->n_cigar2);
  d.direct = direct;
  if (direct) {
    d.pe_r1 = r->pe_r1;
    d.pe_r2 = r->pe_r2;
    d.se[0] = r->se1;
    d.se[1] = r->se2;
    d.ncig[0] = r->n_cigar1;
    d.ncig[1] = r->n_cigar2;
  }
  else {
    d.pe_r1 = m->h_pe_r1;
    d.pe_r2 = m->h_pe_r2;
    for (int e = 0; e < 2; ++e) {
      d.se[e] = m->h_se[e];
      d.ncig[e] = m->h_ncigar[e];
    }
  }
  return d;
}

// Device -> host copy of the results of reads [c0, c0 + n): hit records, CIGAR lengths and the compact
// [n][inline_ops] array holding the first operations of every CIGAR (into the mapper's pinned staging).
int copy_results_async(abg_mapper *m, const ResultDst &d, uint32_t c0, uint32_t n, cudaStream_t st) {
  if (n == 0) return ABG_OK;
  const uint32_t w = std::min(kInlineOps, m->params.cigar_stride);
  const int n_ends = m->paired ? 2 : 1;
  for (int e = 0; e < n_ends; ++e) {
    ABG_CUDA(cudaMemcpyAsync(d.se[e] + c0, m->d_se[e] + c0, (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, st));
    ABG_CUDA(cudaMemcpyAsync(d.ncig[e] + c0, m->d_ncigar[e] + c0, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    ABG_CUDA(cudaMemcpyAsync(m->h_cigar[e] + (size_t)c0 * w, m->d_cigar_inline[e] + (size_t)c0 * w, (size_t)n * w * 4,
                             cudaMemcpyDeviceToHost, st));
  }
  if (m->paired) {
    ABG_CUDA(cudaMemcpyAsync(d.pe_r1 + c0, m->d_pe_r1 + c0, (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, st));
    ABG_CUDA(cudaMemcpyAsync(d.pe_r2 + c0, m->d_pe_r2 + c0, (size_t)n * sizeof(abg_hit), cudaMemcpyDeviceToHost, st));
  }
  return ABG_OK;
}

// Host side of reads [c0, c0 + n) once their copies have landed: staging -> caller for pageable result
// buffers, and the used part of every CIGAR row scattered into the caller's cigar_stride rows.
void scatter_results_part(abg_mapper *m, const ResultDst &d, abg_results *r, uint32_t c0, uint32_t n);

// The CIGAR rows are the one piece of host work per read inside abg_map_batch (a 64-byte line written per read):
// sub-batches of some size are split over a few threads, so that the part of it that cannot overlap the GPU (the
// last sub-batch) stays short.
void scatter_results(abg_mapper *m, const ResultDst &d, abg_results *r, uint32_t c0, uint32_t n) {
  const unsigned hw = std::thread::hardware_concurrency();
  const uint32_t n_thr = n >= 16384u ? std::min<uint32_t>(4u, std::max(1u, hw / 4u)) : 1u;
  if (n_thr <= 1u) {
    scatter_results_part(m, d, r, c0, n);
    return;
  }
  std::vector<std::thread> th;
  const uint32_t per = (n + n_thr - 1) / n_thr;
  for (uint32_t t = 1; t < n_thr; ++t) {
    const uint32_t a = std::min(n, t * per), b = std::min(n, a + per);
    if (b > a) th.emplace_back([=, &d] { scatter_results_part(m, d, r, c0 + a, b - a); });
  }
  scatter_results_part(m, d, r, c0, std::min(n, per));
  for (std::thread &x : th) x.join();
}

void scatter_results_part(abg_mapper *m, const ResultDst &d, abg_results *r, uint32_t c0, uint32_t n) {
  const uint32_t stride = m->params.cigar_stride;
  const uint32_t w = std::min(kInlineOps, stride);
  const int n_ends = m->paired ? 2 : 1;
  abg_hit *const u_se[2] = {r->se1, r->se2};
  uint32_t *const u_cig[2] = {r->cigar1, r->cigar2};
  uint32_t *const u_ncig[2] = {r->n_cigar1, r->n_cigar2};
  for (int e = 0; e < n_ends; ++e) {
    if (!d.direct) {
      std::memcpy(u_se[e] + c0, d.se[e] + c0, (size_t)n * sizeof(abg_hit));
      std::memcpy(u_ncig[e] + c0, d.ncig[e] + c0, (size_t)n * 4);
    }
    if (u_cig[e]) {
      const uint32_t *nc = d.ncig[e] + c0;
      const uint32_t *src = m->h_cigar[e] + (size_t)c0 * w;
      uint32_t *dst = u_cig[e] + (size_t)c0 * stride;
      for (uint32_t i = 0; i < n; ++i, src += w, dst += stride) {
        const uint32_t k = std::min(nc[i], w);
        for (uint32_t t = 0; t < k; ++t) dst[t] = src[t];
      }
    }
  }
  if (m->paired && !d.direct) {
    std::memcpy(r->pe_r1 + c0, d.pe_r1 + c0, (size_t)n * sizeof(abg_hit));
    std::memcpy(r->pe_r2 + c0, d.pe_r2 + c0, (size_t)n * sizeof(abg_hit));
  }
}

// After everything has landed: error flag, CIGARs longer than kInlineOps (rare, fetched one by one).
int finish_results(abg_mapper *m, const ResultDst &d, abg_results *r, uint32_t n) {
  if (m->h_flags[0] != 0u)
    return fail(ABG_ERR_CIGAR_OVERFLOW, "abg_map_batch: a reported CIGAR needed more than cigar_stride operations");
  const uint32_t stride = m->params.cigar_stride;
  const uint32_t w = std::min(kInlineOps, stride);
  uint32_t *const u_cig[2] = {r->cigar1, r->cigar2};
  for (int e = 0; e < (m->paired ? 2 : 1); ++e) {
    if (!u_cig[e]) continue;
    const uint32_t *nc = d.ncig[e];
    size_t n_long = 0;
    for (uint32_t i = 0; i < n; ++i) n_long += nc[i] > w;
    if (n_long == 0) continue;
    if (n_long <= 256) {
      for (uint32_t i = 0; i < n; ++i)
        if (nc[i] > w)  // (a hit that is not reported may have a longer CIGAR than its row: the row is what exists)
          ABG_CUDA(cudaMemcpy(u_cig[e] + (size_t)i * stride, m->d_cigar[e] + (size_t)i * stride,
                              (size_t)std::min(nc[i], stride) * 4, cudaMemcpyDeviceToHost));
    }
    else {  // many long CIGARs (long reads, high indel rates): one bulk copy of the full-stride rows
      std::vector<uint32_t> tmp((size_t)n * stride);
      ABG_CUDA(cudaMemcpy(tmp.data(), m->d_cigar[e], tmp.size() * 4, cudaMemcpyDeviceToHost));
      for (uint32_t i = 0; i < n; ++i)
        if (nc[i] > w)
          std::memcpy(u_cig[e] + (size_t)i * stride, tmp.data() + (size_t)i * stride, (size_t)std::min(nc[i], stride) * 4);
    }
  }
  return ABG_OK;
}

}  // namespace

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
  if (v->window_size != 0u && v->window_size != 12u && v->window_size != 20u)
    return fail(ABG_ERR_INVALID, "abg_index_create: window_size must be 20, or 12 (--enable-short)");
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
  // ABISMAL_B200_VERBOSE=1: where the load time goes (stderr)
  const bool verbose = std::getenv("ABISMAL_B200_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    (void)cudaDeviceSynchronize();
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[abg_index_create] %s: %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
    t_last = t;
  };
  (void)cudaFree(nullptr);
  lap("context");
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
  lap("arrays of the index file to HBM");
  if ((rc = make_bitmap(ix->counter, v->counter_size, &ix->bits)) ||
      (rc = make_bitmap(ix->counter_t, v->counter_size_three, &ix->bits_t)) ||
      (rc = make_bitmap(ix->counter_a, v->counter_size_three, &ix->bits_a))) {
    abg_index_destroy(ix);
    return rc;
  }
  {  // 2-bit genome copy + exception bitmap for the candidate compare
    const uint64_t n2 = (v->genome_words + 1) / 2 + 8;  // spare zero words: look-ahead of the last windows
    const uint64_t nx = (n2 * 32 / 256 + 31) / 32 + 2;
    uint32_t *d_gi = nullptr;  // blocks holding IUPAC codes; needed only while the records are built
    cudaError_t e1 = cudaMalloc(reinterpret_cast<void **>(&ix->g2), n2 * 8);
    cudaError_t e2 = cudaMalloc(reinterpret_cast<void **>(&ix->gx), nx * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc(reinterpret_cast<void **>(&d_gi), nx * 4);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      cudaFree(d_gi);
      abg_index_destroy(ix);
      return fail(ABG_ERR_CUDA, "abg_index_create: out of device memory for the 2-bit genome");
    }
    cudaMemset(ix->gx, 0, nx * 4);
    cudaMemset(d_gi, 0, nx * 4);
    unsigned int *d_iupac = nullptr, h_iupac = 1;
    if (cudaMalloc(reinterpret_cast<void **>(&d_iupac), 4) != cudaSuccess) {
      cudaFree(d_gi);
      abg_index_destroy(ix);
      return fail(ABG_ERR_CUDA, "abg_index_create: out of device memory");
    }
    cudaMemset(d_iupac, 0, 4);
    ab2dev::pack_genome2_kernel<<<prep_grid(8), 256>>>(ix->genome, v->genome_words + 4, n2, ix->g2, ix->gx, d_gi, d_iupac);
    cudaError_t e3 = cudaDeviceSynchronize();
    if (e3 == cudaSuccess) e3 = cudaMemcpy(&h_iupac, d_iupac, 4, cudaMemcpyDeviceToHost);
    cudaFree(d_iupac);
    if (e3 != cudaSuccess) {
      cudaFree(d_gi);
      abg_index_destroy(ix);
      return fail(ABG_ERR_CUDA, std::string("pack_genome2_kernel: ") + cudaGetErrorString(e3));
    }
    ix->dev.g2 = ix->g2;
    ix->dev.gx = ix->gx;
    ix->bytes_extra = n2 * 8 + nx * 4;
    ix->has_iupac = h_iupac != 0;
    // Seed-context records: 4 x 32 bytes per index entry (21 GB at 3.1 Gbp -- HBM is 180 GB).  They are an
    // accelerator of the compare, not a different algorithm: without them (not enough memory,
    // ABISMAL_B200_CTX=0) every candidate takes the direct index + genome gather.  In a genome with IUPAC
    // codes only the entries whose compare window can reach one take that route (sentinel records).
    const char *env = std::getenv("ABISMAL_B200_CTX");
    if (!(env && env[0] == '0')) {
      struct Tab {
        const uint32_t *index;
        uint64_t n;
        uint4 **dst;
      } tabs[3] = {{ix->index, v->index_size, &ix->ctx},
                   {ix->index_t, v->index_size_three, &ix->ctx_t},
                   {ix->index_a, v->index_size_three, &ix->ctx_a}};
      bool ok = true;
      for (const Tab &t : tabs) {
        if (t.n == 0) continue;
        const uint64_t bytes = t.n * (uint64_t)ab2dev::kCtxArrays * 32u;
        if (cudaMalloc(reinterpret_cast<void **>(t.dst), bytes) != cudaSuccess) {
          (void)cudaGetLastError();
          *t.dst = nullptr;
          ok = false;
          break;
        }
        ab2dev::seed_context_kernel<<<prep_grid(16), 256>>>(t.index, t.n, ix->g2, ix->has_iupac ? d_gi : nullptr,
                                                            n2 * 32 / 256 + 1, *t.dst);
        ix->bytes_extra += bytes;
      }
      const cudaError_t e4 = cudaDeviceSynchronize();
      if (e4 != cudaSuccess) {
        cudaFree(d_gi);
        abg_index_destroy(ix);
        return fail(ABG_ERR_CUDA, std::string("seed_context_kernel: ") + cudaGetErrorString(e4));
      }
      if (!ok) {  // all or nothing, so that the memory footprint is predictable
        for (const Tab &t : tabs) {
          if (*t.dst) ix->bytes_extra -= t.n * (uint64_t)ab2dev::kCtxArrays * 32u;
          cudaFree(*t.dst);
          *t.dst = nullptr;
        }
      }
      ix->dev.ctx = ix->ctx;
      ix->dev.ctx_t = ix->ctx_t;
      ix->dev.ctx_a = ix->ctx_a;
      ix->dev.n_ctx = v->index_size;
      ix->dev.n_ctx3 = v->index_size_three;
    }
    cudaFree(d_gi);
  }
  {  // compact two-letter counters, pinned in L2 by the mappers' streams (ABISMAL_B200_CC=0 disables)
    const char *env = std::getenv("ABISMAL_B200_CC");
    if (!(env && env[0] == '0')) {
      const uint64_t n_blocks = (v->counter_size + ab2dev::kCcKeys - 1) / ab2dev::kCcKeys;
      if (cudaMalloc(reinterpret_cast<void **>(&ix->cc), n_blocks * 32) == cudaSuccess) {
        ab2dev::compact_counter_kernel<<<prep_grid(8), 256>>>(ix->counter, v->counter_size, n_blocks, ix->cc);
        const cudaError_t e5 = cudaDeviceSynchronize();
        if (e5 != cudaSuccess) {
          abg_index_destroy(ix);
          return fail(ABG_ERR_CUDA, std::string("compact_counter_kernel: ") + cudaGetErrorString(e5));
        }
        ix->cc_bytes = n_blocks * 32;
        ix->bytes_extra += ix->cc_bytes;
        ix->dev.cc = ix->cc;
        // reserve as much persisting L2 as the device allows for it (a hint; failure only costs speed)
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
          const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)ix->cc_bytes);
          if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) (void)cudaGetLastError();
        }
      }
      else {
        (void)cudaGetLastError();
        ix->cc = nullptr;
      }
    }
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
  ix->dev.window_size = v->window_size ? v->window_size : 20u;
  lap("derived arrays (bitmaps, 2-bit genome, seed-context records, compact counters)");
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
  cudaFree(ix->g2);
  cudaFree(ix->gx);
  cudaFree(ix->cc);
  cudaFree(ix->ctx);
  cudaFree(ix->ctx_t);
  cudaFree(ix->ctx_a);
  delete ix;
}

uint64_t abg_index_device_bytes(const abg_index *ix) { return ix ? ix->bytes + ix->bytes_extra : 0; }

uint32_t abg_index_features(const abg_index *ix) {
  if (!ix) return 0u;
  uint32_t f = 0;
  // the two-letter table always has entries; the three-letter tables may be empty (then they need no records)
  if (ix->ctx != nullptr && (ix->dev.n_ctx3 == 0 || (ix->ctx_t != nullptr && ix->ctx_a != nullptr))) f |= ABG_FEATURE_SEED_CONTEXT;
  if (ix->cc != nullptr) f |= ABG_FEATURE_COMPACT_COUNTERS;
  if (ix->has_iupac) f |= ABG_FEATURE_GENOME_HAS_IUPAC;
  return f;
}

int abg_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(ABG_ERR_INVALID, "abg_host_alloc: null argument");
  *out = nullptr;
  ABG_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return ABG_OK;
}

void abg_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

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
  {
    // sub-batch size of the pipelined abg_map_batch (copies of chunk j+1 overlap the kernel of chunk j)
    const char *e = std::getenv("ABISMAL_B200_CHUNK");
    const long v = e ? std::atol(e) : 0;
    // measured through abg_map_batch at 2^20 pairs (profiles/r02_sweep_v19_e2e_chunks.txt): 32768 -> 102.8 ms,
    // 65536 -> 99.5, 98304 -> 98.5, 131072 -> 99.9 (fewer launch tails against a longer exposed last copy + scatter);
    // smaller batches keep at least four sub-batches of 32768
    uint32_t c = v > 0 ? (uint32_t)v : (max_batch >= 4u * 98304u ? 98304u : 32768u);
    const uint32_t min_c = (max_batch + kMaxChunks - 1) / kMaxChunks;
    m->chunk = std::max(c, std::max(min_c, 1u));
    // sub-batch size of the kernels after the filter (binned seeding); at least `chunk`
    const char *e2 = std::getenv("ABISMAL_B200_CHUNK2");
    m->chunk2 = (e2 && std::atol(e2) > 0) ? (uint32_t)std::atol(e2) : m->chunk;
  }

#define ABG_M(call)                                                                         \
  do {                                                                                      \
    const cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                                \
      abg_mapper_destroy(m);                                                                \
      return fail(ABG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    }                                                                                       \
  } while (0)

  ABG_M(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  ABG_M(cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking));
  ABG_M(cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking));
  ABG_M(cudaStreamCreateWithFlags(&m->s_run[0], cudaStreamNonBlocking));
  ABG_M(cudaStreamCreateWithFlags(&m->s_run[1], cudaStreamNonBlocking));
  if (ix->cc != nullptr) {
    // keep the compact counters resident in L2 while the kernels stream the rest of the index past them
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ix->device) == cudaSuccess && prop.accessPolicyMaxWindowSize > 0) {
      cudaStreamAttrValue av;
      std::memset(&av, 0, sizeof av);
      av.accessPolicyWindow.base_ptr = ix->cc;
      av.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)ix->cc_bytes, (size_t)prop.accessPolicyMaxWindowSize);
      const double avail = (double)prop.persistingL2CacheMaxSize;
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, avail / (double)av.accessPolicyWindow.num_bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      for (cudaStream_t st : {m->stream, m->s_run[0], m->s_run[1]})
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
    }
  }
  ABG_M(cudaEventCreate(&m->ev0));
  ABG_M(cudaEventCreate(&m->ev1));
  ABG_M(cudaEventCreate(&m->ev_ph[0]));
  ABG_M(cudaEventCreate(&m->ev_ph[1]));
  ABG_M(cudaEventCreate(&m->ev_t[0]));
  ABG_M(cudaEventCreate(&m->ev_t[1]));
  ABG_M(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
  for (uint32_t k = 0; k < kMaxChunks; ++k) {
    ABG_M(cudaEventCreateWithFlags(&m->ev_in[k], cudaEventDisableTiming));
    ABG_M(cudaEventCreateWithFlags(&m->ev_k[k], cudaEventDisableTiming));
    ABG_M(cudaEventCreateWithFlags(&m->ev_out[k], cudaEventDisableTiming));
  }

  // launch shape: persistent grid, as many CTAs per SM as shared memory/registers allow
  m->smem = ab2dev::block_smem_bytes(m->ml, m->paired, ab2dev::kLayoutFull);
  m->smem_s = ab2dev::block_smem_bytes(m->ml, m->paired, ab2dev::kLayoutSeed);
  m->smem_a = ab2dev::block_smem_bytes(m->ml, m->paired, ab2dev::kLayoutAlign);
  {
    // register-allocation variant (CTAs per SM the kernel is bounded for); ABISMAL_B200_MINB overrides for tuning
    const char *e = std::getenv("ABISMAL_B200_MINB");
    const int v = e ? std::atoi(e) : 0;
    m->minb = (v >= 2 && v <= 4) ? v : kDefaultMinB;
  }
  m->kernel = m->minb == 2 ? (const void *)ab2dev::map_reads_kernel<2>
            : m->minb == 4 ? (const void *)ab2dev::map_reads_kernel<4>
                           : (const void *)ab2dev::map_reads_kernel<3>;
  ABG_M(raise_smem_cap(ix->device, m->kernel, m->smem));
  int n_sm = 0, per_sm = 0;
  ABG_M(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ix->device));
  ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->kernel, ab2dev::kThreadsPerBlock, m->smem));
  if (per_sm < 1) {
    abg_mapper_destroy(m);
    return fail(ABG_ERR_CUDA, "abg_mapper_create: kernel does not fit on an SM");
  }
  m->grid = n_sm * per_sm;
  {
    // two-phase launch (default); ABISMAL_B200_SPLIT=0 keeps everything in map_reads_kernel
    const char *e = std::getenv("ABISMAL_B200_SPLIT");
    m->split = !(e && std::atoi(e) == 0);
  }
  int grid_max = m->grid;
  if (m->split) {
    int minb_s = m->minb == kDefaultMinB ? kDefaultMinBSeed : m->minb;
    {
      const char *e = std::getenv("ABISMAL_B200_MINB_SEED");  // tuning: the seeding kernel's own register bound
      const int v = e ? std::atoi(e) : 0;
      if (v >= 2 && v <= 6) minb_s = v;
    }
    m->kernel_s = minb_s == 2 ? (const void *)ab2dev::seed_kernel<2>
                : minb_s == 3 ? (const void *)ab2dev::seed_kernel<3>
                : minb_s == 5 ? (const void *)ab2dev::seed_kernel<5>
                : minb_s == 6 ? (const void *)ab2dev::seed_kernel<6>
                              : (const void *)ab2dev::seed_kernel<4>;
    int minb_a = m->minb;
    if (const char *e = std::getenv("ABISMAL_B200_MINB_ALIGN"))  // tuning: the selection kernel's own register bound
      if (std::atoi(e) >= 2 && std::atoi(e) <= 4) minb_a = std::atoi(e);
    m->kernel_a = minb_a == 2 ? (const void *)ab2dev::align_kernel<2>
                : minb_a == 4 ? (const void *)ab2dev::align_kernel<4>
                              : (const void *)ab2dev::align_kernel<3>;
    int per_s = 0, per_a = 0;
    // one shared-memory carve-out for every kernel of the path: CTAs of kernels that ask for different
    // carve-outs cannot share an SM (the overlapped launch co-schedules seed_kernel and align_kernel)
    for (const void *k : {m->kernel, m->kernel_s, m->kernel_a})
      if (cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        (void)cudaGetLastError();
    ABG_M(raise_smem_cap(ix->device, m->kernel_s, m->smem_s));
    ABG_M(raise_smem_cap(ix->device, m->kernel_a, m->smem_a));
    ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_s, m->kernel_s, ab2dev::kThreadsPerBlock, m->smem_s));
    ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_a, m->kernel_a, ab2dev::kThreadsPerBlock, m->smem_a));
    if (per_s < 1 || per_a < 1) {
      abg_mapper_destroy(m);
      return fail(ABG_ERR_CUDA, "abg_mapper_create: kernel does not fit on an SM");
    }
    m->grid_s = n_sm * per_s;
    m->grid_a = n_sm * per_a;
    {
      // overlapped launch: ABISMAL_B200_OVERLAP=1 (default: seed_kernel -> align_kernel back to back).
      // ABISMAL_B200_SEED_BLOCKS / ABISMAL_B200_AUX_BLOCKS: CTAs per SM of seed_kernel and of the align_kernel
      // launch that runs next to it (together they must fit on an SM: registers and shared memory).
      const char *e = std::getenv("ABISMAL_B200_OVERLAP");
      m->overlap = (e && std::atoi(e) != 0) && per_s >= 2;
      if (!m->overlap) {
        const char *es = std::getenv("ABISMAL_B200_SEED_BLOCKS");  // tuning: fewer seeding CTAs per SM
        if (es && std::atoi(es) >= 1) m->grid_s = n_sm * std::min(per_s, std::atoi(es));
      }
      if (m->overlap) {
        const char *es = std::getenv("ABISMAL_B200_SEED_BLOCKS"), *ea = std::getenv("ABISMAL_B200_AUX_BLOCKS");
        const int sb = std::max(1, std::min(per_s, es ? std::atoi(es) : per_s - 1));
        const int ab = std::max(1, std::min(per_a, ea ? std::atoi(ea) : 1));
        m->grid_s = n_sm * sb;
        m->grid_a_aux = n_sm * ab;
        const char *ew = std::getenv("ABISMAL_B200_WAIT_MS");
        const long wms = ew ? std::atol(ew) : 50;
        m->wait_ns = (uint32_t)std::min<long>(4000, std::max<long>(1, wms)) * 1000000u;
        ABG_M(cudaMalloc(&m->d_ready, (size_t)max_batch * sizeof(unsigned int)));
        for (int k = 0; k < 3; ++k) {
          ABG_M(cudaStreamCreateWithFlags(&m->s_aux[k], cudaStreamNonBlocking));
          ABG_M(cudaEventCreateWithFlags(&m->ev_go[k], cudaEventDisableTiming));
          ABG_M(cudaEventCreateWithFlags(&m->ev_aux[k], cudaEventDisableTiming));
        }
      }
    }
    {
      // task-parallel alignment (default for reads up to kDpMaxMl bases; ABISMAL_B200_TASKS=0 keeps every DP in
      // the warp of its pair)
      const char *e = std::getenv("ABISMAL_B200_TASKS");
      m->use_tasks = !(e && std::atoi(e) == 0) && !m->overlap && m->ml <= ab2dev::kDpMaxMl;
    }
    if (m->use_tasks) {
      int minb_e = m->minb;
      if (const char *ee = std::getenv("ABISMAL_B200_MINB_ENUM"))  // tuning: enum_kernel's own register bound
        if (std::atoi(ee) >= 2 && std::atoi(ee) <= 4) minb_e = std::atoi(ee);
      m->kernel_e = minb_e == 2 ? (const void *)ab2dev::enum_kernel<2>
                  : minb_e == 4 ? (const void *)ab2dev::enum_kernel<4>
                                : (const void *)ab2dev::enum_kernel<3>;
      m->smem_d = ab2dev::dp_block_smem_bytes(m->ml);
      int per_e = 0, per_d = 0;
      for (const void *k : {m->kernel_e, (const void *)ab2dev::dp_kernel})
        if (cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
          (void)cudaGetLastError();
      ABG_M(raise_smem_cap(ix->device, m->kernel_e, m->smem_a));
      ABG_M(raise_smem_cap(ix->device, (const void *)ab2dev::dp_kernel, m->smem_d));
      ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_e, m->kernel_e, ab2dev::kThreadsPerBlock, m->smem_a));
      ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_d, (const void *)ab2dev::dp_kernel, ab2dev::kThreadsPerBlock, m->smem_d));
      if (per_e < 1 || per_d < 1) m->use_tasks = false;
      m->grid_e = n_sm * std::max(per_e, 1);
      m->grid_d = n_sm * std::max(per_d, 1);
    }
    {
      // binned seeding (default whenever the index carries seed-context records; ABISMAL_B200_BINS=0: one warp per
      // strand gathers its own records, the round-1 seeding)
      const char *e = std::getenv("ABISMAL_B200_BINS");
      const bool have_ctx = ix->ctx != nullptr && (ix->dev.n_ctx3 == 0 || (ix->ctx_t != nullptr && ix->ctx_a != nullptr));
      m->use_bins = !(e && std::atoi(e) == 0) && !m->overlap && have_ctx && m->ml / 32u + 5u <= 32u;  // (pw <= 32)  // a plane of the read fits one word per lane
    }
    if (m->use_bins) {
      const bool rp = (p->mode & ABG_MODE_RANDOM_PBAT) != 0;
      m->spi = m->paired ? (rp ? 8u : 4u) : (rp ? 4u : 2u);
      // {lo, hi} words per strand: the whole read (replay's match masks) and the five words from q0 / 32 that a
      // tuple's 128 compared bases span (q0 = offset - 96 for offsets >= 128, offset <= length - 25); 40 bytes
      // per strand for 150-base reads -- the filter gathers them at random, the smaller the array the better
      m->pw = std::max(m->ml / 32u, (max_read_len > 152u ? ((max_read_len - 121u) >> 5) : 0u) + 5u);
      const uint64_t n_strands = (uint64_t)max_batch * m->spi;
      const char *ef = std::getenv("ABISMAL_B200_TUPLE_FACTOR");
      const double factor = (ef && std::atof(ef) > 0.0) ? std::atof(ef) : 1.25;
      const uint64_t warps = (uint64_t)n_sm * 8 * ab2dev::kWarpsPerBlock;
      uint64_t cap = (uint64_t)((double)n_strands * (double)(m->ml - 8u) * factor) + warps * ab2dev::kTupleBlock;
      if (const char *ec = std::getenv("ABISMAL_B200_TUPLE_CAP"))  // testing aid: absolute capacity
        if (std::atol(ec) > 0) cap = (uint64_t)std::atol(ec);
      cap = std::min<uint64_t>(cap, 0xfff00000ull) / ab2dev::kTupleBlock * ab2dev::kTupleBlock;
      cap = std::max<uint64_t>(cap, ab2dev::kTupleBlock);
      m->tup_cap = (uint32_t)cap;
      const uint64_t total_rec = (uint64_t)ab2dev::kCtxArrays * (ix->dev.n_ctx + 2 * ix->dev.n_ctx3);
      const char *es = std::getenv("ABISMAL_B200_BIN_SHIFT");
      uint32_t shift = (es && std::atoi(es) >= 10 && std::atoi(es) <= 30) ? (uint32_t)std::atoi(es) : 20u;  // 32 MB of records per bin (the filter does not care between 16 and 32 MB, the scatter prefers fewer bins: 5.7 vs 7.4 ms)
      while ((total_rec >> shift) + 1 > ab2dev::kMaxBins) ++shift;
      m->bin_shift = shift;
      m->n_bins = (uint32_t)(total_rec >> shift) + 1u;
      if (const char *eg = std::getenv("ABISMAL_B200_FILTER_GRAB"))  // tuning: 32 .. 4096 tuples, 0 = static distribution
        if (std::atoi(eg) == 0 || (std::atoi(eg) >= 32 && std::atoi(eg) <= 4096)) m->filter_grab = (uint32_t)std::atoi(eg) / 32u * 32u;
      m->surv_cap = ab2dev::kSurvSlots;
      if (const char *ea = std::getenv("ABISMAL_B200_ACC_CAP"))  // testing aid: 0 sends every strand through the general replay
        if (std::atoi(ea) >= 0 && std::atoi(ea) <= 32) m->acc_cap = (uint32_t)std::atoi(ea);
      if (const char *ec = std::getenv("ABISMAL_B200_SURV_CAP"))  // testing aid: fewer listed survivors per strand
        if (std::atol(ec) > 0) m->surv_cap = std::min<uint32_t>(ab2dev::kSurvSlots, (uint32_t)std::atol(ec));
      m->kernel_h = (const void *)ab2dev::hash_kernel<4>;
      if (cudaFuncSetAttribute(m->kernel_h, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        (void)cudaGetLastError();
      ABG_M(raise_smem_cap(ix->device, m->kernel_h, m->smem_s));
      int per_h = 0, per_f = 0;
      ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_h, m->kernel_h, ab2dev::kThreadsPerBlock, m->smem_s));
      if (const char *ep = std::getenv("ABISMAL_B200_FILTER_PIPE")) m->filter_pipe = std::atoi(ep) != 0;
      if (const char *ek = std::getenv("ABISMAL_B200_FILTER_CURSORS"))
        if (std::atoi(ek) >= 1 && std::atoi(ek) <= (int)kMaxFilterCursors) m->filter_cursors = (uint32_t)std::atoi(ek);
      if (const char *ec2 = std::getenv("ABISMAL_B200_FILTER_CACHE")) m->filter_cache = (uint32_t)std::atoi(ec2) & 3u;
      if (const char *e8 = std::getenv("ABISMAL_B200_FILTER_MINB")) m->filter_minb8 = std::atoi(e8) == 8 && !m->filter_pipe;
      if (m->filter_minb8) ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_f, (const void *)ab2dev::filter_kernel<false, false, 8>, 256, 0));
      else
      ABG_M(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_f, m->filter_pipe ? (const void *)ab2dev::filter_kernel<true, false>
                                                                                   : (const void *)ab2dev::filter_kernel<false, false>, 256, 0));
      // any allocation that fails switches the binned path off (the direct path needs none of this memory)
      bool ok = per_h >= 1 && per_f >= 1;
      auto grab = [&](void **ptr, size_t bytes) {
        if (ok && cudaMalloc(ptr, bytes) != cudaSuccess) {
          (void)cudaGetLastError();
          ok = false;
        }
      };
      grab((void **)&m->d_tup, cap * sizeof(ab2dev::SeedTuple));
      grab((void **)&m->d_tup_b, cap * sizeof(ab2dev::SeedTuple));
      grab((void **)&m->d_planes, n_strands * 2 * m->pw * 4);
      grab((void **)&m->d_strand_flag, n_strands);
      grab((void **)&m->d_bin_hist, (size_t)m->n_bins * n_sm * kScatterCtasPerSmMax * 4);
      grab((void **)&m->d_surv_count, n_strands * 4);
      grab((void **)&m->d_surv, n_strands * m->surv_cap * sizeof(uint2));
      grab((void **)&m->d_bin_work, kBinWorkWords * sizeof(unsigned int));
      if (!ok) {
        for (void *q : {(void *)m->d_tup, (void *)m->d_tup_b, (void *)m->d_planes, (void *)m->d_strand_flag,
                        (void *)m->d_bin_hist, (void *)m->d_surv_count, (void *)m->d_surv, (void *)m->d_bin_work})
          cudaFree(q);
        m->d_tup = m->d_tup_b = nullptr;
        m->d_planes = m->d_bin_hist = m->d_surv_count = nullptr;
        m->d_strand_flag = nullptr;
        m->d_surv = nullptr;
        m->d_bin_work = nullptr;
        m->use_bins = false;
      }
      else {
        m->grid_h = n_sm * per_h;
        // 1024-thread CTAs; every CTA has its own write range in every bin, and the scatter slows down with the
        // number of open ranges (measured: 2 CTAs per SM 32 ms, 1 CTA per SM 27 ms with the payloads still written here)
        m->grid_sc = n_sm;
        if (const char *es2 = std::getenv("ABISMAL_B200_SCATTER_SORT")) m->scatter_sorted = std::atoi(es2) != 0;
        if (m->scatter_sorted) {
          const size_t need = ab2dev::sort_scatter_smem(m->n_bins);
          if (m->n_bins > ab2dev::kSortMaxBins ||
              cudaFuncSetAttribute((const void *)ab2dev::scatter_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) != cudaSuccess) {
            (void)cudaGetLastError();
            m->scatter_sorted = false;
          }
        }
        if (const char *ec = std::getenv("ABISMAL_B200_SCATTER_CTAS"))
          if (std::atoi(ec) >= 1 && std::atoi(ec) <= kScatterCtasPerSmMax) m->grid_sc = n_sm * std::atoi(ec);
        if (m->scatter_sorted) m->grid_sc = n_sm;
        // fewer resident filter warps = a narrower live window of the record array (every warp holds `grab` tuples of
        // the bin-ordered stream): ABISMAL_B200_FILTER_CTAS caps the CTAs per SM
        if (const char *ef2 = std::getenv("ABISMAL_B200_FILTER_CTAS"))
          if (std::atoi(ef2) >= 1) per_f = std::min(per_f, std::atoi(ef2));
        m->grid_f = n_sm * per_f;
        for (int k = 0; k < 3; ++k) ABG_M(cudaEventCreate(&m->ev_b[k]));
        ABG_M(cudaEventCreateWithFlags(&m->ev_bins, cudaEventDisableTiming));
        grid_max = std::max(grid_max, m->grid_h);
      }
    }
    grid_max = std::max(grid_max, std::max(std::max(m->grid_s + m->grid_a_aux, m->grid_a), m->grid_e));
    const bool rpbat = (p->mode & ABG_MODE_RANDOM_PBAT) != 0;
    m->n_pass = m->paired ? (rpbat ? 8u : 4u) : 1u;
    m->set_slots = m->paired ? ab2dev::kSetSlotsPe : ab2dev::kSetSlotsSe;
    ABG_M(cudaMalloc(&m->d_sets, (size_t)max_batch * m->n_pass * (ab2dev::kSetStateWords + m->set_slots) * sizeof(uint64_t)));
    ABG_M(cudaMalloc(&m->d_redo_flag, (size_t)max_batch * sizeof(unsigned int)));
    ABG_M(cudaMalloc(&m->d_redo_list, (size_t)max_batch * sizeof(uint32_t)));
    if (m->paired) {
      // ABISMAL_B200_OVF_PER_ITEM: arena entries per pair (repeat-rich genomes want more: a set grows to 32 768 entries)
      const char *eo = std::getenv("ABISMAL_B200_OVF_PER_ITEM");
      const uint64_t per_item = (eo && std::atol(eo) > 0) ? (uint64_t)std::atol(eo) : kOvfPerItem;
      const uint64_t cap = std::min<uint64_t>((uint64_t)max_batch * per_item + (1u << 20), 0x7fffffffull);
      m->ovf_cap = (uint32_t)cap;
      ABG_M(cudaMalloc(&m->d_set_ovf, cap * sizeof(uint64_t)));
      ABG_M(cudaMalloc(&m->d_task_ovf, cap * sizeof(uint32_t)));
    }
    if (m->use_tasks) {
      // every sub-batch owns the part of the arenas that belongs to its items plus one slack region (task slots
      // and traceback words are handed to the warps of enum_kernel in blocks; a warp's last block stays part empty)
      const uint32_t chunk_now = std::max(m->chunk, (max_batch + kMaxChunks - 1) / kMaxChunks);
      m->n_chunks_max = std::max(1u, (max_batch + chunk_now - 1) / chunk_now);
      const uint32_t warps = (uint32_t)m->grid_e * ab2dev::kWarpsPerBlock;
      m->task_slack = warps * 32u;  // a warp's last block holds at most 31 unused slots (one task per lane at a time)
      m->tb_slack = warps * ab2dev::kTbGrabTasks * ab2dev::tb_sm_words(m->ml);
      const uint32_t *cap = m->paired ? kTaskCapPe : kTaskCapSe;
      uint64_t off = 0;
      // ABISMAL_B200_TASK_SCALE: multiplier on the task-list and traceback capacities per read / pair
      const char *et = std::getenv("ABISMAL_B200_TASK_SCALE");
      const uint32_t tscale = (et && std::atol(et) > 0) ? (uint32_t)std::atol(et) : 1u;
      for (int c = 0; c < 3; ++c) {
        m->task_cap_item[c] = cap[c] * (rpbat ? 2u : 1u) * tscale;
        m->task_class_off[c] = off;
        off += (uint64_t)max_batch * m->task_cap_item[c] + (uint64_t)m->n_chunks_max * m->task_slack;
      }
      m->tb_cap_item = (m->paired ? kTbTasksPe : kTbTasksSe) * (rpbat ? 2u : 1u) * ab2dev::tb_sm_words(m->ml);
      const uint64_t tb_units = (uint64_t)max_batch * m->tb_cap_item + (uint64_t)m->n_chunks_max * m->tb_slack;
      if (off >= 0xffffffffull || tb_units >= 0xffffffffull) m->use_tasks = false;  // ids are 32 bits
      else {
        ABG_M(cudaMalloc(&m->d_tasks, off * sizeof(ab2dev::AlignTask)));
        ABG_M(cudaMalloc(&m->d_task_res, off * sizeof(ab2dev::TaskResult)));
        ABG_M(cudaMalloc(&m->d_task_tb, tb_units * 64u));
        ABG_M(cudaMalloc(&m->d_task_of, (size_t)max_batch * m->n_pass * m->set_slots * sizeof(uint32_t)));
      }
    }
  }
  // two scratch sets: consecutive chunks run on alternating streams and may overlap at their tails
  m->grid_scratch = grid_max;
  const size_t slots = (size_t)grid_max * ab2dev::kWarpsPerBlock * 2;

  m->seq_cap = (size_t)max_batch * max_read_len;
  const uint32_t w = std::min(kInlineOps, stride);
  for (int e = 0; e < n_ends; ++e) {
    ABG_M(cudaMalloc(&m->d_seq[e], m->seq_cap + 16));
    ABG_M(cudaMalloc(&m->d_off[e], ((size_t)max_batch + 1) * 4));
    ABG_M(cudaMalloc(&m->d_se[e], (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_cigar[e], (size_t)max_batch * stride * 4));
    ABG_M(cudaMalloc(&m->d_ncigar[e], (size_t)max_batch * 4));
    ABG_M(cudaMalloc(&m->d_cigar_inline[e], (size_t)max_batch * w * 4));
    // (h_seq, the page-locked staging of pageable read bytes, is allocated when a caller first passes pageable memory:
    //  page-locking max_batch x max_read_len bytes per end costs the front end's start-up more than anything else here)
    ABG_M(cudaMallocHost(&m->h_off[e], ((size_t)max_batch + 1) * 4));
    ABG_M(cudaMallocHost(&m->h_se[e], (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_cigar[e], (size_t)max_batch * w * 4));
    ABG_M(cudaMallocHost(&m->h_ncigar[e], (size_t)max_batch * 4));
  }
  if (m->paired) {
    ABG_M(cudaMalloc(&m->d_pe_r1, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_pe_r2, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_pe_r1, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMallocHost(&m->h_pe_r2, (size_t)max_batch * sizeof(abg_hit)));
    ABG_M(cudaMalloc(&m->d_pe_overflow, slots * 2 * ab2dev::kPeLarge * sizeof(uint64_t)));
    ABG_M(cudaMalloc(&m->d_mem_scr, slots * ab2dev::kPeLarge * sizeof(int16_t)));
    ABG_M(cudaMalloc(&m->d_mem_scr2, slots * ab2dev::kPeLarge * sizeof(int16_t)));
    if (const char *eh = std::getenv("ABISMAL_B200_HEAVY_MIN"))  // set size from which best_pair goes by rows (0: always)
      m->heavy_min = (uint32_t)std::max(0l, std::atol(eh));
  }
  m->tb_words = ab2dev::tb_sm_words(m->ml);
  ABG_M(cudaMalloc(&m->d_tb, slots * 2 * m->tb_words * 32 * sizeof(uint64_t)));
  ABG_M(cudaMalloc(&m->d_work, kWorkWords * sizeof(unsigned int)));
  ABG_M(cudaMallocHost(&m->h_flags, 2 * sizeof(unsigned int)));
  if (m->count_work) ABG_M(cudaMalloc(&m->d_counters, 6 * sizeof(unsigned long long)));
#undef ABG_M
  *out = m;
  return ABG_OK;
}

void abg_mapper_destroy(abg_mapper *m) {
  if (!m) return;
  cudaSetDevice(m->idx->device);
  cudaDeviceSynchronize();
  for (int e = 0; e < 2; ++e) {
    cudaFree(m->d_seq[e]);
    cudaFree(m->d_off[e]);
    cudaFree(m->d_se[e]);
    cudaFree(m->d_cigar[e]);
    cudaFree(m->d_ncigar[e]);
    cudaFree(m->d_cigar_inline[e]);
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
  cudaFree(m->d_mem_scr2);
  cudaFree(m->d_tb);
  cudaFree(m->d_work);
  cudaFree(m->d_ready);
  for (int k = 0; k < 3; ++k) {
    if (m->s_aux[k]) cudaStreamDestroy(m->s_aux[k]);
    if (m->ev_go[k]) cudaEventDestroy(m->ev_go[k]);
    if (m->ev_aux[k]) cudaEventDestroy(m->ev_aux[k]);
  }
  cudaFree(m->d_counters);
  cudaFree(m->d_tup);
  cudaFree(m->d_tup_b);
  cudaFree(m->d_planes);
  cudaFree(m->d_strand_flag);
  cudaFree(m->d_bin_hist);
  cudaFree(m->d_surv_count);
  cudaFree(m->d_surv);
  cudaFree(m->d_bin_work);
  for (cudaEvent_t e : m->ev_b)
    if (e) cudaEventDestroy(e);
  if (m->ev_bins) cudaEventDestroy(m->ev_bins);
  cudaFree(m->d_sets);
  cudaFree(m->d_redo_flag);
  cudaFree(m->d_redo_list);
  cudaFree(m->d_tasks);
  cudaFree(m->d_task_res);
  cudaFree(m->d_task_tb);
  cudaFree(m->d_task_of);
  cudaFree(m->d_set_ovf);
  cudaFree(m->d_task_ovf);
  for (cudaEvent_t e : m->ev_t)
    if (e) cudaEventDestroy(e);
  cudaFreeHost(m->h_flags);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (cudaEvent_t e : m->ev_ph)
    if (e) cudaEventDestroy(e);
  if (m->ev_start) cudaEventDestroy(m->ev_start);
  for (uint32_t k = 0; k < kMaxChunks; ++k) {
    if (m->ev_in[k]) cudaEventDestroy(m->ev_in[k]);
    if (m->ev_k[k]) cudaEventDestroy(m->ev_k[k]);
    if (m->ev_out[k]) cudaEventDestroy(m->ev_out[k]);
  }
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->s_h2d) cudaStreamDestroy(m->s_h2d);
  if (m->s_d2h) cudaStreamDestroy(m->s_d2h);
  if (m->s_run[0]) cudaStreamDestroy(m->s_run[0]);
  if (m->s_run[1]) cudaStreamDestroy(m->s_run[1]);
  delete m;
}

namespace {

int check_batch(const abg_mapper *m, const abg_batch *b, const char *who) {
  if (!m || !b || !b->seq1 || !b->off1) return fail(ABG_ERR_INVALID, std::string(who) + ": null argument");
  if (b->n > m->max_batch) return fail(ABG_ERR_INVALID, std::string(who) + ": batch larger than max_batch");
  if (m->paired && (!b->seq2 || !b->off2)) return fail(ABG_ERR_INVALID, std::string(who) + ": paired mode needs two ends");
  const uint32_t *offs[2] = {b->off1, b->off2};
  for (int e = 0; e < (m->paired ? 2 : 1); ++e)
    if ((size_t)(offs[e][b->n] - offs[e][0]) > m->seq_cap)
      return fail(ABG_ERR_TOO_LONG, std::string(who) + ": batch sequence bytes exceed capacity");
  return ABG_OK;
}

}  // namespace

int abg_mapper_upload(abg_mapper *m, const abg_batch *b) {
  int rc;
  if ((rc = check_batch(m, b, "abg_mapper_upload")) != ABG_OK) return rc;
  ABG_CUDA(cudaSetDevice(m->idx->device));
  const int n_ends = m->paired ? 2 : 1;
  const char *seqs[2] = {b->seq1, b->seq2};
  const uint32_t *offs[2] = {b->off1, b->off2};
  for (int e = 0; e < n_ends; ++e) {
    const uint32_t *off = offs[e];
    const size_t bytes = off[b->n] - off[0];
    if ((rc = stage_offsets(m, off, 0, b->n, m->h_off[e])) != ABG_OK) return rc;
    const char *src = seqs[e] + off[0];
    if (!is_pinned(src)) {
      if (!m->h_seq[e]) ABG_CUDA(cudaMallocHost(&m->h_seq[e], m->seq_cap + 16));
      std::memcpy(m->h_seq[e], src, bytes);
      src = m->h_seq[e];
    }
    ABG_CUDA(cudaMemcpyAsync(m->d_seq[e], src, bytes, cudaMemcpyHostToDevice, m->stream));
    ABG_CUDA(cudaMemcpyAsync(m->d_off[e], m->h_off[e], ((size_t)b->n + 1) * 4, cudaMemcpyHostToDevice, m->stream));
  }
  m->cur_n = b->n;
  return ABG_OK;
}

int abg_mapper_run(abg_mapper *m) {
  if (!m) return fail(ABG_ERR_INVALID, "abg_mapper_run: null mapper");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  ab2dev::KernelParams P;
  fill_params(m, P, 0, m->cur_n, 0);
  ABG_CUDA(cudaMemsetAsync(m->d_work, 0, (2 + kChunkWords) * sizeof(unsigned int), m->stream));
  if (m->d_counters) ABG_CUDA(cudaMemsetAsync(m->d_counters, 0, 6 * sizeof(unsigned long long), m->stream));
  ABG_CUDA(cudaEventRecord(m->ev0, m->stream));
  int rc;
  if ((rc = launch(m, P, m->stream, m->ev_ph)) != ABG_OK) return rc;
  ABG_CUDA(cudaEventRecord(m->ev1, m->stream));
  // the kernels' error flag travels behind them (4 bytes, outside the timed events); abg_mapper_sync reports it
  ABG_CUDA(cudaMemcpyAsync(m->h_flags + 1, m->d_work, sizeof(unsigned int), cudaMemcpyDeviceToHost, m->stream));
  m->timed = true;
  m->run_pending = true;
  return ABG_OK;
}

namespace {
void read_times(abg_mapper *m) {
  m->phase_ms[0] = m->phase_ms[1] = m->phase_ms[2] = 0.f;
  if (!m->timed || cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1) != cudaSuccess) {
    m->last_ms = 0.f;
    (void)cudaGetLastError();  // do not leave a stale error behind
    return;
  }
  m->phase_ms[0] = m->last_ms;
  if (m->split && m->cur_n) {
    float a = 0.f, b = 0.f, c = 0.f;
    if (cudaEventElapsedTime(&a, m->ev0, m->ev_ph[0]) == cudaSuccess && cudaEventElapsedTime(&b, m->ev_ph[0], m->ev_ph[1]) == cudaSuccess &&
        cudaEventElapsedTime(&c, m->ev_ph[1], m->ev1) == cudaSuccess) {
      m->phase_ms[0] = a;
      m->phase_ms[1] = b;
      m->phase_ms[2] = c;
      m->task_ms[0] = m->task_ms[1] = 0.f;
      m->bin_ms[0] = m->bin_ms[1] = m->bin_ms[2] = 0.f;
      if (m->use_bins) {
        float h = 0.f, sc = 0.f, f = 0.f;
        if (cudaEventElapsedTime(&h, m->ev0, m->ev_b[0]) == cudaSuccess && cudaEventElapsedTime(&sc, m->ev_b[0], m->ev_b[1]) == cudaSuccess &&
            cudaEventElapsedTime(&f, m->ev_b[1], m->ev_b[2]) == cudaSuccess) {
          m->bin_ms[0] = h;
          m->bin_ms[1] = sc;
          m->bin_ms[2] = f;
        }
        else
          (void)cudaGetLastError();
      }
      if (m->use_tasks) {
        float e = 0.f, d = 0.f;
        if (cudaEventElapsedTime(&e, m->ev_ph[0], m->ev_t[0]) == cudaSuccess && cudaEventElapsedTime(&d, m->ev_t[0], m->ev_t[1]) == cudaSuccess) {
          m->task_ms[0] = e;
          m->task_ms[1] = d;
        }
        else
          (void)cudaGetLastError();
      }
    }
    else
      (void)cudaGetLastError();
  }
}
}  // namespace

int abg_mapper_sync(abg_mapper *m) {
  if (!m) return fail(ABG_ERR_INVALID, "abg_mapper_sync: null mapper");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  ABG_CUDA(cudaStreamSynchronize(m->stream));
  read_times(m);
  if (m->run_pending) {
    m->run_pending = false;
    if (m->h_flags[1] != 0u)
      return fail(ABG_ERR_CIGAR_OVERFLOW, "abg_mapper_run: a reported CIGAR needed more than cigar_stride operations");
  }
  return ABG_OK;
}

int abg_mapper_download(abg_mapper *m, abg_results *r) {
  if (!m || !r || !r->se1 || !r->n_cigar1) return fail(ABG_ERR_INVALID, "abg_mapper_download: null argument");
  if (m->paired && (!r->pe_r1 || !r->pe_r2 || !r->se2 || !r->n_cigar2))
    return fail(ABG_ERR_INVALID, "abg_mapper_download: paired results need pe_r1/pe_r2/se2/n_cigar2");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  const uint32_t n = m->cur_n;
  const ResultDst d = pick_result_dst(m, r);
  int rc;
  if ((rc = copy_results_async(m, d, 0, n, m->stream)) != ABG_OK) return rc;
  ABG_CUDA(cudaMemcpyAsync(m->h_flags, m->d_work, sizeof(unsigned int), cudaMemcpyDeviceToHost, m->stream));
  if (m->d_counters)
    ABG_CUDA(cudaMemcpyAsync(&m->counters, m->d_counters, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             m->stream));
  ABG_CUDA(cudaStreamSynchronize(m->stream));
  read_times(m);
  scatter_results(m, d, r, 0, n);
  return finish_results(m, d, r, n);
}

// Host buffers in, host buffers out.  The batch is cut into sub-batches that flow through three streams:
// H2D copy of chunk j+1 | kernel of chunk j | D2H copy of chunk j-1, so that only the first copy in and the
// last copy out are exposed.  Pinned caller buffers (abg_host_alloc) are used in place; pageable ones are
// staged through the mapper's pinned buffers chunk by chunk, which overlaps too.
namespace {
int map_batch_impl(abg_mapper *m, const abg_batch *b, abg_results *r);
}

int abg_map_batch(abg_mapper *m, const abg_batch *b, abg_results *r) {
  int rc;
  if ((rc = check_batch(m, b, "abg_map_batch")) != ABG_OK) return rc;
  if (!r || !r->se1 || !r->n_cigar1) return fail(ABG_ERR_INVALID, "abg_map_batch: null results");
  if (m->paired && (!r->pe_r1 || !r->pe_r2 || !r->se2 || !r->n_cigar2))
    return fail(ABG_ERR_INVALID, "abg_map_batch: paired results need pe_r1/pe_r2/se2/n_cigar2");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  rc = map_batch_impl(m, b, r);
  if (rc != ABG_OK) {
    // Nothing may still be in flight when the caller gets its buffers back: copies could be writing into them,
    // and the front end recycles them.  (The message of the first failure is kept.)
    const std::string msg = g_err;
    for (cudaStream_t st : {m->s_h2d, m->s_run[0], m->s_run[1], m->s_d2h, m->s_aux[1], m->s_aux[2]})
      if (st) (void)cudaStreamSynchronize(st);
    (void)cudaGetLastError();
    g_err = msg;
  }
  return rc;
}

namespace {
int map_batch_impl(abg_mapper *m, const abg_batch *b, abg_results *r) {
  int rc;
  const auto t_call = std::chrono::steady_clock::now();
  const uint32_t n = b->n;
  m->cur_n = n;
  m->timed = false;
  const int n_ends = m->paired ? 2 : 1;
  const char *seqs[2] = {b->seq1, b->seq2};
  const uint32_t *offs[2] = {b->off1, b->off2};
  bool seq_pinned[2] = {is_pinned(b->seq1), m->paired && is_pinned(b->seq2)};
  const ResultDst d = pick_result_dst(m, r);
  const uint32_t chunk = std::max(m->chunk, (n + kMaxChunks - 1) / kMaxChunks);
  const uint32_t n_chunks = n ? (n + chunk - 1) / chunk : 0;

  ABG_CUDA(cudaMemsetAsync(m->d_work, 0, kWorkWords * sizeof(unsigned int), m->s_h2d));
  if (m->d_counters) ABG_CUDA(cudaMemsetAsync(m->d_counters, 0, 6 * sizeof(unsigned long long), m->s_h2d));
  const size_t slot_sets = (size_t)m->grid_scratch * ab2dev::kWarpsPerBlock;
  // Binned seeding works on the whole batch (the more strands share a bin, the fewer records come from DRAM):
  // the sub-batches are hashed as they arrive (stream s_run[0]), binned and filtered together, and only then
  // flow through the per-sub-batch kernels, whose results are copied back as they finish.
  const bool bins = m->use_bins && m->split;
  if (bins && n_chunks != 0 && (rc = reset_bins(m, n, m->s_h2d)) != ABG_OK) return rc;
  for (uint32_t j = 0; j < n_chunks; ++j) {
    const uint32_t c0 = j * chunk, c1 = std::min(n, c0 + chunk);
    for (int e = 0; e < n_ends; ++e) {
      const uint32_t *off = offs[e];
      if ((rc = stage_offsets(m, off, c0, c1, m->h_off[e])) != ABG_OK) return rc;
      const size_t o0 = off[c0] - off[0], bytes = off[c1] - off[c0];
      const char *src = seqs[e] + off[c0];
      if (!seq_pinned[e]) {
        if (!m->h_seq[e]) ABG_CUDA(cudaMallocHost(&m->h_seq[e], m->seq_cap + 16));
        std::memcpy(m->h_seq[e] + o0, src, bytes);
        src = m->h_seq[e] + o0;
      }
      if (bytes) ABG_CUDA(cudaMemcpyAsync(m->d_seq[e] + o0, src, bytes, cudaMemcpyHostToDevice, m->s_h2d));
      ABG_CUDA(cudaMemcpyAsync(m->d_off[e] + c0, m->h_off[e] + c0, ((size_t)(c1 - c0) + 1) * 4, cudaMemcpyHostToDevice,
                               m->s_h2d));
    }
    ABG_CUDA(cudaEventRecord(m->ev_in[j], m->s_h2d));
    if (bins) {
      ABG_CUDA(cudaStreamWaitEvent(m->s_run[0], m->ev_in[j], 0));
      ab2dev::KernelParams P;
      fill_params(m, P, c0, c1 - c0, j);
      if ((rc = launch_hash(m, P, m->s_run[0])) != ABG_OK) return rc;
    }
  }
  if (bins && n_chunks != 0) {
    if ((rc = launch_bins(m, m->s_run[0])) != ABG_OK) return rc;
    ABG_CUDA(cudaEventRecord(m->ev_bins, m->s_run[0]));
    ABG_CUDA(cudaStreamWaitEvent(m->s_run[1], m->ev_bins, 0));
  }
  // Binned seeding: the kernels after the filter may take larger sub-batches than the copies and the hashing
  // (every launch has a tail; the copies back still overlap the next sub-batch's kernels)
  const uint32_t chunk2 = bins ? std::max(chunk, m->chunk2) : chunk;
  const uint32_t n_chunks2 = n ? (n + chunk2 - 1) / chunk2 : 0;
  for (uint32_t j = 0; j < n_chunks2; ++j) {
    const uint32_t c0 = j * chunk2, c1 = std::min(n, c0 + chunk2);
    cudaStream_t sr = m->s_run[j & 1];
    if (!bins) ABG_CUDA(cudaStreamWaitEvent(sr, m->ev_in[j], 0));  // (binned: everything was hashed, hence copied, before the bins)
    ab2dev::KernelParams P;
    fill_params(m, P, c0, c1 - c0, j);
    // scratch set of this stream
    if (j & 1) {
      if (P.pe_overflow) P.pe_overflow += slot_sets * 2 * ab2dev::kPeLarge;
      if (P.mem_scr) P.mem_scr += slot_sets * ab2dev::kPeLarge;
      if (P.mem_scr2) P.mem_scr2 += slot_sets * ab2dev::kPeLarge;
      P.tb += slot_sets * 2 * m->tb_words * 32;
    }
    if ((rc = launch(m, P, sr, nullptr, bins)) != ABG_OK) return rc;
    ABG_CUDA(cudaEventRecord(m->ev_k[j], sr));
    ABG_CUDA(cudaStreamWaitEvent(m->s_d2h, m->ev_k[j], 0));
    if ((rc = copy_results_async(m, d, c0, c1 - c0, m->s_d2h)) != ABG_OK) return rc;
    ABG_CUDA(cudaEventRecord(m->ev_out[j], m->s_d2h));
  }
  if (n_chunks == 0) ABG_CUDA(cudaStreamSynchronize(m->s_h2d));
  ABG_CUDA(cudaMemcpyAsync(m->h_flags, m->d_work, sizeof(unsigned int), cudaMemcpyDeviceToHost, m->s_d2h));
  if (m->d_counters)
    ABG_CUDA(cudaMemcpyAsync(&m->counters, m->d_counters, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             m->s_d2h));
  if (std::getenv("ABISMAL_B200_VERBOSE") != nullptr && n_chunks != 0) {
    // coarse timeline of the call (host clock at the moment each stage's event has fired; everything above is
    // enqueued by now)
    const auto t_enq = std::chrono::steady_clock::now();
    auto since = [&](const std::chrono::steady_clock::time_point &t0) {
      return 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    const double enq_ms = 1e3 * std::chrono::duration<double>(t_enq - t_call).count();
    (void)cudaEventSynchronize(m->ev_in[n_chunks - 1]);
    const double h2d_ms = since(t_call);
    double bins_ms = 0.0;
    if (bins) {
      (void)cudaEventSynchronize(m->ev_bins);
      bins_ms = since(t_call);
    }
    (void)cudaEventSynchronize(m->ev_k[n_chunks2 - 1]);
    const double kern_ms = since(t_call);
    (void)cudaEventSynchronize(m->ev_out[n_chunks2 - 1]);
    const double d2h_ms = since(t_call);
    std::fprintf(stderr, "[abg_map_batch] n %u, %u + %u sub-batches: enqueued %.2f ms, copies in done %.2f, hash+bins+filter done %.2f, "
                         "last kernel done %.2f, last copy out done %.2f (ms since the call)\n",
                 n, n_chunks, n_chunks2, enq_ms, h2d_ms, bins_ms, kern_ms, d2h_ms);
  }
  // host side of each sub-batch as it lands, while the GPU works on the later ones
  for (uint32_t j = 0; j < n_chunks2; ++j) {
    const uint32_t c0 = j * chunk2, c1 = std::min(n, c0 + chunk2);
    ABG_CUDA(cudaEventSynchronize(m->ev_out[j]));
    scatter_results(m, d, r, c0, c1 - c0);
  }
  ABG_CUDA(cudaStreamSynchronize(m->s_d2h));
  return finish_results(m, d, r, n);
}
}  // namespace

float abg_mapper_last_kernel_ms(const abg_mapper *m) { return m ? m->last_ms : 0.f; }
void abg_mapper_last_phase_ms(const abg_mapper *m, float out[3]) {
  for (int k = 0; k < 3; ++k) out[k] = m ? m->phase_ms[k] : 0.f;
}
void abg_mapper_last_kernel_times(const abg_mapper *m, float out[5]) {
  for (int k = 0; k < 5; ++k) out[k] = 0.f;
  if (!m) return;
  out[0] = m->phase_ms[0];
  out[1] = m->task_ms[0];
  out[2] = m->task_ms[1];
  out[3] = m->phase_ms[1] - m->task_ms[0] - m->task_ms[1];
  out[4] = m->phase_ms[2];
}
void abg_mapper_last_seed_times(const abg_mapper *m, float out[4]) {
  for (int k = 0; k < 4; ++k) out[k] = 0.f;
  if (!m) return;
  out[0] = m->bin_ms[0];
  out[1] = m->bin_ms[1];
  out[2] = m->bin_ms[2];
  out[3] = m->phase_ms[0] - m->bin_ms[0] - m->bin_ms[1] - m->bin_ms[2];
}
uint32_t abg_mapper_launches_per_run(const abg_mapper *m) {
  return (m && m->cur_n) ? (m->split ? (m->overlap ? 4u : (m->use_tasks ? 5u : 3u)) + (m->use_bins ? 4u : 0u) : 1u) : 0u;
}
int abg_mapper_binned(const abg_mapper *m) { return (m && m->use_bins) ? 1 : 0; }
int abg_mapper_bin_stats(abg_mapper *m, uint64_t out[8]) {
  if (!m || !out) return fail(ABG_ERR_INVALID, "abg_mapper_bin_stats: null argument");
  for (int k = 0; k < 8; ++k) out[k] = 0;
  if (!m->use_bins) return ABG_OK;
  ABG_CUDA(cudaSetDevice(m->idx->device));
  ABG_CUDA(cudaDeviceSynchronize());
  unsigned int w[4];
  ABG_CUDA(cudaMemcpy(w, m->d_bin_work, sizeof w, cudaMemcpyDeviceToHost));
  const size_t ns = (size_t)m->cur_n * m->spi;
  std::vector<uint8_t> flag(ns);
  std::vector<uint32_t> cnt(ns);
  if (ns) {
    ABG_CUDA(cudaMemcpy(flag.data(), m->d_strand_flag, ns, cudaMemcpyDeviceToHost));
    ABG_CUDA(cudaMemcpy(cnt.data(), m->d_surv_count, ns * 4, cudaMemcpyDeviceToHost));
  }
  uint64_t direct = 0, surv = 0;
  for (size_t k = 0; k < ns; ++k) {
    if (flag[k] == 2) continue;
    if (flag[k] == 1 || cnt[k] > m->surv_cap) ++direct;
    else surv += cnt[k];
  }
  out[0] = ns;          // strands of the last batch
  out[1] = direct;      // of them mapped by process_seeds itself
  out[2] = w[1];        // tuples binned
  out[3] = surv;        // prefilter survivors of the listed strands
  out[4] = m->n_bins;
  out[5] = m->tup_cap;
  out[6] = (m->scatter_sorted ? 1u : 0u) | (m->filter_pipe ? 2u : 0u);
  out[7] = m->filter_grab;
  return ABG_OK;
}
uint32_t abg_mapper_chunk(const abg_mapper *m) { return m ? m->chunk : 0u; }

int abg_mapper_last_run_stats(abg_mapper *m, uint32_t out[8]) {
  if (!m || !out) return fail(ABG_ERR_INVALID, "abg_mapper_last_run_stats: null argument");
  ABG_CUDA(cudaSetDevice(m->idx->device));
  unsigned int w[2 + kChunkWords];
  ABG_CUDA(cudaStreamSynchronize(m->stream));
  ABG_CUDA(cudaMemcpy(w, m->d_work, sizeof w, cudaMemcpyDeviceToHost));
  out[0] = w[2 + 3];
  out[1] = w[1];
  out[2] = m->ovf_cap;
  out[3] = w[2 + 5];
  out[4] = w[2 + 6];
  out[5] = w[2 + 7];
  out[6] = w[2 + 8];
  out[7] = w[0];
  return ABG_OK;
}

int abg_mapper_get_counters(const abg_mapper *m, abg_work_counters *out) {
  if (!m || !out) return fail(ABG_ERR_INVALID, "abg_mapper_get_counters: null argument");
  if (!m->count_work) return fail(ABG_ERR_INVALID, "abg_mapper_get_counters: mapper created with count_work == 0");
  *out = m->counters;
  return ABG_OK;
}

}  // extern "C"
