// seed_bins.cuh -- binned seeding: cross-read locality for the seed-context prefilter.
//
// process_seeds (abismal.cpp:1269-1375) examines, per read strand, ~1000 index entries at ~126 seed offsets;
// with one warp per strand every examined entry is a random 32-byte sector of the 20 GB record array, and a
// batch of 2^20 pairs fetches every record ~6.6 times from DRAM.  check_hits / full_compare (:1105-1150) is a
// PURE function of (strand, offset, entry); only the updates of the candidate set are order dependent, and no
// update can ever be accepted above the set's initial cutoff (0.4 x read length: the heaps start with that
// sentinel and cutoffs only tighten).  So the batch is seeded in four kernels:
//
//   hash_kernel     one warp per read / pair, its strands one after the other: encode, hash, probe the counters
//                   (L2 resident), narrow oversized buckets (find_candidates) -- everything of process_seeds up to
//                   the candidate ranges -- and emit one 16-byte TUPLE per (offset, table) with a non-empty range,
//                   plus the strand's 2-bit read planes;
//   count_kernel, bin_prefix_kernel, scatter_sorted_kernel
//                   tuples -> bins by the address of their first record (one bin = 2^bin_shift records, 32 MB),
//                   without global atomics: every CTA has its own write range in every bin, and sorts tiles of
//                   8192 tuples in shared memory so that a bin receives runs of consecutive tuples;
//   filter_kernel   tuples in bin order, 32 per warp at a time: the payload of each (the 128 read bases its records
//                   are compared with) is built in shared memory from the strand's planes, then one lane per
//                   candidate: record against payload, lower bound of the distance, survivors (~3 %) appended to
//                   their strand's list;
//   seed_kernel     (mapper_kernels.cuh, process_binned below) one warp per strand: the deep compare of its
//                   survivors; those a phase could still accept are put into canonical order (offset, two-letter
//                   before three-letter, bucket order) and replayed against the candidate set exactly as
//                   process_seeds orders them.  Strands outside the fast path (reads with N, longer than
//                   kBinMaxLen, survivor or tuple overflow) run process_seeds itself.
#pragma once

#include "mapper_kernels.cuh"

namespace ab2dev {

constexpr uint32_t kBinMaxLen = 1023;    // seed offset field of a tuple: 10 bits
constexpr uint32_t kMaxBins = 8192;      // bin cursors of one scatter CTA live in shared memory
constexpr uint32_t kScatterThreads = 1024;
constexpr uint32_t kTupleBlock = 256;    // tuple slots a warp reserves at a time
constexpr uint32_t kSurvSlots = 128;     // survivors per strand the replay handles (four per lane)
constexpr uint32_t kMaxTupleCount = (1u << 24) - 1u;

__device__ __forceinline__ uint32_t tuple_cnt(const SeedTuple &t) { return t.cnt_hi >> 8; }
__device__ __forceinline__ uint64_t tuple_rec(const SeedTuple &t) { return (uint64_t)t.rec_lo | ((uint64_t)(t.cnt_hi & 255u) << 32); }
__device__ __forceinline__ uint32_t tuple_table(uint32_t meta) { return ((meta >> 27) & 1u) ? (1u + ((meta >> 30) & 1u)) : 0u; }

// ---- hash_kernel ---------------------------------------------------------------------------------------------
struct TupleAlloc {  // warp-uniform cursor into the warp's current block of tuple slots
  uint32_t cur, end;
  bool full;
};

// Writes the tuples of the lanes that have one (`has`), compacted, into the warp's block(s).
__device__ __forceinline__ void emit_tuples(const BinParams &B, TupleAlloc &al, bool has, const SeedTuple &t, int lane) {
  const unsigned m = __ballot_sync(FULL, has);
  if (m == 0u) return;
  const uint32_t n = (uint32_t)__popc(m);
  const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
  const uint32_t avail = al.end - al.cur;
  if (has && rank < avail) B.tup[al.cur + rank] = t;
  if (n <= avail) {
    al.cur += n;
    return;
  }
  // a fresh block for the rest
  uint32_t base = 0;
  if (lane == 0) {  // (tup_cap is a multiple of kTupleBlock; once the buffer is full the cursor stops moving)
    base = *reinterpret_cast<volatile unsigned int *>(B.tup_count);
    if (base < B.tup_cap) base = atomicAdd(B.tup_count, kTupleBlock);
  }
  base = __shfl_sync(FULL, base, 0);
  if (base >= B.tup_cap) {  // out of tuple memory: the strand takes the direct path
    al.cur = al.end = 0;
    al.full = true;
    return;
  }
  if (has && rank >= avail) B.tup[base + (rank - avail)] = t;
  al.cur = base + (n - avail);
  al.end = base + kTupleBlock;
}

// Everything of process_seeds up to the candidate ranges, for pass `strand_code` of `end`: tuples + read planes.
// Returns the strand's flag.
// (The tuple cursor travels by value: by reference it would live in local memory of the caller.)
struct EmitResult {
  TupleAlloc al;
  uint32_t flag;
};
__device__ __noinline__ EmitResult emit_strand(TupleAlloc al, int end, uint32_t strand_code, uint32_t sid) {
  const Warp W;
  const KernelParams &P = params();
  const IndexDev &ix = P.ix;
  const BinParams &B = P.bp;
  const int lane = W.lane;
  build_qcode(end, strand_code);
  al.full = false;
  const uint32_t readlen = W.scal()->len[end];
  const uint8_t *qcode = W.qcode(end);
  const bool g_to_a = ((strand_code & ABG_FLAG_A_RICH) != 0) != ((strand_code & ABG_FLAG_RC) != 0);
  // One pass over the encoded read: the three hash planes (as build_packed leaves them; packed words and match
  // masks are not needed here) and the 2-bit planes of the bases (conversion undone: code 5 is A, code 10 is T)
  // for the scatter and seed kernels.  N anywhere -> direct path.
  {
    uint32_t *p2w = W.plane(0), *p3aw = W.plane(1), *p3bw = W.plane(2);
    const uint32_t hw = W.L.plane_words;
    uint32_t my_lo = 0, my_hi = 0;
    bool has_n = false;
    for (uint32_t w = 0; w < max(B.pw, hw); ++w) {
      const uint32_t i = 32u * w + (uint32_t)lane;
      const uint32_t code = i < readlen ? qcode[i] : 0u;
      const uint32_t t = three_num(g_to_a, code);
      const unsigned m2 = __ballot_sync(FULL, get_bit(code));
      const unsigned ma = __ballot_sync(FULL, t & 1u);
      const unsigned mb = __ballot_sync(FULL, t & 2u);
      const unsigned lo = __ballot_sync(FULL, code == 2u || (code & 8u) != 0u);
      const unsigned hi = __ballot_sync(FULL, code == 4u || (code & 8u) != 0u);
      has_n = has_n || __ballot_sync(FULL, i < readlen && code == 0u) != 0u;
      if (lane == 0 && w < hw) {
        p2w[w] = m2;
        p3aw[w] = ma;
        p3bw[w] = mb;
      }
      if ((uint32_t)lane == w) {
        my_lo = lo;
        my_hi = hi;
      }
    }
    if (lane == 0) W.scal()->packed_key = ~0u;  // the planes no longer belong to what build_packed last built
    __syncwarp();
    // {lo, hi} of 32 bases side by side: the scatter kernel's five-word windows are five 8-byte loads
    uint2 *dst = reinterpret_cast<uint2 *>(B.planes) + (size_t)sid * B.pw;
    if ((uint32_t)lane < B.pw) dst[lane] = make_uint2(my_lo, my_hi);
    if (has_n || readlen > kBinMaxLen) return EmitResult{al, 1u};
  }
  const uint32_t *p2 = W.plane(0), *p3a = W.plane(1), *p3b = W.plane(2);
  const uint32_t *T3 = tab3();
  const uint32_t *counter3 = g_to_a ? ix.counter_a : ix.counter_t;
  const uint32_t *index3 = g_to_a ? ix.index_a : ix.index_t;
  const uint32_t *bits3 = g_to_a ? ix.bits_a : ix.bits_t;
  const uint32_t maxc = P.max_candidates;
  const bool have3 = ix.n_ctx3 != 0;  // entries in the three-letter tables at all
  const uint32_t specific_len = min(readlen - P.window_size, readlen >> 1);
  const uint32_t specific_lim = max(P.window_size, readlen >> 1);
  const uint32_t lim_two = readlen - 25u + 1u;
  const uint32_t n_off = max(specific_lim, lim_two);
  const uint32_t bound = (uint32_t)invalid_hit_diffs(readlen);
  const uint32_t meta_strand = (bound << 18) | (g_to_a ? (1u << 30) : 0u);
  bool bad = false;
  for (uint32_t base_off = 0; base_off < n_off; base_off += 32) {
    const uint32_t i = base_off + (uint32_t)lane;
    const bool active = i < n_off;
    const bool in_spec = i < specific_lim, in_sens = i < lim_two;
    SeedTuple t2, t3;
    bool has2 = false, has3 = false;
    if (active) {
      const uint32_t k = __brev(plane_window(p2, i)) >> 7;
      // (an index without three-letter entries -- random genomes -- has nothing to hash for)
      uint32_t k3 = 0, bw3 = 0;
      if (have3) {
        const uint32_t x0 = plane_window(p3a, i) & 0xffffu, x1 = plane_window(p3b, i) & 0xffffu;
        k3 = T3[x0 & 255u] + T3[256 + (x0 >> 8)] + 2u * (T3[x1 & 255u] + T3[256 + (x1 >> 8)]);
        // the bitmap word of the three-letter bucket is requested before the two-letter probe is decoded (both
        // come from L2; the kernel's longest stalls are these round trips)
        bw3 = bits3 != nullptr ? __ldg(bits3 + (k3 >> 5)) : ~0u;
      }
      uint32_t s2, e2, s3 = 0, e3 = 0;
      probe_two(ix, k, s2, e2);
      if ((bw3 >> (k3 & 31u)) & 1u) {
        s3 = __ldg(counter3 + k3);
        e3 = __ldg(counter3 + k3 + 1);
      }
      const uint32_t d_two = e2 - s2, d_three = e3 - s3;
      // the sensitive phase's bucket rule (abismal.cpp:1351-1370), on the raw bucket sizes
      const bool el2 = d_two != 0u && d_two <= maxc && (d_three == 0u || d_two <= 10u * d_three);
      const bool el3 = d_three != 0u && d_three <= maxc;
      const uint32_t a = min((uint32_t)(kCtxArrays - 1), i >> 5);
      const uint32_t q0 = i - 32u * a;
      const uint32_t meta_off = i | (min(128u, readlen - q0) << 10) | meta_strand;
      {  // two-letter table
        bool spec = false;
        const bool sens = el2 && in_sens;
        uint32_t lo = s2, hi = e2;
        if (in_spec) {
          if (d_two > maxc) {  // the specific phase narrows the bucket (abismal.cpp:1316-1323)
            const SeedRange r = find_candidates(maxc, qcode + i, readlen - i, s2, e2);
            lo = r.low;
            hi = r.high;
            spec = (hi - lo) != 0u && ((hi - lo) <= maxc || r.p >= specific_len);
          }
          else spec = d_two != 0u;  // (an empty bucket comes back from find_candidates as it went in)
        }
        if (spec || sens) {
          const uint32_t cnt = hi - lo;
          if (cnt > kMaxTupleCount) bad = true;
          const uint64_t rec = (uint64_t)a * ix.n_ctx + lo;
          t2.rec_lo = (uint32_t)rec;
          t2.cnt_hi = (uint32_t)(rec >> 32) | (cnt << 8);
          t2.sid = sid;
          t2.meta = meta_off | (spec ? (1u << 28) : 0u) | (sens ? (1u << 29) : 0u);
          has2 = true;
        }
      }
      {  // three-letter table
        bool spec = false;
        const bool sens = el3 && in_sens;
        uint32_t lo = s3, hi = e3;
        if (in_spec) {
          if (d_three > maxc) {
            const SeedRange r = find_candidates_three(index3, g_to_a, maxc, qcode + i, readlen - i, s3, e3);
            lo = r.low;
            hi = r.high;
            spec = (hi - lo) != 0u && ((hi - lo) <= maxc || r.p >= specific_len);
          }
          else spec = d_three != 0u;
        }
        if (spec || sens) {
          const uint32_t cnt = hi - lo;
          if (cnt > kMaxTupleCount) bad = true;
          const uint64_t rec = (uint64_t)a * ix.n_ctx3 + lo;
          t3.rec_lo = (uint32_t)rec;
          t3.cnt_hi = (uint32_t)(rec >> 32) | (cnt << 8);
          t3.sid = sid;
          t3.meta = meta_off | (1u << 27) | (spec ? (1u << 28) : 0u) | (sens ? (1u << 29) : 0u);
          has3 = true;
        }
      }
    }
    __syncwarp();
    emit_tuples(B, al, has2, t2, lane);
    emit_tuples(B, al, has3, t3, lane);
  }
  return EmitResult{al, (__any_sync(FULL, bad) || al.full) ? 1u : 0u};
}

// (end, flags) of strand `pass` of an item: the order process_seeds is called in by map_single_ended[_rand]
// (abismal.cpp:1511-1704) and map_paired_ended[_rand] (:1887-2185)
__device__ __forceinline__ void strand_plan(const KernelParams &P, int pass, int &end, uint32_t &flags) {
  const bool paired = P.mode & ABG_MODE_PAIRED;
  const bool a_rich = P.mode & ABG_MODE_A_RICH;
  const bool rpbat = P.mode & ABG_MODE_RANDOM_PBAT;
  const uint32_t T = 0, A = ABG_FLAG_A_RICH, RC = ABG_FLAG_RC;
  if (!paired) {
    end = 0;
    if (rpbat) flags = pass == 0 ? T : (pass == 1 ? A : (pass == 2 ? (A | RC) : (T | RC)));
    else flags = (a_rich ? A : T) | (pass == 1 ? RC : 0u);
    return;
  }
  const CallPlan cp = call_plan(pass >> 1, rpbat, a_rich);
  end = (pass & 1) ? cp.e2 : cp.e1;
  flags = (pass & 1) ? cp.f2 : cp.f1;
}

template <int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) hash_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const BinParams &B = P.bp;
  const Warp W;
  const int lane = W.lane;
  WarpScalars *S = W.scal();
  TupleAlloc al;
  al.cur = al.end = 0u;
  al.full = false;
  // one read / pair per work item: its strands (2, 4 or 8 passes over one or two ends) share the loaded bases
  const bool paired = (P.mode & ABG_MODE_PAIRED) != 0;
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= P.n) break;
    uint32_t o[2] = {0u, 0u}, len[2] = {0u, 0u};
    o[0] = P.off[0][item];
    len[0] = P.off[0][item + 1] - o[0];
    if (paired) {
      o[1] = P.off[1][item];
      len[1] = P.off[1][item + 1] - o[1];
    }
    if (lane == 0) {
      S->qkey[0] = S->qkey[1] = ~0u;
      S->packed_key = ~0u;
      S->len[0] = len[0];
      S->len[1] = len[1];
    }
    __syncwarp();
    if (len[0] != 0) load_end(W, 0, P.seq[0] + o[0], len[0]);
    if (len[1] != 0) load_end(W, 1, P.seq[1] + o[1], len[1]);
    for (uint32_t pass = 0; pass < B.spi; ++pass) {
      const uint32_t w = item * B.spi + pass;
      int end;
      uint32_t flags;
      strand_plan(P, (int)pass, end, flags);
      uint32_t flag = 2u;
      if (len[end] != 0) {
        const EmitResult er = emit_strand(al, end, flags, B.sid_base + w);
        al = er.al;
        flag = er.flag;
      }
      if (lane == 0) B.strand_flag[B.sid_base + w] = (uint8_t)flag;
    }
  }
  // the unused rest of the warp's last block: empty tuples
  for (uint32_t k = al.cur + (uint32_t)lane; k < al.end; k += 32) B.tup[k] = SeedTuple{0u, 0u, 0u, 0u};
}

// ---- scatter_kernel / filter_kernel ----------------------------------------------------------------------------
struct FilterParams {
  const SeedTuple *tup;
  const unsigned int *tup_count;
  uint32_t tup_cap;
  const uint32_t *planes;
  uint32_t pw;
  uint32_t bin_shift, n_bins;
  uint64_t rec_base[3];
  uint32_t *bin_hist;      // in: tuples per bin; bin_prefix_kernel turns it into the bins' write cursors
  uint32_t *n_binned;      // out of bin_prefix_kernel: tuples in all bins
  SeedTuple *tup_b;        // binned tuples
  const uint4 *ctx[3];     // record arrays of the three tables
  uint64_t n_tab[3];       // entries per table (records of array a start at a * n_tab)
  uint32_t *surv_count;
  uint2 *surv;
  uint32_t surv_cap;
  unsigned int *work;      // filter_kernel's work cursor
  uint32_t grab;           // tuples a filter warp takes per work-cursor atomic (a multiple of 32)
  uint32_t n_cursors;      // interleaved work cursors of the filter (F.work[0 .. n_cursors))
  uint32_t cache;          // tuning: bit 0 planes loaded evict-first, bit 1 records loaded with the L2 evict-last hint
};

// Binning without global atomics: CTA c of count_kernel and of scatter_kernel owns the same contiguous range of
// the tuple array.  count_kernel leaves the CTA's tuples per bin in bin_hist[bin * n_cta + c]; one exclusive
// prefix sum over that (bin-major) matrix gives every (bin, CTA) its own write range, so the scatter CTAs hand
// out positions from shared-memory cursors.  (One global atomicAdd per tuple was measured at 45 ms for 2^29
// tuples, tools/micro/gather_bench4.cu.)
__device__ __forceinline__ void cta_range(const FilterParams &F, uint64_t &i0, uint64_t &i1) {
  const uint64_t n = min(*F.tup_count, F.tup_cap);
  uint64_t per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + kScatterThreads - 1) / kScatterThreads * kScatterThreads;
  i0 = min(n, per * blockIdx.x);
  i1 = min(n, i0 + per);
}
__device__ __forceinline__ SeedTuple load_tuple(const SeedTuple *p, uint4 &raw) {
  raw = __ldg(reinterpret_cast<const uint4 *>(p));
  SeedTuple t;
  t.rec_lo = raw.x; t.cnt_hi = raw.y; t.sid = raw.z; t.meta = raw.w;
  return t;
}

__global__ void __launch_bounds__(kScatterThreads) count_kernel(FilterParams F) {
  extern __shared__ uint32_t s_bins[];
  for (uint32_t b = threadIdx.x; b < F.n_bins; b += blockDim.x) s_bins[b] = 0u;
  __syncthreads();
  uint64_t i0, i1;
  cta_range(F, i0, i1);
  for (uint64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint4 raw;
    const SeedTuple t = load_tuple(F.tup + i, raw);
    if (tuple_cnt(t) == 0u) continue;
    atomicAdd(s_bins + (uint32_t)((F.rec_base[tuple_table(t.meta)] + tuple_rec(t)) >> F.bin_shift), 1u);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < F.n_bins; b += blockDim.x) F.bin_hist[(size_t)b * gridDim.x + blockIdx.x] = s_bins[b];
}

// exclusive prefix sum of the n_bins x n_cta matrix in place (one block)
__global__ void bin_prefix_kernel(FilterParams F, uint32_t n_cta) {
  __shared__ uint32_t part[1024];
  const uint32_t n = F.n_bins * n_cta;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t b0 = min(n, threadIdx.x * per), b1 = min(n, b0 + per);
  uint32_t sum = 0;
  for (uint32_t b = b0; b < b1; ++b) sum += F.bin_hist[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (uint32_t k = 0; k < blockDim.x; ++k) {
      const uint32_t v = part[k];
      part[k] = acc;
      acc += v;
    }
    *F.n_binned = acc;
  }
  __syncthreads();
  uint32_t acc = part[threadIdx.x];
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t v = F.bin_hist[b];
    F.bin_hist[b] = acc;
    acc += v;
  }
}

__global__ void __launch_bounds__(kScatterThreads, 2) scatter_kernel(FilterParams F) {
  extern __shared__ uint32_t s_bins[];
  for (uint32_t b = threadIdx.x; b < F.n_bins; b += blockDim.x) s_bins[b] = F.bin_hist[(size_t)b * gridDim.x + blockIdx.x];
  __syncthreads();
  uint64_t i0, i1;
  cta_range(F, i0, i1);
  // tuples only: their payloads (the 128 read bases a tuple's records are compared with) are built by the filter
  // kernel from the strand's planes, so the scatter writes 16 bytes per tuple and its write frontier (one open
  // line per bin and CTA) stays in L2.  The next tuple is in flight while this one is placed.
  const uint4 *src = reinterpret_cast<const uint4 *>(F.tup);
  uint64_t i = i0 + threadIdx.x;
  uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
  if (i < i1) nxt = __ldcs(src + i);
  for (; i < i1; i += blockDim.x) {
    const uint4 raw = nxt;
    if (i + blockDim.x < i1) nxt = __ldcs(src + i + blockDim.x);
    if ((raw.y >> 8) == 0u) continue;
    const uint64_t g = F.rec_base[tuple_table(raw.w)] + ((uint64_t)raw.x | ((uint64_t)(raw.y & 255u) << 32));
    const uint32_t at = atomicAdd(s_bins + (uint32_t)(g >> F.bin_shift), 1u);
    reinterpret_cast<uint4 *>(F.tup_b)[at] = raw;
  }
}

// The same scatter through a shared-memory tile: the CTA sorts kSortTile tuples by bin in shared memory (counting
// sort: shared-memory atomics give the rank, one block scan the offsets) and writes them out in bin order, so
// the tuples of one bin leave as one run of consecutive addresses (~7 tuples = 107 bytes per bin and tile at
// 1227 bins) instead of one 16-byte store per tuple into 1227 open lines.  Same (bin, CTA) write ranges as
// scatter_kernel.  Shared memory: the tile + three words per bin (n_bins <= kSortMaxBins).
constexpr uint32_t kSortPerThread = 8, kSortTile = kSortPerThread * kScatterThreads, kSortMaxBins = 2048;
__host__ __device__ __forceinline__ size_t sort_scatter_smem(uint32_t n_bins) { return (size_t)kSortTile * 16u + 3u * (size_t)n_bins * 4u + 32u * 4u; }

__global__ void __launch_bounds__(kScatterThreads, 1) scatter_sorted_kernel(FilterParams F) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint4 *tile = reinterpret_cast<uint4 *>(s_raw);
  uint32_t *cursor = reinterpret_cast<uint32_t *>(s_raw + (size_t)kSortTile * 16u);
  uint32_t *hist = cursor + F.n_bins, *off = hist + F.n_bins, *wsum = off + F.n_bins;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  for (uint32_t b = tid; b < F.n_bins; b += blockDim.x) {
    cursor[b] = F.bin_hist[(size_t)b * gridDim.x + blockIdx.x];
    hist[b] = 0u;
  }
  __syncthreads();
  uint64_t i0, i1;
  cta_range(F, i0, i1);
  const uint4 *src = reinterpret_cast<const uint4 *>(F.tup);
  const uint32_t per = (F.n_bins + kScatterThreads - 1u) / kScatterThreads;  // bins per thread in the scan
  for (uint64_t base = i0; base < i1; base += kSortTile) {
    uint4 raw[kSortPerThread];
    uint32_t bin[kSortPerThread], rk[kSortPerThread];
#pragma unroll
    for (uint32_t k = 0; k < kSortPerThread; ++k) {
      const uint64_t i = base + (uint64_t)k * kScatterThreads + tid;
      raw[k] = i < i1 ? __ldcs(src + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (uint32_t k = 0; k < kSortPerThread; ++k) {
      bin[k] = ~0u;
      rk[k] = 0u;
      if ((raw[k].y >> 8) != 0u) {
        const uint64_t g = F.rec_base[tuple_table(raw[k].w)] + ((uint64_t)raw[k].x | ((uint64_t)(raw[k].y & 255u) << 32));
        bin[k] = (uint32_t)(g >> F.bin_shift);
        rk[k] = atomicAdd(hist + bin[k], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of hist -> off
    uint32_t sum = 0;
    for (uint32_t q = 0; q < per; ++q) {
      const uint32_t b = tid * per + q;
      sum += b < F.n_bins ? hist[b] : 0u;
    }
    const uint32_t incl = warp_incl_scan_add(sum, (int)lane);
    if (lane == 31u) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0u) {
      const uint32_t v = wsum[lane];
      const uint32_t iv = warp_incl_scan_add(v, (int)lane);
      wsum[lane] = iv - v;
    }
    __syncthreads();
    uint32_t acc = wsum[wid] + incl - sum;
    for (uint32_t q = 0; q < per; ++q) {
      const uint32_t b = tid * per + q;
      if (b < F.n_bins) {
        off[b] = acc;
        acc += hist[b];
      }
    }
    __syncthreads();
    const uint32_t total = off[F.n_bins - 1u] + hist[F.n_bins - 1u];
#pragma unroll
    for (uint32_t k = 0; k < kSortPerThread; ++k)
      if (bin[k] != ~0u) tile[off[bin[k]] + rk[k]] = raw[k];
    __syncthreads();
    for (uint32_t j = tid; j < total; j += kScatterThreads) {
      const uint4 t = tile[j];
      const uint64_t g = F.rec_base[tuple_table(t.w)] + ((uint64_t)t.x | ((uint64_t)(t.y & 255u) << 32));
      const uint32_t b = (uint32_t)(g >> F.bin_shift);
      reinterpret_cast<uint4 *>(F.tup_b)[cursor[b] + (j - off[b])] = t;
    }
    __syncthreads();
    for (uint32_t b = tid; b < F.n_bins; b += blockDim.x) {
      cursor[b] += hist[b];
      hist[b] = 0u;
    }
    __syncthreads();
  }
}

// mismatch bits of 32 bases: genome planes (glo, ghi) against read planes (rlo, rhi), both complemented for
// g_to_a strands, so that the one conversion rule left is "read T (11) also matches genome C (01)"
__device__ __forceinline__ uint32_t mismatch_bits(uint32_t glo, uint32_t ghi, uint32_t rlo, uint32_t rhi) {
  const uint32_t diff = (glo ^ rlo) | (ghi ^ rhi);
  return diff & ~(rlo & rhi & glo & ~ghi);
}
__device__ __forceinline__ uint32_t low_mask(int k) {  // k low bits set, k in [0, 32]
  uint32_t r;
  asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(k));  // shifts of 32 and more give 0
  return r - 1u;
}

// a record with the L2 evict-last hint (the records of a bin are what the filter wants to keep in L2)
__device__ __forceinline__ void load_ctx_keep(const uint4 *p, uint32_t (&w)[8], uint64_t policy) {
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p), "l"(policy));
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p + 1), "l"(policy));
}

// One candidate per lane: record against payload -> lower bound of the distance -> survivor entry.
__device__ __forceinline__ void filter_candidate(const FilterParams &F, const uint4 h, uint64_t rec, const uint32_t (&w)[8],
                                                 const uint4 pa, const uint4 pb) {
  const uint32_t meta = h.w;
  const bool sentinel = (w[0] & w[1] & w[2] & w[3] & w[4] & w[5] & w[6] & w[7]) == ~0u;
  const uint32_t x = ((meta >> 30) & 1u) ? ~0u : 0u;
  const uint32_t m0 = mismatch_bits(w[0] ^ x, w[1] ^ x, pa.x, pa.y), m1 = mismatch_bits(w[2] ^ x, w[3] ^ x, pa.z, pa.w);
  const uint32_t m2 = mismatch_bits(w[4] ^ x, w[5] ^ x, pb.x, pb.y), m3 = mismatch_bits(w[6] ^ x, w[7] ^ x, pb.z, pb.w);
  const int nb = (int)((meta >> 10) & 255u);  // compared bases of the record: 1..128
  int lb;
  if (nb >= 64)
    lb = __popc(m0) + __popc(m1) + __popc(m2 & low_mask(min(nb - 64, 32))) + __popc(m3 & low_mask(max(nb - 96, 0)));
  else
    lb = __popc(m0 & low_mask(min(nb, 32))) + __popc(m1 & low_mask(max(nb - 32, 0)));
  const int bound = (int)((meta >> 18) & 511u);
  if (lb <= bound || sentinel) {
    const uint32_t tab = tuple_table(meta);
    const uint32_t off = meta & 1023u;
    const uint32_t a = min((uint32_t)(kCtxArrays - 1), off >> 5);
    const uint32_t entry = (uint32_t)(rec - (uint64_t)a * F.n_tab[tab]);
    const uint32_t at = atomicAdd(F.surv_count + h.z, 1u);
    if (at < F.surv_cap) F.surv[(size_t)h.z * F.surv_cap + at] = make_uint2(entry, off | ((meta >> 27) & 7u) << 10);
  }
}

// PIPE: the records of the next 32 candidates travel to shared memory (cp.async, no registers) while this round's
// are compared: the kernel waits on the record gathers, one dependent round trip per 32 candidates.
template <bool PIPE, bool KEEP, int MINB = 6>
__global__ void __launch_bounds__(256, MINB) filter_kernel(FilterParams F) {
  __shared__ uint4 s_hdr[8][32];
  __shared__ uint4 s_pay[8][32][2];  // per tuple: {lo, hi} plane words of the 128 read bases its records are compared with
  __shared__ uint32_t s_excl[8][32];
  __shared__ uint4 s_ctx[PIPE ? 8 : 1][2][PIPE ? 32 : 1][2];  // PIPE: two stages of one record per lane
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = *F.n_binned;
  uint64_t keep_policy = 0;
  if (KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
  // Work distribution.  grab != 0: a warp takes `grab` tuples at a time from a global cursor.  grab == 0: static,
  // warp w of the grid takes the groups of 32 tuples w, w + W, w + 2 W, ...: no atomics, and the grid's live
  // window of the tuple array (hence of the record array: tuples are in bin order) is as narrow as it gets.
  const uint32_t n_warps = gridDim.x * 8u, warp_id = blockIdx.x * 8u + (uint32_t)wid;
  for (uint64_t round = 0;; ++round) {
    uint32_t g0 = 0, g1;
    if (F.grab != 0u) {
      // n_cursors interleaved work cursors (cursor c hands out the pieces c, c + C, c + 2 C, ... of `grab` tuples):
      // the atomics of the grid spread over C addresses, so pieces can be small -- a narrow live window of the
      // record array -- without the one-address atomic rate becoming the limit
      const uint32_t c = warp_id % F.n_cursors;
      uint32_t k = 0;
      if (lane == 0) k = atomicAdd(F.work + c, 1u);
      k = __shfl_sync(FULL, k, 0);
      const uint64_t at = ((uint64_t)k * F.n_cursors + c) * F.grab;
      if (at >= n) break;
      g0 = (uint32_t)at;
      g1 = (uint32_t)min((uint64_t)n, at + F.grab);
    }
    else {
      const uint64_t at = (round * n_warps + warp_id) * 32u;
      if (at >= n) break;
      g0 = (uint32_t)at;
      g1 = min(n, g0 + 32u);
    }
    for (uint32_t w0 = g0; w0 < g1; w0 += 32) {
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (w0 + (uint32_t)lane < n) raw = __ldcs(reinterpret_cast<const uint4 *>(F.tup_b) + w0 + lane);
      const uint32_t tot = raw.y >> 8;
      const uint32_t incl = warp_incl_scan_add(tot, lane);
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      const uint32_t excl = incl - tot;
      __syncwarp();
      s_hdr[wid][lane] = raw;
      s_excl[wid][lane] = excl;
      if (tot != 0u) {
        // the 128 read bases from q0 = offset - 32 a (a = min(offset / 32, 3)): five {lo, hi} words of the strand's
        // planes, shifted.  g_to_a strands are compared complemented (A<->T, C<->G in the 2-bit code): "read A
        // also matches genome G" becomes "read T also matches genome C", the rule of the other strands.
        const uint32_t off = raw.w & 1023u;
        const uint32_t q0 = off - 32u * min((uint32_t)(kCtxArrays - 1), off >> 5);
        const uint32_t sh = q0 & 31u;
        const uint2 *pl = reinterpret_cast<const uint2 *>(F.planes) + (size_t)raw.z * F.pw + (q0 >> 5);
        uint2 v[5];
        if (F.cache & 1u) {
#pragma unroll
          for (int c = 0; c < 5; ++c) v[c] = __ldcs(pl + c);
        }
        else {
#pragma unroll
          for (int c = 0; c < 5; ++c) v[c] = __ldg(pl + c);
        }
        const uint32_t x = ((raw.w >> 30) & 1u) ? ~0u : 0u;
        s_pay[wid][lane][0] = make_uint4(__funnelshift_r(v[0].x, v[1].x, sh) ^ x, __funnelshift_r(v[0].y, v[1].y, sh) ^ x,
                                         __funnelshift_r(v[1].x, v[2].x, sh) ^ x, __funnelshift_r(v[1].y, v[2].y, sh) ^ x);
        s_pay[wid][lane][1] = make_uint4(__funnelshift_r(v[2].x, v[3].x, sh) ^ x, __funnelshift_r(v[2].y, v[3].y, sh) ^ x,
                                         __funnelshift_r(v[3].x, v[4].x, sh) ^ x, __funnelshift_r(v[3].y, v[4].y, sh) ^ x);
      }
      __syncwarp();
      // owner tuple of candidate c0 + lane: tuples that end at or before c0, plus the tuples that start inside
      // (c0, c0 + lane] (every binned tuple has at least one candidate); -> the candidate's record, or false
      auto locate = [&](uint32_t c0, uint32_t &o, uint64_t &rec) -> bool {
        const uint32_t st = excl - c0;
        const uint32_t heads = __reduce_or_sync(FULL, (tot != 0u && st - 1u < 31u) ? (1u << st) : 0u);
        const uint32_t first = (uint32_t)__popc(__ballot_sync(FULL, incl <= c0));
        const uint32_t cidx = c0 + (uint32_t)lane;
        o = (first + (uint32_t)__popc(heads & ((2u << lane) - 1u) & ~1u)) & 31u;
        if (cidx >= total) return false;
        const uint4 h = s_hdr[wid][o];
        rec = ((uint64_t)h.x | ((uint64_t)(h.y & 255u) << 32)) + (cidx - s_excl[wid][o]);
        return true;
      };
      if (!PIPE) {
        for (uint32_t c0 = 0; c0 < total; c0 += 32) {
          uint32_t o;
          uint64_t rec = 0;
          if (!locate(c0, o, rec)) continue;
          const uint4 h = s_hdr[wid][o];
          uint32_t w[8];
          if (KEEP) load_ctx_keep(F.ctx[tuple_table(h.w)] + 2 * rec, w, keep_policy);
          else load_ctx(F.ctx[tuple_table(h.w)] + 2 * rec, w);
          filter_candidate(F, h, rec, w, s_pay[wid][o][0], s_pay[wid][o][1]);
        }
      }
      else {
        auto fetch = [&](int stage, uint32_t o_, uint64_t rec_) {
          const uint4 *src = F.ctx[tuple_table(s_hdr[wid][o_].w)] + 2 * rec_;
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_ctx[wid][stage][lane][0]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 1) : "memory");
        };
        uint32_t o = 0;
        uint64_t rec = 0;
        bool ok = total != 0u && locate(0u, o, rec);
        if (ok) fetch(0, o, rec);
        asm volatile("cp.async.commit_group;" ::: "memory");
        int stage = 0;
        for (uint32_t c0 = 0; c0 < total; c0 += 32, stage ^= 1) {
          uint32_t o_n = 0;
          uint64_t rec_n = 0;
          bool ok_n = false;
          if (c0 + 32u < total) {  // (warp-uniform)
            ok_n = locate(c0 + 32u, o_n, rec_n);
            if (ok_n) fetch(stage ^ 1, o_n, rec_n);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 1;" ::: "memory");  // this round's record (committed one group earlier) has landed
          if (ok) {
            const uint4 x = s_ctx[wid][stage][lane][0], y = s_ctx[wid][stage][lane][1];
            const uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            filter_candidate(F, s_hdr[wid][o], rec, w, s_pay[wid][o][0], s_pay[wid][o][1]);
          }
          ok = ok_n;
          o = o_n;
          rec = rec_n;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
  }
}

// ---- seed_kernel side: the survivors of one strand against its candidate set ---------------------------------
// The read's bases, encoded for the pass, when something needs them after all (the direct path, exact_compare)
__device__ __forceinline__ void ensure_encoded(const Warp &W, int end, uint32_t strand_code, const char *seq) {
  WarpScalars *S = W.scal();
  __syncwarp();
  if (S->loaded[end] == 0u) {
    load_end(W, end, seq, S->len[end]);
    if (W.lane == 0) S->loaded[end] = 1u;
    __syncwarp();
  }
  build_qcode(end, strand_code);
  build_packed(end, strand_code);
}

// The general replay of a strand's n <= kSurvSlots listed survivors (the match masks are in place): all of them
// are ranked into the canonical order (offset, two-letter before three-letter, entry), deep-compared, and
// replayed by one lane.  process_binned takes this path when more than 32 survivors could still be accepted
// (reads in repeats); otherwise it ranks and replays only those.
__device__ __noinline__ void replay_all(int set_id, int end, uint32_t strand_code, uint32_t sid, const char *seq, int n,
                                        int n_words, int bound, const uint32_t *index3) {
  const Warp W;
  const BinParams &B = params().bp;
  const int lane = W.lane;
  uint64_t *keys = reinterpret_cast<uint64_t *>(W.log_pos());  // [kSurvSlots] (the survivor log's region)
  uint8_t *perm = reinterpret_cast<uint8_t *>(W.stage());       // [kSurvSlots] (the staging region of replay_hits)
  __syncwarp();
  for (int k = lane; k < n; k += 32) {
    const uint2 e = __ldcg(B.surv + (size_t)sid * B.surv_cap + k);
    const uint64_t off = e.y & 1023u, is3 = (e.y >> 10) & 1u, fl = (e.y >> 11) & 3u;
    keys[k] = (off << 36) | (is3 << 35) | ((uint64_t)e.x << 3) | fl;  // unique per candidate: fl never decides
  }
  __syncwarp();
  for (int k = lane; k < n; k += 32) {
    const uint64_t mine = keys[k];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += keys[j] < mine;
    perm[rank] = (uint8_t)k;
  }
  __syncwarp();
  uint32_t *res = reinterpret_cast<uint32_t *>(keys);  // slot k: {pos, d | pm << 16 | spec << 30 | sens << 31}
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    const bool valid = k < n;
    const uint64_t key = valid ? keys[k] : 0ull;
    const uint32_t entry = (uint32_t)(key >> 3), is3 = (uint32_t)(key >> 35) & 1u, off = (uint32_t)(key >> 36) & 1023u;
    Deep1 r = compare_deep_one(index3, n_words, bound, valid, entry, off | (is3 << 31), true);
    const bool deferred = valid && r.pm == kDeferredExact;
    if (__any_sync(FULL, deferred)) {  // a window with N / IUPAC codes: the exact compare needs the packed read
      ensure_encoded(W, end, strand_code, seq);
      if (deferred) {
        int mx = 0;
        r.d = exact_compare(r.pos, n_words, bound, &mx);
        r.pm = mx;
      }
    }
    __syncwarp();
    if (valid) {
      res[2 * k] = r.pos;
      res[2 * k + 1] = (uint32_t)(uint16_t)(int16_t)r.d | ((uint32_t)max(0, min(r.pm, 0x3fff)) << 16) | ((uint32_t)(key & 3u) << 30);
    }
  }
  __syncwarp();
  if (lane == 0) {  // one lane mutates the set (see replay_hits)
    CandSet cs;
    cs.load(W, set_id);
    cs.set_specific();
    for (int q = 0; q < n && !cs.sure_ambig; ++q) {
      const int k = perm[q];
      const uint32_t m = res[2 * k + 1];
      if (((m >> 30) & 1u) == 0u) continue;  // bit 30: examined by the specific phase (tuple bit 28)
      if ((int)((m >> 16) & 0x3fffu) <= cs.cutoff) cs.update(true, (int)(int16_t)(m & 0xffffu), strand_code, res[2 * k]);
    }
    if (cs.should_do_sensitive()) {
      cs.set_sensitive();
      for (int q = 0; q < n && !cs.sure_ambig; ++q) {
        const int k = perm[q];
        const uint32_t m = res[2 * k + 1];
        if (((m >> 31) & 1u) == 0u) continue;  // bit 31: examined by the sensitive phase (tuple bit 29)
        if ((int)((m >> 16) & 0x3fffu) <= cs.cutoff) cs.update(true, (int)(int16_t)(m & 0xffffu), strand_code, res[2 * k]);
      }
    }
    cs.store_one_lane(W, set_id);
  }
  __syncwarp();
}

// Equivalent of process_seeds(set_id, end, strand_code) for strand `sid` when its survivors were listed.  The
// match masks of the deep compare come from the 2-bit planes hash_kernel stored (no N in the read here), so the
// read is not loaded and encoded a second time.
__device__ __noinline__ void process_binned(int set_id, int end, uint32_t strand_code, uint32_t sid, const char *seq) {
  const Warp W;
  const KernelParams &P = params();
  const BinParams &B = P.bp;
  const int lane = W.lane;
  WarpScalars *S = W.scal();
  const uint32_t n_raw = __ldcg(B.surv_count + sid);
  if (__ldcg(B.strand_flag + sid) != 0u || n_raw > min(B.surv_cap, kSurvSlots)) {
    ensure_encoded(W, end, strand_code, seq);
    process_seeds(set_id, end, strand_code);
    return;
  }
  const int n = (int)n_raw;
  const uint32_t readlen = S->len[end];
  const int n_words = (int)((readlen + 15) / 16);
  const int bound = invalid_hit_diffs(readlen);
  const bool g_to_a = ((strand_code & ABG_FLAG_A_RICH) != 0) != ((strand_code & ABG_FLAG_RC) != 0);
  const uint32_t *index3 = g_to_a ? P.ix.index_a : P.ix.index_t;
  // ---- match masks (what build_packed derives from the encoded read): bit j of masks(X)[c] <=> read base 32c+j
  // matches genome base X; a converted base also matches its partner; the 0xF tail of the last packed word
  // matches everything, positions past it nothing
  uint32_t lo = 0, hi = 0;
  if ((uint32_t)lane < B.pw) {
    const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(B.planes) + (size_t)sid * B.pw + lane);
    lo = v.x;
    hi = v.y;
  }
  __syncwarp();
  if ((uint32_t)lane < W.L.mask_words) {
    const int r = (int)readlen - 32 * lane, t = 16 * n_words - 32 * lane;
    const uint32_t real = r >= 32 ? ~0u : (r <= 0 ? 0u : ((1u << r) - 1u));
    const uint32_t tail = (t >= 32 ? ~0u : (t <= 0 ? 0u : ((1u << t) - 1u))) & ~real;
    const uint32_t isA = ~lo & ~hi & real, isC = lo & ~hi & real, isG = ~lo & hi & real, isT = lo & hi & real;
    W.masks(0)[lane] = isA | tail;
    W.masks(1)[lane] = isC | (g_to_a ? 0u : isT) | tail;
    W.masks(2)[lane] = isG | (g_to_a ? isA : 0u) | tail;
    W.masks(3)[lane] = isT | tail;
  }
  if (lane == 0) S->packed_key = ~0u;  // the masks no longer belong to what build_packed last built
  // ---- deep compare of every survivor (pure: any order).  Only the few that a phase could still accept go on:
  // the specific phase starts from cutoff = good_cutoff, the sensitive phase from the heap's top, and within a
  // phase cutoffs only tighten (CandSet::update), so pm above both can never be accepted -- on a random genome
  // that is everything but the read's true locus (seen once per indexed offset).
  const CandState *st0 = W.cs(set_id);
  const int c_spec = st0->good_cutoff;
  // (within the specific phase the heap's top can rise to good_cutoff when it starts below it: a full SE heap
  // pops its top for a worse hit that is still <= cutoff)
  const int c_sens = max(heap_of(W, set_id).get(0).diffs(), c_spec);
  uint64_t *acc_key = reinterpret_cast<uint64_t *>(W.log_pos());  // [32] canonical-order keys of the kept ones
  uint64_t *acc_res = acc_key + 32;                                // [32] {pos, d | pm << 16 | spec << 30 | sens << 31}
  uint64_t *sorted = acc_res + 32;                                 // [32] acc_res in canonical order
  int n_acc = 0;
  __syncwarp();
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    const bool valid = k < n;
    uint2 e = make_uint2(0u, 0u);
    if (valid) e = __ldcg(B.surv + (size_t)sid * B.surv_cap + k);
    const uint32_t off = e.y & 1023u, is3 = (e.y >> 10) & 1u, fl = (e.y >> 11) & 3u;
    Deep1 r = compare_deep_one(index3, n_words, bound, valid, e.x, off | (is3 << 31), true);
    const bool deferred = valid && r.pm == kDeferredExact;
    if (__any_sync(FULL, deferred)) {  // a window with N / IUPAC codes: the exact compare needs the packed read
      ensure_encoded(W, end, strand_code, seq);
      if (deferred) {
        int mx = 0;
        r.d = exact_compare(r.pos, n_words, bound, &mx);
        r.pm = mx;
      }
    }
    const bool keep = valid && (((fl & 1u) != 0u && r.pm <= c_spec) || ((fl & 2u) != 0u && r.pm <= c_sens));
    const unsigned km = __ballot_sync(FULL, keep);
    const int at = n_acc + __popc(km & ((1u << lane) - 1u));
    if (keep && at < 32) {
      // unique per candidate (fl never decides): offset, two-letter before three-letter, bucket order
      acc_key[at] = ((uint64_t)off << 36) | ((uint64_t)is3 << 35) | ((uint64_t)e.x << 3) | fl;
      acc_res[at] = (uint64_t)r.pos | ((uint64_t)((uint32_t)(uint16_t)(int16_t)r.d | ((uint32_t)max(0, min(r.pm, 0x3fff)) << 16) | (fl << 30)) << 32);
    }
    n_acc += __popc(km);
  }
  if (n_acc > (int)B.acc_cap) {  // many acceptable candidates (repeats): every survivor is ranked and replayed
    replay_all(set_id, end, strand_code, sid, seq, n, n_words, bound, index3);
    return;
  }
  __syncwarp();
  if (lane < n_acc) {
    const uint64_t mine = acc_key[lane];
    int rank = 0;
    for (int j = 0; j < n_acc; ++j) rank += acc_key[j] < mine;
    sorted[rank] = acc_res[lane];
  }
  __syncwarp();
  if (lane == 0) {  // one lane mutates the set (see replay_hits); the phases move the cutoff even without updates
    CandSet cs;
    cs.load(W, set_id);
    cs.set_specific();
    for (int q = 0; q < n_acc && !cs.sure_ambig; ++q) {
      const uint64_t v = sorted[q];
      const uint32_t m = (uint32_t)(v >> 32);
      if (((m >> 30) & 1u) == 0u) continue;  // bit 30: examined by the specific phase (tuple bit 28)
      if ((int)((m >> 16) & 0x3fffu) <= cs.cutoff) cs.update(true, (int)(int16_t)(m & 0xffffu), strand_code, (uint32_t)v);
    }
    if (cs.should_do_sensitive()) {
      cs.set_sensitive();
      for (int q = 0; q < n_acc && !cs.sure_ambig; ++q) {
        const uint64_t v = sorted[q];
        const uint32_t m = (uint32_t)(v >> 32);
        if (((m >> 31) & 1u) == 0u) continue;  // bit 31: examined by the sensitive phase (tuple bit 29)
        if ((int)((m >> 16) & 0x3fffu) <= cs.cutoff) cs.update(true, (int)(int16_t)(m & 0xffffu), strand_code, (uint32_t)v);
      }
    }
    cs.store_one_lane(W, set_id);
  }
  __syncwarp();
}

}  // namespace ab2dev
