// mapper_kernels.cuh -- the sm_100a device code of the `abismal map` hot path.
//
// One persistent kernel, one WARP per read (SE) or read pair (PE), fetched from
// a global work counter.  A warp carries its read through the whole path the
// reference runs per read between load_reads and format_*:
//
//   encode/pack            prep_read / pack_read            abismal.cpp:1377-1426
//   seed hash + lookup     process_seeds, get_1bit_hash...  abismal.cpp:1269-1375
//   bucket narrowing       find_candidates[_three]          abismal.cpp:1163-1259
//   packed compare         check_hits / full_compare        abismal.cpp:1105-1150
//   candidate sets         se_candidates / pe_candidates    abismal.cpp:334-449, 775-863
//   banded alignment       AbismalAlign::align              AbismalAlign.hpp:320-386
//   CIGAR / NM / position  build_cigar_len_and_pos          AbismalAlign.hpp:388-440
//   SE selection           align_se_candidates              abismal.cpp:1435-1497
//   PE mating + selection  best_pair/best_single/...        abismal.cpp:1715-1885
//
// The reference's results are order dependent (cutoffs tighten as hits arrive,
// heaps evict, ties break by iteration order), so only PURE quantities are
// computed in parallel and the order-dependent state machine is replayed in the
// reference's sequence on warp-uniform state:
//   * seed hashes come from three bit planes of the encoded read (two-letter
//     bit, three-letter digit bits) built with ballots; a lane extracts its
//     25-bit key with a funnel shift + brev and its base-3 key with four table
//     look-ups, then gathers its two counter pairs (4 loads in flight per lane);
//   * candidates are enumerated in the reference's canonical order (offset,
//     two-letter bucket before three-letter bucket, bucket order) and compared
//     kCand per lane at a time: all index gathers are issued together, then all
//     first-stage genome gathers, so a lane keeps up to 5*kCand independent
//     loads in flight instead of one dependent chain (the HBM-gather hot spot);
//   * survivors (ballot) are replayed IN ORDER against the candidate set with
//     libstdc++'s heap routines restated verbatim, so cutoff tightening,
//     evictions, sure_ambig exits and the specific->sensitive gate match;
//   * banded DP runs as a wavefront: lane l owns band columns 2l and 2l+1 and,
//     at iteration T, row T-l; one shuffle per cell column replaces the serial
//     from_left scan.  Traceback arrows are packed 4 bits per iteration into a
//     per-lane 64-bit shift register and flushed every 16 iterations.
//
// All per-warp state that device functions share lives in SHARED memory (the
// kernel parameters are copied there too), so no function takes a context
// struct by reference: nothing is forced into local memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#include "abismal_b200.h"

namespace ab2dev {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerBlock = 8;
constexpr int kThreadsPerBlock = kWarpsPerBlock * 32;
constexpr int kSeMax = 50;          // se_candidates::max_size
constexpr int kSeSlots = 64;        // padded to a power of two for the sort
constexpr int kPeSmall = 32;        // pe_candidates::max_size_small
constexpr int kPeLarge = 32 << 10;  // pe_candidates::max_size_large
constexpr int kPeSmemSlots = 128;   // PE heap entries kept in shared memory
constexpr int kMaxDiffs = 32767;
#ifndef ABG_KCAND
#define ABG_KCAND 1
#endif
constexpr int kCand = ABG_KCAND;    // candidates per lane per compare round (the memory system saturates at ~120
                                    // sectors in flight per SM: one 5-word gather per lane is plenty)
constexpr int kLogCap = 128;          // survivor-log entries per pass (specific -> sensitive reuse)
constexpr int kLogMaxLen = 1023;     // reads longer than this do not use the log (9-bit offset field)
constexpr int kTbLanesSm = 8;       // traceback words of lanes < 8 (band <= 16) stay in shared memory
constexpr int kTbCacheMaxCands = 4; // record traceback while scoring when a set has at most this many candidates
constexpr int kSetStateWords = 4;   // u64 words of scalar state in front of a stored candidate set
constexpr int kSetSlotsPe = 32;     // pe_candidates::max_size_small: a PE set that grew beyond it is redone
constexpr int kSetSlotsSe = 64;     // >= se_candidates::max_size

struct IndexDev {
  const uint64_t *genome;
  const uint32_t *counter, *counter_t, *counter_a;
  const uint32_t *index, *index_t, *index_a;
  // bit k set <=> bucket k of the table is non-empty (counter[k+1] != counter[k]); small enough to stay in L2.
  // A null pointer means "probe the counters directly" (dense tables, where the bitmap would save nothing).
  const uint32_t *bits, *bits_t, *bits_a;
  // 2-bit copy of the genome for the candidate compare: word w holds bases 32w..32w+31, bit j of the low half =
  // code bit 0 and of the high half = code bit 1 of base 32w+j (A0 C1 G2 T3; anything else reads as A and is
  // flagged).  gx: bit b set <=> the 256-base block b holds a base that is not A/C/G/T (N, IUPAC); candidates
  // whose window touches a flagged block take the exact 4-bit compare instead.
  const uint64_t *g2;
  const uint32_t *gx;
  // Seed-context records (prefilter of the candidate compare): for entry j of an index table, record
  // ctx[a * n_ctx + j] (32 bytes = one sector, g2 word format) holds the 128 genome bases starting at
  // entry - 32 a, a = 0..kCtxArrays-1.  A candidate found at seed offset i reads ONE record (a = min(i / 32,
  // kCtxArrays - 1)) instead of its index entry plus a random genome window; buckets are contiguous in every
  // array, so the candidates of a bucket share DRAM pages.  Null = not built (no memory, or disabled): every
  // candidate then takes the direct compare.
  const uint4 *ctx, *ctx_t, *ctx_a;
  uint64_t n_ctx, n_ctx3;
  // Compact two-letter counters: block b (32 bytes = one sector) = {counter[28 b], 28 one-byte bucket sizes},
  // 38 MB instead of 134 MB so that the table can be pinned in L2 (persisting access window) and a probe
  // costs no HBM access.  A block with a bucket of 255 or more entries has base ~0u: probe the full table.
  const uint4 *cc;
  uint32_t max_candidates;
  uint32_t window_size;  // seed::window_size the index was built with (20, or 12)
};
constexpr int kCtxArrays = 4;
constexpr uint32_t kCcKeys = 28;

// One banded alignment handed from enum_kernel to dp_kernel (align_tasks.cuh)
struct AlignTask {
  uint32_t t_pos;     // candidate position (centre diagonal of the band)
  uint32_t item;      // read / pair of the sub-batch
  uint32_t meta;      // bit 0: end; bit 1: reverse-complemented; bit 2: a-rich flag; bits 8..15: band width (0 = empty slot)
  uint32_t tb_index;  // traceback words at task_tb + 8 * tb_index, or kNoTask when not recorded
};
struct TaskResult {   // AlnOut of the task
  int16_t score, row, col, bw;
};
constexpr uint32_t kNoTask = 0xffffffffu;

// Binned seeding (seed_bins.cuh): the (strand, seed offset) probes of a whole batch as 16-byte tuples, binned by
// the address range of their seed-context records so that a slice of the records is L2-resident while every
// probe that touches it is filtered.
struct SeedTuple {
  uint32_t rec_lo;   // record number a * n_entries + first entry of the range, within the tuple's table (low 32 bits)
  uint32_t cnt_hi;   // bits 0..7: high bits of the record number; bits 8..31: candidates in the range
  uint32_t sid;      // strand: (read pair, pass) of the batch
  uint32_t meta;     // seed offset [0,10) | compared bases of the record [10,18) | bound [18,27) | three-letter table 27
                     // | examined by the specific phase 28 | by the sensitive phase 29 | g_to_a 30
};
struct BinParams {
  SeedTuple *tup;             // emission order (hash_kernel); nullptr = binned seeding off
  unsigned int *tup_count;    // slots handed out (in blocks; unused slots hold empty tuples)
  uint32_t tup_cap;
  uint32_t pw;                // words per read plane
  uint32_t *planes;           // [strand][pw][2]: 2-bit codes of the strand's read (A0 C1 G2 T3), {lo, hi} plane words of 32 bases
  uint8_t *strand_flag;       // [strand]: 0 = survivors listed, 1 = take the direct path (process_seeds), 2 = empty read
  uint32_t *bin_hist;         // (scatter kernels) tuples per (bin, scatter CTA)
  uint32_t bin_shift, n_bins; // bin = global record number >> bin_shift
  uint64_t rec_base[3];       // global record number of the first record of ctx, ctx_t, ctx_a
  uint32_t *surv_count;       // [strand] prefilter survivors found (may exceed surv_cap: then the strand takes the direct path)
  uint2 *surv;                // [strand][surv_cap]: {entry number in its index table, offset | table << 10 | spec << 11 | sens << 12}
  uint32_t surv_cap;
  uint32_t acc_cap;           // survivors a strand may keep for the short replay (<= 32); more: all are ranked and replayed
  uint32_t spi;               // strands per read / pair
  uint32_t sid_base;          // strand number of the first strand of this launch (sub-batches of one batch)
  uint32_t pad;
};

struct KernelParams {
  IndexDev ix;
  BinParams bp;
  // batch
  uint32_t n;
  const char *seq[2];
  const uint32_t *off[2];
  // results
  abg_hit *pe_r1, *pe_r2, *se[2];
  uint32_t *cigar[2];
  uint32_t *n_cigar[2];
  uint32_t *cigar_inline[2];  // [n][inline_ops]: the first operations of every CIGAR, the part copied back in bulk
  uint32_t cigar_stride, inline_ops;
  // params
  uint32_t mode, allow_ambig, min_dist, max_dist, max_candidates;
  double valid_frac;
  // per-warp-slot scratch
  uint32_t ml;            // padded max read length (multiple of 32)
  uint64_t *pe_overflow;  // [slots][2][kPeLarge]
  int16_t *mem_scr;       // [slots][kPeLarge]: best_pair's memo / alignment scores of set 2's partners
  int16_t *mem_scr2;      // [slots][kPeLarge]: alignment scores of set 3 (row-parallel best_pair)
  uint32_t heavy_min;     // sets with at least this many entries take the row-parallel best_pair
  uint64_t *tb;           // [slots][2][tb_words][32]
  uint32_t tb_words;      // 64-bit traceback words per lane per alignment
  unsigned int *work_counter;
  unsigned int *error_flag;
  unsigned long long *counters;  // abg_work_counters layout, or nullptr
  // two-phase launch (seed_kernel -> align_kernel -> map_reads_kernel on the redo list); null / 0 otherwise
  uint64_t *sets;            // [n][n_pass][kSetStateWords + set_slots]: candidate sets as the seed phase leaves them
  uint32_t set_slots;        // heap entries stored per set
  uint32_t n_pass;           // stored sets per read / pair
  unsigned int *redo_flag;   // [n]: the pair has a set that does not fit set_slots -> mapped by map_reads_kernel
  uint32_t *redo_list;       // items flagged by the seed phase, in no particular order
  unsigned int *redo_count;
  const uint32_t *item_list; // map_reads_kernel only: map item_list[0 .. *n_items_ptr) instead of 0 .. n
  const unsigned int *n_items_ptr;
  uint32_t window_size;      // seed::window_size of the index: 20, or 12 (--enable-short)
  uint32_t layout_kind;      // kLayoutFull / kLayoutSeed / kLayoutAlign: which regions the warp's shared memory holds
  uint32_t slot_base;        // first per-warp scratch slot of this launch (kernels that run concurrently get disjoint slots)
  // Overlapped launch: align_kernel runs next to seed_kernel and consumes pairs as their sets complete.
  unsigned int *ready;       // [n]: stored sets of the item so far (seed_kernel increments, release); null = stream order
  uint32_t ready_need;       // sets per item
  uint32_t wait_ns;          // how long an align warp waits for an item before leaving it to the redo kernel
  // task-parallel alignment (align_tasks.cuh); tasks == nullptr: every alignment runs in the warp of its pair
  AlignTask *tasks;          // three lists: class c (groups of 8 << c lanes) at task_base[c], task_cap[c] slots
  TaskResult *task_res;      // [task id]
  uint64_t *task_tb;         // traceback arena, addressed in units of 8 words
  uint32_t *task_of;         // [n][n_pass][set_slots]: task id of entry j of the sorted stored set, or kNoTask
  unsigned int *task_count;  // [0..2] slots handed out per class, [3] traceback units handed out
  unsigned int *task_cursor; // [0..2] dp_kernel's work cursors
  uint32_t task_base[3], task_cap[3];
  uint32_t task_tb_cap;      // units of 8 words
  // stored candidate sets larger than set_slots entries (repeats): the entries beyond live in an arena that the
  // seeding kernel allocates from; task_ovf holds their task ids at the same offsets
  uint64_t *set_ovf;
  uint32_t *task_ovf;
  unsigned int *ovf_count;
  uint32_t ovf_cap;
};

// ---- 64-bit view of se_element {int16 diffs; uint16 flags; uint32 pos} -------
struct Hit {
  uint64_t w;
  __device__ __forceinline__ Hit() : w(0) {}
  __device__ __forceinline__ explicit Hit(uint64_t x) : w(x) {}
  __device__ __forceinline__ Hit(int diffs, uint32_t flags, uint32_t pos)
    : w((uint64_t)(uint16_t)diffs | ((uint64_t)(flags & 0xffffu) << 16) | ((uint64_t)pos << 32)) {}
  __device__ __forceinline__ int diffs() const { return (int)(int16_t)(w & 0xffffu); }
  __device__ __forceinline__ uint32_t flags() const { return (uint32_t)(w >> 16) & 0xffffu; }
  __device__ __forceinline__ uint32_t pos() const { return (uint32_t)(w >> 32); }
  __device__ __forceinline__ bool empty() const { return pos() == 0; }
  __device__ __forceinline__ bool ambig() const { return (w >> 16) & ABG_FLAG_AMBIG; }
  __device__ __forceinline__ void set_ambig() { w |= (uint64_t)ABG_FLAG_AMBIG << 16; }
  __device__ __forceinline__ void set_diffs(int d) { w = (w & ~0xffffull) | (uint64_t)(uint16_t)d; }
  __device__ __forceinline__ void set_pos(uint32_t p) { w = (w & 0xffffffffull) | ((uint64_t)p << 32); }
  __device__ __forceinline__ void reset() {  // se_element::reset(): flags kept
    set_pos(0);
    set_diffs(kMaxDiffs);
  }
  // sort key of prepare_for_alignments / prepare_for_mating: (pos, flags)
  __device__ __forceinline__ uint64_t key() const { return ((uint64_t)pos() << 16) | flags(); }
};

// ---- thresholds: evaluated in double exactly as the reference writes them ----
__device__ __forceinline__ int frac_of(double f, uint32_t x) {  // static_cast<score_t>(f * x)
  return (int)(int16_t)__double2int_rz(__dmul_rn(f, (double)x));
}
__device__ __forceinline__ int invalid_hit_diffs(uint32_t readlen) { return frac_of(0.4, readlen); }
// min_len = ReadLoader::min_read_length = key_weight + window_size - 1 (abismal.cpp:212-213): 44, or 36
__device__ __forceinline__ bool valid_len(uint32_t aln_len, uint32_t readlen, uint32_t min_len) {  // abismal.cpp:307-314
  const double min_aln_frac = __dsub_rn(1.0, 0.4);
  const uint32_t a = (uint32_t)__double2uint_rz(__dmul_rn(min_aln_frac, (double)readlen));
  return aln_len >= (a > min_len ? a : min_len);
}

// ---- shared memory ----------------------------------------------------------------
// block: [KernelParams | base-3 tables | warp 0 | warp 1 | ...]
struct CandState {  // scalar state of one se_candidates / pe_candidates
  int sz, cutoff, good_cutoff, capacity;
  int sure_ambig, is_pe;
  uint64_t best;  // SE only
  uint32_t ovf;   // stored PE set: 1 + offset of its entries beyond set_slots in KernelParams::set_ovf, 0 = none
  uint32_t pad;
};

struct AlnOut {
  int score, row, col, bw;
};

struct TbKey {  // which alignment the traceback words of a slot belong to
  uint32_t pos, key, valid;
  int score, row, col, bw;
  uint32_t tb_index;  // kNoTask: the words are in the slot's own storage; else in the task arena (align_tasks.cuh)
};

struct WarpScalars {
  uint32_t len[2];
  uint32_t qkey[2];     // which (flags) is encoded in qcode[e]; ~0u = none
  uint32_t packed_key;  // which (end, flags) is in packed/planes; ~0u = none
  uint32_t loaded[2];   // binned seeding: base[e] holds the read's bases (the replay loads them only when needed)
  uint32_t pad;
  TbKey tbk[2];
  unsigned long long cnt[6];
};

constexpr int kParamBytes = 768;
constexpr int kTab3Bytes = 2 * 256 * 4;
static_assert(sizeof(KernelParams) <= kParamBytes, "KernelParams must fit its shared-memory slot");

__host__ __device__ __forceinline__ uint32_t tb_sm_words(uint32_t ml) { return (ml + 64u + 32u + 15u) / 16u; }

// Shared-memory regions per warp.  The three kernels keep different subsets (layout_kind): the seeding kernel
// needs no traceback words, no reference bytes and one candidate set; the alignment kernel needs no log, masks
// or planes.
constexpr uint32_t kLayoutFull = 0, kLayoutSeed = 1, kLayoutAlign = 2;
// Smaller instruction footprint (the kernels stall on instruction fetch: L0 ~6 KB, L1.5 32 KB against 100-160 KB
// of code): one wavefront DP function with a run-time traceback flag, a rolled reference-staging loop,
// out-of-line align() and deep compare.
#ifndef ABG_SMALL_CODE
#define ABG_SMALL_CODE 1  // measured: seed_kernel 96.0 vs 100.5 ms, align_kernel 47.7 vs 50.5 ms per 2^20 pairs
#endif

struct WarpLayout {
  uint32_t o_packed, o_masks, o_planes, o_se, o_pe, o_cs, o_scal, o_tb, o_log, o_elig, o_base, o_qcode, o_refb, o_stage, total;
  uint32_t plane_words, elig_words, mask_words;
};

__host__ __device__ __forceinline__ WarpLayout warp_layout(uint32_t ml, bool paired, uint32_t kind) {
  WarpLayout L;
  const bool seed = kind != kLayoutAlign, aln = kind != kLayoutSeed;
  uint32_t o = 0;
  L.o_packed = o;
  o += ml / 2;                               // packed read, ml/16 u64
  L.o_se = o;
  if (aln || !paired) o += 2u * kSeSlots * 8u;  // two SE heaps
  L.o_pe = o;
  if (paired) o += (aln ? 2u : 1u) * kPeSmemSlots * 8u;   // PE heap heads (the seeding kernel fills one set)
  L.o_cs = o;
  o += 4u * (uint32_t)sizeof(CandState);
  L.o_scal = o;
  o += (uint32_t)sizeof(WarpScalars);
  L.o_tb = o;
  if (aln) o += 2u * tb_sm_words(ml) * kTbLanesSm * 8u;  // traceback words of 2 slots
  L.o_log = o;
  if (seed) o += 2u * kLogCap * 4u;          // survivor log: positions + packed (d, pm, table, offset)
  L.o_stage = o;
  if (seed) o += 32u * 12u;                  // survivors of one compare round, staged for the single-lane heap replay
  L.elig_words = ml / 64u + 1u;              // specific offsets are < readlen / 2
  L.o_elig = o;
  if (seed) o += 2u * L.elig_words * 4u;
  L.mask_words = ml / 32u + 1u;
  L.o_masks = o;
  if (seed) o += 4u * L.mask_words * 4u;     // read-vs-{A,C,G,T} match masks per 32-base chunk
  L.plane_words = ml / 32u + 2u;
  L.o_planes = o;
  if (seed) o += 3u * L.plane_words * 4u;
  L.o_base = o;
  o += 2u * ml;
  L.o_qcode = o;
  o += 2u * (ml + 32u);
  L.o_refb = o;
  if (aln) o += ml + 64u + 32u;              // reference bytes of the DP
  L.total = (o + 15u) & ~15u;
  return L;
}

__host__ __device__ __forceinline__ size_t block_smem_bytes(uint32_t ml, bool paired, uint32_t kind) {
  return (size_t)kParamBytes + kTab3Bytes + (size_t)warp_layout(ml, paired, kind).total * kWarpsPerBlock;
}

extern __shared__ __align__(16) unsigned char smem_raw[];

__device__ __forceinline__ const KernelParams &params() { return *reinterpret_cast<const KernelParams *>(smem_raw); }
__device__ __forceinline__ const uint32_t *tab3() { return reinterpret_cast<const uint32_t *>(smem_raw + kParamBytes); }

struct Warp {  // view of this warp's shared memory; rebuilt (cheaply) inside every function
  unsigned char *base_ptr;
  WarpLayout L;
  int lane;
  __device__ __forceinline__ Warp() {
    const KernelParams &P = params();
    L = warp_layout(P.ml, (P.mode & ABG_MODE_PAIRED) != 0, P.layout_kind);
    base_ptr = smem_raw + kParamBytes + kTab3Bytes + (size_t)L.total * (threadIdx.x >> 5);
    lane = threadIdx.x & 31;
  }
  __device__ __forceinline__ uint64_t *packed() const { return reinterpret_cast<uint64_t *>(base_ptr + L.o_packed); }
  __device__ __forceinline__ uint64_t *se_heap(int i) const { return reinterpret_cast<uint64_t *>(base_ptr + L.o_se) + i * kSeSlots; }
  __device__ __forceinline__ uint64_t *pe_heap(int i) const { return reinterpret_cast<uint64_t *>(base_ptr + L.o_pe) + i * kPeSmemSlots; }
  __device__ __forceinline__ CandState *cs(int id) const { return reinterpret_cast<CandState *>(base_ptr + L.o_cs) + id; }
  __device__ __forceinline__ WarpScalars *scal() const { return reinterpret_cast<WarpScalars *>(base_ptr + L.o_scal); }
  __device__ __forceinline__ uint64_t *tb_sm(int slot) const {
    return reinterpret_cast<uint64_t *>(base_ptr + L.o_tb) + (size_t)slot * tb_sm_words(params().ml) * kTbLanesSm;
  }
  __device__ __forceinline__ uint32_t *plane(int k) const { return reinterpret_cast<uint32_t *>(base_ptr + L.o_planes) + k * L.plane_words; }
  __device__ __forceinline__ uint8_t *base(int e) const { return base_ptr + L.o_base + (size_t)e * params().ml; }
  __device__ __forceinline__ uint8_t *qcode(int e) const { return base_ptr + L.o_qcode + (size_t)e * (params().ml + 32u); }
  __device__ __forceinline__ uint8_t *refb() const { return base_ptr + L.o_refb; }
  __device__ __forceinline__ uint32_t *masks(int k) const { return reinterpret_cast<uint32_t *>(base_ptr + L.o_masks) + k * L.mask_words; }
  __device__ __forceinline__ uint32_t *log_pos() const { return reinterpret_cast<uint32_t *>(base_ptr + L.o_log); }
  __device__ __forceinline__ uint32_t *log_meta() const { return log_pos() + kLogCap; }
  __device__ __forceinline__ uint32_t *stage() const { return reinterpret_cast<uint32_t *>(base_ptr + L.o_stage); }
  __device__ __forceinline__ uint32_t *elig(int k) const { return reinterpret_cast<uint32_t *>(base_ptr + L.o_elig) + k * L.elig_words; }
  __device__ __forceinline__ size_t slot() const {
    return (size_t)params().slot_base + (size_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  }
  __device__ __forceinline__ uint64_t *tb_gm(int s) const {
    const KernelParams &P = params();
    return P.tb + (slot() * 2 + s) * (size_t)P.tb_words * 32;
  }
};

// candidate set ids: 0,1 = se_candidates of end 1 / end 2; 2,3 = the two pe_candidates of a map_fragments call
struct HeapRef {
  uint64_t *sm;
  uint64_t *gm;
  int cap_sm;
  __device__ __forceinline__ Hit get(int i) const { return Hit(i < cap_sm ? sm[i] : gm[i]); }
  __device__ __forceinline__ void set(int i, Hit h) const {
    if (i < cap_sm) sm[i] = h.w;
    else gm[i] = h.w;
  }
};

__device__ __forceinline__ HeapRef heap_of(const Warp &W, int id) {
  if (id < 2) return HeapRef{W.se_heap(id), nullptr, kSeSlots};
  uint64_t *ov = params().pe_overflow + W.slot() * (size_t)2 * kPeLarge;
  return HeapRef{W.pe_heap(id - 2), ov + (size_t)(id - 2) * kPeLarge, kPeSmemSlots};
}

// ---- candidate set: register working copy of CandState + heap, warp-uniform ------------
struct CandSet {
  HeapRef v;
  int sz, cutoff, good_cutoff, capacity;
  bool sure_ambig, is_pe;
  Hit best;  // SE only

  __device__ __forceinline__ void load(const Warp &W, int id) {
    v = heap_of(W, id);
    const CandState *s = W.cs(id);
    sz = s->sz;
    cutoff = s->cutoff;
    good_cutoff = s->good_cutoff;
    capacity = s->capacity;
    sure_ambig = s->sure_ambig != 0;
    is_pe = s->is_pe != 0;
    best = Hit(s->best);
  }
  __device__ __forceinline__ void store(const Warp &W, int id) const {  // same values from every lane
    CandState *s = W.cs(id);
    __syncwarp();
    if (W.lane == 0) {
      s->sz = sz;
      s->cutoff = cutoff;
      s->good_cutoff = good_cutoff;
      s->capacity = capacity;
      s->sure_ambig = sure_ambig;
      s->is_pe = is_pe;
      s->best = best.w;
    }
    __syncwarp();
  }

  __device__ __forceinline__ void store_one_lane(const Warp &W, int id) const {  // caller: one lane, syncs around it
    CandState *s = W.cs(id);
    s->sz = sz;
    s->cutoff = cutoff;
    s->good_cutoff = good_cutoff;
    s->capacity = capacity;
    s->sure_ambig = sure_ambig;
    s->is_pe = is_pe;
    s->best = best.w;
  }

  __device__ __forceinline__ bool full() const { return sz == (is_pe ? capacity : kSeMax); }

  __device__ void sift_up(int hole, Hit val) const {  // std::__push_heap, top = 0
    int parent = (hole - 1) / 2;
    while (hole > 0) {
      const Hit p = v.get(parent);
      if (!(p.diffs() < val.diffs())) break;
      v.set(hole, p);
      hole = parent;
      parent = (hole - 1) / 2;
    }
    v.set(hole, val);
  }
  __device__ void heap_pop(int n) const {  // std::pop_heap(v, v + n)
    if (n <= 1) return;
    const Hit val = v.get(n - 1);
    v.set(n - 1, v.get(0));
    const int len = n - 1;
    int hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      Hit c = v.get(child);
      const Hit c1 = v.get(child - 1);
      if (c.diffs() < c1.diffs()) {
        --child;
        c = c1;
      }
      v.set(hole, c);
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      v.set(hole, v.get(child - 1));
      hole = child - 1;
    }
    sift_up(hole, val);
  }

  __device__ void reset_se(uint32_t readlen) {  // se_candidates::reset(readlen) :417-427
    is_pe = false;
    best = Hit(invalid_hit_diffs(readlen), 0, 0);
    v.set(0, Hit(invalid_hit_diffs(readlen), 0, 0));
    cutoff = invalid_hit_diffs(readlen);
    good_cutoff = (int)(int16_t)(readlen / 10u);
    sure_ambig = false;
    sz = 1;
    capacity = kSeMax;
  }
  __device__ void reset_se_noarg() {  // se_candidates::reset() :406-415
    best.reset();
    v.set(0, Hit(kMaxDiffs, 0, 0));
    cutoff = kMaxDiffs;
    sure_ambig = false;
    sz = 1;
  }
  __device__ void reset_pe(uint32_t readlen) {  // pe_candidates::reset :778-787
    is_pe = true;
    best = Hit(0);
    v.set(0, Hit(invalid_hit_diffs(readlen), 0, 0));
    sure_ambig = false;
    cutoff = invalid_hit_diffs(readlen);
    good_cutoff = (int)(int16_t)(readlen / 10u);
    sz = 1;
    capacity = kPeSmall;
  }
  __device__ __forceinline__ void set_specific() { cutoff = good_cutoff; }
  __device__ __forceinline__ void set_sensitive() { cutoff = v.get(0).diffs(); }
  __device__ __forceinline__ bool should_do_sensitive() const {
    return is_pe ? (capacity == kPeSmall || cutoff > good_cutoff) : (sz != kSeMax || cutoff > good_cutoff);
  }
  __device__ __forceinline__ bool should_align() const { return sz != kPeLarge || cutoff != 0; }

  // se_candidates::update :394-404 / pe_candidates::update :824-842
  __device__ __forceinline__ void update(bool specific, int d, uint32_t flags, uint32_t pos) {
    if (!is_pe) {
      if (d == 0) {
        if (best.empty()) best = Hit(0, flags, pos);
        else if (pos != best.pos() || flags != best.flags()) best.set_ambig();
      }
      else {
        if (sz == kSeMax) {
          heap_pop(sz);
          v.set(sz - 1, Hit(d, flags, pos));
        }
        else v.set(sz++, Hit(d, flags, pos));
        sift_up(sz - 1, Hit(d, flags, pos));
      }
      sure_ambig = best.ambig() && best.diffs() == 0;
      const int top = v.get(0).diffs();
      cutoff = specific ? min(cutoff, top) : top;
    }
    else {
      if (sz == capacity) {
        if (specific && capacity != kPeLarge && d <= good_cutoff) ++capacity;
        else {
          heap_pop(sz);
          --sz;
        }
      }
      v.set(sz++, Hit(d, flags, pos));
      sift_up(sz - 1, Hit(d, flags, pos));
      const int top = v.get(0).diffs();
      cutoff = specific ? min(cutoff, top) : top;
      sure_ambig = (sz == capacity) && cutoff == 0;
    }
  }
};

__device__ __forceinline__ uint32_t get_bit(uint32_t nt) { return (nt & 5u) == 0u; }
__device__ __forceinline__ uint32_t three_num(bool g_to_a, uint32_t nt) {
  return g_to_a ? ((((nt & 8u) != 0u) << 1) | ((nt & 2u) != 0u)) : ((((nt & 4u) != 0u) << 1) | ((nt & 1u) != 0u));
}
__device__ __forceinline__ uint32_t three_fast(bool g_to_a, uint32_t nt) { return g_to_a ? (nt & 10u) : (nt & 5u); }
__device__ __forceinline__ uint32_t genome_base(const uint64_t *g, uint64_t pos) {
  return (uint32_t)(__ldg(g + (pos >> 4)) >> ((pos & 15u) << 2)) & 15u;
}

// Load one end of a read/pair: ASCII -> one-hot nibble (A1 C2 G4 T8, else 0).
__device__ __forceinline__ void load_end(const Warp &W, int end, const char *s, uint32_t n) {
  uint8_t *b = W.base(end);
  for (uint32_t i = W.lane; i < n; i += 32) {
    const char ch = s[i];
    uint8_t x = 0;
    if (ch == 'A' || ch == 'a') x = 1;
    else if (ch == 'C' || ch == 'c') x = 2;
    else if (ch == 'G' || ch == 'g') x = 4;
    else if (ch == 'T' || ch == 't') x = 8;
    b[i] = x;
  }
  __syncwarp();
}

// prep_read for the pass identified by `flags` on `end` into qcode[end]:
// orientation by the rc bit, encoding by a_rich XOR rc (abismal.cpp:1463-1465).
__device__ __noinline__ void build_qcode(int end, uint32_t flags) {
  const Warp W;
  WarpScalars *S = W.scal();
  const uint32_t key = flags & (ABG_FLAG_RC | ABG_FLAG_A_RICH);
  if (S->qkey[end] == key) return;
  __syncwarp();
  const bool rc = flags & ABG_FLAG_RC;
  const bool enc_a = ((flags & ABG_FLAG_A_RICH) != 0) != rc;
  const uint32_t n = S->len[end];
  const uint8_t *b = W.base(end);
  uint8_t *q = W.qcode(end);
  for (uint32_t i = W.lane; i < n + 32; i += 32) {
    uint32_t code = 0;
    if (i < n) {
      uint32_t x = rc ? b[n - 1 - i] : b[i];
      if (rc) x = (__brev(x) >> 28);  // complement of a one-hot nibble = bit reversal (A1<->T8, C2<->G4)
      code = enc_a ? (x == 1u ? 5u : x) : (x == 8u ? 10u : x);
    }
    q[i] = (uint8_t)code;
  }
  if (W.lane == 0) S->qkey[end] = key;
  __syncwarp();
}

// pack_read (abismal.cpp:1393-1426) + the three hash bit planes for qcode[end]
__device__ __noinline__ void build_packed(int end, uint32_t flags) {
  const Warp W;
  WarpScalars *S = W.scal();
  const uint32_t key = ((uint32_t)end << 16) | (flags & (ABG_FLAG_RC | ABG_FLAG_A_RICH));
  if (S->packed_key == key) return;
  __syncwarp();
  const bool g_to_a = ((flags & ABG_FLAG_A_RICH) != 0) != ((flags & ABG_FLAG_RC) != 0);
  const uint32_t n = S->len[end];
  const uint8_t *q = W.qcode(end);
  uint64_t *packed = W.packed();
  const uint32_t nw = (n + 15) / 16;
  for (uint32_t w = W.lane; w < nw; w += 32) {
    uint64_t word = 0;
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) {
      const uint32_t i = 16 * w + j;
      const uint64_t nib = i < n ? q[i] : 0xFull;  // tail matches anything :1424-1425
      word |= nib << (4 * j);
    }
    packed[w] = word;
  }
  // planes: bit p of word w describes base 32w+p; bases past the end encode as 0 (zero padding)
  uint32_t *p2 = W.plane(0), *p3a = W.plane(1), *p3b = W.plane(2);
  const uint32_t pw = W.L.plane_words;
  for (uint32_t w = 0; w < pw; ++w) {
    const uint32_t i = 32 * w + W.lane;
    const uint32_t code = i < n ? q[i] : 0u;
    const uint32_t t = three_num(g_to_a, code);
    const unsigned m2 = __ballot_sync(FULL, get_bit(code));
    const unsigned ma = __ballot_sync(FULL, t & 1u);
    const unsigned mb = __ballot_sync(FULL, t & 2u);
    if (W.lane == 0) {
      p2[w] = m2;
      p3a[w] = ma;
      p3b[w] = mb;
    }
  }
  // match masks against the 2-bit genome: bit j of masks(X)[c] <=> read base 32c+j matches genome base X
  // (code & one-hot(X)); the 0xF tail of the last packed word matches everything, positions past it nothing.
  {
    uint32_t *mA = W.masks(0), *mC = W.masks(1), *mG = W.masks(2), *mT = W.masks(3);
    const uint32_t n_tail = 16u * nw;
    for (uint32_t c = 0; c < W.L.mask_words; ++c) {
      const uint32_t i = 32 * c + W.lane;
      const uint32_t code = i < n ? q[i] : (i < n_tail ? 0xFu : 0u);
      const unsigned a = __ballot_sync(FULL, code & 1u), cc = __ballot_sync(FULL, code & 2u);
      const unsigned g = __ballot_sync(FULL, code & 4u), t = __ballot_sync(FULL, code & 8u);
      if (W.lane == 0) {
        mA[c] = a;
        mC[c] = cc;
        mG[c] = g;
        mT[c] = t;
      }
    }
  }
  if (W.lane == 0) S->packed_key = key;
  __syncwarp();
}

// std::lower_bound over idx[low, high): first entry whose genome base at
// entry + p does not satisfy pred(base) < val
template <class F>
__device__ __forceinline__ uint32_t lower_bound_idx(const uint32_t *idx, uint32_t low, uint32_t high, F less_than) {
  int len = (int)(high - low);
  uint32_t first = low;
  while (len > 0) {
    const int half = len >> 1;
    const uint32_t mid = first + (uint32_t)half;
    if (less_than(__ldg(idx + mid))) {
      first = mid + 1;
      len = len - half - 1;
    }
    else len = half;
  }
  return first;
}

struct SeedRange {  // narrowed bucket [low, high) and the seed length reached
  uint32_t low, high, p;
};

// A read shorter than 25 + its seed offset never meets `p == read_lim`; the extension then stops inside the
// genome's 32767-base end padding at the latest (only reachable on read lengths the reference handles by
// reading past its buffers).
constexpr uint32_t kMaxExtend = 32000u;

// find_candidates<25> (abismal.cpp:1163-1194); `read_start` = qcode + i
__device__ __noinline__ SeedRange find_candidates(uint32_t maxc, const uint8_t *read_start, uint32_t read_lim,
                                                  uint32_t low, uint32_t high) {
  const IndexDev &ix = params().ix;
  uint32_t p = 25;
  uint32_t prev_low = low, prev_high = high;
  for (; p != read_lim && p < kMaxExtend && (high - low) > maxc; ++p) {
    prev_low = low;
    prev_high = high;
    const uint32_t first_1 = lower_bound_idx(ix.index, low, high, [&](uint32_t e) {
      return get_bit(genome_base(ix.genome, (uint64_t)e + p)) < 1u;
    });
    // bases past the end of the read are 0 (the reference reads whatever follows its vector there: the
    // specific phase of reads shorter than 2 * window_size + 8 runs offsets whose 25-mer overhangs the read)
    const uint32_t the_bit = get_bit(p < read_lim ? read_start[p] : 0u);
    high = the_bit ? high : first_1;
    low = the_bit ? first_1 : low;
  }
  if (low == high) {
    --p;
    low = prev_low;
    high = prev_high;
  }
  return SeedRange{low, high, p};
}

// find_candidates_three<16, conv> (abismal.cpp:1214-1259)
__device__ __noinline__ SeedRange find_candidates_three(const uint32_t *index3, bool g_to_a, uint32_t maxc,
                                                        const uint8_t *read_start, uint32_t max_size, uint32_t low,
                                                        uint32_t high) {
  const IndexDev &ix = params().ix;
  uint32_t p = 16;
  uint32_t prev_low = low, prev_high = high;
  const uint32_t v1 = g_to_a ? 2u : 1u, v2 = g_to_a ? 8u : 4u;
  for (; p != max_size && p < kMaxExtend && (high - low) > maxc; ++p) {
    prev_low = low;
    prev_high = high;
    const uint32_t first_1 = lower_bound_idx(index3, low, high, [&](uint32_t e) {
      return three_fast(g_to_a, genome_base(ix.genome, (uint64_t)e + p)) < v1;
    });
    const uint32_t first_2 = lower_bound_idx(index3, low, high, [&](uint32_t e) {
      return three_fast(g_to_a, genome_base(ix.genome, (uint64_t)e + p)) < v2;
    });
    const uint32_t the_num = three_fast(g_to_a, p < max_size ? read_start[p] : 0u);
    const uint32_t old_low = low, old_high = high;
    high = (the_num == 0u) ? first_1 : ((the_num == v1) ? first_2 : old_high);
    low = (the_num == 0u) ? old_low : ((the_num == v1) ? first_1 : first_2);
  }
  if (low == high) {
    --p;
    low = prev_low;
    high = prev_high;
  }
  return SeedRange{low, high, p};
}

__device__ __forceinline__ uint32_t warp_incl_scan_add(uint32_t x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// 32 plane bits starting at base i (bit j of the result describes base i + j)
__device__ __forceinline__ uint32_t plane_window(const uint32_t *pl, uint32_t i) {
  const uint32_t w = i >> 5;
  return __funnelshift_r(pl[w], pl[w + 1], i & 31u);
}

// packed word w of the genome window starting at base the_pos: 16 bases from word g0 (low) and g1
__device__ __forceinline__ uint64_t window_word(uint64_t g0, uint64_t g1, uint32_t off) {
  return (g0 >> off) | ((g1 << (63u - off)) << 1);
}

// full_compare (abismal.cpp:1105-1122) on the 4-bit genome, word by word.  It stops at the first word after
// which the running distance exceeds the cutoff, and a word can contribute a NEGATIVE amount (multi-bit IUPAC
// genome codes give popcounts > 16): a candidate is accepted iff the MAXIMUM prefix sum pm is <= the cutoff,
// and then d is the full sum.  Used only for candidates whose window holds a base that is not A/C/G/T.
__device__ __noinline__ int exact_compare(uint32_t the_pos, int n_words, int bound, int *pm_out) {
  const Warp W;
  const uint64_t *packed = W.packed();
  const uint64_t *g = params().ix.genome + (the_pos >> 4);
  const uint32_t off = (the_pos & 15u) << 2;
  int d = 0, pm = 0;
  uint64_t cur = __ldg(g);
  for (int w = 0; w < n_words && pm <= bound; ++w) {
    const uint64_t nxt = __ldg(g + w + 1);
    d += 16 - __popcll(packed[w] & window_word(cur, nxt, off));
    pm = max(pm, d);
    cur = nxt;
  }
  *pm_out = pm;
  return d;
}

// 32 match bits of chunk c: genome code planes (lo, hi) of the 32 bases select among the read's four masks
__device__ __forceinline__ uint32_t match_bits(uint32_t lo, uint32_t hi, uint32_t mA, uint32_t mC, uint32_t mG,
                                               uint32_t mT) {
  const uint32_t x0 = (lo & mC) | (~lo & mA);  // hi == 0: A or C
  const uint32_t x1 = (lo & mT) | (~lo & mG);  // hi == 1: G or T
  return (hi & x1) | (~hi & x0);
}

// One compare chunk: KC candidates per lane, candidate (k, lane) = canonical index c0 + 32k + lane, compared
// against the 2-bit genome.  NC0 = 32-base chunks compared in the first stage (NC0 + 1 words gathered at once
// per candidate, all before any use); later stages add two chunks at a time while the distance is within the
// bound.  Every base pair contributes 0 or 1 here, so the running distance is monotone and pm == d.
// One 32-byte seed-context record (one sector)
__device__ __forceinline__ void load_ctx(const uint4 *p, uint32_t (&w)[8]) {
#ifdef ABG_LDG256
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
#else
  const uint4 x = __ldg(p), y = __ldg(p + 1);
  w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
  w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
#endif
}

constexpr int kDeferredExact = -2;  // pm of a candidate whose window needs exact_compare, left to the caller
// Deep part of a compare chunk: index entries, 2-bit genome windows (staged, with early exit at the bound) and
// the exact 4-bit compare for windows that hold N / IUPAC codes, for the candidates still `valid`.
template <int KC, int NC0>
__device__ __forceinline__ void compare_deep(const IndexDev &ix, const uint32_t *__restrict__ index3,
                                             const uint32_t *mA, const uint32_t *mC, const uint32_t *mG,
                                             const uint32_t *mT, int n_words, int bound, const bool (&valid)[KC],
                                             const uint32_t (&slot)[KC], const uint32_t (&sub)[KC], int (&d)[KC],
                                             int (&pm)[KC], uint32_t (&the_pos)[KC], uint32_t &n_entry,
                                             uint32_t &n_word, bool defer_exact = false) {
  const int n_bases = 16 * n_words;
  const int n_chunks = (n_bases + 31) >> 5;
  // ---- index gathers of the candidates the records did not reject (all of them without records) ----
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    uint32_t entry = 0;
    if (valid[k]) entry = (sub[k] >> 31) ? __ldg(index3 + slot[k]) : __ldg(ix.index + slot[k]);
    the_pos[k] = entry - (sub[k] & 0x7fffffffu);
  }

  // ---- stage 0: NC0 + 1 words of the 2-bit genome per candidate + its exception bits ----
  uint64_t g[KC][NC0 + 1];
  bool exc[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const uint64_t *gp = ix.g2 + (the_pos[k] >> 5);
#pragma unroll
    for (int j = 0; j <= NC0; ++j) g[k][j] = (valid[k] && j <= n_chunks) ? __ldg(gp + j) : 0ull;
    exc[k] = false;
    if (valid[k]) {
      const uint32_t b0 = the_pos[k] >> 8, b1 = (the_pos[k] + (uint32_t)n_bases - 1u) >> 8;
      if (b1 - b0 <= 1u && (b0 >> 5) == (b1 >> 5))  // the usual case: one probe of the (L2-resident) bitmap
        exc[k] = ((__ldg(ix.gx + (b0 >> 5)) >> (b0 & 31u)) & (1u | (1u << (b1 - b0)))) != 0u;
      else
        for (uint32_t b = b0; b <= b1; ++b) exc[k] = exc[k] || ((__ldg(ix.gx + (b >> 5)) >> (b & 31u)) & 1u);
    }
  }
  uint64_t carry[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const uint32_t sh = the_pos[k] & 31u;
    int dd = 0;
#pragma unroll
    for (int c = 0; c < NC0; ++c)
      if (c < n_chunks) {
        const uint32_t lo = __funnelshift_r((uint32_t)g[k][c], (uint32_t)g[k][c + 1], sh);
        const uint32_t hi = __funnelshift_r((uint32_t)(g[k][c] >> 32), (uint32_t)(g[k][c + 1] >> 32), sh);
        dd += min(32, n_bases - 32 * c) - __popc(match_bits(lo, hi, mA[c], mC[c], mG[c], mT[c]));
      }
    d[k] = dd;
    pm[k] = valid[k] ? dd : (1 << 30);
    carry[k] = g[k][NC0];
    if (valid[k]) {
      n_entry += 1;
      n_word += (uint32_t)min(NC0, n_chunks);
    }
  }
  // ---- later stages: two more chunks for every candidate still within the bound ----
  for (int c = NC0; c < n_chunks; c += 2) {
    uint64_t a[KC], b[KC];
    bool alive[KC], any = false;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      alive[k] = pm[k] <= bound && !exc[k];
      any = any || alive[k];
      const uint64_t *gp = ix.g2 + (the_pos[k] >> 5) + c;
      a[k] = alive[k] ? __ldg(gp + 1) : 0ull;
      b[k] = (alive[k] && c + 1 < n_chunks) ? __ldg(gp + 2) : 0ull;
    }
    if (!any) break;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (alive[k]) {
        const uint32_t sh = the_pos[k] & 31u;
        uint32_t lo = __funnelshift_r((uint32_t)carry[k], (uint32_t)a[k], sh);
        uint32_t hi = __funnelshift_r((uint32_t)(carry[k] >> 32), (uint32_t)(a[k] >> 32), sh);
        int dd = d[k] + min(32, n_bases - 32 * c) - __popc(match_bits(lo, hi, mA[c], mC[c], mG[c], mT[c]));
        if (c + 1 < n_chunks) {
          lo = __funnelshift_r((uint32_t)a[k], (uint32_t)b[k], sh);
          hi = __funnelshift_r((uint32_t)(a[k] >> 32), (uint32_t)(b[k] >> 32), sh);
          dd += min(32, n_bases - 32 * (c + 1)) -
                __popc(match_bits(lo, hi, mA[c + 1], mC[c + 1], mG[c + 1], mT[c + 1]));
        }
        d[k] = dd;
        pm[k] = dd;
        carry[k] = b[k];
        n_word += (c + 1 < n_chunks) ? 2u : 1u;
      }
    }
  }
  // ---- windows holding N / IUPAC bases: the exact 4-bit compare (rare) ----
#pragma unroll
  for (int k = 0; k < KC; ++k)
    if (exc[k]) {
      if (defer_exact) {  // the caller has no packed read yet: it runs exact_compare itself (kDeferredExact)
        pm[k] = kDeferredExact;
        continue;
      }
      int mx = 0;
      d[k] = exact_compare(the_pos[k], n_words, bound, &mx);
      pm[k] = mx;
    }
}

// Out-of-line deep compare of one candidate per lane, everything by value (survivors of the prefilter are rare;
// keeping this out of the gather loop keeps the loop small, and by-value keeps its state in registers).
struct Deep1 {
  int d, pm;
  uint32_t pos, n_entry, n_word;
};
__device__ __noinline__ Deep1 compare_deep_one(const uint32_t *__restrict__ index3, int n_words, int bound, bool valid_in,
                                               uint32_t slot_in, uint32_t sub_in, bool defer_exact = false) {
  const Warp W;
  const bool valid[1] = {valid_in};
  const uint32_t slot[1] = {slot_in}, sub[1] = {sub_in};
  int d[1] = {0}, pm[1] = {1 << 30};
  uint32_t the_pos[1] = {0u};
  Deep1 r;
  r.n_entry = 0;
  r.n_word = 0;
  compare_deep<1, 4>(params().ix, index3, W.masks(0), W.masks(1), W.masks(2), W.masks(3), n_words, bound, valid, slot, sub,
                     d, pm, the_pos, r.n_entry, r.n_word, defer_exact);
  r.d = d[0];
  r.pm = pm[0];
  r.pos = the_pos[0];
  return r;
}

template <int KC, int NC0>
__device__ __forceinline__ void compare_chunk(const IndexDev &ix, const uint32_t *__restrict__ index3,
                                              const uint4 *__restrict__ ctx3,
                                              const uint32_t *mA, const uint32_t *mC, const uint32_t *mG,
                                              const uint32_t *mT, int n_words, int bound, uint32_t c0,
                                              uint32_t total, uint32_t base_off, uint32_t incl, uint32_t tot,
                                              uint32_t n2, uint32_t s2, uint32_t s3, int lane, int (&d)[KC],
                                              int (&pm)[KC], uint32_t (&the_pos)[KC], uint32_t (&sub)[KC],
                                              uint32_t &n_entry,
                                              uint32_t &n_word) {
  bool valid[KC];
  const int n_bases = 16 * n_words;          // compared positions incl. the 0xF tail of the last packed word
  const int n_chunks = (n_bases + 31) >> 5;
  // ---- owners; seed-context prefilter (one sector per candidate) ----
  uint32_t slot[KC];  // position of the candidate's entry in its index table
  bool any_left = false;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const uint32_t cidx = c0 + 32u * k + lane;
    valid[k] = cidx < total;
    // owner lane = number of lanes whose inclusive sum is <= cidx
    int o = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const uint32_t vv = __shfl_sync(FULL, incl, (o + s - 1) & 31);
      if (vv <= cidx) o += s;
    }
    o &= 31;
    const uint32_t o_incl = __shfl_sync(FULL, incl, o);
    const uint32_t o_tot = __shfl_sync(FULL, tot, o);
    const uint32_t o_n2 = __shfl_sync(FULL, n2, o);
    const uint32_t o_s2 = __shfl_sync(FULL, s2, o);
    const uint32_t o_s3 = __shfl_sync(FULL, s3, o);
    const uint32_t r = cidx - (o_incl - o_tot);
    const bool three = valid[k] && r >= o_n2;
    slot[k] = three ? o_s3 + (r - o_n2) : o_s2 + r;
    // seed offset of the candidate; bit 31 = it came from the three-letter bucket
    sub[k] = (base_off + (uint32_t)o) | (three ? 0x80000000u : 0u);
    d[k] = 0;
    pm[k] = 1 << 30;
    the_pos[k] = 0;
    const uint4 *tab = three ? ctx3 : ix.ctx;
    if (valid[k] && tab != nullptr) {
      // Lower bound of the distance from the 128 genome bases of the record: every compared position adds 0
      // or 1 (entries whose window could hold an IUPAC code carry the sentinel record; an N reads as A here and
      // really mismatches), so more than `bound` mismatches on a subset of the positions already rejects the
      // candidate.
      const uint32_t i_off = base_off + (uint32_t)o;
      const uint32_t a = min((uint32_t)(kCtxArrays - 1), i_off >> 5);
      const uint32_t q0 = i_off - 32u * a;  // read position of the record's first base
      uint32_t w[8];
      load_ctx(tab + 2 * ((uint64_t)a * (three ? ix.n_ctx3 : ix.n_ctx) + slot[k]), w);
      // all-ones = sentinel of an entry near an IUPAC code (seed_context_kernel): no lower bound, compare exactly
      const bool sentinel = (w[0] & w[1] & w[2] & w[3] & w[4] & w[5] & w[6] & w[7]) == ~0u;
      int lb = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int qb = (int)q0 + 32 * c;
        const int nb = n_bases - qb;
        if (nb > 0) {
          const uint32_t wi = (uint32_t)qb >> 5, sh = (uint32_t)qb & 31u;
          const uint32_t a_ = __funnelshift_r(mA[wi], mA[wi + 1], sh), c_ = __funnelshift_r(mC[wi], mC[wi + 1], sh);
          const uint32_t g_ = __funnelshift_r(mG[wi], mG[wi + 1], sh), t_ = __funnelshift_r(mT[wi], mT[wi + 1], sh);
          const uint32_t mm = ~match_bits(w[2 * c], w[2 * c + 1], a_, c_, g_, t_);
          lb += __popc(nb >= 32 ? mm : (mm & ((1u << nb) - 1u)));
        }
      }
      if (lb > bound && !sentinel) {
        valid[k] = false;
        n_entry += 1;
        n_word += 4;
      }
    }
    any_left = any_left || valid[k];
  }
  if (!__any_sync(FULL, any_left)) return;
#if ABG_SMALL_CODE
  if (KC == 1) {  // out of line, by value: the gather loop stays small
    const Deep1 r = compare_deep_one(index3, n_words, bound, valid[0], slot[0], sub[0]);
    d[0] = r.d;
    pm[0] = r.pm;
    the_pos[0] = r.pos;
    n_entry += r.n_entry;
    n_word += r.n_word;
    return;
  }
#endif
  compare_deep<KC, NC0>(ix, index3, mA, mC, mG, mT, n_words, bound, valid, slot, sub, d, pm, the_pos, n_entry, n_word);
}

// Ordered replay of the survivors of one compare round against candidate set `set_id`
// (check_hits' `if (diffs <= res.cutoff) res.update(...)`, abismal.cpp:1146-1147, in bucket order).
// Kept out of line: survivors are rare, and the heap code must not bloat the gather loop.
// The candidate sets are mutated by ONE lane (between warp syncs): the heap routines read and write the same
// shared-memory words many times over, which is only well defined within a single thread.
__device__ __noinline__ void replay_hits(int set_id, uint32_t strand_code, unsigned mask, int d, int pm, uint32_t pos) {
  const Warp W;
  uint32_t *stg = W.stage();  // [32][3]: d, pm, pos of the survivors in lane (= canonical) order
  __syncwarp();
  if ((mask >> W.lane) & 1u) {
    const int k = __popc(mask & ((1u << W.lane) - 1u));
    stg[3 * k] = (uint32_t)d;
    stg[3 * k + 1] = (uint32_t)pm;
    stg[3 * k + 2] = pos;
  }
  __syncwarp();
  if (W.lane == 0) {
    const int n = __popc(mask);
    CandSet res;
    res.load(W, set_id);
    for (int k = 0; k < n && !res.sure_ambig; ++k)
      if ((int)stg[3 * k + 1] <= res.cutoff)
        res.update(true, (int)stg[3 * k], strand_code, stg[3 * k + 2]);  // check_hits<.., true> in both phases
    res.store_one_lane(W, set_id);
  }
  __syncwarp();
}

// survivor-log entry: d (11 bits) | pm (11 bits) | three-letter table (1 bit) | seed offset (9 bits)
__device__ __forceinline__ uint32_t log_pack(int d, int pm, uint32_t is3, uint32_t off) {
  return (uint32_t)d | ((uint32_t)pm << 11) | (is3 << 22) | (off << 23);
}

// Sensitive phase for the seed offsets the specific phase has already examined: every candidate of a
// sensitive-eligible bucket at those offsets whose distance is within the (looser) sensitive cutoff was logged,
// in canonical order, by the specific phase; replay them instead of gathering the same words again.
__device__ __noinline__ void replay_log(int set_id, uint32_t strand_code, int n_log) {
  const Warp W;
  const uint32_t *lp = W.log_pos(), *lm = W.log_meta();
  const uint32_t *e2 = W.elig(0), *e3 = W.elig(1);
  __syncwarp();
  if (W.lane == 0) {  // one lane mutates the set (see replay_hits)
    CandSet res;
    res.load(W, set_id);
    for (int k = 0; k < n_log && !res.sure_ambig; ++k) {
      const uint32_t m = lm[k];
      const uint32_t off = m >> 23, is3 = (m >> 22) & 1u;
      const uint32_t ew = (is3 ? e3 : e2)[off >> 5];
      if (((ew >> (off & 31u)) & 1u) == 0u) continue;  // bucket not examined by the sensitive phase
      const int d = (int)(m & 2047u), pm = (int)((m >> 11) & 2047u);
      if (pm <= res.cutoff) res.update(true, d, strand_code, lp[k]);
    }
    res.store_one_lane(W, set_id);
  }
  __syncwarp();
}

// 2 adjacent counters of bucket k; the L2-resident emptiness bitmap (when present) filters out empty buckets
// without touching HBM.
__device__ __forceinline__ void probe(const uint32_t *__restrict__ counter, const uint32_t *__restrict__ bits,
                                      uint32_t k, uint32_t &s, uint32_t &e) {
  s = 0;
  e = 0;
  if (bits != nullptr && ((__ldg(bits + (k >> 5)) >> (k & 31u)) & 1u) == 0u) return;
  s = __ldg(counter + k);
  e = __ldg(counter + k + 1);
}

// Two-letter bucket [s, e) of key k through the compact counters when they exist
__device__ __forceinline__ void probe_two(const IndexDev &ix, uint32_t k, uint32_t &s, uint32_t &e) {
  if (ix.cc == nullptr) {
    probe(ix.counter, ix.bits, k, s, e);
    return;
  }
  const uint32_t b = k / kCcKeys, t = k - b * kCcKeys;
  const uint4 x = __ldg(ix.cc + 2 * (size_t)b), y = __ldg(ix.cc + 2 * (size_t)b + 1);
  if (x.x == ~0u) {
    s = __ldg(ix.counter + k);
    e = __ldg(ix.counter + k + 1);
    return;
  }
  // 28 one-byte bucket sizes in 7 words; bucket t = byte (t & 3) of word t >> 2.  Kept in registers (no array)
  const uint32_t tw = t >> 2, tb = 8u * (t & 3u);
  uint32_t before = 0;
  before += tw > 0u ? __vsadu4(x.y, 0u) : 0u;
  before += tw > 1u ? __vsadu4(x.z, 0u) : 0u;
  before += tw > 2u ? __vsadu4(x.w, 0u) : 0u;
  before += tw > 3u ? __vsadu4(y.x, 0u) : 0u;
  before += tw > 4u ? __vsadu4(y.y, 0u) : 0u;
  before += tw > 5u ? __vsadu4(y.z, 0u) : 0u;
  uint32_t cur = x.y;
  cur = tw == 1u ? x.z : cur;
  cur = tw == 2u ? x.w : cur;
  cur = tw == 3u ? y.x : cur;
  cur = tw == 4u ? y.y : cur;
  cur = tw == 5u ? y.z : cur;
  cur = tw == 6u ? y.w : cur;
  before += __vsadu4(cur & ((1u << tb) - 1u), 0u);  // bytes of the bucket's own word that precede it
  const uint32_t cnt = (cur >> tb) & 255u;
  s = x.x + before;
  e = s + cnt;
  if (cnt == 0u) s = e = 0u;  // what probe() reports for an empty bucket
}

// process_seeds (abismal.cpp:1269-1375) for pass `strand_code` of `end` into candidate set `set_id`
__device__ __noinline__ void process_seeds(int set_id, int end, uint32_t strand_code) {
  const Warp W;
  const KernelParams &P = params();
  const IndexDev &ix = P.ix;
  const int lane = W.lane;
  build_qcode(end, strand_code);
  build_packed(end, strand_code);
  const uint32_t readlen = W.scal()->len[end];
  const uint8_t *qcode = W.qcode(end);
  const uint32_t *mA = W.masks(0), *mC = W.masks(1), *mG = W.masks(2), *mT = W.masks(3);
  const uint32_t *p2 = W.plane(0), *p3a = W.plane(1), *p3b = W.plane(2);
  const uint32_t *T3 = tab3();
  const bool g_to_a = ((strand_code & ABG_FLAG_A_RICH) != 0) != ((strand_code & ABG_FLAG_RC) != 0);
  const uint32_t *counter3 = g_to_a ? ix.counter_a : ix.counter_t;
  const uint32_t *index3 = g_to_a ? ix.index_a : ix.index_t;
  const uint32_t *bits3 = g_to_a ? ix.bits_a : ix.bits_t;
  const uint4 *ctx3 = g_to_a ? ix.ctx_a : ix.ctx_t;
  const uint32_t maxc = P.max_candidates;
  const int n_words = (int)((readlen + 15) / 16);
  uint32_t c_lookup = 0, c_entry = 0, c_word = 0;  // per pass: far below 2^32
  volatile CandState *st = W.cs(set_id);  // the set's scalar state stays in shared memory
  const volatile uint64_t *heap0 = heap_of(W, set_id).sm;
  uint32_t *log_pos = W.log_pos(), *log_meta = W.log_meta();
  uint32_t *elig2 = W.elig(0), *elig3 = W.elig(1);

  const uint32_t specific_len = min(readlen - P.window_size, readlen >> 1);  // seed::window_size, abismal.cpp:1302-1305
  const uint32_t specific_lim = max(P.window_size, readlen >> 1);
  const uint32_t lim_two = readlen - 25u + 1u;
  int n_log = 0;
  bool log_ok = readlen <= (uint32_t)kLogMaxLen;

  for (int phase = 0; phase < 2; ++phase) {
    const bool specific = phase == 0;
    uint32_t first_off = 0;
    {
      CandSet res;
      res.load(W, set_id);
      if (specific) res.set_specific();
      else {
        if (!res.should_do_sensitive()) break;
        res.set_sensitive();
      }
      res.store(W, set_id);
    }
    if (!specific && log_ok) {
      if (n_log > 0) replay_log(set_id, strand_code, n_log);
      first_off = specific_lim;
    }
    const uint32_t n_off = specific ? specific_lim : lim_two;
    for (uint32_t base_off = first_off; base_off < n_off && !st->sure_ambig; base_off += 32) {
      const uint32_t i = base_off + lane;
      const bool active = i < n_off;
      uint32_t s2 = 0, e2 = 0, s3 = 0, e3 = 0, n2 = 0, n3 = 0;
      bool el2 = false, el3 = false;
      if (active) {
        // get_1bit_hash / get_base_3_hash at offset i (rolling == direct), AbismalIndex.hpp:285-305
        const uint32_t k = __brev(plane_window(p2, i)) >> 7;
        const uint32_t x0 = plane_window(p3a, i) & 0xffffu, x1 = plane_window(p3b, i) & 0xffffu;
        const uint32_t k3 = T3[x0 & 255u] + T3[256 + (x0 >> 8)] + 2u * (T3[x1 & 255u] + T3[256 + (x1 >> 8)]);
        probe_two(ix, k, s2, e2);
        probe(counter3, bits3, k3, s3, e3);
        {  // the sensitive phase's bucket rule (abismal.cpp:1351-1370), on the raw bucket sizes
          const uint32_t d_two = e2 - s2, d_three = e3 - s3;
          el2 = d_two != 0u && d_two <= maxc && (d_three == 0u || d_two <= 10u * d_three);
          el3 = d_three != 0u && d_three <= maxc;
        }
        if (specific) {
          uint32_t l_two = 25, l_three = 16;
          if (e2 - s2 > maxc || e2 == s2) {
            const SeedRange r = find_candidates(maxc, qcode + i, readlen - i, s2, e2);
            s2 = r.low;
            e2 = r.high;
            l_two = r.p;
          }
          if (e3 - s3 > maxc || e3 == s3) {
            const SeedRange r = find_candidates_three(index3, g_to_a, maxc, qcode + i, readlen - i, s3, e3);
            s3 = r.low;
            e3 = r.high;
            l_three = r.p;
          }
          const uint32_t d_two = e2 - s2, d_three = e3 - s3;
          n2 = (d_two <= maxc || l_two >= specific_len) ? d_two : 0u;
          n3 = (d_three <= maxc || l_three >= specific_len) ? d_three : 0u;
        }
        else {
          n2 = el2 ? e2 - s2 : 0u;
          n3 = el3 ? e3 - s3 : 0u;
        }
        c_lookup += 2;
      }
      __syncwarp();
      if (specific) {
        // the sensitive phase only visits offsets below lim_two (reads shorter than 2 * window_size + 8 have
        // specific offsets beyond it): their logged survivors must not be replayed
        const bool in_sens = i < lim_two;
        const unsigned m2 = __ballot_sync(FULL, el2 && in_sens), m3 = __ballot_sync(FULL, el3 && in_sens);
        if (lane == 0) {
          elig2[base_off >> 5] = m2;
          elig3[base_off >> 5] = m3;
        }
      }
      const uint32_t tot = n2 + n3;
      const uint32_t incl = warp_incl_scan_add(tot, lane);
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      bool stop = false;
      for (uint32_t c0 = 0; c0 < total && !stop; c0 += 32u * kCand) {
        const int cutoff = st->cutoff;
        // In the specific phase compare against the looser bound the sensitive phase may use later (the heap
        // top only ever decreases) and log what survives it, unless the set is already full (then the
        // sensitive phase is unlikely to run and deep compares would be wasted).
        bool deep = false;
        int bound = cutoff;
        if (specific && log_ok) {
          const int cap_now = st->is_pe ? (int)st->capacity : (int)kSeMax;
          if (st->sz == cap_now) log_ok = false;
          else {
            deep = true;
            bound = Hit(heap0[0]).diffs();
          }
        }
        int d[kCand], pm[kCand];
        uint32_t the_pos[kCand];
        uint32_t sub[kCand];
        compare_chunk<kCand, 4>(ix, index3, ctx3, mA, mC, mG, mT, n_words, bound, c0, total, base_off, incl, tot, n2, s2,
                                s3, lane, d, pm, the_pos, sub, c_entry, c_word);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kCand; ++k) {
          if (deep) {
            const unsigned ml = __ballot_sync(FULL, pm[k] <= bound);
            if (ml != 0u) {
              const int at = n_log + __popc(ml & ((1u << lane) - 1u));
              n_log += __popc(ml);
              if (n_log > kLogCap) log_ok = false;
              else if (pm[k] <= bound) {
                log_pos[at] = the_pos[k];
                log_meta[at] = log_pack(d[k], pm[k], sub[k] >> 31, sub[k] & 0x7fffffffu);
              }
            }
          }
          const unsigned mask = __ballot_sync(FULL, pm[k] <= cutoff);
          if (mask != 0u && !stop) {
            replay_hits(set_id, strand_code, mask, d[k], pm[k], the_pos[k]);
            stop = st->sure_ambig != 0;
          }
        }
      }
    }
    __syncwarp();
  }
  if (P.counters != nullptr) {
    unsigned long long *cnt = W.scal()->cnt;
    for (int d = 16; d >= 1; d >>= 1) {
      c_lookup += __shfl_xor_sync(FULL, c_lookup, d);
      c_entry += __shfl_xor_sync(FULL, c_entry, d);
      c_word += __shfl_xor_sync(FULL, c_word, d);
    }
    if (lane == 0) {
      cnt[0] += c_lookup;
      cnt[1] += c_entry;
      cnt[2] += c_word;
    }
    __syncwarp();
  }
}

// ---- sort by (pos, flags) + unique: prepare_for_alignments / prepare_for_mating ----
__device__ __noinline__ void sort_unique(int set_id) {
  const Warp W;
  const int lane = W.lane;
  const HeapRef v = heap_of(W, set_id);
  CandState *st = W.cs(set_id);
  const int sz = st->sz;
  int n2 = 1;
  while (n2 < sz) n2 <<= 1;
  if (n2 < 2) return;
  __syncwarp();
  for (int i = sz + lane; i < n2; i += 32) v.set(i, Hit(~0ull));  // pad with +inf keys
  __syncwarp();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int l = i | j;
        const Hit a = v.get(i), b = v.get(l);
        const bool asc = (i & k) == 0;
        if ((a.key() > b.key()) == asc) {
          v.set(i, b);
          v.set(l, a);
        }
      }
      __syncwarp();
    }
  }
  // std::unique on (pos, flags)
  int out = 0;
  for (int c0 = 0; c0 < sz; c0 += 32) {
    const int i = c0 + lane;
    Hit h, hp;
    bool keep = false;
    if (i < sz) {
      h = v.get(i);
      keep = true;
      if (i > 0) {
        hp = v.get(i - 1);
        keep = h.key() != hp.key();
      }
    }
    __syncwarp();
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) v.set(out + __popc(m & ((1u << lane) - 1u)), h);
    out += __popc(m);
    __syncwarp();
  }
  if (lane == 0) st->sz = out;
  __syncwarp();
}

// ---- banded alignment --------------------------------------------------------------
__device__ __forceinline__ int band_width(int diffs, int max_diffs) {  // AbismalAlign.hpp:333-334
  const int want = 2 * min(diffs, max_diffs) + 1;
  return want < 0 ? 61 : min(61, want);
}

// AbismalAlign::align<do_traceback> (AbismalAlign.hpp:320-386) for diffs != 0, as a wavefront.
// Lane l owns band columns 2l (value A) and 2l+1 (value B); at iteration T it is on table row
// i = T - l (reference base t_beg + i - 1).  Dependencies of cell (i, j):
//   diag  (i-1, j)   own value of the previous iteration
//   above (i-1, j+1) A: own B of the previous iteration; B: lane l+1's A of THIS iteration (its row is i-1)
//   left  (i, j-1)   A: lane l-1's B of the previous iteration (its row was i); B: own A of this iteration
// Arrow precedence on ties is left (I) > above (D) > diag (M), the reference's write order; the result is
// the first maximum in row-major order (std::max_element).
template <bool TB>
__device__ __noinline__ void align_wave(int tb_slot, int end, int bw, int q_sz, uint32_t t_pos, AlnOut *out,
                                        bool store_tb = true) {
  const Warp W;
  const KernelParams &P = params();
  const int lane = W.lane;
  const uint8_t *q = W.qcode(end);
  uint8_t *refb = W.refb();
  const uint32_t t_beg = t_pos - (uint32_t)((bw - 1) / 2);
  const int t_shift = q_sz + bw;
  // stage the reference bases of rows 1 .. t_shift-1 as bytes: refb[i - 1] = genome base t_beg + i - 1
  {
    const uint32_t w0 = t_beg >> 4;
    const int n_ref = t_shift - 1;
    const int nw = (int)(((t_beg + (uint32_t)n_ref - 1u) >> 4) - w0) + 1;
    const int shift0 = (int)(t_beg & 15u);
    __syncwarp();
#if ABG_SMALL_CODE
    (void)w0; (void)nw; (void)shift0;
    for (int r = lane; r < n_ref; r += 32) {  // one base per lane per step: tiny code, the words come from L1
      const uint32_t gpos = t_beg + (uint32_t)r;
      refb[r] = (uint8_t)((__ldg(P.ix.genome + (gpos >> 4)) >> (4u * (gpos & 15u))) & 15u);
    }
#else
    for (int k = lane; k < nw; k += 32) {
      const uint64_t word = __ldg(P.ix.genome + w0 + k);
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int r = 16 * k + n - shift0;
        if (r >= 0 && r < n_ref) refb[r] = (uint8_t)((word >> (4 * n)) & 15u);
      }
    }
#endif
    __syncwarp();
  }
  const int nl = (bw + 1) >> 1;  // lanes in use
  const int n_iter = (t_shift - 1) + (nl - 1);
  // Cell (i, j) of the band holds query base qi = i + j - bw; it exists iff j < bw and 0 <= qi < q_sz (that is
  // the reference's [left, right) of row i).  Cells that do not exist hold 0, as in the zero-filled table, so
  // an `above` or `left` taken from one is -4 and never wins against v >= 0; the one existing cell the
  // reference leaves out is `above` for the last query base (j + 1 == right on the rows below q_sz).
  // At iteration T this lane is on row i = T - lane: qi(A) = T + lane - bw, qi(B) = qi(A) + 1.
  const unsigned limA = 2 * lane < bw ? (unsigned)q_sz : 0u;
  const unsigned limB = 2 * lane + 1 < bw ? (unsigned)q_sz : 0u;
  const int up_mask = lane > 0 ? -1 : 0;  // lane 0 has no left neighbour for its A column
  uint64_t *tbs = W.tb_sm(tb_slot) + lane;           // [word][kTbLanesSm]
  uint64_t *tbg = W.tb_gm(tb_slot) + lane;           // [word][32]
  int A = 0, B = 0;
  int best = 0, best_at = 0;  // best_at = 2 T + (column B)
  uint64_t tbw = 0;
  int qi = 1 + lane - bw;     // qi(A) of iteration T = 1
  const uint8_t *rp = refb - lane;  // rp[T - 1] = reference base of this lane's row
  uint32_t qa = (unsigned)qi < limA ? (uint32_t)q[qi] : 0u;
  for (int T = 1; T <= n_iter; ++T, ++qi) {
    const bool okA = (unsigned)qi < limA, okB = (unsigned)(qi + 1) < limB;
    const uint32_t ref = (okA || okB) ? (uint32_t)rp[T - 1] : 0u;
    const uint32_t qb = (unsigned)(qi + 1) < (unsigned)q_sz ? (uint32_t)q[qi + 1] : 0u;
    const int left_in = __shfl_up_sync(FULL, B, 1) & up_mask;
    // column A
    int diag = A + ((qa & ref) ? 2 : -3);
    int v = max(diag, 0);
    int cA = diag >= 0 ? 0 : 3;
    {
      const int above = (qi + 1 < q_sz ? B : 0) - 4;
      if (above >= v) {
        v = above;
        cA = 2;
      }
      const int left = left_in - 4;
      if (left >= v) {
        v = left;
        cA = 1;
      }
    }
    const int newA = okA ? v : 0;
    if (TB) cA = newA > 0 ? cA : 3;
    if (newA > best) {
      best = newA;
      best_at = 2 * T;
    }
    const int a_down = __shfl_down_sync(FULL, newA, 1);
    // column B
    diag = B + ((qb & ref) ? 2 : -3);
    v = max(diag, 0);
    int cB = diag >= 0 ? 0 : 3;
    {
      const int above = (qi + 2 < q_sz ? a_down : 0) - 4;
      if (above >= v) {
        v = above;
        cB = 2;
      }
      const int left = newA - 4;
      if (left >= v) {
        v = left;
        cB = 1;
      }
    }
    const int newB = okB ? v : 0;
    if (TB) cB = newB > 0 ? cB : 3;
    if (newB > best) {
      best = newB;
      best_at = 2 * T + 1;
    }
    A = newA;
    B = newB;
    qa = qb;
    if (TB) {
      // 4 bits per iteration, iteration T at bits 4 (T & 15) of word T >> 4
      tbw = (tbw >> 4) | ((uint64_t)(uint32_t)(cA | (cB << 2)) << 60);
      if ((T & 15) == 15 || T == n_iter) {
        if ((T & 15) != 15) tbw >>= 4 * (15 - (T & 15));
        if (lane < nl && store_tb) {
          if (lane < kTbLanesSm) tbs[(T >> 4) * kTbLanesSm] = tbw;
          else tbg[(size_t)(T >> 4) * 32] = tbw;
        }
        tbw = 0;
      }
    }
  }
  // first maximum in row-major order (std::max_element)
  int bv = best, br = (best_at >> 1) - lane, bc = 2 * lane + (best_at & 1);
  if (best == 0) br = 0, bc = 0;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const int ov = __shfl_xor_sync(FULL, bv, d);
    const int orow = __shfl_xor_sync(FULL, br, d);
    const int oc = __shfl_xor_sync(FULL, bc, d);
    const bool take = ov > bv || (ov == bv && (orow < br || (orow == br && oc < bc)));
    if (take) {
      bv = ov;
      br = orow;
      bc = oc;
    }
  }
  out->score = bv;
  out->row = br;
  out->col = bc;
  out->bw = bw;
  __syncwarp();
}

// Returns the alignment score of qcode[end] (encoded for `flags`) against the genome at t_pos.
// record_tb: also keep the traceback words in `tb_slot` (so that a later traceback of the very
// same alignment can skip the DP); need_tb: the caller will read the traceback right away.
// `out` is meaningful only when diffs != 0.
#if ABG_SMALL_CODE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
int align(bool record_tb, bool need_tb, int tb_slot, int end, uint32_t flags, int diffs,
                                     int max_diffs, int q_sz, uint32_t t_pos, AlnOut &out, uint32_t task_id = kNoTask) {
  if (diffs == 0) return 2 * q_sz;  // AbismalAlign.hpp:329-330
  const Warp W;
  const int bw = band_width(diffs, max_diffs);
  TbKey *tk = &W.scal()->tbk[tb_slot];
  const uint32_t key = ((uint32_t)end << 16) | (flags & (ABG_FLAG_RC | ABG_FLAG_A_RICH));
  if (tk->valid && tk->pos == t_pos && tk->key == key && tk->bw == bw) {
    out.score = tk->score;
    out.row = tk->row;
    out.col = tk->col;
    out.bw = bw;
    return out.score;
  }
  if (task_id != kNoTask) {
    // dp_kernel has run this very alignment (same read strand, position and band): take its result
    const KernelParams &P = params();
    const uint2 rr = __ldcg(reinterpret_cast<const uint2 *>(P.task_res + task_id));
    const uint32_t tbi = __ldcg(&P.tasks[task_id].tb_index);
    if ((int)(int16_t)(rr.y >> 16) == bw && (tbi != kNoTask || !need_tb)) {
      out.score = (int)(int16_t)(rr.x & 0xffffu);
      out.row = (int)(int16_t)(rr.x >> 16);
      out.col = (int)(int16_t)(rr.y & 0xffffu);
      out.bw = bw;
      if (tbi != kNoTask) {  // its traceback words are on record: the slot now describes this alignment
        __syncwarp();
        if (W.lane == 0) {
          tk->pos = t_pos;
          tk->key = key;
          tk->score = out.score;
          tk->row = out.row;
          tk->col = out.col;
          tk->bw = bw;
          tk->tb_index = tbi;
          tk->valid = 1;
        }
        __syncwarp();
      }
      return out.score;
    }
  }
  build_qcode(end, flags);
  const bool tb = record_tb || need_tb;
  if (params().counters != nullptr && W.lane == 0) {
    W.scal()->cnt[3] += 1;
    W.scal()->cnt[4] += (unsigned long long)(q_sz + bw);
  }
  __syncwarp();
  if (tb && W.lane == 0) tk->valid = 0;
#if ABG_SMALL_CODE
  align_wave<true>(tb_slot, end, bw, q_sz, t_pos, &out, tb);
#else
  if (tb) align_wave<true>(tb_slot, end, bw, q_sz, t_pos, &out);
  else align_wave<false>(tb_slot, end, bw, q_sz, t_pos, &out);
#endif
  if (tb && W.lane == 0) {
    tk->pos = t_pos;
    tk->key = key;
    tk->score = out.score;
    tk->row = out.row;
    tk->col = out.col;
    tk->bw = bw;
    tk->tb_index = kNoTask;
    tk->valid = 1;
  }
  __syncwarp();
  return out.score;
}

struct CigarOut {
  uint32_t *ops;     // results array of this read (cigar_stride slots)
  uint32_t stride;
  uint32_t n;        // ops in the CIGAR (may exceed stride -> overflow)
  uint32_t ref_len;  // cigar_rseq_ops
};

__device__ __forceinline__ void cigar_default(CigarOut &cg, uint32_t len, int lane) {  // make_default_cigar
  if (lane == 0 && cg.stride > 0) cg.ops[0] = len << 4;
  cg.n = 1;
  cg.ref_len = len;
  __syncwarp();
}

// build_cigar_len_and_pos + get_traceback (AbismalAlign.hpp:388-440, :166-193)
// followed by simple_aln::edit_distance (:73-89).  Uniform across the warp.
__device__ __noinline__ int build_cigar(int tb_slot, int diffs, int a_score, int a_row, int a_col, int bw, int q_sz,
                                        int scr_for_nm, uint32_t *ops, uint32_t stride, uint32_t *n_out,
                                        uint32_t *ref_len_out, uint32_t *len_out, uint32_t *t_pos_io) {
  const Warp W;
  const int lane = W.lane;
  int ins = 0, del = 0;
  uint32_t len;
  if (diffs == 0 || a_score == 0) {
    if (lane == 0 && stride > 0) ops[0] = (uint32_t)q_sz << 4;
    *n_out = 1;
    *ref_len_out = (uint32_t)q_sz;
    len = (uint32_t)q_sz;
    __syncwarp();
  }
  else {
    const uint64_t *tbs = W.tb_sm(tb_slot);
    const uint64_t *tbg = W.tb_gm(tb_slot);
    // words of an alignment dp_kernel ran: [block of 16 iterations][lane of its group of 8 / 16 / 32]
    const uint32_t tbi = W.scal()->tbk[tb_slot].tb_index;
    const uint64_t *tba = tbi != kNoTask ? params().task_tb + (size_t)tbi * 8u : nullptr;
    const int tbG = bw <= 16 ? 8 : (bw <= 32 ? 16 : 32);
    int row = a_row, col = a_col;
    const int clip_bottom = (q_sz + (bw - 1)) - (row + col);
    __syncwarp();
    // traceback word holding cell (r, cc): lane l = cc / 2, iteration T = r + l
    const auto word_at = [&](int T, int l) -> uint64_t {
      if (tba != nullptr) return __ldcg(tba + (size_t)(T >> 4) * tbG + l);
      return l < kTbLanesSm ? tbs[(T >> 4) * kTbLanesSm + l] : *(volatile const uint64_t *)(tbg + (size_t)(T >> 4) * 32 + l);
    };
    const auto code_at = [&](int r, int cc) -> int {
      if (cc < 0 || cc >= bw || r <= 0) return 3;
      const int l = cc >> 1, T = r + l;
      return (int)(word_at(T, l) >> (4 * (T & 15) + 2 * (cc & 1))) & 3;
    };
    uint32_t n_ops = 0, ref_len = 0;
    const auto emit = [&](uint32_t n, int op) {
      if (n_ops < stride && lane == 0) ops[n_ops] = (n << 4) | (uint32_t)op;
      ++n_ops;
      if (op == 1) ins += (int)(n & 0xffu);  // abismal_bam_cigar_oplen returns uint8_t
      if (op == 2) del += (int)(n & 0xffu);
      if (op == 0 || op == 2) ref_len += n;
    };
    int prev_arrow = code_at(row, col);
    if (prev_arrow == 3) prev_arrow = 0;  // cannot happen: the best cell is positive
    {
      const bool is_del = prev_arrow == 2, is_ins = prev_arrow == 1;
      row -= !is_ins;
      col -= is_ins;
      col += is_del;
    }
    uint32_t n = 1;
    for (;;) {
      if (col < 0 || col >= bw || row <= 0) break;
      const int l = col >> 1, T = row + l;
      const uint64_t word = word_at(T, l);
      const int pos = T & 15;
      const int arrow = (int)(word >> (4 * pos + 2 * (col & 1))) & 3;
      if (arrow == 3) break;  // table[row][col] <= 0
      if (arrow == 0) {
        // a run of diagonal arrows stays in this lane's column: count them inside the word
        // (codes of this column for iterations pos, pos-1, ..., 0, most recent first)
        const uint64_t m = ((word >> (2 * (col & 1))) & 0x3333333333333333ull) << (4 * (15 - pos));
        int run = m == 0 ? pos + 1 : (__clzll((long long)m) >> 2);  // leading zero codes
        if (run > row) run = row;
        if (prev_arrow != 0) {
          emit(n, prev_arrow);
          n = 0;
        }
        n += (uint32_t)run;
        row -= run;
        prev_arrow = 0;
      }
      else {
        const bool is_del = arrow == 2, is_ins = arrow == 1;
        row -= !is_ins;
        col -= is_ins;
        col += is_del;
        if (arrow != prev_arrow) {
          emit(n, prev_arrow);
          n = 0;
        }
        ++n;
        prev_arrow = arrow;
      }
    }
    emit(n, prev_arrow);
    const int clip_top = (row + col) - (bw - 1);
    if (clip_top > 0) {
      if (n_ops < stride && lane == 0) ops[n_ops] = ((uint32_t)clip_top << 4) | 4u;
      ++n_ops;
    }
    __syncwarp();
    // reverse in place
    const uint32_t m = min(n_ops, stride);
    if (n_ops <= stride) {
      for (uint32_t k = lane; k < m / 2; k += 32) {
        const uint32_t x = ops[k], y = ops[m - 1 - k];
        ops[k] = y;
        ops[m - 1 - k] = x;
      }
    }
    __syncwarp();
    if (clip_bottom > 0) {
      if (n_ops < stride && lane == 0) ops[n_ops] = ((uint32_t)clip_bottom << 4) | 4u;
      ++n_ops;
    }
    __syncwarp();
    *n_out = n_ops;
    *ref_len_out = ref_len;
    len = (uint32_t)(q_sz - clip_bottom - clip_top);
    const uint32_t t_beg = *t_pos_io - (uint32_t)((bw - 1) / 2);
    *t_pos_io = t_beg + (uint32_t)row;
  }
  *len_out = len;
  // edit_distance(scr, len, cigar): same promotions as the reference (unsigned quotient)
  if (scr_for_nm == 0) return (int)(int16_t)len;
  const int A = (int)(int16_t)(scr_for_nm + 4 * (ins + del));
  const uint32_t num = 2u * (len - (uint32_t)ins) - (uint32_t)A;
  const int mism = (int)(int16_t)(num / 5u);
  return (int)(int16_t)(mism + ins + del);
}

__device__ __forceinline__ int build_cigar(int tb_slot, int diffs, const AlnOut &a, int q_sz, int scr_for_nm,
                                           CigarOut &cg, uint32_t &len, uint32_t &t_pos) {
  uint32_t n = 0, ref_len = 0, l = 0, p = t_pos;
  const int nm = build_cigar(tb_slot, diffs, a.score, a.row, a.col, a.bw, q_sz, scr_for_nm, cg.ops, cg.stride, &n,
                             &ref_len, &l, &p);
  cg.n = n;
  cg.ref_len = ref_len;
  len = l;
  t_pos = p;
  return nm;
}

// task id of entry j of a stored set (tof: its set_slots ids; entries beyond: the overflow arena at the set's offset)
__device__ __forceinline__ uint32_t task_id_of(const uint32_t *tof, int j, uint32_t ovf) {
  if (tof == nullptr) return kNoTask;
  const KernelParams &P = params();
  if ((uint32_t)j < P.set_slots) return __ldcg(tof + j);
  return ovf != 0u ? __ldcg(P.task_ovf + (ovf - 1u) + ((uint32_t)j - P.set_slots)) : kNoTask;
}

__device__ __forceinline__ bool same_pos(uint32_t a, uint32_t b) { return (a > b ? a - b : b - a) <= 3u; }

// align_se_candidates (abismal.cpp:1435-1497) on se set `end`
// task_of: task ids of the entries of the (already sorted + uniqued) set, or nullptr: align in the warp
__device__ __noinline__ uint64_t align_se_candidates(int end, double cutoff, uint32_t *ops, uint32_t stride,
                                                     uint32_t *cg_n, uint32_t *cg_ref_len, uint64_t best_in,
                                                     const uint32_t *task_of = nullptr) {
  const Warp W;
  const uint32_t readlen_u = W.scal()->len[end];
  const int readlen = (int)(int16_t)readlen_u;
  const int max_diffs = frac_of(cutoff, (uint32_t)readlen);
  const int max_scr = (int)(int16_t)(2 * readlen);
  Hit best(best_in);
  CandState *st = W.cs(end);
  if (!Hit(st->best).empty()) {
    best = Hit(st->best);
    if (W.lane == 0 && stride > 0) ops[0] = (uint32_t)readlen << 4;  // make_default_cigar
    *cg_n = 1;
    *cg_ref_len = (uint32_t)readlen;
    __syncwarp();
    return best.w;
  }
  int best_scr = 0;
  uint32_t best_pos = 0, best_task = kNoTask;
  if (task_of == nullptr) sort_unique(end);  // enum_kernel has done it otherwise
  const HeapRef v = heap_of(W, end);
  int it = 0;
  const int lim = st->sz;
  for (; it != lim && v.get(it).empty(); ++it) {
  }
  const bool record_tb = (lim - it) <= kTbCacheMaxCands;
  const int invalid = frac_of(0.4, (uint32_t)readlen);
  const int q_sz = (int)readlen_u;
  AlnOut ao;
  for (; it != lim; ++it) {
    const Hit h = v.get(it);
    if (h.diffs() < invalid) {
      const uint32_t cand_pos = h.pos();
      const uint32_t tid = task_id_of(task_of, it, 0u);
      const int cand_scr =
        (int)(int16_t)align(record_tb, false, 0, end, h.flags(), h.diffs(), max_diffs, q_sz, cand_pos, ao, tid);
      if (cand_scr > best_scr) {
        best = h;
        best_scr = cand_scr;
        best_pos = cand_pos;
        best_task = tid;
      }
      else if (cand_scr == best_scr && (cand_scr == max_scr ? cand_pos != best_pos : !same_pos(cand_pos, best_pos)))
        best.set_ambig();
    }
  }
  if (best.pos() != 0) {
    ao.score = 0;
    align(false, true, 0, end, best.flags(), best.diffs(), max_diffs, q_sz, best.pos(), ao, best_task);
    uint32_t len = 0, pos = best.pos();
    CigarOut cg{ops, stride, 0u, 0u};
    const int nm = build_cigar(0, best.diffs(), ao, q_sz, best_scr, cg, len, pos);
    *cg_n = cg.n;
    *cg_ref_len = cg.ref_len;
    best.set_pos(pos);
    best.set_diffs(nm);
    if (!(valid_len(len, (uint32_t)readlen, params().window_size + 24u) && nm <= frac_of(cutoff, (uint32_t)readlen)))
      best.reset();
  }
  else best.reset();
  return best.w;
}

// pe_element (abismal.cpp:547-622)
struct PeBest {
  int aln_score, max_aln_score;
  Hit r1, r2;
  __device__ void reset(uint32_t l1, uint32_t l2) {
    aln_score = 0;
    r1 = Hit(invalid_hit_diffs(l1), 0, 0);
    r2 = Hit(invalid_hit_diffs(l2), 0, 0);
    max_aln_score = (int)(int16_t)((int)(int16_t)(2 * l1) + (int)(int16_t)(2 * l2));
  }
  __device__ void reset() {
    aln_score = 0;
    r1.reset();
    r2.reset();
  }
  __device__ bool update(int scr, Hit s1, Hit s2) {
    const int rd = r1.diffs() + r2.diffs();
    const int sd = s1.diffs() + s2.diffs();
    if (scr > aln_score || (scr == aln_score && sd < rd)) {
      r1 = s1;
      r2 = s2;
      aln_score = scr;
      return true;
    }
    if (scr == aln_score && sd == rd) r1.set_ambig();
    return false;
  }
  __device__ bool sure_ambig() const { return r1.ambig() && aln_score == max_aln_score; }
  __device__ bool should_report(bool allow_ambig) const { return !r1.empty() && (allow_ambig || !r1.ambig()); }
  __device__ int diffs() const { return (int)(int16_t)(r1.diffs() + r2.diffs()); }
};

// ---- best_pair on large candidate sets (repeats): the sweep by rows ----------------------------------------
// best_pair (abismal.cpp:1722-1831) walks the position-sorted sets with two pointers: row j2 meets the entries
// j1 in [lo(j2), hi(j2)) of the other set (lo, hi non-decreasing in j2), in the order (j2, j1), and folds every
// pair into pe_element::update.  That fold depends on the order only through
//   * the winner: the FIRST pair, in that order, whose key (pair score, then fewer diffs) beats the state the
//     call started with and is not beaten later -- i.e. the first pair holding the best key;
//   * the ambiguity flag: set when another pair (or the starting state) holds the same key;
//   * `scr1`: on a memo hit the reference records the score of the set-1 alignment it ran most recently, not
//     the partner's.  An alignment runs when its entry is met for the first time (j1 at
//     or beyond every earlier row's hi) or when its memoised score is 0.
// so 32 rows are evaluated at a time, one per lane, from precomputed scores, and combined by a warp reduction.
// The early exit on sure_ambig only skips pairs that cannot change the state.
constexpr int kScoreNone = -32768;

struct PairPick {
  int scr1, scr2;
  uint32_t pos1, pos2, t1, t2;
};

// alignment scores of the entries of set `id` into scr[]: 2 * len for exact matches, the task's result, or -- for
// needed entries without a task (list full) -- the banded DP in this warp
template <class F>
__device__ __forceinline__ void score_set(const Warp &W, int id, int end, uint32_t flags, int max_diffs, int readlen,
                                          const uint32_t *tof, uint32_t ovf, int16_t *scr, F needed) {
  const KernelParams &P = params();
  const HeapRef v = heap_of(W, id);
  const int n = W.cs(id)->sz;
  AlnOut ao;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + W.lane;
    Hit h(0);
    int sc = kScoreNone;
    bool want = false;
    if (j < n) {
      h = v.get(j);
      if (!h.empty()) {
        if (h.diffs() == 0) sc = (int)(int16_t)(2 * readlen);
        else {
          const uint32_t t = task_id_of(tof, j, ovf);
          if (t != kNoTask) {
            const uint2 rr = __ldcg(reinterpret_cast<const uint2 *>(P.task_res + t));
            if ((int)(int16_t)(rr.y >> 16) == band_width(h.diffs(), max_diffs)) sc = (int)(int16_t)(rr.x & 0xffffu);
          }
          if (sc == kScoreNone) want = needed(h);
        }
      }
    }
    unsigned m = __ballot_sync(FULL, want);
    while (m != 0u) {
      const int k = __ffs(m) - 1;
      m &= m - 1u;
      const Hit hk(__shfl_sync(FULL, h.w, k));
      const int a = (int)(int16_t)align(false, false, end, end, flags, hk.diffs(), max_diffs, readlen, hk.pos(), ao, kNoTask);
      if (W.lane == k) sc = a;
    }
    if (j < n) scr[j] = (int16_t)sc;
  }
  __syncwarp();
}

__device__ __noinline__ PairPick best_pair_rows(bool swap_ends, int e1, uint32_t flags1, uint32_t flags2, PeBest *best_io,
                                                const uint32_t *tof1, const uint32_t *tof2) {
  const Warp W;
  const KernelParams &P = params();
  const int lane = W.lane;
  const int e2 = 1 - e1;
  PeBest best = *best_io;
  const HeapRef v1 = heap_of(W, 2), v2 = heap_of(W, 3);
  const int n1 = W.cs(2)->sz, n2 = W.cs(3)->sz;
  int16_t *scr1 = P.mem_scr + W.slot() * (size_t)kPeLarge, *scr2 = P.mem_scr2 + W.slot() * (size_t)kPeLarge;
  const uint32_t readlen1 = W.scal()->len[e1], readlen2 = W.scal()->len[e2];
  const int max_diffs1 = frac_of(P.valid_frac, readlen1), max_diffs2 = frac_of(P.valid_frac, readlen2);
  const uint32_t min_dist = P.min_dist, max_dist = P.max_dist;
  const uint32_t ovf1 = W.cs(2)->ovf, ovf2 = W.cs(3)->ovf;
  PairPick pick = PairPick{0, 0, 0u, 0u, kNoTask, kNoTask};
  int s1 = 0, s2 = 0;  // leading empties (position 0 sorts first)
  for (int j0 = 0; j0 < n1; j0 += 32) s1 += __popc(__ballot_sync(FULL, j0 + lane < n1 && v1.get(j0 + lane).empty()));
  for (int j0 = 0; j0 < n2; j0 += 32) s2 += __popc(__ballot_sync(FULL, j0 + lane < n2 && v2.get(j0 + lane).empty()));
  // first j1 in [s1, n1) with pos1 + add >= lim / pos1 + add > lim (32-bit arithmetic as in best_pair)
  auto first_ge = [&](int lo, uint32_t add, uint32_t lim) {
    int hi = n1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (v1.get(mid).pos() + add >= lim) hi = mid;
      else lo = mid + 1;
    }
    return lo;
  };
  auto first_gt = [&](int lo, uint32_t add, uint32_t lim) {
    int hi = n1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (v1.get(mid).pos() + add > lim) hi = mid;
      else lo = mid + 1;
    }
    return lo;
  };
  // scores of the entries that have a partner (the others are never asked about)
  score_set(W, 3, e2, flags2, max_diffs2, (int)readlen2, tof2, ovf2, scr2, [&](Hit h) {
    const uint32_t lim = h.pos() + readlen2;
    const int lo = first_ge(s1, max_dist, lim);
    return lo < n1 && v1.get(lo).pos() + min_dist <= lim;
  });
  score_set(W, 2, e1, flags1, max_diffs1, (int)readlen1, tof1, ovf1, scr1, [&](Hit h) {
    const uint32_t p1 = h.pos();
    int lo = s2, hi = n2;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (v2.get(mid).pos() + readlen2 >= p1 + min_dist) hi = mid;
      else lo = mid + 1;
    }
    return lo < n2 && p1 + max_dist >= v2.get(lo).pos() + readlen2;
  });
  __threadfence_block();
  int carry_hi = 0;        // hi of the last non-empty row so far: entries below it have been met
  int carry_exec = 0;      // score of the set-1 alignment run most recently (`scr1` of the reference)
  for (int r0 = s2; r0 < n2 && !best.sure_ambig(); r0 += 32) {
    const int j2 = r0 + lane;
    const bool row = j2 < n2;
    Hit h2(0);
    int lo = 0, hi = 0, sc2 = 0;
    if (row) {
      h2 = v2.get(j2);
      const uint32_t lim = h2.pos() + readlen2;
      lo = first_ge(s1, max_dist, lim);
      hi = first_gt(lo, min_dist, lim);
      sc2 = (int)*(volatile int16_t *)(scr2 + j2);
    }
    const bool nonempty = row && lo < hi;
    // hi of the nearest non-empty row below this lane (hi is non-decreasing over the rows)
    int prev_hi = nonempty ? hi : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, prev_hi, d);
      if (lane >= d) prev_hi = max(prev_hi, o);
    }
    const int chunk_hi = __shfl_sync(FULL, prev_hi, 31);
    prev_hi = __shfl_up_sync(FULL, prev_hi, 1);
    prev_hi = max(carry_hi, lane == 0 ? 0 : prev_hi);
    // the row: best key, the first j1 holding it, how many hold it, zero scores in front of it, last run alignment
    long long kbest = LLONG_MIN;
    int jbest = -1, nbest = 0, last_zero = -1;
    bool zero_before = false;
    if (nonempty) {
      for (int j1 = lo; j1 < hi; ++j1) {
        const Hit h1 = v1.get(j1);
        const int m1 = (int)*(volatile int16_t *)(scr1 + j1);
        const int pair_scr = (int)(int16_t)(sc2 + m1);
        const long long key = (long long)pair_scr * (1ll << 20) - (long long)(h1.diffs() + h2.diffs());
        if (key > kbest) {
          kbest = key;
          jbest = j1;
          nbest = 1;
          zero_before = last_zero >= 0;
        }
        else if (key == kbest) ++nbest;
        if (m1 == 0) last_zero = j1;
      }
    }
    // A pair score equal to max_aln_score can end the reference's sweep in the middle of the chunk (sure_ambig):
    // such chunks (rare: both ends match exactly) are folded pair by pair, in order, from the same scores.
    int smax = nonempty ? (int)((kbest + (1ll << 19)) >> 20) : INT_MIN;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) smax = max(smax, __shfl_xor_sync(FULL, smax, d));
    if (max(smax, best.aln_score) >= best.max_aln_score) {
      int seen_hi = carry_hi, cur_scr1 = carry_exec;
      bool stop = false;
      for (int l = 0; l < 32 && !stop; ++l) {
        const int lo_l = __shfl_sync(FULL, lo, l), hi_l = __shfl_sync(FULL, hi, l);
        if (!(__shfl_sync(FULL, (int)nonempty, l) != 0)) continue;
        const Hit g2(__shfl_sync(FULL, h2.w, l));
        const int gsc2 = __shfl_sync(FULL, sc2, l);
        for (int j1 = lo_l; j1 < hi_l; ++j1) {
          if (best.sure_ambig()) {
            stop = true;
            break;
          }
          const Hit g1 = v1.get(j1);
          const int m1 = (int)*(volatile int16_t *)(scr1 + j1);
          if (j1 >= seen_hi || m1 == 0) cur_scr1 = m1;  // this alignment runs now
          const int pair_scr = (int)(int16_t)(gsc2 + m1);
          if (swap_ends ? best.update(pair_scr, g2, g1) : best.update(pair_scr, g1, g2)) {
            pick.scr1 = cur_scr1;
            pick.scr2 = gsc2;
            pick.pos1 = g1.pos();
            pick.pos2 = g2.pos();
            pick.t1 = task_id_of(tof1, j1, ovf1);
            pick.t2 = task_id_of(tof2, r0 + l, ovf2);
          }
        }
        seen_hi = max(seen_hi, hi_l);
      }
      carry_exec = cur_scr1;
      carry_hi = max(carry_hi, chunk_hi);
      if (stop) break;
      continue;
    }
    bool exec_any = false;
    int exec_scr = 0;
    if (nonempty) {
      if (hi > prev_hi) {  // the row's last entry is met for the first time
        exec_any = true;
        exec_scr = (int)*(volatile int16_t *)(scr1 + hi - 1);
      }
      else if (last_zero >= 0) exec_any = true;  // exec_scr = 0
    }
    // ---- the chunk's first pair with the best key, and how many pairs hold that key
    long long kmax = kbest;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) kmax = max(kmax, __shfl_xor_sync(FULL, kmax, d));
    const unsigned holders = __ballot_sync(FULL, nonempty && kbest == kmax);
    if (holders != 0u) {
      const int wl = __ffs(holders) - 1;
      int count = (nonempty && kbest == kmax) ? nbest : 0;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) count += __shfl_xor_sync(FULL, count, d);
      const int rd = best.r1.diffs() + best.r2.diffs();
      const long long kcur = (long long)best.aln_score * (1ll << 20) - (long long)rd;
      if (kmax > kcur) {
        // replacement by pair (row wl, its jbest); later holders of the key make it ambiguous
        const int wj1 = __shfl_sync(FULL, jbest, wl);
        const int wj2 = r0 + wl;
        const Hit w1 = v1.get(wj1), w2 = v2.get(wj2);
        const int wm1 = (int)*(volatile int16_t *)(scr1 + wj1), wsc2 = __shfl_sync(FULL, sc2, wl);
        best.aln_score = (int)(int16_t)(wsc2 + wm1);
        best.r1 = swap_ends ? w2 : w1;
        best.r2 = swap_ends ? w1 : w2;
        if (count >= 2) best.r1.set_ambig();
        // scr1 at that moment
        const int w_prev_hi = __shfl_sync(FULL, prev_hi, wl);
        const bool w_zero_before = __shfl_sync(FULL, (int)zero_before, wl) != 0;
        int stale;
        if (wj1 >= w_prev_hi || wm1 == 0) stale = wm1;
        else if (w_zero_before) stale = 0;
        else {
          const unsigned below = __ballot_sync(FULL, exec_any) & ((1u << wl) - 1u);
          stale = below != 0u ? __shfl_sync(FULL, exec_scr, 31 - __clz(below)) : carry_exec;
        }
        pick.scr1 = stale;
        pick.scr2 = wsc2;
        pick.pos1 = w1.pos();
        pick.pos2 = w2.pos();
        pick.t1 = task_id_of(tof1, wj1, ovf1);
        pick.t2 = task_id_of(tof2, wj2, ovf2);
      }
      else if (kmax == kcur) best.r1.set_ambig();
    }
    const unsigned execs = __ballot_sync(FULL, exec_any);
    if (execs != 0u) carry_exec = __shfl_sync(FULL, exec_scr, 31 - __clz(execs));
    carry_hi = max(carry_hi, chunk_hi);
  }
  *best_io = best;
  return pick;
}

// best_pair<swap_ends> (abismal.cpp:1722-1831) over pe sets 2 (end e1, un-reversed) and 3 (end e2, reversed).
// The traceback slot of an end is its end number.
// tof1 / tof2: task ids of the entries of sets 2 / 3 (enum_kernel), or nullptr
__device__ __noinline__ void best_pair(bool swap_ends, int e1, uint32_t flags1, uint32_t flags2, CigarOut *cgs,
                                       PeBest *best_io, const uint32_t *tof1 = nullptr, const uint32_t *tof2 = nullptr) {
  const Warp W;
  const KernelParams &P = params();
  const int e2 = 1 - e1;
  PeBest best = *best_io;
  const HeapRef v1 = heap_of(W, 2), v2 = heap_of(W, 3);
  const int j1_end = W.cs(2)->sz, j2_end = W.cs(3)->sz;
  int16_t *mem_scr = P.mem_scr + W.slot() * (size_t)kPeLarge;
  int j1 = 0, j2 = 0;
  __syncwarp();
  for (int k = W.lane; k < j1_end; k += 32) mem_scr[k] = 0;
  __syncwarp();
  const uint32_t readlen1 = W.scal()->len[e1], readlen2 = W.scal()->len[e2];
  const int max_diffs1 = frac_of(P.valid_frac, readlen1);
  const int max_diffs2 = frac_of(P.valid_frac, readlen2);
  const uint32_t min_dist = P.min_dist, max_dist = P.max_dist;
  int scr1 = 0, best_scr1 = 0, best_scr2 = 0;
  uint32_t best_pos1 = 0, best_pos2 = 0, best_t1 = kNoTask, best_t2 = kNoTask;
  const uint32_t ovf1 = W.cs(2)->ovf, ovf2 = W.cs(3)->ovf;
  AlnOut ao;

  const bool by_rows = tof1 != nullptr && tof2 != nullptr && (uint32_t)max(j1_end, j2_end) >= P.heavy_min;
  if (by_rows) {
    const PairPick pk = best_pair_rows(swap_ends, e1, flags1, flags2, &best, tof1, tof2);
    if (pk.pos1 != 0u) {
      best_scr1 = pk.scr1;
      best_scr2 = pk.scr2;
      best_pos1 = pk.pos1;
      best_pos2 = pk.pos2;
      best_t1 = pk.t1;
      best_t2 = pk.t2;
    }
    j2 = j2_end;  // the sweep below has nothing left to do
  }
  for (; j1 != j1_end && v1.get(j1).empty(); ++j1) {
  }
  for (; j2 != j2_end && v2.get(j2).empty(); ++j2) {
  }
  const bool rec1 = (j1_end - j1) <= kTbCacheMaxCands, rec2 = (j2_end - j2) <= kTbCacheMaxCands;
  for (; j2 != j2_end && !best.sure_ambig(); ++j2) {
    const Hit s2 = v2.get(j2);
    int scr2 = 0;
    const uint32_t lim = s2.pos() + readlen2;
    for (; (j1 == j1_end) || (j1 != 0 && v1.get(j1).pos() + max_dist >= lim); --j1) {
    }
    for (; j1 != j1_end && v1.get(j1).pos() + max_dist < lim; ++j1) {
    }
    for (; j1 != j1_end && !best.sure_ambig(); ++j1) {
      const Hit s1 = v1.get(j1);
      if (!(s1.pos() + min_dist <= lim)) break;
      const uint32_t t2 = task_id_of(tof2, j2, ovf2), t1 = task_id_of(tof1, j1, ovf1);
      if (scr2 == 0)
        scr2 = (int)(int16_t)align(rec2, false, e2, e2, flags2, s2.diffs(), max_diffs2, (int)readlen2, s2.pos(), ao, t2);
      int m1 = *(volatile int16_t *)(mem_scr + j1);
      if (m1 == 0) {
        scr1 = (int)(int16_t)align(rec1, false, e1, e1, flags1, s1.diffs(), max_diffs1, (int)readlen1, s1.pos(), ao, t1);
        m1 = scr1;
        __syncwarp();
        if (W.lane == 0) mem_scr[j1] = (int16_t)scr1;
        __syncwarp();
      }
      const int pair_scr = (int)(int16_t)(scr2 + m1);
      if (swap_ends ? best.update(pair_scr, s2, s1) : best.update(pair_scr, s1, s2)) {
        best_scr1 = scr1;  // stale on a memo hit, as in the reference
        best_scr2 = scr2;
        best_pos1 = s1.pos();
        best_pos2 = s2.pos();
        best_t1 = t1;
        best_t2 = t2;
      }
    }
  }
  if (best_pos1 != 0) {
    Hit s1 = swap_ends ? best.r2 : best.r1;
    Hit s2 = swap_ends ? best.r1 : best.r2;
    uint32_t len1 = 0, len2 = 0;
    ao.score = 0;
    align(false, true, e1, e1, flags1, s1.diffs(), max_diffs1, (int)readlen1, best_pos1, ao, best_t1);
    int nm = build_cigar(e1, s1.diffs(), ao, (int)readlen1, best_scr1, cgs[e1], len1, best_pos1);
    s1.set_pos(best_pos1);
    s1.set_diffs(nm);
    ao.score = 0;
    align(false, true, e2, e2, flags2, s2.diffs(), max_diffs2, (int)readlen2, best_pos2, ao, best_t2);
    nm = build_cigar(e2, s2.diffs(), ao, (int)readlen2, best_scr2, cgs[e2], len2, best_pos2);
    s2.set_pos(best_pos2);
    s2.set_diffs(nm);
    const uint32_t frag_end = best_pos2 + len2;
    if (frag_end >= best_pos1 + min_dist && frag_end <= best_pos1 + max_dist) {
      best.r1 = swap_ends ? s2 : s1;
      best.r2 = swap_ends ? s1 : s2;
    }
    else best.reset();
  }
  *best_io = best;
}

// best_single (abismal.cpp:1715-1720): feed every PE candidate of set `pe_id` into SE set `se_id`
__device__ __noinline__ void best_single(int pe_id, int se_id) {
  const Warp W;
  const HeapRef pv = heap_of(W, pe_id);
  const int n = W.cs(pe_id)->sz;
  __syncwarp();
  // one lane mutates the set (see replay_hits); the entries reach it 32 at a time through the warp (large sets
  // live in HBM: one coalesced load per 32 entries instead of one dependent load per entry)
  CandSet res;
  res.load(W, se_id);
  for (int i0 = 0; i0 < n; i0 += 32) {
    const Hit h = i0 + W.lane < n ? pv.get(i0 + W.lane) : Hit(0);
    const int m = min(32, n - i0);
    for (int k = 0; k < m; ++k) {
      const Hit hk(__shfl_sync(FULL, h.w, k));
      if (W.lane == 0 && !res.sure_ambig) res.update(false, hk.diffs(), hk.flags(), hk.pos());
    }
    if (__shfl_sync(FULL, (int)res.sure_ambig, 0) != 0) break;
  }
  __syncwarp();
  if (W.lane == 0) res.store_one_lane(W, se_id);
  __syncwarp();
}

// format_se's test for writing a record (abismal.cpp:486-487)
__device__ __forceinline__ bool hit_reported(Hit h, bool allow_ambig) { return !h.empty() && (allow_ambig || !h.ambig()); }

__device__ __forceinline__ abg_hit to_abg(Hit h) {
  abg_hit r;
  r.diffs = (int16_t)h.diffs();
  r.flags = (uint16_t)h.flags();
  r.pos = h.pos();
  return r;
}

__device__ __forceinline__ void reset_set(const Warp &W, int id, int kind, uint32_t readlen) {
  __syncwarp();
  if (W.lane == 0) {
    CandSet s;
    s.v = heap_of(W, id);
    s.capacity = kSeMax;
    s.good_cutoff = 0;
    if (kind == 0) s.reset_se(readlen);
    else if (kind == 1) s.reset_pe(readlen);
    else {
      s.load(W, id);
      s.reset_se_noarg();
    }
    s.store_one_lane(W, id);
    W.cs(id)->ovf = 0u;
  }
  __syncwarp();
}

// bit k of `bits` = bucket k is non-empty; *n_set += number of non-empty buckets (n = counter_size)
__global__ void bucket_bitmap_kernel(const uint32_t *__restrict__ counter, uint64_t n, uint32_t *bits,
                                     unsigned long long *n_set) {
  const uint64_t n_round = (n + 31) & ~(uint64_t)31;
  unsigned long long local = 0;
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_round; k += (uint64_t)gridDim.x * blockDim.x) {
    const bool nonempty = k < n && counter[k + 1] != counter[k];
    const unsigned m = __ballot_sync(FULL, nonempty);
    if ((threadIdx.x & 31) == 0) {
      bits[k >> 5] = m;
      local += __popc(m);
    }
  }
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_set, local);
}

// 2-bit copy of the 4-bit genome + exception bits (see IndexDev::g2 / gx).  One thread per 32 bases.
// gi (same geometry as gx): the block holds a MULTI-bit code (IUPAC R, Y, ...; not N): such a base can match with
// popcount 2-3 in full_compare, i.e. contribute a negative amount, which the records' lower bound cannot express.
__global__ void pack_genome2_kernel(const uint64_t *__restrict__ genome, uint64_t n_words4, uint64_t n_words2,
                                    uint64_t *g2, uint32_t *gx, uint32_t *gi, unsigned int *iupac) {
  bool multi = false;  // a multi-bit (IUPAC) code was seen: *iupac = 1
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words2; w += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = 0;
    bool bad = false, multi_here = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t x = (2 * w + h) < n_words4 ? genome[2 * w + h] : 0ull;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t nib = (uint32_t)(x >> (4 * j)) & 15u;
        uint32_t code = 0;
        if (nib == 2u) code = 1;
        else if (nib == 4u) code = 2;
        else if (nib == 8u) code = 3;
        else if (nib != 1u) bad = true;
        multi_here = multi_here || (nib & (nib - 1u)) != 0u;
        lo |= (code & 1u) << (16 * h + j);
        hi |= (code >> 1) << (16 * h + j);
      }
    }
    g2[w] = (uint64_t)lo | ((uint64_t)hi << 32);
    if (bad) atomicOr(gx + (w >> 8), 1u << ((w >> 3) & 31u));  // block = (32 w) >> 8 = w >> 3
    if (multi_here) atomicOr(gi + (w >> 8), 1u << ((w >> 3) & 31u));
    multi = multi || multi_here;
  }
  if (multi) atomicOr(iupac, 1u);
}

// Compact two-letter counters (IndexDev::cc): one thread per block of kCcKeys keys.
__global__ void compact_counter_kernel(const uint32_t *__restrict__ counter, uint64_t n_keys, uint64_t n_blocks,
                                       uint4 *cc) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_blocks; b += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const uint64_t k0 = b * kCcKeys;
    uint32_t prev = counter[k0];
    w[0] = prev;
    bool big = false;
    for (uint32_t t = 0; t < kCcKeys; ++t) {
      const uint64_t k = k0 + t;
      uint32_t c = 0;
      if (k < n_keys) {
        const uint32_t nxt = counter[k + 1];
        c = nxt - prev;
        prev = nxt;
      }
      big = big || c >= 255u;
      w[1 + (t >> 2)] |= (c & 255u) << (8u * (t & 3u));
    }
    if (big) w[0] = ~0u;
    cc[2 * b] = make_uint4(w[0], w[1], w[2], w[3]);
    cc[2 * b + 1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// Seed-context records of one index table (IndexDev::ctx): one thread per (array a, entry j) writes the 128
// genome bases starting at index[j] - 32 a in g2 word format.  The writes stream; the reads are one random
// 40-byte window of the 2-bit genome per record (once per index load).
//
// gi != nullptr (the genome holds IUPAC codes): an entry whose compare window can reach a block with such a code
// (any read of up to kCtxReach bases at any seed offset) gets the all-ones SENTINEL record instead, which the
// prefilter never rejects -- the candidate then takes the exact 4-bit compare.  A real window of 128 T's reads
// as the sentinel too, which only costs it the prefilter.
constexpr uint32_t kCtxReach = 4096u + 64u;  // longest admitted read + the look-ahead of the last packed word
__global__ void seed_context_kernel(const uint32_t *__restrict__ index, uint64_t n, const uint64_t *__restrict__ g2,
                                    const uint32_t *__restrict__ gi, uint64_t n_blocks, uint4 *ctx) {
  const uint64_t total = n * (uint64_t)kCtxArrays;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t a = (uint32_t)(t / n);
    const uint32_t e = index[t - (uint64_t)a * n];
    uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    bool near_iupac = false;
    if (gi != nullptr) {
      const uint64_t b0 = (e > kCtxReach ? e - kCtxReach : 0u) >> 8;
      const uint64_t b1 = min(((uint64_t)e + kCtxReach) >> 8, n_blocks - 1);
      for (uint64_t wd = b0 >> 5; wd <= (b1 >> 5); ++wd) {
        uint32_t m = __ldg(gi + wd);
        if (wd == (b0 >> 5)) m &= ~0u << (b0 & 31u);
        if (wd == (b1 >> 5)) m &= ~0u >> (31u - (uint32_t)(b1 & 31u));
        near_iupac = near_iupac || m != 0u;
      }
    }
    if (near_iupac) {
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = ~0u;
    }
    else if (e >= 32u * a) {  // always: index entries lie beyond the 32 767-base padding
      const uint32_t p = e - 32u * a, sh = p & 31u;
      const uint64_t *gp = g2 + (p >> 5);
      uint64_t cur = __ldg(gp);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint64_t nxt = __ldg(gp + c + 1);
        w[2 * c] = __funnelshift_r((uint32_t)cur, (uint32_t)nxt, sh);
        w[2 * c + 1] = __funnelshift_r((uint32_t)(cur >> 32), (uint32_t)(nxt >> 32), sh);
        cur = nxt;
      }
    }
    ctx[2 * t] = make_uint4(w[0], w[1], w[2], w[3]);
    ctx[2 * t + 1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// kernel parameters + base-3 hash tables to shared memory (every device function reads them there)
__device__ __forceinline__ void block_prologue(const KernelParams &Pin) {
  {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(&Pin);
  uint32_t *dst = reinterpret_cast<uint32_t *>(smem_raw);
  for (uint32_t k = threadIdx.x; k < sizeof(KernelParams) / 4; k += blockDim.x) dst[k] = src[k];
  uint32_t *T3 = reinterpret_cast<uint32_t *>(smem_raw + kParamBytes);
  for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) {
    // bit j of x0/x1 is base j of the 16-base window; digit weight of base j is 3^(15-j)
    uint32_t lo = 0, hi = 0, w = 1;
    for (int j = 7; j >= 0; --j) {  // hi table: bases 8..15 (bit j -> base 8+j, weight 3^(7-j))
      if (b & (1u << j)) hi += w;
      w *= 3u;
    }
    for (int j = 7; j >= 0; --j) {  // lo table: bases 0..7 (weight 3^(15-j)), w continues at 3^8
      if (b & (1u << j)) lo += w;
      w *= 3u;
    }
    T3[b] = lo;
    T3[256 + b] = hi;
  }
}
__syncthreads();
}

// map_fragments instantiations of call `call` of a pair: which end is mapped un-reversed
// (e1, flags f1) and which reversed (e2, flags f2)
struct CallPlan {
  bool first_is_r1, swap_ends;
  uint32_t f1, f2;
  int e1, e2;
};
__device__ __forceinline__ CallPlan call_plan(int call, bool rpbat, bool a_rich) {
  CallPlan c;
  c.first_is_r1 = (call & 1) == 0;
  c.swap_ends = !c.first_is_r1;
  bool enc_a_call;  // encoding shared by both ends of this call
  if (rpbat) enc_a_call = (call == 1 || call == 2);
  else enc_a_call = a_rich ? (call == 0) : (call == 1);
  c.f1 = enc_a_call ? ABG_FLAG_A_RICH : 0u;                      // un-reversed read: a_rich bit == encoding
  c.f2 = (enc_a_call ? 0u : ABG_FLAG_A_RICH) | ABG_FLAG_RC;      // reversed read: a_rich bit == !encoding
  c.e1 = c.first_is_r1 ? 0 : 1;
  c.e2 = 1 - c.e1;
  return c;
}

// candidate set `id` <-> its stored form (kSetStateWords of state, then the first `slots` heap entries; a paired-end
// set with more entries keeps the rest in the overflow arena).  alloc: reserve arena space when the set needs
// it and has none yet (the seeding kernel; later stores reuse the reservation, sets only shrink).  Returns
// false when the set does not fit (arena full): the caller leaves the pair to the redo kernel.
__device__ __forceinline__ bool store_set(const Warp &W, int id, uint64_t *dst, int slots, bool alloc = false) {
  __syncwarp();
  const KernelParams &P = params();
  CandState *st = W.cs(id);
  const HeapRef v = heap_of(W, id);
  const int sz = st->sz;
  uint32_t ovf = st->is_pe ? st->ovf : 0u;
  bool ok = true;
  if (sz > slots) {
    if (ovf == 0u && alloc && st->is_pe && P.set_ovf != nullptr) {
      uint32_t off = 0;
      if (W.lane == 0) off = atomicAdd(P.ovf_count, (unsigned int)(sz - slots));
      off = __shfl_sync(FULL, off, 0);
      if (off + (uint32_t)(sz - slots) <= P.ovf_cap) ovf = off + 1u;
    }
    ok = ovf != 0u;
  }
  __syncwarp();
  if (W.lane == 0) {
    st->ovf = ovf;
    dst[0] = (uint64_t)(uint32_t)st->sz | ((uint64_t)(uint32_t)st->cutoff << 32);
    dst[1] = (uint64_t)(uint32_t)st->good_cutoff | ((uint64_t)(uint32_t)st->capacity << 32);
    dst[2] = (uint64_t)(uint32_t)st->sure_ambig | ((uint64_t)(uint32_t)st->is_pe << 32);
    dst[3] = st->is_pe ? (uint64_t)ovf : st->best;
  }
  for (int i = W.lane; i < sz && i < slots; i += 32) dst[kSetStateWords + i] = v.get(i).w;
  if (ovf != 0u) {
    uint64_t *o = P.set_ovf + (ovf - 1u);
    for (int i = slots + W.lane; i < sz; i += 32) o[i - slots] = v.get(i).w;
  }
  __syncwarp();
  return ok;
}
__device__ __forceinline__ void load_set(const Warp &W, int id, const uint64_t *src) {
  __syncwarp();
  const KernelParams &P = params();
  CandState *st = W.cs(id);
  const HeapRef v = heap_of(W, id);
  // __ldcg: the sets may have been written by seed_kernel on another SM while this kernel was already
  // running (overlapped launch); a neighbouring item's 128-byte line in L1 could predate them.
  const uint64_t w0 = __ldcg(src), w1 = __ldcg(src + 1), w2 = __ldcg(src + 2), w3 = __ldcg(src + 3);
  const int sz = (int)(uint32_t)w0;
  const int is_pe = (int)(uint32_t)(w2 >> 32);
  const uint32_t ovf = is_pe ? (uint32_t)w3 : 0u;
  if (W.lane == 0) {
    st->sz = sz;
    st->cutoff = (int)(uint32_t)(w0 >> 32);
    st->good_cutoff = (int)(uint32_t)w1;
    st->capacity = (int)(uint32_t)(w1 >> 32);
    st->sure_ambig = (int)(uint32_t)w2;
    st->is_pe = is_pe;
    st->best = is_pe ? 0ull : w3;
    st->ovf = ovf;
  }
  const int slots = (int)P.set_slots;
  for (int i = W.lane; i < sz && i < slots; i += 32) v.set(i, Hit(__ldcg(src + kSetStateWords + i)));
  if (ovf != 0u) {
    const uint64_t *o = P.set_ovf + (ovf - 1u);
    for (int i = slots + W.lane; i < sz; i += 32) v.set(i, Hit(__ldcg(o + (i - slots))));
  }
  __syncwarp();
}

__device__ __forceinline__ uint64_t *stored_set(const KernelParams &P, unsigned item, int pass) {
  return P.sets + ((size_t)item * P.n_pass + (size_t)pass) * (size_t)(kSetStateWords + P.set_slots);
}

// One read / pair through the path.  FROM_SETS: the candidate sets come from seed_kernel (stored_set) instead
// of being computed here by process_seeds; everything after seeding is the same code.
template <bool FROM_SETS>
__device__ __forceinline__ void map_one(const Warp &W, unsigned int item) {
  const KernelParams &P = params();
  const int lane = W.lane;
  const bool paired = P.mode & ABG_MODE_PAIRED;
  const bool a_rich = P.mode & ABG_MODE_A_RICH;
  const bool rpbat = P.mode & ABG_MODE_RANDOM_PBAT;
  WarpScalars *S = W.scal();
  const uint32_t T = 0, A = ABG_FLAG_A_RICH, RC = ABG_FLAG_RC;
  if (lane == 0) {
    S->qkey[0] = S->qkey[1] = ~0u;
    S->packed_key = ~0u;
    S->tbk[0].valid = S->tbk[1].valid = 0;
  }
  __syncwarp();

  if (!paired) {
    // map_single_ended<conv> / map_single_ended_rand (abismal.cpp:1511-1704)
    const uint32_t o0 = P.off[0][item], len = P.off[0][item + 1] - o0;
    uint32_t cg_n = 0, cg_ref = 0;
    Hit best(kMaxDiffs, 0, 0);
    if (lane == 0) S->len[0] = len;
    __syncwarp();
    if (len != 0) {
      load_end(W, 0, P.seq[0] + o0, len);
      if (FROM_SETS) load_set(W, 0, stored_set(P, item, 0));
      else if (rpbat) {
        reset_set(W, 0, 0, len);
        process_seeds(0, 0, T);
        process_seeds(0, 0, A);
        process_seeds(0, 0, A | RC);
        process_seeds(0, 0, T | RC);
      }
      else {
        reset_set(W, 0, 0, len);
        const uint32_t cv = a_rich ? A : T;
        process_seeds(0, 0, cv);
        process_seeds(0, 0, cv | RC);
      }
      const uint32_t *tof = (FROM_SETS && P.tasks != nullptr) ? P.task_of + (size_t)item * P.n_pass * P.set_slots : nullptr;
      best = Hit(align_se_candidates(0, P.valid_frac, P.cigar[0] + (size_t)item * P.cigar_stride, P.cigar_stride,
                                     &cg_n, &cg_ref, best.w, tof));
    }
    if (lane == 0) {
      P.se[0][item] = to_abg(best);
      P.n_cigar[0][item] = cg_n;
      // only a CIGAR the caller will read is an error when it does not fit (rejected hits keep their length)
      if (cg_n > P.cigar_stride && hit_reported(best, P.allow_ambig != 0u)) atomicExch(P.error_flag, 1u);
    }
    __syncwarp();
    if ((uint32_t)lane < P.inline_ops)
      P.cigar_inline[0][(size_t)item * P.inline_ops + lane] = P.cigar[0][(size_t)item * P.cigar_stride + lane];
  }
  else {
    // map_paired_ended<conv> / map_paired_ended_rand (abismal.cpp:1887-2185)
    uint32_t len0, len1;
    {
      const uint32_t oa = P.off[0][item], ob = P.off[1][item];
      len0 = P.off[0][item + 1] - oa;
      len1 = P.off[1][item + 1] - ob;
      if (lane == 0) {
        S->len[0] = len0;
        S->len[1] = len1;
      }
      __syncwarp();
      if (len0 != 0) load_end(W, 0, P.seq[0] + oa, len0);
      if (len1 != 0) load_end(W, 1, P.seq[1] + ob, len1);
    }
    CigarOut cg[2] = {{P.cigar[0] + (size_t)item * P.cigar_stride, P.cigar_stride, 0u, 0u},
                      {P.cigar[1] + (size_t)item * P.cigar_stride, P.cigar_stride, 0u, 0u}};
    reset_set(W, 0, 0, len0);
    reset_set(W, 1, 0, len1);
    PeBest best;
    best.reset(len0, len1);
    Hit best_se0(invalid_hit_diffs(len0), 0, 0), best_se1(invalid_hit_diffs(len1), 0, 0);
    bool any_success = false;
    const int n_calls = rpbat ? 4 : 2;
    for (int call = 0; call < n_calls; ++call) {
      const CallPlan cp = call_plan(call, rpbat, a_rich);
      const bool swap_ends = cp.swap_ends;
      const uint32_t f1 = cp.f1, f2 = cp.f2;
      const int e1 = cp.e1, e2 = cp.e2;
      const uint32_t l1 = cp.first_is_r1 ? len0 : len1, l2 = cp.first_is_r1 ? len1 : len0;
      reset_set(W, 2, 1, l1);
      reset_set(W, 3, 1, l2);
      if (l1 == 0 && l2 == 0) continue;
      any_success = true;
      if (FROM_SETS) {
        load_set(W, 2, stored_set(P, item, 2 * call));
        load_set(W, 3, stored_set(P, item, 2 * call + 1));
      }
      else {
        if (l1 != 0) process_seeds(2, e1, f1);
        if (l2 != 0) process_seeds(3, e2, f2);
      }
      // select_maps (abismal.cpp:1833-1847)
      CandSet t0, t1;
      t0.load(W, 2);
      t1.load(W, 3);
      if (t0.should_align() && t1.should_align()) {
        if (FROM_SETS && P.tasks != nullptr) {  // enum_kernel has sorted + uniqued the stored sets and emitted the tasks
          const uint32_t *tof = P.task_of + ((size_t)item * P.n_pass + 2u * call) * P.set_slots;
          best_pair(swap_ends, e1, f1, f2, cg, &best, tof, tof + P.set_slots);
        }
        else {
          sort_unique(2);
          sort_unique(3);
          best_pair(swap_ends, e1, f1, f2, cg, &best);
        }
      }
      best_single(2, e1);
      best_single(3, e2);
    }
    if (!any_success) {
      best.reset();
      reset_set(W, 0, 2, 0);
      reset_set(W, 1, 2, 0);
    }
    {  // valid_pair (abismal.cpp:624-631)
      const uint32_t al1 = cg[0].ref_len, al2 = cg[1].ref_len;
      const uint32_t min_len = P.window_size + 24u;
      const bool ok = valid_len(al1, len0, min_len) && valid_len(al2, len1, min_len) &&
                      best.diffs() <= frac_of(P.valid_frac, al1 + al2);
      if (!ok) best.reset();
    }
    if (!best.should_report(P.allow_ambig != 0u)) {
      const double half = P.valid_frac / 2.0;
      best_se0 = Hit(align_se_candidates(0, half, cg[0].ops, cg[0].stride, &cg[0].n, &cg[0].ref_len, best_se0.w));
      best_se1 = Hit(align_se_candidates(1, half, cg[1].ops, cg[1].stride, &cg[1].n, &cg[1].ref_len, best_se1.w));
    }
    if (lane == 0) {
      P.pe_r1[item] = to_abg(best.r1);
      P.pe_r2[item] = to_abg(best.r2);
      P.se[0][item] = to_abg(best_se0);
      P.se[1][item] = to_abg(best_se1);
      P.n_cigar[0][item] = cg[0].n;
      P.n_cigar[1][item] = cg[1].n;
      const bool amb_ok = P.allow_ambig != 0u;
      const bool pair_out = best.should_report(amb_ok);
      if ((cg[0].n > cg[0].stride && (pair_out || hit_reported(best_se0, amb_ok))) ||
          (cg[1].n > cg[1].stride && (pair_out || hit_reported(best_se1, amb_ok))))
        atomicExch(P.error_flag, 1u);
    }
    __syncwarp();
    if ((uint32_t)lane < P.inline_ops) {
      P.cigar_inline[0][(size_t)item * P.inline_ops + lane] = cg[0].ops[lane];
      P.cigar_inline[1][(size_t)item * P.inline_ops + lane] = cg[1].ops[lane];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void flush_counters(const Warp &W) {
  const KernelParams &P = params();
  const WarpScalars *S = W.scal();
  if (P.counters != nullptr && W.lane == 0) {
    atomicAdd(P.counters + 0, S->cnt[0]);  // n_lookup
    atomicAdd(P.counters + 1, S->cnt[1]);  // n_entry
    atomicAdd(P.counters + 2, S->cnt[1]);  // n_cmp == n_entry
    atomicAdd(P.counters + 3, S->cnt[2]);  // n_word
    atomicAdd(P.counters + 4, S->cnt[3]);  // n_align
    atomicAdd(P.counters + 5, S->cnt[4]);  // n_dpref
  }
}

// MINB = resident CTAs per SM the register allocation is bounded for (2: <=128 regs, 3: <=80, 4: <=64)
template <int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) map_reads_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const Warp W;
  const int lane = W.lane;
  if (lane < 6) W.scal()->cnt[lane] = 0;
  __syncwarp();
  const unsigned int n_items = P.n_items_ptr != nullptr ? *P.n_items_ptr : P.n;
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= n_items) break;
    if (P.item_list != nullptr) item = P.item_list[item];
    map_one<false>(W, item);
  }
  flush_counters(W);
}

// Poll with a relaxed GPU-scope load: ld.acquire would add a CCTL.IVALL (the whole L1 of the SM invalidated) to
// every poll, which wrecks the seeding kernel running on the same SM.  Everything read after the flag is read
// with L1-bypassing loads (__ldcg), issued after the branch on the flag resolves.
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_cg(const unsigned int *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Publish one stored set of `item` (all lanes: their stores first, then one release increment)
__device__ __forceinline__ void publish_set(const KernelParams &P, unsigned int item, int lane) {
  if (P.ready == nullptr) return;
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicAdd(P.ready + item, 1u);
}

// Phase 2 of the two-phase launch: everything after seeding (sort/unique, banded alignment, mating, selection,
// CIGARs) for the reads / pairs whose sets seed_kernel stored; flagged pairs are left to map_reads_kernel.
template <int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) align_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const Warp W;
  const int lane = W.lane;
  if (lane < 6) W.scal()->cnt[lane] = 0;
  __syncwarp();
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= P.n) break;
    if (P.ready != nullptr) {
      // overlapped launch: wait until seed_kernel has stored every set of this item (acquire)
      unsigned int ok = 1;
      if (lane == 0 && ld_acquire(P.ready + item) < P.ready_need) {
        const unsigned long long t0 = global_ns();
        for (;;) {
          __nanosleep(1000);
          if (ld_acquire(P.ready + item) >= P.ready_need) break;
          if (global_ns() - t0 > (unsigned long long)P.wait_ns) {
            ok = 0;
            break;
          }
        }
      }
      ok = __shfl_sync(FULL, ok, 0);
      if (!ok) {
        // The seeding of this item is not coming (its kernel is not resident, or extremely slow): leave the
        // item to the redo kernel, which maps it from scratch after both kernels, and stop waiting here.
        if (lane == 0 && atomicExch(P.redo_flag + item, 1u) == 0u) P.redo_list[atomicAdd(P.redo_count, 1u)] = item;
        break;
      }
    }
    if (ld_relaxed_cg(P.redo_flag + item) != 0) continue;
    map_one<true>(W, item);
  }
  flush_counters(W);
}

// seed_bins.cuh: process_seeds for a strand whose prefilter survivors the binned kernels have listed
__device__ void process_binned(int set_id, int end, uint32_t strand_code, uint32_t sid, const char *seq);

// Phase 1 of the two-phase launch: seeding only.  Paired: one work item per (pair, call, side) -- the passes
// of a pair are independent (every map_fragments call starts from reset pe_candidates) -- so a warp hashes,
// probes and compares ONE read strand and stores the resulting set.  Single-end: the passes of a read feed
// the same se_candidates in order, so the work item is the read.
template <int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) seed_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const Warp W;
  const int lane = W.lane;
  const bool paired = P.mode & ABG_MODE_PAIRED;
  const bool a_rich = P.mode & ABG_MODE_A_RICH;
  const bool rpbat = P.mode & ABG_MODE_RANDOM_PBAT;
  WarpScalars *S = W.scal();
  if (lane < 6) S->cnt[lane] = 0;
  __syncwarp();
  const uint32_t T = 0, A = ABG_FLAG_A_RICH, RC = ABG_FLAG_RC;
  const unsigned int n_work = paired ? P.n * P.n_pass : P.n;
  for (;;) {
    unsigned int w = 0;
    if (lane == 0) w = atomicAdd(P.work_counter, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= n_work) break;
    if (lane == 0) {
      S->qkey[0] = S->qkey[1] = ~0u;
      S->packed_key = ~0u;
    }
    __syncwarp();
    if (!paired) {
      const unsigned int item = w;
      const uint32_t o0 = P.off[0][item], len = P.off[0][item + 1] - o0;
      if (lane == 0) S->len[0] = len;
      __syncwarp();
      if (len == 0) {  // skipped read: nothing stored, but the alignment kernel must not wait for it
        publish_set(P, item, lane);
        continue;
      }
      const bool binned = P.bp.tup != nullptr;
      if (!binned) load_end(W, 0, P.seq[0] + o0, len);
      else if (lane == 0) S->loaded[0] = 0u;
      reset_set(W, 0, 0, len);
      if (binned) {  // binned seeding: the strands in the order of strand_plan (seed_bins.cuh)
        const uint32_t s0 = P.bp.sid_base + item * P.bp.spi;
        const char *sq = P.seq[0] + o0;
        if (rpbat) {
          process_binned(0, 0, T, s0, sq);
          process_binned(0, 0, A, s0 + 1, sq);
          process_binned(0, 0, A | RC, s0 + 2, sq);
          process_binned(0, 0, T | RC, s0 + 3, sq);
        }
        else {
          const uint32_t cv = a_rich ? A : T;
          process_binned(0, 0, cv, s0, sq);
          process_binned(0, 0, cv | RC, s0 + 1, sq);
        }
      }
      else if (rpbat) {
        process_seeds(0, 0, T);
        process_seeds(0, 0, A);
        process_seeds(0, 0, A | RC);
        process_seeds(0, 0, T | RC);
      }
      else {
        const uint32_t cv = a_rich ? A : T;
        process_seeds(0, 0, cv);
        process_seeds(0, 0, cv | RC);
      }
      store_set(W, 0, stored_set(P, item, 0), (int)P.set_slots);
      publish_set(P, item, lane);
    }
    else {
      const unsigned int item = w / P.n_pass;
      const int pass = (int)(w % P.n_pass);
      const CallPlan cp = call_plan(pass >> 1, rpbat, a_rich);
      const int end = (pass & 1) ? cp.e2 : cp.e1;
      const uint32_t flags = (pass & 1) ? cp.f2 : cp.f1;
      const uint32_t o = P.off[end][item], len = P.off[end][item + 1] - o;
      if (lane == 0) S->len[end] = len;
      __syncwarp();
      reset_set(W, 2, 1, len);
      if (len != 0) {
        if (P.bp.tup != nullptr) {
          if (lane == 0) S->loaded[end] = 0u;
          process_binned(2, end, flags, P.bp.sid_base + w, P.seq[end] + o);
        }
        else {
          load_end(W, end, P.seq[end] + o, len);
          process_seeds(2, end, flags);
        }
      }
      if (!store_set(W, 2, stored_set(P, item, pass), (int)P.set_slots, true)) {
        // grew beyond the stored form and the overflow arena is full (or absent): the whole pair is redone
        if (lane == 0 && atomicExch(P.redo_flag + item, 1u) == 0u) P.redo_list[atomicAdd(P.redo_count, 1u)] = item;
      }
      publish_set(P, item, lane);
    }
  }
  flush_counters(W);
}

}  // namespace ab2dev
