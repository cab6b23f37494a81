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
// Parallel decomposition inside the warp (the results are order dependent in
// the reference, so only PURE quantities are computed in parallel):
//   * 32 seed offsets at a time: each lane hashes one offset and gathers its two
//     counter pairs; bucket sizes are prefix-summed across the warp so that
//   * 32 candidates at a time (in the reference's canonical order: offset,
//     two-letter bucket before three-letter bucket, bucket order) each get one
//     lane doing the index gather + packed-genome gather + popcount compare;
//   * survivors (ballot) are replayed IN ORDER against the candidate set with
//     libstdc++'s heap routines restated verbatim, executed redundantly by all
//     lanes on warp-uniform state (same-value writes), so the cutoff tightening,
//     evictions, sure_ambig exits and the specific->sensitive gate match the
//     reference bit for bit;
//   * banded DP: lanes are band columns (1 or 2 per lane), rows are sequential,
//     the serial from_left recurrence is a warp max-plus prefix scan, traceback
//     arrows go to 2-bit planes written with ballots.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

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
constexpr int kNegInf = -(1 << 28);
constexpr uint32_t kHashMaskThree = 43046721u;

struct IndexDev {
  const uint64_t *genome;
  const uint32_t *counter, *counter_t, *counter_a;
  const uint32_t *index, *index_t, *index_a;
  uint32_t max_candidates;
};

struct KernelParams {
  IndexDev ix;
  // batch
  uint32_t n;
  const char *seq[2];
  const uint32_t *off[2];
  // results
  abg_hit *pe_r1, *pe_r2, *se[2];
  uint32_t *cigar[2];
  uint32_t *n_cigar[2];
  uint32_t cigar_stride;
  // params
  uint32_t mode, allow_ambig, min_dist, max_dist, max_candidates;
  double valid_frac;
  // per-warp-slot scratch
  uint32_t ml;  // padded max read length (multiple of 32)
  uint64_t *pe_overflow;  // [slots][2][kPeLarge]
  int16_t *mem_scr;       // [slots][kPeLarge]
  uint32_t *tb;           // [slots][tb_rows][4]
  uint32_t tb_rows;
  unsigned int *work_counter;
  unsigned int *error_flag;
  unsigned long long *counters;  // abg_work_counters layout, or nullptr
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

// heap storage: first `cap_sm` entries in shared memory, the rest in global
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

// ---- thresholds: evaluated in double exactly as the reference writes them ----
__device__ __forceinline__ int frac_of(double f, uint32_t x) {  // static_cast<score_t>(f * x)
  return (int)(int16_t)__double2int_rz(__dmul_rn(f, (double)x));
}
__device__ __forceinline__ int invalid_hit_diffs(uint32_t readlen) { return frac_of(0.4, readlen); }
__device__ __forceinline__ bool valid_len(uint32_t aln_len, uint32_t readlen) {  // abismal.cpp:307-314
  const double min_aln_frac = __dsub_rn(1.0, 0.4);
  const uint32_t a = (uint32_t)__double2uint_rz(__dmul_rn(min_aln_frac, (double)readlen));
  return aln_len >= (a > 44u ? a : 44u);
}

// ---- candidate set: se_candidates or pe_candidates, warp-uniform state -------
struct CandSet {
  HeapRef v;
  int sz, cutoff, good_cutoff, capacity;
  bool sure_ambig, is_pe;
  Hit best;  // SE only

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
  __device__ void update(bool specific, int d, uint32_t flags, uint32_t pos) {
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

// ---- per-warp context ----------------------------------------------------------
struct WarpCtx {
  const KernelParams *P;
  int lane;
  // shared memory
  uint8_t *base[2];   // one-hot base codes of each end, FASTQ orientation
  uint8_t *qcode;     // current pass: bisulfite-encoded read, zero padded
  uint64_t *packed;   // current pass: pack_read
  uint64_t *refw;     // DP: staged genome words
  uint32_t len[2];
  uint32_t cur_key;   // which (end, flags) is in qcode/packed; ~0u = none
  // global scratch
  uint32_t *tb;
  int16_t *mem_scr;
  // counters (lane-local partial sums, reduced at the end)
  unsigned long long c_lookup, c_entry, c_word, c_align, c_dpref;
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
__device__ void load_end(WarpCtx &c, int end, const char *s, uint32_t n) {
  c.len[end] = n;
  for (uint32_t i = c.lane; i < n; i += 32) {
    const char ch = s[i];
    uint8_t b = 0;
    if (ch == 'A' || ch == 'a') b = 1;
    else if (ch == 'C' || ch == 'c') b = 2;
    else if (ch == 'G' || ch == 'g') b = 4;
    else if (ch == 'T' || ch == 't') b = 8;
    c.base[end][i] = b;
  }
  __syncwarp();
}

// prep_read + pack_read for the pass identified by `flags` on `end`:
// orientation by the rc bit, encoding by a_rich XOR rc (abismal.cpp:1463-1465).
__device__ void build_pass(WarpCtx &c, int end, uint32_t flags) {
  const uint32_t key = ((uint32_t)end << 16) | (flags & (ABG_FLAG_RC | ABG_FLAG_A_RICH));
  if (c.cur_key == key) return;
  c.cur_key = key;
  const bool rc = flags & ABG_FLAG_RC;
  const bool enc_a = ((flags & ABG_FLAG_A_RICH) != 0) != rc;
  const uint32_t n = c.len[end];
  const uint8_t *b = c.base[end];
  __syncwarp();
  for (uint32_t i = c.lane; i < n + 32; i += 32) {
    uint32_t code = 0;
    if (i < n) {
      uint32_t x = rc ? b[n - 1 - i] : b[i];
      if (rc) x = ((x & 1u) << 3) | ((x & 2u) << 1) | ((x & 4u) >> 1) | ((x & 8u) >> 3);  // complement
      code = enc_a ? (x == 1u ? 5u : x) : (x == 8u ? 10u : x);
    }
    c.qcode[i] = (uint8_t)code;
  }
  __syncwarp();
  const uint32_t nw = (n + 15) / 16;
  for (uint32_t w = c.lane; w < nw; w += 32) {
    uint64_t word = 0;
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) {
      const uint32_t i = 16 * w + j;
      const uint64_t nib = i < n ? c.qcode[i] : 0xFull;  // tail matches anything :1424-1425
      word |= nib << (4 * j);
    }
    c.packed[w] = word;
  }
  __syncwarp();
}

// full_compare (abismal.cpp:1105-1122): exact distance if it is <= cutoff,
// otherwise some value > cutoff (early exit).
__device__ __forceinline__ int full_compare(const uint64_t *__restrict__ genome, uint32_t the_pos,
                                            const uint64_t *packed, int n_words, int cutoff, int &words_seen) {
  const uint64_t *g = genome + (the_pos >> 4);
  const uint32_t off = (the_pos & 15u) << 2;
  int d = 0;
  uint64_t cur = __ldg(g);
  int w = 0;
  for (; w < n_words && d <= cutoff; ++w) {
    const uint64_t nxt = __ldg(g + w + 1);
    const uint64_t gw = (cur >> off) | ((nxt << (63u - off)) << 1);
    d += 16 - __popcll(packed[w] & gw);
    cur = nxt;
  }
  words_seen = w;
  return d;
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

// find_candidates<25> (abismal.cpp:1163-1194); `read_start` = qcode + i
__device__ uint32_t find_candidates(const IndexDev &ix, uint32_t maxc, const uint8_t *read_start, uint32_t read_lim,
                                    uint32_t &low, uint32_t &high) {
  uint32_t p = 25;
  uint32_t prev_low = low, prev_high = high;
  for (; p != read_lim && (high - low) > maxc; ++p) {
    prev_low = low;
    prev_high = high;
    const uint32_t first_1 = lower_bound_idx(ix.index, low, high, [&](uint32_t e) {
      return get_bit(genome_base(ix.genome, (uint64_t)e + p)) < 1u;
    });
    const uint32_t the_bit = get_bit(read_start[p]);
    high = the_bit ? high : first_1;
    low = the_bit ? first_1 : low;
  }
  if (low == high) {
    --p;
    low = prev_low;
    high = prev_high;
  }
  return p;
}

// find_candidates_three<16, conv> (abismal.cpp:1214-1259)
__device__ uint32_t find_candidates_three(const IndexDev &ix, const uint32_t *index3, bool g_to_a, uint32_t maxc,
                                          const uint8_t *read_start, uint32_t max_size, uint32_t &low,
                                          uint32_t &high) {
  uint32_t p = 16;
  uint32_t prev_low = low, prev_high = high;
  const uint32_t v1 = g_to_a ? 2u : 1u, v2 = g_to_a ? 8u : 4u;
  for (; p != max_size && (high - low) > maxc; ++p) {
    prev_low = low;
    prev_high = high;
    const uint32_t first_1 = lower_bound_idx(index3, low, high, [&](uint32_t e) {
      return three_fast(g_to_a, genome_base(ix.genome, (uint64_t)e + p)) < v1;
    });
    const uint32_t first_2 = lower_bound_idx(index3, low, high, [&](uint32_t e) {
      return three_fast(g_to_a, genome_base(ix.genome, (uint64_t)e + p)) < v2;
    });
    const uint32_t the_num = three_fast(g_to_a, read_start[p]);
    const uint32_t old_low = low, old_high = high;
    high = (the_num == 0u) ? first_1 : ((the_num == v1) ? first_2 : old_high);
    low = (the_num == 0u) ? old_low : ((the_num == v1) ? first_1 : first_2);
  }
  if (low == high) {
    --p;
    low = prev_low;
    high = prev_high;
  }
  return p;
}

__device__ __forceinline__ uint32_t warp_incl_scan_add(uint32_t x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// process_seeds (abismal.cpp:1269-1375) for the pass currently in qcode/packed
__device__ __noinline__ void process_seeds(WarpCtx &c, uint32_t strand_code, uint32_t readlen, CandSet &res) {
  const KernelParams &P = *c.P;
  const IndexDev &ix = P.ix;
  const int lane = c.lane;
  const bool g_to_a = ((strand_code & ABG_FLAG_A_RICH) != 0) != ((strand_code & ABG_FLAG_RC) != 0);
  const uint32_t *counter3 = g_to_a ? ix.counter_a : ix.counter_t;
  const uint32_t *index3 = g_to_a ? ix.index_a : ix.index_t;
  const uint32_t maxc = P.max_candidates;
  const int n_words = (int)((readlen + 15) / 16);
  const bool count = P.counters != nullptr;

  const uint32_t specific_len = min(readlen - 20u, readlen >> 1);
  const uint32_t specific_lim = max(20u, readlen >> 1);
  const uint32_t lim_two = readlen - 25u + 1u;

  for (int phase = 0; phase < 2; ++phase) {
    const bool specific = phase == 0;
    if (specific) res.set_specific();
    else {
      if (!res.should_do_sensitive()) return;
      res.set_sensitive();
    }
    const uint32_t n_off = specific ? specific_lim : lim_two;
    for (uint32_t base_off = 0; base_off < n_off && !res.sure_ambig; base_off += 32) {
      const uint32_t i = base_off + lane;
      const bool active = i < n_off;
      uint32_t s2 = 0, e2 = 0, s3 = 0, e3 = 0, n2 = 0, n3 = 0;
      if (active) {
        // get_1bit_hash / get_base_3_hash at offset i (rolling == direct)
        const uint8_t *r = c.qcode + i;
        uint32_t k = 0, k3 = 0;
#pragma unroll
        for (int j = 0; j < 25; ++j) k = (k << 1) | get_bit(r[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) k3 = k3 * 3u + three_num(g_to_a, r[j]);
        s2 = __ldg(ix.counter + k);
        e2 = __ldg(ix.counter + k + 1);
        s3 = __ldg(counter3 + k3);
        e3 = __ldg(counter3 + k3 + 1);
        if (specific) {
          uint32_t l_two = 24, l_three = 15;
          if (e2 - s2 > maxc || e2 == s2) l_two = find_candidates(ix, maxc, r, readlen - i, s2, e2);
          else l_two = 25;
          if (e3 - s3 > maxc || e3 == s3)
            l_three = find_candidates_three(ix, index3, g_to_a, maxc, r, readlen - i, s3, e3);
          else l_three = 16;
          const uint32_t d_two = e2 - s2, d_three = e3 - s3;
          n2 = (d_two <= maxc || l_two >= specific_len) ? d_two : 0u;
          n3 = (d_three <= maxc || l_three >= specific_len) ? d_three : 0u;
        }
        else {
          const uint32_t d_two = e2 - s2, d_three = e3 - s3;
          n2 = (d_two != 0u && d_two <= maxc && (d_three == 0u || d_two <= 10u * d_three)) ? d_two : 0u;
          n3 = (d_three != 0u && d_three <= maxc) ? d_three : 0u;
        }
        if (count) c.c_lookup += 2;
      }
      __syncwarp();
      const uint32_t tot = n2 + n3;
      const uint32_t incl = warp_incl_scan_add(tot, lane);
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      for (uint32_t c0 = 0; c0 < total && !res.sure_ambig; c0 += 32) {
        const uint32_t cidx = c0 + lane;
        const bool valid = cidx < total;
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
        int d = kMaxDiffs;
        uint32_t the_pos = 0;
        const int cutoff = res.cutoff;
        if (valid) {
          const uint32_t r = cidx - (o_incl - o_tot);
          const uint32_t entry = (r < o_n2) ? __ldg(ix.index + o_s2 + r) : __ldg(index3 + o_s3 + (r - o_n2));
          the_pos = entry - (base_off + (uint32_t)o);
          int words = 0;
          d = full_compare(ix.genome, the_pos, c.packed, n_words, cutoff, words);
          if (count) {
            c.c_entry += 1;
            c.c_word += (unsigned long long)words;
          }
        }
        __syncwarp();
        unsigned mask = __ballot_sync(FULL, valid && d <= cutoff);
        while (mask != 0u && !res.sure_ambig) {
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          const int dd = __shfl_sync(FULL, d, l);
          const uint32_t pp = __shfl_sync(FULL, the_pos, l);
          if (dd <= res.cutoff) res.update(true, dd, strand_code, pp);
        }
      }
    }
  }
}

// ---- sort by (pos, flags) + unique: prepare_for_alignments / prepare_for_mating ----
__device__ __noinline__ void sort_unique(const HeapRef &v, int &sz, int lane) {
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
  sz = out;
}

// ---- banded alignment --------------------------------------------------------------
struct AlnOut {
  int score, row, col, bw;
};

__device__ __forceinline__ int band_width(int diffs, int max_diffs) {  // AbismalAlign.hpp:333-334
  const int want = 2 * min(diffs, max_diffs) + 1;
  return want < 0 ? 61 : min(61, want);
}

// AbismalAlign::align<do_traceback> (AbismalAlign.hpp:320-386); diffs != 0.
// CPL = band columns per lane.  Query = c.qcode (length q_sz).
template <int CPL>
__device__ void align_rows(WarpCtx &c, bool do_tb, int bw, int q_sz, uint32_t t_pos, AlnOut &out) {
  const int lane = c.lane;
  const uint32_t t_beg = t_pos - (uint32_t)((bw - 1) / 2);
  const int t_shift = q_sz + bw;
  const uint32_t w0 = t_beg >> 4;
  const int nw = (int)(((t_beg + (uint32_t)t_shift - 2u) >> 4) - w0) + 1;
  __syncwarp();
  for (int k = lane; k < nw; k += 32) c.refw[k] = __ldg(c.P->ix.genome + w0 + k);
  if (do_tb && lane < 4) c.tb[lane] = 0xffffffffu;  // row 0: every cell is "stop"
  __syncwarp();

  int prev[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) prev[k] = 0;
  int best = 0, best_row = 0, best_col = 0;
  const uint8_t *q = c.qcode;

  for (int i = 1; i < t_shift; ++i) {
    const int left = i < bw ? bw - i : 0;
    const int right = min(bw, t_shift - i);
    const uint32_t gp = t_beg + (uint32_t)i - 1u;
    const uint32_t ref = (uint32_t)(c.refw[(gp >> 4) - w0] >> ((gp & 15u) << 2)) & 15u;
    // prev[j + 1] of the last column of this lane lives in the next lane
    const int nxt_lane_first = __shfl_down_sync(FULL, prev[0], 1);
    int val[CPL], arrow[CPL], u[CPL];
    bool in[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = lane * CPL + k;
      in[k] = j >= left && j < right;
      int v = 0, a = 3;
      if (in[k]) {
        const uint32_t qb = q[i + j - bw];
        const int diag = prev[k] + ((qb & ref) ? 2 : -3);
        v = max(0, diag);
        a = (v == diag) ? 0 : 3;
        if (j + 1 < right) {
          const int above = ((k + 1 < CPL) ? prev[(k + 1) % CPL] : nxt_lane_first) - 4;
          v = max(v, above);
          if (v == above) a = 2;
        }
      }
      val[k] = v;
      arrow[k] = a;
      u[k] = in[k] ? v + 4 * j : kNegInf;
    }
    // from_left: cur[j] = max(val[j], cur[j-1] - 4)  ==  prefix-max of (val[j] + 4j) - 4j
    int lane_max = u[0];
#pragma unroll
    for (int k = 1; k < CPL; ++k) lane_max = max(lane_max, u[k]);
    int incl = lane_max;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl = max(incl, y);
    }
    int run = __shfl_up_sync(FULL, incl, 1);  // prefix max over all earlier columns
    if (lane == 0) run = kNegInf;
    unsigned code_bits = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = lane * CPL + k;
      int cur = 0, code = 3;
      if (in[k]) {
        const int U = max(run, u[k]);
        cur = U - 4 * j;
        int a = arrow[k];
        if (u[k] <= run) a = 1;  // cur[j] == cur[j-1] - 4  => I (precedence I > D > M)
        code = cur > 0 ? a : 3;
        if (cur > best) {
          best = cur;
          best_row = i;
          best_col = j;
        }
        run = U;
      }
      prev[k] = cur;
      code_bits |= (unsigned)code << (2 * k);
    }
    if (do_tb) {
#pragma unroll
      for (int pl = 0; pl < 2 * CPL; ++pl) {
        const unsigned m = __ballot_sync(FULL, (code_bits >> pl) & 1u);
        if (lane == pl) c.tb[(size_t)i * 4 + pl] = m;
      }
    }
  }
  // first maximum in row-major order (std::max_element)
  int bv = best, br = best_row, bc = best_col;
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
  out.score = bv;
  out.row = br;
  out.col = bc;
  out.bw = bw;
  __syncwarp();
}

// returns the alignment score; out is meaningful only when diffs != 0
__device__ __noinline__ int align(WarpCtx &c, bool do_tb, int diffs, int max_diffs, int q_sz, uint32_t t_pos,
                                  AlnOut &out) {
  if (diffs == 0) return 2 * q_sz;  // AbismalAlign.hpp:329-330
  const int bw = band_width(diffs, max_diffs);
  if (c.P->counters != nullptr && c.lane == 0) {
    c.c_align += 1;
    c.c_dpref += (unsigned long long)(q_sz + bw);
  }
  if (bw <= 32) align_rows<1>(c, do_tb, bw, q_sz, t_pos, out);
  else align_rows<2>(c, do_tb, bw, q_sz, t_pos, out);
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
__device__ __noinline__ int build_cigar(WarpCtx &c, int diffs, const AlnOut &a, int q_sz, int scr_for_nm,
                                        CigarOut &cg, uint32_t &len, uint32_t &t_pos) {
  const int lane = c.lane;
  int ins = 0, del = 0;
  if (diffs == 0 || a.score == 0) {
    cigar_default(cg, (uint32_t)q_sz, lane);
    len = (uint32_t)q_sz;
  }
  else {
    const int bw = a.bw;
    const int cpl = bw <= 32 ? 1 : 2;
    int row = a.row, col = a.col;
    const int clip_bottom = (q_sz + (bw - 1)) - (row + col);
    const auto code_at = [&](int r, int cc) -> int {
      if (cc < 0 || cc >= bw || r <= 0) return 3;
      const int ln = cc / cpl, k = cc % cpl;
      const uint32_t p0 = c.tb[(size_t)r * 4 + 2 * k], p1 = c.tb[(size_t)r * 4 + 2 * k + 1];
      return (int)((p0 >> ln) & 1u) | ((int)((p1 >> ln) & 1u) << 1);
    };
    uint32_t n_ops = 0, ref_len = 0;
    const auto emit = [&](uint32_t n, int op) {
      if (n_ops < cg.stride && lane == 0) cg.ops[n_ops] = (n << 4) | (uint32_t)op;
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
      const int arrow = code_at(row, col);
      if (arrow == 3) break;  // table[row][col] <= 0
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
    emit(n, prev_arrow);
    const int clip_top = (row + col) - (bw - 1);
    if (clip_top > 0) {
      if (n_ops < cg.stride && lane == 0) cg.ops[n_ops] = ((uint32_t)clip_top << 4) | 4u;
      ++n_ops;
    }
    __syncwarp();
    // reverse in place
    const uint32_t m = min(n_ops, cg.stride);
    if (n_ops <= cg.stride) {
      for (uint32_t k = lane; k < m / 2; k += 32) {
        const uint32_t x = cg.ops[k], y = cg.ops[m - 1 - k];
        cg.ops[k] = y;
        cg.ops[m - 1 - k] = x;
      }
    }
    __syncwarp();
    if (clip_bottom > 0) {
      if (n_ops < cg.stride && lane == 0) cg.ops[n_ops] = ((uint32_t)clip_bottom << 4) | 4u;
      ++n_ops;
    }
    __syncwarp();
    cg.n = n_ops;
    cg.ref_len = ref_len;
    len = (uint32_t)(q_sz - clip_bottom - clip_top);
    const uint32_t t_beg = t_pos - (uint32_t)((bw - 1) / 2);
    t_pos = t_beg + (uint32_t)row;
  }
  // edit_distance(scr, len, cigar): same promotions as the reference (unsigned quotient)
  if (scr_for_nm == 0) return (int)(int16_t)len;
  const int A = (int)(int16_t)(scr_for_nm + 4 * (ins + del));
  const uint32_t num = 2u * (len - (uint32_t)ins) - (uint32_t)A;
  const int mism = (int)(int16_t)(num / 5u);
  return (int)(int16_t)(mism + ins + del);
}

__device__ __forceinline__ bool same_pos(uint32_t a, uint32_t b) { return (a > b ? a - b : b - a) <= 3u; }

// align_se_candidates (abismal.cpp:1435-1497)
__device__ __noinline__ void align_se_candidates(WarpCtx &c, int end, uint32_t readlen_u, double cutoff,
                                                 CandSet &res, Hit &best, CigarOut &cg) {
  const int readlen = (int)(int16_t)readlen_u;
  const int max_diffs = frac_of(cutoff, (uint32_t)readlen);
  const int max_scr = (int)(int16_t)(2 * readlen);
  if (!res.best.empty()) {
    best = res.best;
    cigar_default(cg, (uint32_t)readlen, c.lane);
    return;
  }
  int best_scr = 0;
  uint32_t best_pos = 0;
  sort_unique(res.v, res.sz, c.lane);
  int it = 0;
  const int lim = res.sz;
  for (; it != lim && res.v.get(it).empty(); ++it) {
  }
  const int invalid = frac_of(0.4, (uint32_t)readlen);
  AlnOut ao;
  for (; it != lim; ++it) {
    const Hit h = res.v.get(it);
    if (h.diffs() < invalid) {
      build_pass(c, end, h.flags());
      const uint32_t cand_pos = h.pos();
      const int cand_scr = (int)(int16_t)align(c, false, h.diffs(), max_diffs, (int)c.len[end], cand_pos, ao);
      if (cand_scr > best_scr) {
        best = h;
        best_scr = cand_scr;
        best_pos = cand_pos;
      }
      else if (cand_scr == best_scr && (cand_scr == max_scr ? cand_pos != best_pos : !same_pos(cand_pos, best_pos)))
        best.set_ambig();
    }
  }
  if (best.pos() != 0) {
    build_pass(c, end, best.flags());
    ao.score = 0;
    align(c, true, best.diffs(), max_diffs, (int)c.len[end], best.pos(), ao);
    uint32_t len = 0, pos = best.pos();
    const int nm = build_cigar(c, best.diffs(), ao, (int)c.len[end], best_scr, cg, len, pos);
    best.set_pos(pos);
    best.set_diffs(nm);
    if (!(valid_len(len, (uint32_t)readlen) && nm <= frac_of(cutoff, (uint32_t)readlen))) best.reset();
  }
  else best.reset();
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

// best_pair<swap_ends> (abismal.cpp:1722-1831).  e1/e2 = which end of the pair
// plays "1" (un-reversed) / "2" (reversed) in this map_fragments call.
__device__ __noinline__ void best_pair(WarpCtx &c, bool swap_ends, int e1, uint32_t flags1, int e2, uint32_t flags2,
                                       const CandSet &res1, const CandSet &res2, CigarOut &cg1, CigarOut &cg2,
                                       PeBest &best) {
  const KernelParams &P = *c.P;
  const int j1_end = res1.sz, j2_end = res2.sz;
  int j1 = 0, j2 = 0;
  __syncwarp();
  for (int k = c.lane; k < res1.sz; k += 32) c.mem_scr[k] = 0;
  __syncwarp();
  const uint32_t readlen1 = c.len[e1], readlen2 = c.len[e2];
  const int max_diffs1 = frac_of(P.valid_frac, readlen1);
  const int max_diffs2 = frac_of(P.valid_frac, readlen2);
  const uint32_t min_dist = P.min_dist, max_dist = P.max_dist;
  int scr1 = 0, best_scr1 = 0, best_scr2 = 0;
  uint32_t best_pos1 = 0, best_pos2 = 0;
  AlnOut ao;

  for (; j1 != j1_end && res1.v.get(j1).empty(); ++j1) {
  }
  for (; j2 != j2_end && res2.v.get(j2).empty(); ++j2) {
  }
  for (; j2 != j2_end && !best.sure_ambig(); ++j2) {
    const Hit s2 = res2.v.get(j2);
    int scr2 = 0;
    const uint32_t lim = s2.pos() + readlen2;
    for (; (j1 == j1_end) || (j1 != 0 && res1.v.get(j1).pos() + max_dist >= lim); --j1) {
    }
    for (; j1 != j1_end && res1.v.get(j1).pos() + max_dist < lim; ++j1) {
    }
    for (; j1 != j1_end && !best.sure_ambig(); ++j1) {
      const Hit s1 = res1.v.get(j1);
      if (!(s1.pos() + min_dist <= lim)) break;
      if (scr2 == 0) {
        build_pass(c, e2, flags2);
        scr2 = (int)(int16_t)align(c, false, s2.diffs(), max_diffs2, (int)readlen2, s2.pos(), ao);
      }
      int m1 = c.mem_scr[j1];
      if (m1 == 0) {
        build_pass(c, e1, flags1);
        scr1 = (int)(int16_t)align(c, false, s1.diffs(), max_diffs1, (int)readlen1, s1.pos(), ao);
        m1 = scr1;
        __syncwarp();
        if (c.lane == 0) c.mem_scr[j1] = (int16_t)scr1;
        __syncwarp();
      }
      const int pair_scr = (int)(int16_t)(scr2 + m1);
      if (swap_ends ? best.update(pair_scr, s2, s1) : best.update(pair_scr, s1, s2)) {
        best_scr1 = scr1;  // stale on a memo hit, as in the reference
        best_scr2 = scr2;
        best_pos1 = s1.pos();
        best_pos2 = s2.pos();
      }
    }
  }
  if (best_pos1 != 0) {
    Hit s1 = swap_ends ? best.r2 : best.r1;
    Hit s2 = swap_ends ? best.r1 : best.r2;
    uint32_t len1 = 0, len2 = 0;
    build_pass(c, e1, flags1);
    ao.score = 0;
    align(c, true, s1.diffs(), max_diffs1, (int)readlen1, best_pos1, ao);
    int nm = build_cigar(c, s1.diffs(), ao, (int)readlen1, best_scr1, cg1, len1, best_pos1);
    s1.set_pos(best_pos1);
    s1.set_diffs(nm);
    build_pass(c, e2, flags2);
    ao.score = 0;
    align(c, true, s2.diffs(), max_diffs2, (int)readlen2, best_pos2, ao);
    nm = build_cigar(c, s2.diffs(), ao, (int)readlen2, best_scr2, cg2, len2, best_pos2);
    s2.set_pos(best_pos2);
    s2.set_diffs(nm);
    const uint32_t frag_end = best_pos2 + len2;
    if (frag_end >= best_pos1 + min_dist && frag_end <= best_pos1 + max_dist) {
      best.r1 = swap_ends ? s2 : s1;
      best.r2 = swap_ends ? s1 : s2;
    }
    else best.reset();
  }
}

// best_single (abismal.cpp:1715-1720)
__device__ void best_single(const CandSet &pres, CandSet &res) {
  for (int i = 0; i != pres.sz && !res.sure_ambig; ++i) {
    const Hit h = pres.v.get(i);
    res.update(false, h.diffs(), h.flags(), h.pos());
  }
}

__device__ __forceinline__ abg_hit to_abg(Hit h) {
  abg_hit r;
  r.diffs = (int16_t)h.diffs();
  r.flags = (uint16_t)h.flags();
  r.pos = h.pos();
  return r;
}

__device__ __forceinline__ size_t smem_per_warp(uint32_t ml, bool paired) {
  size_t b = 0;
  b += 2 * (size_t)ml;                                        // base[2]
  b += (size_t)ml + 32;                                       // qcode
  b += (size_t)ml / 2;                                        // packed
  b += ((size_t)(ml + 64) / 16 + 2) * 8;                      // refw
  b += (size_t)2 * kSeSlots * 8;                              // two SE sets
  if (paired) b += (size_t)2 * kPeSmemSlots * 8;              // two PE heaps (head)
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(kThreadsPerBlock) map_reads_kernel(const KernelParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool paired = P.mode & ABG_MODE_PAIRED;
  const bool a_rich = P.mode & ABG_MODE_A_RICH;
  const bool rpbat = P.mode & ABG_MODE_RANDOM_PBAT;
  const size_t per_warp = smem_per_warp(P.ml, paired);
  unsigned char *sp = smem_raw + per_warp * warp;
  const size_t slot = (size_t)blockIdx.x * kWarpsPerBlock + warp;

  WarpCtx c;
  c.P = &P;
  c.lane = lane;
  // 8-byte aligned arrays first
  c.packed = reinterpret_cast<uint64_t *>(sp);
  sp += (size_t)P.ml / 2;
  c.refw = reinterpret_cast<uint64_t *>(sp);
  sp += ((size_t)(P.ml + 64) / 16 + 2) * 8;
  uint64_t *se_sm0 = reinterpret_cast<uint64_t *>(sp);
  sp += (size_t)kSeSlots * 8;
  uint64_t *se_sm1 = reinterpret_cast<uint64_t *>(sp);
  sp += (size_t)kSeSlots * 8;
  uint64_t *pe_sm0 = nullptr, *pe_sm1 = nullptr;
  if (paired) {
    pe_sm0 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)kPeSmemSlots * 8;
    pe_sm1 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)kPeSmemSlots * 8;
  }
  c.base[0] = sp;
  sp += P.ml;
  c.base[1] = sp;
  sp += P.ml;
  c.qcode = sp;
  c.tb = P.tb + slot * (size_t)P.tb_rows * 4;
  c.mem_scr = P.mem_scr ? P.mem_scr + slot * (size_t)kPeLarge : nullptr;
  c.c_lookup = c.c_entry = c.c_word = c.c_align = c.c_dpref = 0;

  CandSet se0, se1, pe0, pe1;
  se0.v = HeapRef{se_sm0, nullptr, kSeSlots};
  se1.v = HeapRef{se_sm1, nullptr, kSeSlots};
  if (paired) {
    uint64_t *ov = P.pe_overflow + slot * (size_t)2 * kPeLarge;
    pe0.v = HeapRef{pe_sm0, ov, kPeSmemSlots};
    pe1.v = HeapRef{pe_sm1, ov + kPeLarge, kPeSmemSlots};
  }

  const uint32_t T = 0, A = ABG_FLAG_A_RICH, RC = ABG_FLAG_RC;

  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= P.n) break;
    c.cur_key = ~0u;

    if (!paired) {
      // map_single_ended<conv> / map_single_ended_rand (abismal.cpp:1511-1704)
      const uint32_t o0 = P.off[0][item], len = P.off[0][item + 1] - o0;
      CigarOut cg{P.cigar[0] + (size_t)item * P.cigar_stride, P.cigar_stride, 0u, 0u};
      Hit best(kMaxDiffs, 0, 0);
      if (len != 0) {
        load_end(c, 0, P.seq[0] + o0, len);
        se0.reset_se(len);
        uint32_t passes[4];
        int n_pass;
        if (rpbat) {
          passes[0] = T; passes[1] = A; passes[2] = A | RC; passes[3] = T | RC;
          n_pass = 4;
        }
        else {
          const uint32_t cv = a_rich ? A : T;
          passes[0] = cv; passes[1] = cv | RC;
          n_pass = 2;
        }
        for (int p = 0; p < n_pass; ++p) {
          build_pass(c, 0, passes[p]);
          process_seeds(c, passes[p], len, se0);
        }
        align_se_candidates(c, 0, len, P.valid_frac, se0, best, cg);
      }
      if (lane == 0) {
        P.se[0][item] = to_abg(best);
        P.n_cigar[0][item] = cg.n;
        if (cg.n > cg.stride) atomicExch(P.error_flag, 1u);
      }
    }
    else {
      // map_paired_ended<conv> / map_paired_ended_rand (abismal.cpp:1887-2185)
      uint32_t len[2];
      for (int e = 0; e < 2; ++e) {
        const uint32_t o = P.off[e][item];
        len[e] = P.off[e][item + 1] - o;
        c.len[e] = len[e];
        if (len[e] != 0) load_end(c, e, P.seq[e] + o, len[e]);
      }
      CigarOut cg[2] = {{P.cigar[0] + (size_t)item * P.cigar_stride, P.cigar_stride, 0u, 0u},
                        {P.cigar[1] + (size_t)item * P.cigar_stride, P.cigar_stride, 0u, 0u}};
      CandSet *res_se[2] = {&se0, &se1};
      se0.reset_se(len[0]);
      se1.reset_se(len[1]);
      PeBest best;
      best.reset(len[0], len[1]);
      Hit best_se[2] = {Hit(invalid_hit_diffs(len[0]), 0, 0), Hit(invalid_hit_diffs(len[1]), 0, 0)};
      bool any_success = false;
      const int n_calls = rpbat ? 4 : 2;
      for (int call = 0; call < n_calls; ++call) {
        // map_fragments instantiations, SURVEY appendix C
        const bool first_is_r1 = (call & 1) == 0;
        const bool swap_ends = !first_is_r1;
        bool enc_a_call;  // encoding shared by both ends of this call
        if (rpbat) enc_a_call = (call == 1 || call == 2);
        else enc_a_call = a_rich ? (call == 0) : (call == 1);
        const uint32_t f1 = enc_a_call ? A : T;         // un-reversed read: a_rich bit == encoding
        const uint32_t f2 = (enc_a_call ? T : A) | RC;  // reversed read: a_rich bit == !encoding
        const int e1 = first_is_r1 ? 0 : 1, e2 = 1 - e1;
        pe0.reset_pe(len[e1]);
        pe1.reset_pe(len[e2]);
        if (len[e1] == 0 && len[e2] == 0) continue;
        any_success = true;
        if (len[e1] != 0) {
          build_pass(c, e1, f1);
          process_seeds(c, f1, len[e1], pe0);
        }
        if (len[e2] != 0) {
          build_pass(c, e2, f2);
          process_seeds(c, f2, len[e2], pe1);
        }
        // select_maps (abismal.cpp:1833-1847)
        if (pe0.should_align() && pe1.should_align()) {
          sort_unique(pe0.v, pe0.sz, lane);
          sort_unique(pe1.v, pe1.sz, lane);
          best_pair(c, swap_ends, e1, f1, e2, f2, pe0, pe1, cg[e1], cg[e2], best);
        }
        best_single(pe0, *res_se[e1]);
        best_single(pe1, *res_se[e2]);
      }
      if (!any_success) {
        best.reset();
        se0.reset_se_noarg();
        se1.reset_se_noarg();
      }
      {  // valid_pair (abismal.cpp:624-631)
        const uint32_t al1 = cg[0].ref_len, al2 = cg[1].ref_len;
        const bool ok = valid_len(al1, len[0]) && valid_len(al2, len[1]) &&
                        best.diffs() <= frac_of(P.valid_frac, al1 + al2);
        if (!ok) best.reset();
      }
      if (!best.should_report(P.allow_ambig != 0u)) {
        const double half = P.valid_frac / 2.0;
        align_se_candidates(c, 0, len[0], half, se0, best_se[0], cg[0]);
        align_se_candidates(c, 1, len[1], half, se1, best_se[1], cg[1]);
      }
      if (lane == 0) {
        P.pe_r1[item] = to_abg(best.r1);
        P.pe_r2[item] = to_abg(best.r2);
        P.se[0][item] = to_abg(best_se[0]);
        P.se[1][item] = to_abg(best_se[1]);
        P.n_cigar[0][item] = cg[0].n;
        P.n_cigar[1][item] = cg[1].n;
        if (cg[0].n > cg[0].stride || cg[1].n > cg[1].stride) atomicExch(P.error_flag, 1u);
      }
    }
    __syncwarp();
  }

  if (P.counters != nullptr) {
    unsigned long long v[5] = {c.c_lookup, c.c_entry, c.c_word, c.c_align, c.c_dpref};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned long long x = v[k];
      for (int d = 16; d >= 1; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
      v[k] = x;
    }
    if (lane == 0) {
      atomicAdd(P.counters + 0, v[0]);  // n_lookup
      atomicAdd(P.counters + 1, v[1]);  // n_entry
      atomicAdd(P.counters + 2, v[1]);  // n_cmp == n_entry
      atomicAdd(P.counters + 3, v[2]);  // n_word
      atomicAdd(P.counters + 4, v[3]);  // n_align
      atomicAdd(P.counters + 5, v[4]);  // n_dpref
    }
  }
}

}  // namespace ab2dev
