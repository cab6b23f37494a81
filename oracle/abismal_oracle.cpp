/* abismal_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Sequential CPU restatement of the read-mapping hot path of
 * smithlabcode/abismal v3.3.0 (reference paths below are relative to
 * /root/reference).  It is written from the reference's behaviour, function
 * by function, with the reference location cited at each step; it is used as
 * the checker in tests/ and as the "port" CPU baseline in bench.py.
 *
 * Parity status: PINNED (see abismal_oracle.h).  tests/test_oracle_vs_ref.py
 * compares the SAM produced through this restatement with the SAM of the
 * unmodified reference binary (oracle/_ref/abismal), whose own outputs match
 * the reference's golden md5s (data/md5sum.txt) 16/16.
 *
 * Deliberate deviations (unobservable or undefined in the reference):
 *  - reads of 44..47 bases (36..47 with window 12) make the reference read past
 *    the end of the encoded read while rolling its hash (src/abismal.cpp:1308,
 *    1333) and while extending over-full buckets (find_candidates, :1163-1259);
 *    here the bytes past the end are defined to be 0, however far.
 *  - the heap sentinel's stale `flags` field (se_element::reset keeps it,
 *    src/abismal.cpp:286-290) is set to 0; sentinels have pos == 0 and are
 *    skipped wherever flags would be looked at.
 */
#include "abismal_oracle.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

using score_t = int16_t;
using flags_t = uint16_t;

thread_local std::string g_err;

struct Hit {  // se_element, src/abismal.cpp:224-233
  score_t diffs;
  flags_t flags;
  uint32_t pos;
};
static_assert(sizeof(Hit) == sizeof(abg_hit), "layout");

constexpr score_t kMaxDiffs = 32767;          // se_element::MAX_DIFFS :229
constexpr double kInvalidHitFrac = 0.4;       // :228
constexpr uint32_t kKeyWeight = 25;           // AbismalIndex.hpp:68
constexpr uint32_t kKeyWeightThree = 16;      // :69
// seed::window_size, AbismalIndex.hpp:73-77: 20, or 12 when the reference is configured with --enable-short.
// Taken from the index (abg_index_view::window_size) at the start of every abo_map_batch call.
thread_local uint32_t kWindow = 20;
constexpr uint32_t kHashMask = (1u << 25) - 1;    // :82
constexpr uint32_t kHashMaskThree = 43046721u;    // 3^16, :88
thread_local uint32_t kMinReadLen = kKeyWeight + 20 - 1;  // abismal.cpp:212-213
constexpr uint32_t kSeMax = 50;               // se_candidates::max_size :448
constexpr uint32_t kPeSmall = 32;             // pe_candidates::max_size_small :861
constexpr uint32_t kPeLarge = 32u << 10;      // :862
constexpr size_t kMaxOffDiag = 30;            // AbismalAlign.hpp:133
constexpr int kMatch = 2, kMismatch = -3, kIndel = -4;  // AbismalAlign.hpp:51-53
constexpr int kOpM = 0, kOpI = 1, kOpD = 2, kOpS = 4;   // abismal_cigar_utils.hpp

inline bool hit_empty(const Hit &h) { return h.pos == 0; }
inline bool hit_ambig(const Hit &h) { return h.flags & ABG_FLAG_AMBIG; }
inline bool hit_rc(const Hit &h) { return h.flags & ABG_FLAG_RC; }
inline bool hit_a_rich(const Hit &h) { return h.flags & ABG_FLAG_A_RICH; }
inline void hit_reset(Hit &h) {  // se_element::reset() :286-290 (flags kept)
  h.pos = 0;
  h.diffs = kMaxDiffs;
}
inline void hit_reset(Hit &h, uint32_t readlen) {  // :292-296
  hit_reset(h);
  h.diffs = static_cast<score_t>(kInvalidHitFrac * readlen);
}

/* ---- libstdc++ heap primitives, comparator = diffs (SURVEY appendix D;
 *      bits/stl_heap.h __push_heap / __adjust_heap as of GCC 13) ---------- */
void sift_up(Hit *v, long hole, long top, Hit val) {
  long parent = (hole - 1) / 2;
  while (hole > top && v[parent].diffs < val.diffs) {
    v[hole] = v[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[hole] = val;
}
void heap_push(Hit *v, long n) {  // std::push_heap(v, v + n)
  sift_up(v, n - 1, 0, v[n - 1]);
}
void heap_pop(Hit *v, long n) {  // std::pop_heap(v, v + n)
  if (n <= 1) return;
  const Hit val = v[n - 1];
  v[n - 1] = v[0];
  const long len = n - 1;
  long hole = 0, child = 0;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (v[child].diffs < v[child - 1].diffs) --child;
    v[hole] = v[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[hole] = v[child - 1];
    hole = child - 1;
  }
  sift_up(v, hole, 0, val);
}

/* ---- se_candidates, src/abismal.cpp:334-449 ------------------------------ */
struct SeSet {
  bool sure_ambig = false;
  score_t good_cutoff = 0, cutoff = 0;
  uint32_t sz = 1;
  Hit best{kMaxDiffs, 0, 0};
  Hit v[kSeMax];

  SeSet() {
    for (auto &h : v) h = Hit{kMaxDiffs, 0, 0};
  }
  bool full() const { return sz == kSeMax; }
  bool has_exact_match() const { return !hit_empty(best); }
  bool good_diff(score_t d) const { return d <= good_cutoff; }
  bool should_do_sensitive() const { return !full() || !good_diff(cutoff); }  // :367-370
  void set_specific() { cutoff = good_cutoff; }                               // :372-375
  void set_sensitive() { cutoff = v[0].diffs; }                               // :377-380
  void update(bool specific, score_t d, flags_t s, uint32_t p) {              // :394-404
    if (d == 0) {  // update_exact_match :347-355
      const Hit cand{0, s, p};
      if (hit_empty(best)) best = cand;
      else if (cand.pos != best.pos || cand.flags != best.flags) best.flags |= ABG_FLAG_AMBIG;
    }
    else {  // update_cand :382-392
      if (full()) {
        heap_pop(v, sz);
        v[sz - 1] = Hit{d, s, p};
      }
      else v[sz++] = Hit{d, s, p};
      heap_push(v, sz);
    }
    sure_ambig = hit_ambig(best) && best.diffs == 0;
    cutoff = specific ? std::min(cutoff, v[0].diffs) : v[0].diffs;
  }
  void reset() {  // :406-415
    hit_reset(best);
    hit_reset(v[0]);
    v[0].flags = 0;
    cutoff = v[0].diffs;
    sure_ambig = false;
    sz = 1;
  }
  void reset(uint32_t readlen) {  // :417-427
    hit_reset(best, readlen);
    hit_reset(v[0], readlen);
    v[0].flags = 0;
    cutoff = v[0].diffs;
    good_cutoff = static_cast<score_t>(readlen / 10u);
    sure_ambig = false;
    sz = 1;
  }
  void prepare_for_alignments() {  // :430-439
    std::sort(v, v + sz, [](const Hit &a, const Hit &b) {
      return a.pos < b.pos || (a.pos == b.pos && a.flags < b.flags);
    });
    sz = static_cast<uint32_t>(
      std::unique(v, v + sz,
                  [](const Hit &a, const Hit &b) { return a.pos == b.pos && a.flags == b.flags; }) -
      v);
  }
};

/* ---- pe_candidates, src/abismal.cpp:775-863 ------------------------------ */
struct PeSet {
  bool sure_ambig = false;
  score_t cutoff = 0, good_cutoff = 0;
  uint32_t sz = 1, capacity = kPeSmall;
  std::vector<Hit> v;
  PeSet() : v(kPeLarge, Hit{kMaxDiffs, 0, 0}) {}
  void reset(uint32_t readlen) {  // :778-787
    hit_reset(v[0], readlen);
    v[0].flags = 0;
    sure_ambig = false;
    cutoff = v[0].diffs;
    good_cutoff = static_cast<score_t>(readlen / 10);
    sz = 1;
    capacity = kPeSmall;
  }
  void set_specific() { cutoff = good_cutoff; }
  void set_sensitive() { cutoff = v[0].diffs; }
  bool should_align() const { return sz != kPeLarge || cutoff != 0; }  // :799-802
  bool full() const { return sz == capacity; }
  bool good_diff(score_t d) const { return d <= good_cutoff; }
  bool should_do_sensitive() const { return capacity == kPeSmall || !good_diff(cutoff); }  // :819-822
  void update(bool specific, score_t d, flags_t s, uint32_t p) {  // :824-842
    if (full()) {
      if (specific && capacity != kPeLarge && good_diff(d)) ++capacity;
      else {
        heap_pop(v.data(), sz);
        --sz;
      }
    }
    v[sz++] = Hit{d, s, p};
    heap_push(v.data(), sz);
    cutoff = specific ? std::min(cutoff, v[0].diffs) : v[0].diffs;
    sure_ambig = full() && cutoff == 0;
  }
  void prepare_for_mating() {  // :844-852
    std::sort(v.begin(), v.begin() + sz, [](const Hit &a, const Hit &b) { return a.pos < b.pos; });
    sz = static_cast<uint32_t>(
      std::unique(v.begin(), v.begin() + sz,
                  [](const Hit &a, const Hit &b) { return a.pos == b.pos && a.flags == b.flags; }) -
      v.begin());
  }
};

/* ---- pe_element, src/abismal.cpp:547-622 --------------------------------- */
struct PeBest {
  score_t aln_score = 0, max_aln_score = 0;
  Hit r1{kMaxDiffs, 0, 0}, r2{kMaxDiffs, 0, 0};
  score_t diffs() const { return static_cast<score_t>(r1.diffs + r2.diffs); }  // :550-553 (wraps)
  void reset(uint32_t l1, uint32_t l2) {                                       // :555-561
    aln_score = 0;
    hit_reset(r1, l1);
    hit_reset(r2, l2);
    r1.flags = r2.flags = 0;
    max_aln_score = static_cast<score_t>(static_cast<score_t>(kMatch * l1) +
                                         static_cast<score_t>(kMatch * l2));
  }
  void reset() {  // :563-568
    aln_score = 0;
    hit_reset(r1);
    hit_reset(r2);
  }
  bool update(score_t scr, const Hit &s1, const Hit &s2) {  // :570-587
    const int rd = r1.diffs + r2.diffs;
    const int sd = s1.diffs + s2.diffs;
    if (scr > aln_score || (scr == aln_score && sd < rd)) {
      r1 = s1;
      r2 = s2;
      aln_score = scr;
      return true;
    }
    if (scr == aln_score && sd == rd) r1.flags |= ABG_FLAG_AMBIG;
    return false;
  }
  bool ambig() const { return hit_ambig(r1); }
  bool empty() const { return hit_empty(r1); }
  bool should_report(bool allow_ambig) const { return !empty() && (allow_ambig || !ambig()); }
  bool sure_ambig() const { return ambig() && aln_score == max_aln_score; }
};

/* ---- index view ---------------------------------------------------------- */
}  // namespace

struct abo_index {
  abg_index_view v;
};

namespace {

inline uint8_t genome_base(const uint64_t *g, uint64_t pos) {  // genome_four_bit_itr::operator* :198-201
  return static_cast<uint8_t>((g[pos >> 4] >> ((pos & 15u) << 2)) & 15u);
}
inline uint32_t get_bit(uint8_t nt) { return (nt & 5) == 0; }  // AbismalIndex.hpp:255-258
inline uint32_t three_num(bool g_to_a, uint8_t nt) {           // :260-269
  return g_to_a ? ((((nt & 8) != 0) << 1) | ((nt & 2) != 0)) : ((((nt & 4) != 0) << 1) | ((nt & 1) != 0));
}
inline uint32_t three_fast(bool g_to_a, uint8_t nt) {  // abismal.cpp:1196-1203
  return g_to_a ? (nt & 10) : (nt & 5);
}

/* encode_base_t_rich / encode_base_a_rich, dna_four_bit_bisulfite.hpp:32-57 */
inline uint8_t encode_base(bool a_rich_enc, char c) {
  switch (c) {
    case 'A': case 'a': return a_rich_enc ? 5 : 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 4;
    case 'T': case 't': return a_rich_enc ? 8 : 10;
    default: return 0;
  }
}
inline char comp_base(char c) {  // revcomp_inplace, common.hpp:28-36
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
  }
}

/* A read in one orientation and one bisulfite encoding (prep_read :1377-1386
 * + pack_read :1393-1426).  `code` is zero-padded past the end. */
struct EncRead {
  uint32_t len = 0;
  std::vector<uint8_t> code;
  std::vector<uint64_t> packed;
  void build(const char *s, uint32_t n, bool rc, bool a_rich_enc) {
    len = n;
    code.assign(n + 64, 0);
    for (uint32_t i = 0; i < n; ++i) {
      const char c = rc ? comp_base(s[n - 1 - i]) : s[i];
      code[i] = encode_base(a_rich_enc, c);
    }
    const uint32_t w = (n + 15) / 16;
    packed.assign(w, 0);
    for (uint32_t i = 0; i < 16 * w; ++i) {
      const uint64_t nib = i < n ? code[i] : 0xFull;  // tail matches anything :1424-1425
      packed[i >> 4] |= nib << ((i & 15) << 2);
    }
  }
};

/* full_compare, src/abismal.cpp:1105-1122 (with its early exit) */
inline score_t full_compare(score_t cutoff, const uint64_t *rd, uint32_t n_words, uint32_t offset,
                            const uint64_t *g, uint64_t *words_seen) {
  score_t d = 0;
  uint32_t w = 0;
  for (; d <= cutoff && w != n_words; ++w) {
    const uint64_t gw = (g[w] >> offset) | ((g[w + 1] << (63 - offset)) << 1);
    d = static_cast<score_t>(d + 16 - __builtin_popcountll(rd[w] & gw));
  }
  *words_seen += w;
  return d;
}

/* check_hits, src/abismal.cpp:1124-1150 */
template <class RS>
void check_hits(const abg_index_view &ix, flags_t strand_code, uint32_t offset, const EncRead &r,
                const uint32_t *idx, uint32_t start, uint32_t end, RS &res, abg_work_counters *wc) {
  for (; start != end && !res.sure_ambig; ++start) {
    const uint32_t the_pos = idx[start] - offset;
    uint64_t words = 0;
    const score_t d = full_compare(res.cutoff, r.packed.data(), static_cast<uint32_t>(r.packed.size()),
                                   (the_pos & 15u) << 2, ix.genome + (the_pos >> 4), &words);
    if (wc) {
      wc->n_entry += 1;
      wc->n_cmp += 1;
      wc->n_word += words;
    }
    if (d <= res.cutoff) res.update(true, d, strand_code, the_pos);
  }
}

/* std::lower_bound over idx[low, high) with predicate pred(entry) == "less
 * than value" (bits/stl_algobase.h __lower_bound) */
template <class Pred>
uint32_t lower_bound_idx(const uint32_t *idx, uint32_t low, uint32_t high, Pred less_than) {
  long len = static_cast<long>(high) - static_cast<long>(low);
  uint32_t first = low;
  while (len > 0) {
    const long half = len >> 1;
    const uint32_t mid = first + static_cast<uint32_t>(half);
    if (less_than(idx[mid])) {
      first = mid + 1;
      len = len - half - 1;
    }
    else len = half;
  }
  return first;
}

/* bound of the seed extension for reads that never meet p == read_lim (stays inside the end padding) */
constexpr uint32_t kMaxExtend = 32000u;

/* find_candidates<25>, src/abismal.cpp:1163-1194 */
uint32_t find_candidates(const abg_index_view &ix, uint32_t max_candidates, const uint8_t *read_start,
                         uint32_t read_lim, const uint32_t *idx, uint32_t &low, uint32_t &high) {
  uint32_t p = kKeyWeight;
  uint32_t prev_low = low, prev_high = high;
  for (; p != read_lim && p < kMaxExtend && (high - low) > max_candidates; ++p) {
    prev_low = low;
    prev_high = high;
    const uint32_t first_1 = lower_bound_idx(idx, low, high, [&](uint32_t e) {
      return get_bit(genome_base(ix.genome, static_cast<uint64_t>(e) + p)) < 1u;
    });
    const uint32_t the_bit = get_bit(p < read_lim ? read_start[p] : 0);  // past the end: 0 (see the header)
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

/* find_candidates_three<16, conv>, src/abismal.cpp:1214-1259 */
uint32_t find_candidates_three(const abg_index_view &ix, bool g_to_a, uint32_t max_candidates,
                               const uint8_t *read_start, uint32_t max_size, const uint32_t *idx,
                               uint32_t &low, uint32_t &high) {
  uint32_t p = kKeyWeightThree;
  uint32_t prev_low = low, prev_high = high;
  const uint32_t v1 = g_to_a ? 2 : 1, v2 = g_to_a ? 8 : 4;
  for (; p != max_size && p < kMaxExtend && (high - low) > max_candidates; ++p) {
    prev_low = low;
    prev_high = high;
    const auto less = [&](uint32_t val) {
      return [&, val](uint32_t e) {
        return three_fast(g_to_a, genome_base(ix.genome, static_cast<uint64_t>(e) + p)) < val;
      };
    };
    const uint32_t first_1 = lower_bound_idx(idx, low, high, less(v1));
    const uint32_t first_2 = lower_bound_idx(idx, low, high, less(v2));
    const uint32_t the_num = three_fast(g_to_a, p < max_size ? read_start[p] : 0);
    const uint32_t mid_val = g_to_a ? 2u : 1u;
    const uint32_t old_low = low, old_high = high;
    high = (the_num == 0) ? first_1 : ((the_num == mid_val) ? first_2 : old_high);
    low = (the_num == 0) ? old_low : ((the_num == mid_val) ? first_1 : first_2);
  }
  if (low == high) {
    --p;
    low = prev_low;
    high = prev_high;
  }
  return p;
}

/* process_seeds, src/abismal.cpp:1269-1375.  strand_code carries the rc and
 * a-rich bits (get_strand_code :130-134); the three-letter table follows
 * get_conv_type (:1261-1267). */
template <class RS>
void process_seeds(const abg_index_view &ix, uint32_t max_candidates, flags_t strand_code,
                   const EncRead &r, RS &res, abg_work_counters *wc) {
  const bool g_to_a = (((strand_code & ABG_FLAG_A_RICH) != 0) != ((strand_code & ABG_FLAG_RC) != 0));
  const uint32_t *counter3 = g_to_a ? ix.counter_a : ix.counter_t;
  const uint32_t *index3 = g_to_a ? ix.index_a : ix.index_t;
  const uint32_t readlen = r.len;
  const uint8_t *code = r.code.data();

  const auto hash_two = [&](uint32_t &k) {  // get_1bit_hash, AbismalIndex.hpp:285-294
    k = 0;
    for (uint32_t j = 0; j < kKeyWeight; ++j) k = (k << 1) | get_bit(code[j]);
  };
  const auto hash_three = [&](uint32_t &k) {  // get_base_3_hash :296-305
    k = 0;
    for (uint32_t j = 0; j < kKeyWeightThree; ++j) k = (k * 3 + three_num(g_to_a, code[j])) % kHashMaskThree;
  };

  uint32_t k = 0, k3 = 0;
  hash_two(k);
  hash_three(k3);

  const uint32_t specific_len = std::min(readlen - kWindow, readlen >> 1);
  const uint32_t specific_lim = std::max(kWindow, readlen >> 1);

  res.set_specific();
  for (uint32_t i = 0; i < specific_lim && !res.sure_ambig; ++i) {
    uint32_t s = ix.counter[k], e = ix.counter[k + 1];
    const uint32_t l_two = find_candidates(ix, max_candidates, code + i, readlen - i, ix.index, s, e);
    const uint32_t d_two = e - s;
    uint32_t s3 = counter3[k3], e3 = counter3[k3 + 1];
    const uint32_t l_three =
      find_candidates_three(ix, g_to_a, max_candidates, code + i, readlen - i, index3, s3, e3);
    const uint32_t d_three = e3 - s3;
    if (wc) wc->n_lookup += 2;
    if (d_two <= max_candidates || l_two >= specific_len)
      check_hits(ix, strand_code, i, r, ix.index, s, e, res, wc);
    if (d_three <= max_candidates || l_three >= specific_len)
      check_hits(ix, strand_code, i, r, index3, s3, e3, res, wc);
    k = ((k << 1) | get_bit(code[i + kKeyWeight])) & kHashMask;                      // shift_hash_key :271-274
    k3 = (k3 * 3 + three_num(g_to_a, code[i + kKeyWeightThree])) % kHashMaskThree;   // shift_three_key :276-281
  }

  if (!res.should_do_sensitive()) return;

  hash_two(k);
  hash_three(k3);
  res.set_sensitive();

  const uint32_t lim_two = readlen - kKeyWeight + 1;
  constexpr uint32_t kMinFoldSize = 10;
  for (uint32_t i = 0; i < lim_two && !res.sure_ambig; ++i) {
    const uint32_t s = ix.counter[k], e = ix.counter[k + 1];
    const uint32_t d_two = e - s;
    const uint32_t s3 = counter3[k3], e3 = counter3[k3 + 1];
    const uint32_t d_three = e3 - s3;
    if (wc) wc->n_lookup += 2;
    if (d_two != 0 && d_two <= max_candidates && (d_three == 0 || d_two <= kMinFoldSize * d_three))
      check_hits(ix, strand_code, i, r, ix.index, s, e, res, wc);
    if (d_three != 0 && d_three <= max_candidates)
      check_hits(ix, strand_code, i, r, index3, s3, e3, res, wc);
    k = ((k << 1) | get_bit(code[i + kKeyWeight])) & kHashMask;
    k3 = (k3 * 3 + three_num(g_to_a, code[i + kKeyWeightThree])) % kHashMaskThree;
  }
}

/* ---- AbismalAlign, src/AbismalAlign.hpp ---------------------------------- */
struct Aligner {
  const uint64_t *genome;
  std::vector<score_t> table;
  std::vector<int8_t> tb;
  uint32_t q_sz = 0;
  abg_work_counters *wc = nullptr;

  static size_t bandwidth(score_t diffs, score_t max_diffs) {  // :333-334
    const size_t full = 2 * kMaxOffDiag + 1;
    const size_t want = static_cast<size_t>(2 * std::min(diffs, max_diffs) + 1);
    return std::min(full, want);
  }

  /* align<do_traceback>, :320-386 */
  score_t align(bool do_tb, score_t diffs, score_t max_diffs, const uint8_t *q, uint32_t qlen,
                uint32_t t_pos) {
    q_sz = qlen;
    if (diffs == 0) return static_cast<score_t>(kMatch * q_sz);
    const size_t bw = bandwidth(diffs, max_diffs);
    const size_t t_shift = q_sz + bw;
    const size_t n_cells = t_shift * bw;
    if (table.size() < n_cells) table.resize(n_cells);
    if (tb.size() < n_cells) tb.resize(n_cells);
    std::fill_n(table.begin(), n_cells, 0);
    if (do_tb) std::fill_n(tb.begin(), n_cells, -1);
    const size_t t_beg = t_pos - ((bw - 1) / 2);
    if (wc) {
      wc->n_align += 1;
      wc->n_dpref += t_shift;
    }
    for (size_t i = 1; i < t_shift; ++i) {
      const size_t left = i < bw ? bw - i : 0;
      const size_t right = std::min(bw, t_shift - i);
      score_t *prev = table.data() + (i - 1) * bw;
      score_t *cur = prev + bw;
      int8_t *tcur = tb.data() + i * bw;
      const size_t qoff = i > bw ? i - bw : 0;
      const uint8_t ref = genome_base(genome, t_beg + i - 1);
      for (size_t j = left; j < right; ++j) {  // from_diag :233-243 / :266-281
        const score_t s =
          static_cast<score_t>(((q[qoff + (j - left)] & ref) == 0 ? kMismatch : kMatch) + prev[j]);
        if (s > cur[j]) cur[j] = s;
        if (do_tb && cur[j] == s) tcur[j] = kOpM;
      }
      for (size_t j = left; j + 1 < right; ++j) {  // from_above :245-252 / :283-294
        const score_t s = static_cast<score_t>(prev[j + 1] + kIndel);
        if (s > cur[j]) cur[j] = s;
        if (do_tb && cur[j] == s) tcur[j] = kOpD;
      }
      for (size_t j = left + 1; j < right; ++j) {  // from_left :256-263 / :296-307
        const score_t s = static_cast<score_t>(cur[j - 1] + kIndel);
        if (s > cur[j]) cur[j] = s;
        if (do_tb && cur[j] == s) tcur[j] = kOpI;
      }
    }
    return *std::max_element(table.begin(), table.begin() + n_cells);  // get_best_score :221-226
  }

  /* build_cigar_len_and_pos, :388-440 + get_traceback :166-193 */
  void build_cigar(score_t diffs, score_t max_diffs, std::vector<uint32_t> &cigar, uint32_t &len,
                   uint32_t &t_pos) {
    const size_t bw = bandwidth(diffs, max_diffs);
    const size_t n_cells = (q_sz + bw) * bw;
    if (table.size() < n_cells) table.resize(n_cells);  // reference reads its preallocated table
    const auto best = std::max_element(table.begin(), table.begin() + n_cells);
    const size_t cell = static_cast<size_t>(best - table.begin());
    size_t row = cell / bw, col = cell % bw;
    const score_t r = *best;
    if (r == 0 || diffs == 0) {
      cigar.assign(1, q_sz << 4);
      len = q_sz;
      return;
    }
    const size_t clip_bottom = (q_sz + (bw - 1)) - (row + col);
    cigar.clear();
    int8_t prev_arrow = tb[row * bw + col];
    const auto step = [&](int8_t a) {
      const bool is_del = a == kOpD, is_ins = a == kOpI;
      row -= !is_ins;
      col -= is_ins;
      col += is_del;
    };
    step(prev_arrow);
    uint32_t n = 1;
    while (table[row * bw + col] > 0) {
      const int8_t arrow = tb[row * bw + col];
      step(arrow);
      if (arrow != prev_arrow) {
        cigar.push_back((n << 4) | static_cast<uint32_t>(prev_arrow));
        n = 0;
      }
      ++n;
      prev_arrow = arrow;
    }
    cigar.push_back((n << 4) | static_cast<uint32_t>(prev_arrow));
    const size_t clip_top = (row + col) - (bw - 1);
    if (clip_top > 0) cigar.push_back((static_cast<uint32_t>(clip_top) << 4) | kOpS);
    std::reverse(cigar.begin(), cigar.end());
    if (clip_bottom > 0) cigar.push_back((static_cast<uint32_t>(clip_bottom) << 4) | kOpS);
    len = static_cast<uint32_t>(q_sz - clip_bottom - clip_top);
    const size_t t_beg = t_pos - ((bw - 1) / 2);
    t_pos = static_cast<uint32_t>(t_beg + row);
  }
};

/* simple_aln::edit_distance, AbismalAlign.hpp:73-89 (same integer promotions:
 * the quotient is computed in unsigned arithmetic) */
score_t edit_distance(score_t scr, uint32_t len, const std::vector<uint32_t> &cigar) {
  if (scr == 0) return static_cast<score_t>(len);
  int ins_i = 0, del_i = 0;
  for (uint32_t c : cigar) {
    const uint8_t op = c & 0xf;
    const uint8_t l = static_cast<uint8_t>(c >> 4);  // abismal_bam_cigar_oplen returns uint8_t
    if (op == kOpI) ins_i += l;
    if (op == kOpD) del_i += l;
  }
  const score_t ins = static_cast<score_t>(ins_i), del = static_cast<score_t>(del_i);
  const score_t A = static_cast<score_t>(scr - kIndel * (ins + del));
  const uint32_t num = static_cast<uint32_t>(kMatch) * (len - static_cast<uint32_t>(static_cast<int32_t>(ins))) -
                       static_cast<uint32_t>(static_cast<int32_t>(A));
  const score_t mism = static_cast<score_t>(num / static_cast<uint32_t>(kMatch - kMismatch));
  return static_cast<score_t>(mism + ins + del);
}

uint32_t cigar_rseq_ops(const std::vector<uint32_t> &cig) {  // abismal.cpp:451-462
  uint32_t t = 0;
  for (uint32_t c : cig) {
    const uint32_t op = c & 0xf;
    if ((0x3C1A7u >> (op << 1)) & 2u) t += c >> 4;
  }
  return t;
}

inline score_t valid_diffs_cutoff(uint32_t readlen, double cutoff) {  // :301-305
  return static_cast<score_t>(cutoff * readlen);
}
inline bool valid_len(uint32_t aln_len, uint32_t readlen) {  // :307-314
  static const double min_aln_frac = 1.0 - kInvalidHitFrac;
  return aln_len >= std::max(kMinReadLen, static_cast<uint32_t>(min_aln_frac * readlen));
}
inline bool same_pos(uint32_t a, uint32_t b) {  // :1428-1433
  return (a > b ? a - b : b - a) <= 3;
}

/* The four encodings of one end: index = (rc ? 2 : 0) | (a_rich_enc ? 1 : 0). */
struct EndEncodings {
  EncRead e[4];
  bool built[4] = {false, false, false, false};
  const char *seq = nullptr;
  uint32_t len = 0;
  void set(const char *s, uint32_t n) {
    seq = s;
    len = n;
    for (bool &b : built) b = false;
  }
  /* the read as the pass with `flags` sees it: orientation by the rc bit,
   * encoding by a_rich XOR rc (abismal.cpp:1463-1465 / Appendix C) */
  const EncRead &for_flags(flags_t flags) {
    const bool rc = flags & ABG_FLAG_RC;
    const bool enc_a = ((flags & ABG_FLAG_A_RICH) != 0) != rc;
    const int k = (rc ? 2 : 0) | (enc_a ? 1 : 0);
    if (!built[k]) {
      e[k].build(seq, len, rc, enc_a);
      built[k] = true;
    }
    return e[k];
  }
};

/* align_se_candidates, src/abismal.cpp:1435-1497.  `readlen` is the length of
 * the read buffer the reference passes as pread_t. */
void align_se_candidates(EndEncodings &enc, uint32_t readlen_u, double cutoff, SeSet &res, Hit &best,
                         std::vector<uint32_t> &cigar, Aligner &aln) {
  const score_t readlen = static_cast<score_t>(readlen_u);
  const score_t max_diffs = valid_diffs_cutoff(readlen, cutoff);
  const score_t max_scr = static_cast<score_t>(kMatch * readlen);
  if (res.has_exact_match()) {
    best = res.best;
    cigar.assign(1, static_cast<uint32_t>(readlen) << 4);
    return;
  }
  score_t best_scr = 0;
  uint32_t best_pos = 0;
  res.prepare_for_alignments();
  uint32_t it = 0;
  const uint32_t lim = res.sz;
  for (; it != lim && hit_empty(res.v[it]); ++it) {
  }
  for (; it != lim; ++it) {
    const Hit &h = res.v[it];
    if (h.diffs < static_cast<score_t>(kInvalidHitFrac * readlen)) {  // valid_hit :323-326
      const EncRead &q = enc.for_flags(h.flags);
      const uint32_t cand_pos = h.pos;
      const score_t cand_scr = aln.align(false, h.diffs, max_diffs, q.code.data(), q.len, cand_pos);
      if (cand_scr > best_scr) {
        best = h;
        best_scr = cand_scr;
        best_pos = cand_pos;
      }
      else if (cand_scr == best_scr &&
               (cand_scr == max_scr ? cand_pos != best_pos : !same_pos(cand_pos, best_pos)))
        best.flags |= ABG_FLAG_AMBIG;
    }
  }
  if (best.pos != 0) {
    const EncRead &q = enc.for_flags(best.flags);
    aln.align(true, best.diffs, max_diffs, q.code.data(), q.len, best.pos);
    uint32_t len = 0;
    aln.build_cigar(best.diffs, max_diffs, cigar, len, best.pos);
    best.diffs = edit_distance(best_scr, len, cigar);
    if (!(valid_len(len, readlen) && best.diffs <= valid_diffs_cutoff(readlen, cutoff)))  // check_valid :316-321
      hit_reset(best);
  }
  else hit_reset(best);
}

/* best_single, src/abismal.cpp:1715-1720 */
void best_single(const PeSet &pres, SeSet &res) {
  for (uint32_t i = 0; i != pres.sz && !res.sure_ambig; ++i)
    res.update(false, pres.v[i].diffs, pres.v[i].flags, pres.v[i].pos);
}

/* best_pair<swap_ends>, src/abismal.cpp:1722-1831 */
void best_pair(bool swap_ends, uint32_t min_dist, uint32_t max_dist, double valid_frac, const PeSet &res1,
               const PeSet &res2, const EncRead &pread1, const EncRead &pread2,
               std::vector<uint32_t> &cigar1, std::vector<uint32_t> &cigar2, std::vector<score_t> &mem_scr1,
               Aligner &aln, PeBest &best) {
  const long j1_beg = 0, j1_end = res1.sz, j2_end = res2.sz;
  long j1 = 0, j2 = 0;
  std::fill_n(mem_scr1.begin(), res1.sz, 0);
  const uint32_t readlen1 = pread1.len, readlen2 = pread2.len;
  const score_t max_diffs1 = valid_diffs_cutoff(readlen1, valid_frac);
  const score_t max_diffs2 = valid_diffs_cutoff(readlen2, valid_frac);
  score_t scr1 = 0, best_scr1 = 0, best_scr2 = 0;
  uint32_t best_pos1 = 0, best_pos2 = 0;
  Hit s1{kMaxDiffs, 0, 0}, s2{kMaxDiffs, 0, 0};
  const Hit *v1 = res1.v.data(), *v2 = res2.v.data();

  for (; j1 != j1_end && hit_empty(v1[j1]); ++j1) {
  }
  for (; j2 != j2_end && hit_empty(v2[j2]); ++j2) {
  }
  for (; j2 != j2_end && !best.sure_ambig(); ++j2) {
    s2 = v2[j2];
    score_t scr2 = 0;
    const uint32_t lim = s2.pos + readlen2;
    for (; (j1 == j1_end) || (j1 != j1_beg && v1[j1].pos + max_dist >= lim); --j1) {
    }
    for (; j1 != j1_end && v1[j1].pos + max_dist < lim; ++j1) {
    }
    for (; j1 != j1_end && v1[j1].pos + min_dist <= lim && !best.sure_ambig(); ++j1) {
      s1 = v1[j1];
      if (scr2 == 0) scr2 = aln.align(false, v2[j2].diffs, max_diffs2, pread2.code.data(), pread2.len, s2.pos);
      if (mem_scr1[j1] == 0) {
        scr1 = aln.align(false, v1[j1].diffs, max_diffs1, pread1.code.data(), pread1.len, s1.pos);
        mem_scr1[j1] = scr1;
      }
      const score_t pair_scr = static_cast<score_t>(scr2 + mem_scr1[j1]);
      if (swap_ends ? best.update(pair_scr, s2, s1) : best.update(pair_scr, s1, s2)) {
        best_scr1 = scr1;  // NB: scr1 is stale on a memo hit
        best_scr2 = scr2;
        best_pos1 = v1[j1].pos;
        best_pos2 = v2[j2].pos;
      }
    }
  }
  if (best_pos1 != 0) {
    s1 = swap_ends ? best.r2 : best.r1;
    s2 = swap_ends ? best.r1 : best.r2;
    uint32_t len1 = 0;
    aln.align(true, s1.diffs, max_diffs1, pread1.code.data(), pread1.len, best_pos1);
    aln.build_cigar(s1.diffs, max_diffs1, cigar1, len1, best_pos1);
    s1.pos = best_pos1;
    s1.diffs = edit_distance(best_scr1, len1, cigar1);
    uint32_t len2 = 0;
    aln.align(true, s2.diffs, max_diffs2, pread2.code.data(), pread2.len, best_pos2);
    aln.build_cigar(s2.diffs, max_diffs2, cigar2, len2, best_pos2);
    s2.pos = best_pos2;
    s2.diffs = edit_distance(best_scr2, len2, cigar2);
    const uint32_t frag_end = best_pos2 + len2;
    if (frag_end >= best_pos1 + min_dist && frag_end <= best_pos1 + max_dist) {
      best.r1 = swap_ends ? s2 : s1;
      best.r2 = swap_ends ? s1 : s2;
    }
    else best.reset();
  }
}

struct PairCall {  // one map_fragments instantiation
  bool first_is_r1;
  flags_t flags_first, flags_second;
  bool swap_ends;
};

struct Scratch {
  SeSet se, se2;
  PeSet pe1, pe2;
  std::vector<score_t> mem_scr1;
  Aligner aln;
  EndEncodings enc[2];
  std::vector<uint32_t> cig[2];
  Scratch() : mem_scr1(kPeLarge, 0) {}
};

void store_cigar(const std::vector<uint32_t> &cig, uint32_t *dst, uint32_t *n_dst, uint32_t i,
                 uint32_t stride, bool &overflow) {
  if (!n_dst) return;
  n_dst[i] = static_cast<uint32_t>(cig.size());
  if (cig.size() > stride) overflow = true;
  if (dst)
    for (uint32_t k = 0; k < cig.size() && k < stride; ++k) dst[static_cast<size_t>(i) * stride + k] = cig[k];
}

inline abg_hit to_abg(const Hit &h) { return abg_hit{h.diffs, h.flags, h.pos}; }

}  // namespace

extern "C" {

const char *abo_last_error(void) { return g_err.c_str(); }

int abo_index_create(const abg_index_view *view, abo_index **out) {
  if (!view || !out || !view->genome || !view->counter || !view->counter_t || !view->counter_a) {
    g_err = "abo_index_create: null argument";
    return ABG_ERR_INVALID;
  }
  *out = new abo_index{*view};
  return ABG_OK;
}

void abo_index_destroy(abo_index *idx) { delete idx; }

int abo_map_batch(const abo_index *idx, const abg_params *params, const abg_batch *batch,
                  abg_results *results, abg_work_counters *wc) {
  if (!idx || !params || !batch || !results) {
    g_err = "abo_map_batch: null argument";
    return ABG_ERR_INVALID;
  }
  const abg_index_view &ix = idx->v;
  if (ix.window_size != 0 && ix.window_size != 12 && ix.window_size != 20) {
    g_err = "abo_map_batch: window_size must be 12 or 20";
    return ABG_ERR_INVALID;
  }
  kWindow = ix.window_size ? ix.window_size : 20u;
  kMinReadLen = kKeyWeight + kWindow - 1;
  const bool paired = params->mode & ABG_MODE_PAIRED;
  const bool a_rich = params->mode & ABG_MODE_A_RICH;
  const bool rpbat = params->mode & ABG_MODE_RANDOM_PBAT;
  const uint32_t maxc = params->max_candidates ? params->max_candidates : ix.max_candidates;
  const uint32_t stride = params->cigar_stride;
  const double valid_frac = params->valid_frac;
  if (paired && (!batch->seq2 || !batch->off2)) {
    g_err = "abo_map_batch: paired mode without second end";
    return ABG_ERR_INVALID;
  }
  static thread_local Scratch *sc = nullptr;
  if (!sc) sc = new Scratch();
  sc->aln.genome = ix.genome;
  sc->aln.wc = wc;
  bool overflow = false;

  const flags_t T = 0, A = ABG_FLAG_A_RICH, RC = ABG_FLAG_RC;
  if (!paired) {
    /* map_single_ended<conv> :1511-1600 / map_single_ended_rand :1602-1704 */
    flags_t passes[4];
    int n_pass = 0;
    if (rpbat) {
      passes[0] = T; passes[1] = A; passes[2] = A | RC; passes[3] = T | RC;
      n_pass = 4;
    }
    else {
      const flags_t c = a_rich ? A : T;
      passes[0] = c; passes[1] = c | RC;
      n_pass = 2;
    }
    for (uint32_t i = 0; i < batch->n; ++i) {
      const char *s = batch->seq1 + batch->off1[i];
      const uint32_t len = batch->off1[i + 1] - batch->off1[i];
      SeSet &res = sc->se;
      res.reset(len);
      Hit best{kMaxDiffs, 0, 0};
      std::vector<uint32_t> &cig = sc->cig[0];
      cig.clear();
      if (len != 0) {
        sc->enc[0].set(s, len);
        for (int p = 0; p < n_pass; ++p)
          process_seeds(ix, maxc, passes[p], sc->enc[0].for_flags(passes[p]), res, wc);
        align_se_candidates(sc->enc[0], len, valid_frac, res, best, cig, sc->aln);
      }
      results->se1[i] = to_abg(best);
      store_cigar(cig, results->cigar1, results->n_cigar1, i, stride, overflow);
    }
  }
  else {
    /* map_paired_ended<conv> :1887-2029 / map_paired_ended_rand :2031-2185 */
    PairCall calls[4];
    int n_calls = 0;
    if (rpbat) {
      calls[0] = {true, T, static_cast<flags_t>(A | RC), false};
      calls[1] = {false, A, static_cast<flags_t>(T | RC), true};
      calls[2] = {true, A, static_cast<flags_t>(T | RC), false};
      calls[3] = {false, T, static_cast<flags_t>(A | RC), true};
      n_calls = 4;
    }
    else if (!a_rich) {
      calls[0] = {true, T, static_cast<flags_t>(A | RC), false};
      calls[1] = {false, A, static_cast<flags_t>(T | RC), true};
      n_calls = 2;
    }
    else {
      calls[0] = {true, A, static_cast<flags_t>(T | RC), false};
      calls[1] = {false, T, static_cast<flags_t>(A | RC), true};
      n_calls = 2;
    }
    for (uint32_t i = 0; i < batch->n; ++i) {
      const char *s[2] = {batch->seq1 + batch->off1[i], batch->seq2 + batch->off2[i]};
      const uint32_t len[2] = {batch->off1[i + 1] - batch->off1[i], batch->off2[i + 1] - batch->off2[i]};
      SeSet *res_se[2] = {&sc->se, &sc->se2};
      res_se[0]->reset(len[0]);
      res_se[1]->reset(len[1]);
      PeBest best;
      best.reset(len[0], len[1]);
      Hit best_se[2];
      for (int e = 0; e < 2; ++e) {
        best_se[e] = Hit{kMaxDiffs, 0, 0};
        hit_reset(best_se[e], len[e]);
        sc->enc[e].set(s[e], len[e]);
        sc->cig[e].clear();
      }
      bool any_success = false;
      for (int c = 0; c < n_calls; ++c) {
        /* map_fragments :1849-1885: "1" is the un-reversed read, "2" the reversed */
        const int e1 = calls[c].first_is_r1 ? 0 : 1, e2 = 1 - e1;
        PeSet &res1 = sc->pe1, &res2 = sc->pe2;
        res1.reset(len[e1]);
        res2.reset(len[e2]);
        if (len[e1] == 0 && len[e2] == 0) continue;
        any_success = true;
        if (len[e1] != 0)
          process_seeds(ix, maxc, calls[c].flags_first, sc->enc[e1].for_flags(calls[c].flags_first), res1, wc);
        if (len[e2] != 0)
          process_seeds(ix, maxc, calls[c].flags_second, sc->enc[e2].for_flags(calls[c].flags_second), res2, wc);
        /* select_maps :1833-1847 */
        if (res1.should_align() && res2.should_align()) {
          res1.prepare_for_mating();
          res2.prepare_for_mating();
          /* an empty end has no candidates: best_pair then never aligns, so the
           * (stale, in the reference) read buffer of that end is irrelevant */
          static const EncRead empty_read;
          const EncRead &p1 = len[e1] ? sc->enc[e1].for_flags(calls[c].flags_first) : empty_read;
          const EncRead &p2 = len[e2] ? sc->enc[e2].for_flags(calls[c].flags_second) : empty_read;
          best_pair(calls[c].swap_ends, params->min_dist, params->max_dist, valid_frac, res1, res2, p1, p2,
                    sc->cig[e1], sc->cig[e2], sc->mem_scr1, sc->aln, best);
        }
        best_single(res1, *res_se[e1]);
        best_single(res2, *res_se[e2]);
      }
      if (!any_success) {  // :1981-1985
        best.reset();
        res_se[0]->reset();
        res_se[1]->reset();
      }
      /* valid_pair :624-631, :1987-1989 */
      {
        const uint32_t al1 = cigar_rseq_ops(sc->cig[0]), al2 = cigar_rseq_ops(sc->cig[1]);
        const bool ok = valid_len(al1, len[0]) && valid_len(al2, len[1]) &&
                        best.diffs() <= static_cast<score_t>(valid_frac * (al1 + al2));
        if (!ok) best.reset();
      }
      if (!best.should_report(params->allow_ambig)) {  // :1991-1999
        for (int e = 0; e < 2; ++e)
          align_se_candidates(sc->enc[e], len[e], valid_frac / 2.0, *res_se[e], best_se[e], sc->cig[e],
                              sc->aln);
      }
      results->pe_r1[i] = to_abg(best.r1);
      results->pe_r2[i] = to_abg(best.r2);
      results->se1[i] = to_abg(best_se[0]);
      results->se2[i] = to_abg(best_se[1]);
      store_cigar(sc->cig[0], results->cigar1, results->n_cigar1, i, stride, overflow);
      store_cigar(sc->cig[1], results->cigar2, results->n_cigar2, i, stride, overflow);
    }
  }
  if (overflow) {
    g_err = "abo_map_batch: a CIGAR exceeded cigar_stride";
    return ABG_ERR_CIGAR_OVERFLOW;
  }
  return ABG_OK;
}

}  // extern "C"
