// align_tasks.cuh -- task-parallel banded alignment between seeding and selection.
//
// The banded local alignments of AbismalAlign::align (AbismalAlign.hpp:320-386) are pure functions of
// (read end, strand/conversion, genome position, band width).  Which of them a read or pair needs is known
// before any score is: align_se_candidates (abismal.cpp:1435-1497) aligns every unique candidate, best_pair
// (:1722-1831) every candidate that has a concordant partner (its early exits only skip work).  So instead of
// running them one after another inside the warp that owns the pair,
//
//   enum_kernel   one warp per read / pair: sorts + uniques the stored candidate sets (prepare_for_alignments /
//                 prepare_for_mating), finds the candidates the selection can ask about and emits one TASK
//                 per alignment into three lists by band width (<= 16, <= 32, <= 61 columns);
//   dp_kernel     one GROUP of 8 / 16 / 32 lanes per task (4 / 2 / 1 alignments per warp at a time): nothing
//                 but the wavefront loop, so the code is L0-resident and the lanes are busy; scores, best
//                 cells and traceback words go to HBM;
//   align_kernel  (mapper_kernels.cuh) replays best_pair / align_se_candidates in the reference's order and
//                 takes the scores and tracebacks from the tasks instead of running the DP.  Alignments no task
//                 was emitted for (single-end fallback of a pair, arena full) still run in the warp.
#pragma once

#include "mapper_kernels.cuh"

namespace ab2dev {

// ---- the DP of one task on a group of G lanes ------------------------------------------------------------
// Same recurrence, tie-breaks and traceback layout as align_wave (mapper_kernels.cuh); differences: the loop
// runs in blocks of 16 anti-diagonals from T = 0 with zero-padded query / reference bytes around the staged
// region (no tail cases, no guards on the loads), the best cell is tracked as one packed key, and the
// three-way maximum uses the fused add-max instructions.
//
// Shared memory of a group: [64 zero bytes | query codes, ml + 160 | 32 pad | reference bytes, ml + 128]
__host__ __device__ __forceinline__ uint32_t dp_q_bytes(uint32_t ml) { return 64u + ml + 160u; }
__host__ __device__ __forceinline__ uint32_t dp_ref_bytes(uint32_t ml) { return 32u + ml + 128u; }
__host__ __device__ __forceinline__ uint32_t dp_group_bytes(uint32_t ml) { return dp_q_bytes(ml) + dp_ref_bytes(ml); }
constexpr int kDpGroupsPerWarp = 4;  // G = 8
__host__ __device__ __forceinline__ size_t dp_block_smem_bytes(uint32_t ml) {
  return (size_t)kParamBytes + kTab3Bytes + (size_t)dp_group_bytes(ml) * kDpGroupsPerWarp * kWarpsPerBlock;
}
constexpr uint32_t kDpMaxMl = 512;  // longer reads keep the in-warp DP (shared memory per group grows with ml)

template <int G>
__device__ __forceinline__ void dp_tasks(const KernelParams &P, const AlignTask *tasks, uint32_t n_tasks, uint32_t first,
                                         unsigned char *warp_smem) {
  constexpr int GROUPS = 32 / G;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G, l = lane % G;
  const uint32_t ml = P.ml;
  unsigned char *gs = warp_smem + (size_t)grp * dp_group_bytes(ml) * (kDpGroupsPerWarp / GROUPS);
  uint8_t *qz = gs;                       // 64 zero bytes, then the query codes
  uint8_t *q = gs + 64;
  uint8_t *rz = gs + dp_q_bytes(ml);      // 32 pad bytes, then the reference bytes
  uint8_t *refb = rz + 32;

  const uint32_t tix = first + (uint32_t)grp;
  const bool have = tix < n_tasks;
  AlignTask t;
  t.t_pos = 0; t.item = 0; t.meta = 0; t.tb_index = kNoTask;
  if (have) t = tasks[tix];
  const int end = (int)(t.meta & 1u);
  const bool rc = (t.meta >> 1) & 1u, a_rich = (t.meta >> 2) & 1u;
  const int bw = have ? (int)((t.meta >> 8) & 255u) : 0;  // 0: a slot its warp did not fill
  uint32_t q_sz = 0;
  // ---- stage the query (prep_read for this strand / conversion) and the reference window -------------
  __syncwarp();
  {
    const uint32_t qb = dp_q_bytes(ml), rb = dp_ref_bytes(ml);
    for (uint32_t i = l * 4u; i < qb; i += G * 4u) *reinterpret_cast<uint32_t *>(qz + i) = 0u;
    for (uint32_t i = l * 4u; i < rb; i += G * 4u) *reinterpret_cast<uint32_t *>(rz + i) = 0u;
  }
  __syncwarp();
  if (bw != 0) {
    const uint32_t o = P.off[end][t.item];
    q_sz = P.off[end][t.item + 1] - o;
    const char *s = P.seq[end] + o;
    const bool enc_a = a_rich != rc;
    for (uint32_t i = l; i < q_sz; i += G) {
      const char ch = s[rc ? q_sz - 1u - i : i];
      uint32_t x = 0;
      if (ch == 'A' || ch == 'a') x = 1;
      else if (ch == 'C' || ch == 'c') x = 2;
      else if (ch == 'G' || ch == 'g') x = 4;
      else if (ch == 'T' || ch == 't') x = 8;
      if (rc) x = __brev(x) >> 28;
      q[i] = (uint8_t)(enc_a ? (x == 1u ? 5u : x) : (x == 8u ? 10u : x));
    }
    const uint32_t t_beg = t.t_pos - (uint32_t)((bw - 1) / 2);
    const int n_ref = (int)q_sz + bw - 1;
    const uint32_t w0 = t_beg >> 4;
    const int nw = (int)(((t_beg + (uint32_t)n_ref - 1u) >> 4) - w0) + 1;
    const int shift0 = (int)(t_beg & 15u);
    for (int k = l; k < nw; k += G) {
      const uint64_t word = __ldg(P.ix.genome + w0 + k);
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int r = 16 * k + n - shift0;
        if (r >= 0 && r < n_ref) refb[r] = (uint8_t)((word >> (4 * n)) & 15u);
      }
    }
  }
  __syncwarp();
  const int nl = (bw + 1) >> 1;
  int n_iter = bw != 0 ? ((int)q_sz + bw - 1) + (nl - 1) : 0;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) n_iter = max(n_iter, __shfl_xor_sync(FULL, n_iter, d));  // warp-uniform loop bound
  const int n_blk = (n_iter >> 4) + 1;  // T = 0 .. 16 n_blk - 1 covers 1 .. n_iter

  const unsigned limA = 2 * l < bw ? q_sz : 0u;
  const unsigned limB = 2 * l + 1 < bw ? q_sz : 0u;
  const int up_mask = l > 0 ? -1 : 0, down_mask = l < G - 1 ? -1 : 0;  // no neighbour beyond the group
  const int aboveA_lim = (int)q_sz - 1, aboveB_lim = (int)q_sz - 2;  // `above` exists for qi < q_sz - 1 (A) / qi < q_sz - 2 (B)
  const bool rec = bw != 0 && t.tb_index != kNoTask && l < nl;
  uint64_t *tbp = P.task_tb + (size_t)t.tb_index * 8u + l;   // [block][G]
  int A = 0, B = 0;
  int best = 0;                         // (value << 14) + (16383 - (2 T + column B))
  int kbase = 16383;                    // 16383 - 2 T
  int qi = l - bw;                      // qi(A) at T = 0
  const uint8_t *rp = refb - 1 - l;     // rp[T] = reference base of this lane's row at iteration T
  const uint8_t *qp = q + qi + 1;       // qp[T] = query base of column B at iteration T
  uint32_t qa = (uint32_t)q[qi];        // zero padding in front: qi >= -61
  for (int blk = 0; blk < n_blk; ++blk) {
    uint32_t tw[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t w = 0;
#pragma unroll 4
      for (int k = 0; k < 8; ++k) {
        const uint32_t ref = (uint32_t)rp[0];
        const uint32_t qb = (uint32_t)qp[0];
        const int left_in = __shfl_up_sync(FULL, B, 1, G) & up_mask;
        // column A
        const int diagA = A + ((qa & ref) ? 2 : -3);
        const int aboveA = (qi < aboveA_lim ? B : 0) - 4;
        const int leftA = left_in - 4;
        int vA = __vimax3_s32_relu(diagA, aboveA, leftA);
        vA = (unsigned)qi < limA ? vA : 0;
        uint32_t cA = leftA == vA ? 1u : (aboveA == vA ? 2u : 0u);
        cA = vA > 0 ? cA : 3u;
        best = max(best, vA * 16384 + kbase);
        const int a_down = __shfl_down_sync(FULL, vA, 1, G) & down_mask;
        // column B
        const int diagB = B + ((qb & ref) ? 2 : -3);
        const int aboveB = (qi < aboveB_lim ? a_down : 0) - 4;
        const int leftB = vA - 4;
        int vB = __vimax3_s32_relu(diagB, aboveB, leftB);
        vB = (unsigned)(qi + 1) < limB ? vB : 0;
        uint32_t cB = leftB == vB ? 1u : (aboveB == vB ? 2u : 0u);
        cB = vB > 0 ? cB : 3u;
        best = max(best, vB * 16384 + kbase - 1);
        A = vA;
        B = vB;
        qa = qb;
        w = (w >> 4) | ((cA | (cB << 2)) << 28);
        ++qi;
        ++rp;
        ++qp;
        kbase -= 2;
      }
      tw[h] = w;
    }
    if (rec) tbp[(size_t)blk * G] = (uint64_t)tw[0] | ((uint64_t)tw[1] << 32);
  }
  // first maximum in row-major order within the group (std::max_element)
  int bv = best >> 14;
  const int bT2 = 16383 - (best & 16383);  // 2 T + column B
  int br = (bT2 >> 1) - l, bc = 2 * l + (bT2 & 1);
  if (bv == 0) br = 0, bc = 0;
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) {
    const int ov = __shfl_xor_sync(FULL, bv, d, G);
    const int orow = __shfl_xor_sync(FULL, br, d, G);
    const int oc = __shfl_xor_sync(FULL, bc, d, G);
    const bool take = ov > bv || (ov == bv && (orow < br || (orow == br && oc < bc)));
    if (take) {
      bv = ov;
      br = orow;
      bc = oc;
    }
  }
  if (bw != 0 && l == 0) {
    TaskResult r;
    r.score = (int16_t)bv;
    r.row = (int16_t)br;
    r.col = (int16_t)bc;
    r.bw = (int16_t)bw;
    P.task_res[(tasks - P.tasks) + tix] = r;
  }
  __syncwarp();
}

// One persistent launch over the three task lists of a sub-batch.
__global__ void __launch_bounds__(kThreadsPerBlock, 4) dp_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const int lane = threadIdx.x & 31;
  unsigned char *warp_smem = smem_raw + kParamBytes + kTab3Bytes +
                             (size_t)dp_group_bytes(P.ml) * kDpGroupsPerWarp * (threadIdx.x >> 5);
  // class c: groups of 8 << c lanes, tasks [task_base[c], task_base[c] + n_tasks[c]) of P.tasks
  for (int c = 0; c < 3; ++c) {
    const uint32_t n = min(__ldcg(P.task_count + c), P.task_cap[c]);
    const AlignTask *list = P.tasks + P.task_base[c];
    const uint32_t per_warp = 4u >> c;
    for (;;) {
      uint32_t first = 0;
      if (lane == 0) first = atomicAdd(P.task_cursor + c, per_warp);
      first = __shfl_sync(FULL, first, 0);
      if (first >= n) break;
      if (c == 0) dp_tasks<8>(P, list, n, first, warp_smem);
      else if (c == 1) dp_tasks<16>(P, list, n, first, warp_smem);
      else dp_tasks<32>(P, list, n, first, warp_smem);
    }
  }
}

// ---- enumeration ------------------------------------------------------------------------------------------
// Task slots and traceback words are handed out to a warp in blocks (one atomic per block, not per task):
// warp-uniform cursors.  A block's slots are cleared when it is reserved, so the slots a warp leaves unused
// are empty tasks (band 0) that dp_kernel skips.
constexpr uint32_t kTaskBlock = 8;     // task slots a warp reserves at a time
constexpr uint32_t kTbGrabTasks = 4;   // traceback words it reserves at a time, in tasks of the first class
struct TaskAlloc {
  uint32_t cur[3], end[3];
  uint32_t tb_cur, tb_end;
  __device__ __forceinline__ void init() {
    cur[0] = cur[1] = cur[2] = end[0] = end[1] = end[2] = 0u;
    tb_cur = tb_end = 0u;
  }
};

// Appends the tasks of the lanes that have one (`want`) to the list of their class; returns the task id
// (position in P.tasks) or kNoTask when the list is full.  want_tb: also reserve traceback words.
__device__ __forceinline__ uint32_t emit_task(const KernelParams &P, TaskAlloc &al, bool want, int bw, uint32_t t_pos,
                                              uint32_t item, uint32_t meta, bool want_tb, int lane) {
  uint32_t id = kNoTask;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int lo = c == 0 ? 0 : (c == 1 ? 17 : 33), hi = c == 0 ? 16 : (c == 1 ? 32 : 61);
    const bool mine = want && bw >= lo && bw <= hi;
    const unsigned m = __ballot_sync(FULL, mine);
    if (m == 0u) continue;
    const uint32_t n = (uint32_t)__popc(m);
    if (al.cur[c] + n > al.end[c]) {  // a fresh block (the rest of the old one stays empty)
      const uint32_t grab = max(n, kTaskBlock);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.task_count + c, grab);
      base = __shfl_sync(FULL, base, 0);
      const uint32_t lim = min(base + grab, P.task_cap[c]);
      al.cur[c] = min(base, lim);
      al.end[c] = lim;
      AlignTask *blk = P.tasks + P.task_base[c];
      for (uint32_t k = al.cur[c] + lane; k < lim; k += 32) blk[k] = AlignTask{0u, 0u, 0u, kNoTask};
      __syncwarp();
    }
    const uint32_t tbw = P.tb_words << c;  // traceback words of one task (tb_words blocks x 8 << c lanes), in units of 8 words
    const unsigned mt = __ballot_sync(FULL, mine && want_tb);
    const uint32_t need = (uint32_t)__popc(mt) * tbw;
    if (need != 0u && al.tb_cur + need > al.tb_end) {
      const uint32_t grab = max(need, kTbGrabTasks * P.tb_words);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.task_count + 3, grab);
      base = __shfl_sync(FULL, base, 0);
      const uint32_t lim = min(base + grab, P.task_tb_cap);
      al.tb_cur = min(base, lim);
      al.tb_end = lim;
    }
    if (mine) {
      const uint32_t k = al.cur[c] + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (k < al.end[c]) {
        AlignTask t;
        t.t_pos = t_pos;
        t.item = item;
        t.meta = meta | ((uint32_t)bw << 8);
        t.tb_index = kNoTask;
        if (want_tb) {
          const uint32_t tb = al.tb_cur + (uint32_t)__popc(mt & ((1u << lane) - 1u)) * tbw;
          if (tb + tbw <= al.tb_end) t.tb_index = tb;
        }
        id = P.task_base[c] + k;
        P.tasks[id] = t;
      }
    }
    al.cur[c] = min(al.cur[c] + n, al.end[c]);
    al.tb_cur = min(al.tb_cur + need, al.tb_end);
  }
  return id;
}

// Tasks of the entries of one sorted + uniqued candidate set that need_fn selects (entries with diffs == 0
// need no DP, AbismalAlign.hpp:329-330).  meta_fn(hit): end / strand / conversion bits of the task.
template <class F, class M>
__device__ __forceinline__ void enumerate_set(const Warp &W, TaskAlloc &al, int set_id, int max_diffs, uint32_t item,
                                              uint32_t *task_of, uint32_t slots, bool want_tb, F need_fn, M meta_fn) {
  const KernelParams &P = params();
  const HeapRef v = heap_of(W, set_id);
  const int sz = W.cs(set_id)->sz;
  const uint32_t ovf = W.cs(set_id)->ovf;
  const int n = max(sz, (int)slots);
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + W.lane;
    bool want = false;
    Hit h;
    if (j < sz) {
      h = v.get(j);
      want = !h.empty() && h.diffs() != 0 && need_fn(j, h);
    }
    const int bw = want ? band_width(h.diffs(), max_diffs) : 0;
    const uint32_t id = emit_task(P, al, want, bw, h.pos(), item, want ? meta_fn(h) : 0u, want_tb, W.lane);
    if (j < (int)slots) task_of[j] = id;
    else if (j < sz && ovf != 0u) P.task_ovf[(ovf - 1u) + (uint32_t)(j - (int)slots)] = id;
  }
}

__device__ __forceinline__ uint32_t task_meta(int end, uint32_t flags) {
  return (uint32_t)end | ((flags & ABG_FLAG_RC) ? 2u : 0u) | ((flags & ABG_FLAG_A_RICH) ? 4u : 0u);
}

// first index in [lo, hi) of the position-sorted set whose pos + add >= lim (32-bit arithmetic as in best_pair)
__device__ __forceinline__ int first_reaching(const HeapRef &v, int lo, int hi, uint32_t add, uint32_t lim) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v.get(mid).pos() + add >= lim) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ uint32_t *task_slots(const KernelParams &P, unsigned item, int pass) {
  return P.task_of + ((size_t)item * P.n_pass + (size_t)pass) * P.set_slots;
}

template <int MINB>
__global__ void __launch_bounds__(kThreadsPerBlock, MINB) enum_kernel(const __grid_constant__ KernelParams Pin) {
  block_prologue(Pin);
  const KernelParams &P = params();
  const Warp W;
  const int lane = W.lane;
  const bool paired = P.mode & ABG_MODE_PAIRED;
  const bool a_rich = P.mode & ABG_MODE_A_RICH;
  const bool rpbat = P.mode & ABG_MODE_RANDOM_PBAT;
  WarpScalars *S = W.scal();
  TaskAlloc al;
  al.init();
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(P.work_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= P.n) break;
    if (__ldcg(P.redo_flag + item) != 0) continue;
    if (!paired) {
      const uint32_t len = P.off[0][item + 1] - P.off[0][item];
      uint32_t *tof = task_slots(P, item, 0);
      if (len == 0) continue;
      uint64_t *ss = stored_set(P, item, 0);
      load_set(W, 0, ss);
      // align_se_candidates (abismal.cpp:1435-1497): nothing to align when an exact match exists
      const bool exact = !Hit(W.cs(0)->best).empty();
      sort_unique(0);
      const HeapRef v = heap_of(W, 0);
      const int sz = W.cs(0)->sz;
      int n_real = 0;
      for (int j0 = 0; j0 < sz; j0 += 32) n_real += __popc(__ballot_sync(FULL, j0 + lane < sz && !v.get(j0 + lane).empty()));
      const int invalid = invalid_hit_diffs(len);
      const int max_diffs = frac_of(P.valid_frac, (uint32_t)(int)(int16_t)len);
      enumerate_set(W, al, 0, max_diffs, item, tof, P.set_slots, n_real <= kTbCacheMaxCands,
                    [&](int, Hit h) { return !exact && h.diffs() < invalid; },
                    [&](Hit h) { return task_meta(0, h.flags()); });  // strand / conversion differ per candidate
      store_set(W, 0, ss, (int)P.set_slots);
      continue;
    }
    const uint32_t len0 = P.off[0][item + 1] - P.off[0][item], len1 = P.off[1][item + 1] - P.off[1][item];
    if (lane == 0) {
      S->len[0] = len0;
      S->len[1] = len1;
    }
    __syncwarp();
    const int n_calls = rpbat ? 4 : 2;
    for (int call = 0; call < n_calls; ++call) {
      const CallPlan cp = call_plan(call, rpbat, a_rich);
      const uint32_t l1 = cp.first_is_r1 ? len0 : len1, l2 = cp.first_is_r1 ? len1 : len0;
      uint32_t *tof1 = task_slots(P, item, 2 * call), *tof2 = task_slots(P, item, 2 * call + 1);
      for (uint32_t j = lane; j < P.set_slots; j += 32) {
        tof1[j] = kNoTask;
        tof2[j] = kNoTask;
      }
      if (l1 == 0 && l2 == 0) continue;
      uint64_t *ss1 = stored_set(P, item, 2 * call), *ss2 = stored_set(P, item, 2 * call + 1);
      load_set(W, 2, ss1);
      load_set(W, 3, ss2);
      CandSet t0, t1;
      t0.load(W, 2);
      t1.load(W, 3);
      if (!(t0.should_align() && t1.should_align())) continue;
      sort_unique(2);
      sort_unique(3);
      const HeapRef v1 = heap_of(W, 2), v2 = heap_of(W, 3);
      const int n1 = W.cs(2)->sz, n2 = W.cs(3)->sz;
      // leading empties (position 0 sorts first)
      int e1 = 0, e2 = 0;
      for (int j0 = 0; j0 < n1; j0 += 32) e1 += __popc(__ballot_sync(FULL, j0 + lane < n1 && v1.get(j0 + lane).empty()));
      for (int j0 = 0; j0 < n2; j0 += 32) e2 += __popc(__ballot_sync(FULL, j0 + lane < n2 && v2.get(j0 + lane).empty()));
      const uint32_t min_dist = P.min_dist, max_dist = P.max_dist;
      const int max_diffs1 = frac_of(P.valid_frac, l1), max_diffs2 = frac_of(P.valid_frac, l2);
      const bool rec1 = (n1 - e1) <= kTbCacheMaxCands, rec2 = (n2 - e2) <= kTbCacheMaxCands;
      // best_pair evaluates (j1, j2) iff  pos1 + max_dist >= lim  and  pos1 + min_dist <= lim,  lim = pos2 + l2
      // (the empties of set 1 can be reached by its rewinding j1, but never pass the first test there: lim
      //  exceeds max_dist for every genome position beyond the padding)
      const uint32_t meta1 = task_meta(cp.e1, cp.f1), meta2 = task_meta(cp.e2, cp.f2);
      enumerate_set(W, al, 3, max_diffs2, item, tof2, P.set_slots, rec2, [&](int, Hit h) {
        const uint32_t lim = h.pos() + l2;
        const int j1 = first_reaching(v1, e1, n1, max_dist, lim);
        return j1 < n1 && v1.get(j1).pos() + min_dist <= lim;
      }, [&](Hit) { return meta2; });
      enumerate_set(W, al, 2, max_diffs1, item, tof1, P.set_slots, rec1, [&](int, Hit h) {
        // some j2 with pos1 + min_dist <= pos2 + l2 <= pos1 + max_dist: the first one reaching the lower end
        const uint32_t p1 = h.pos();
        int lo = e2, hi = n2;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (v2.get(mid).pos() + l2 >= p1 + min_dist) hi = mid;
          else lo = mid + 1;
        }
        return lo < n2 && p1 + max_dist >= v2.get(lo).pos() + l2;
      }, [&](Hit) { return meta1; });
      store_set(W, 2, ss1, (int)P.set_slots);
      store_set(W, 3, ss2, (int)P.set_slots);
    }
  }
}

}  // namespace ab2dev
