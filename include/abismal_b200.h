/* abismal_b200.h -- C ABI of the B200-native `abismal map` hot path.
 *
 * The reference (smithlabcode/abismal v3.3.0) has no plugin/FFI seam: its hot
 * path is a set of static functions inside src/abismal.cpp.  This ABI is cut
 * at the one place a seam exists, the body of the per-batch worker loop:
 *
 *   reference caller                          what it hands over / gets back
 *   ----------------------------------------  ---------------------------------
 *   map_single_ended      abismal.cpp:1549-1581   batch of reads -> se_element + CIGAR
 *   map_single_ended_rand abismal.cpp:1645-1685   (same, four strand passes)
 *   map_paired_ended      abismal.cpp:1947-2002   batch of pairs -> pe_element,
 *   map_paired_ended_rand abismal.cpp:2091-2158     2 x se_element, 2 x CIGAR
 *
 * i.e. everything between ReadLoader::load_reads (:164) and
 * format_se/format_pe/select_output (:481/:648/:1073).  Plain pointers and
 * sizes only; no C++ or torch types cross this boundary.  All functions
 * return 0 on success and a negative code on failure (abg_last_error() holds
 * the message); nothing throws across the ABI.
 *
 * There is NO CPU fallback behind these symbols: abg_index_create fails when
 * no CUDA device is usable.
 */
#ifndef ABISMAL_B200_H
#define ABISMAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------- */
#define ABG_OK 0
#define ABG_ERR_INVALID (-1)      /* bad argument */
#define ABG_ERR_CUDA (-2)         /* CUDA runtime failure / no device */
#define ABG_ERR_TOO_LONG (-3)     /* read longer than the mapper was created for */
#define ABG_ERR_CIGAR_OVERFLOW (-4) /* a CIGAR needed more than cigar_stride ops */

/* ---- mode bits (abismal.cpp:2468-2483 dispatch) -------------------------- */
#define ABG_MODE_PAIRED 1u      /* two FASTQ files */
#define ABG_MODE_A_RICH 2u      /* SE: -A or -P; PE: -P (conv == a_rich) */
#define ABG_MODE_RANDOM_PBAT 4u /* -R: map_*_rand drivers */

/* ---- flag bits carried in abg_hit.flags ---------------------------------- */
#define ABG_FLAG_RC 0x10u      /* samflags::read_rc        common.hpp:114 */
#define ABG_FLAG_AMBIG 0x100u  /* samflags::secondary_aln  common.hpp:118 */
#define ABG_FLAG_A_RICH 0x1000u /* bsflags::read_is_a_rich  abismal.cpp:130-134 */

/* Mirrors se_element (abismal.cpp:224-233): 8 bytes, same field order.
 * pos == 0 means "empty" (se_element::empty, :276-279); after alignment
 * `diffs` holds NM (edit distance), before it holds the Hamming distance. */
typedef struct abg_hit {
  int16_t diffs;
  uint16_t flags;
  uint32_t pos;
} abg_hit;

/* Host view of the arrays of an AbismalIndex file (AbismalIndex.hpp:160-189,
 * on-disk order AbismalIndex.cpp:1082-1146).  Used only during
 * abg_index_create, which copies everything to HBM once. */
typedef struct abg_index_view {
  const uint64_t *genome;   /* 4-bit packed genome, 16 bases per word      */
  uint64_t genome_words;    /* ceil(genome_size / 16)                      */
  uint64_t genome_size;     /* bases incl. the two 32767-base N paddings   */
  const uint32_t *counter;  /* counter_size + 1 entries (2^25 + 1)         */
  uint64_t counter_size;
  const uint32_t *counter_t; /* counter_size_three + 1 entries (3^16 + 1)  */
  const uint32_t *counter_a;
  uint64_t counter_size_three;
  const uint32_t *index;    /* index_size genome positions                 */
  uint64_t index_size;
  const uint32_t *index_t;  /* index_size_three genome positions each      */
  const uint32_t *index_a;
  uint64_t index_size_three;
  uint32_t max_candidates;  /* value stored in the index file (default 100) */
  uint32_t window_size;     /* seed::window_size of the index file (AbismalIndex.hpp:73-77): 20, or 12 for
                             * an index built by a reference configured with --enable-short; 0 means 20 */
} abg_index_view;

/* Per-run knobs that the reference keeps in globals mutated by the CLI
 * (abismal.cpp:2321-2348, :2448-2452). */
typedef struct abg_params {
  uint32_t mode;           /* ABG_MODE_* bits                                */
  uint32_t allow_ambig;    /* -a                                             */
  uint32_t min_dist;       /* -l, pe_element::min_dist (default 32)          */
  uint32_t max_dist;       /* -L, pe_element::max_dist (default 3000)        */
  double valid_frac;       /* -m, se_element::valid_frac (default 0.1)       */
  uint32_t max_candidates; /* -c, 0 = use the index's value                  */
  uint32_t cigar_stride;   /* u32 CIGAR slots per read in abg_results (>= 4) */
} abg_params;

/* One batch of reads as the reference's ReadLoader (abismal.cpp:164-201)
 * leaves them: upper-case sequence, Ns trimmed from both ends, reads with
 * fewer than 25 + window_size - 1 (= 44, or 36 with window 12) non-N bases emptied (length 0).  Read i of end e occupies
 * seq_e[off_e[i] .. off_e[i+1]).  seq2/off2 are NULL for single-end. */
typedef struct abg_batch {
  uint32_t n;           /* reads (SE) or pairs (PE)                          */
  uint32_t reserved;
  const char *seq1;
  const uint32_t *off1; /* n + 1 offsets                                     */
  const char *seq2;
  const uint32_t *off2;
} abg_batch;

/* Caller-owned output arrays, n entries each (cigar*: n * cigar_stride).
 *   SE:  se1[i]   = bests[i]  (abismal.cpp:1574-1579, before format_se)
 *        cigar1   = r[i].cig
 *   PE:  pe_r1/pe_r2 = bests[i].r1/.r2 after valid_pair (abismal.cpp:1987-1989)
 *        se1/se2  = bests_se1[i]/bests_se2[i] (:1991-1999; reset when the
 *                   SE fallback did not run)
 *        cigar1/2 = r1[i].cig / r2[i].cig as select_output (:2000) sees them
 * Unused pointers may be NULL (pe_*, se2, cigar2, n_cigar2 in SE mode). */
typedef struct abg_results {
  abg_hit *pe_r1;
  abg_hit *pe_r2;
  abg_hit *se1;
  abg_hit *se2;
  uint32_t *cigar1;   /* BAM-encoded ops: len << 4 | op (M=0 I=1 D=2 S=4)   */
  uint32_t *cigar2;
  uint32_t *n_cigar1; /* ops used per read (0 = empty CIGAR)                */
  uint32_t *n_cigar2;
} abg_results;

/* Work counters of the last abg_map_batch call, in the reference's units
 * (SURVEY.md section 8d).  Filled only when the mapper was created with
 * count_work != 0 (slower kernel variant). */
typedef struct abg_work_counters {
  uint64_t n_lookup;  /* counter probes (2 adjacent u32 each)                 */
  uint64_t n_entry;   /* index entries scanned by check_hits                   */
  uint64_t n_cmp;     /* full_compare calls                                    */
  uint64_t n_word;    /* packed words compared incl. early exit               */
  uint64_t n_align;   /* banded alignments run (score-only + traceback)       */
  uint64_t n_dpref;   /* reference bases read by align (q_sz + bw per call)   */
} abg_work_counters;

typedef struct abg_index abg_index;   /* index resident in HBM of one GPU  */
typedef struct abg_mapper abg_mapper; /* stream + staging + scratch        */

const char *abg_last_error(void);
int abg_device_count(void);

/* Replaces AbismalIndex::read's std::vector members as the thing the hot
 * path reads (abismal.cpp:1518-1527): copies the seven arrays to `device`. */
int abg_index_create(const abg_index_view *view, int device, abg_index **out);
void abg_index_destroy(abg_index *idx);
uint64_t abg_index_device_bytes(const abg_index *idx);
/* Which derived arrays the index holds next to the file's seven (they only accelerate the same lookups; results
 * never depend on them): bit set = present.  GENOME_HAS_IUPAC: the genome holds multi-bit codes, the candidates
 * whose compare window can reach one take the exact 4-bit full_compare (abismal.cpp:1105-1122) instead of the
 * seed-context prefilter. */
#define ABG_FEATURE_SEED_CONTEXT 1u
#define ABG_FEATURE_COMPACT_COUNTERS 2u
#define ABG_FEATURE_GENOME_HAS_IUPAC 4u
uint32_t abg_index_features(const abg_index *idx);

/* max_batch: largest abg_batch.n; max_read_len: longest read accepted. */
int abg_mapper_create(abg_index *idx, const abg_params *params, uint32_t max_batch,
                      uint32_t max_read_len, int count_work, abg_mapper **out);
void abg_mapper_destroy(abg_mapper *m);

/* Page-locked host memory for batch and result buffers.  Buffers from here are
 * DMA'd in place; pageable buffers work too but are staged through the
 * mapper's own pinned memory (one extra host copy). */
int abg_host_alloc(size_t bytes, void **out);
void abg_host_free(void *p);

/* Host buffers in, host buffers out.  Internally the batch is cut into
 * sub-batches (abg_mapper_chunk reads) pipelined over three streams, so the
 * H2D copy of one sub-batch, the kernel of the previous and the D2H copy of the
 * one before overlap.  Of every CIGAR only the operations actually used are
 * written to results->cigar*; the rest of each cigar_stride row is untouched. */
int abg_map_batch(abg_mapper *m, const abg_batch *batch, abg_results *results);
uint32_t abg_mapper_chunk(const abg_mapper *m);

/* Split form used to time the device part alone: upload once, run many. */
int abg_mapper_upload(abg_mapper *m, const abg_batch *batch);
int abg_mapper_run(abg_mapper *m);          /* async launch on the mapper's stream */
int abg_mapper_sync(abg_mapper *m);
int abg_mapper_download(abg_mapper *m, abg_results *results);
/* CUDA-event time of the most recent abg_mapper_run (after abg_mapper_sync). */
float abg_mapper_last_kernel_ms(const abg_mapper *m);
/* The same run split by kernel (CUDA events on the launching stream between the launches):
 * out[0] seed_kernel, out[1] align_kernel, out[2] map_reads_kernel over the redo list.
 * Single-kernel mode (ABISMAL_B200_SPLIT=0): out[0] = the whole run. */
void abg_mapper_last_phase_ms(const abg_mapper *m, float out[3]);
/* All five kernels of a run: out[0] seed_kernel, out[1] enum_kernel, out[2] dp_kernel, out[3] align_kernel,
 * out[4] map_reads_kernel over the redo list (out[1] = out[2] = 0 without the task-parallel alignment). */
void abg_mapper_last_kernel_times(const abg_mapper *m, float out[5]);
/* Binned seeding (the default when the index carries seed-context records): out[0] hash_kernel, out[1]
 * scatter_kernel (with the bins' prefix sum), out[2] filter_kernel, out[3] seed_kernel of the last abg_mapper_run;
 * their sum is out[0] of abg_mapper_last_kernel_times.  Without binning out[3] is the whole seeding. */
void abg_mapper_last_seed_times(const abg_mapper *m, float out[4]);
/* 1 when the mapper seeds through the binned kernels, 0 when every strand gathers its own records. */
int abg_mapper_binned(const abg_mapper *m);
/* Diagnostics of the last batch (binned seeding): out[0] strands, out[1] strands that took the direct path
 * (reads with N, survivor or tuple overflow), out[2] tuples binned, out[3] prefilter survivors, out[4] bins,
 * out[5] tuple capacity, out[6] kernel variants in use (bit 0 tile-sorted scatter, bit 1 pipelined filter),
 * out[7] tuples a filter warp takes per work-cursor step (0 = static distribution). */
int abg_mapper_bin_stats(abg_mapper *m, uint64_t out[8]);
uint32_t abg_mapper_launches_per_run(const abg_mapper *m);
int abg_mapper_get_counters(const abg_mapper *m, abg_work_counters *out);
/* Diagnostics of the most recent abg_mapper_run (after abg_mapper_sync): out[0] pairs left to the redo kernel,
 * out[1] / out[2] entries used / available in the arena of stored candidate sets beyond 32 entries,
 * out[3..5] alignment tasks emitted per band class, out[6] traceback units handed out, out[7] kernel error flag. */
int abg_mapper_last_run_stats(abg_mapper *m, uint32_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* ABISMAL_B200_H */
