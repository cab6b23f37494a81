/* abismal_b200_index.h -- C ABI of the GPU AbismalIndex builder (the
 * "next" row 8f-3 of SURVEY.md: `abismal idx`).
 *
 * Replaces AbismalIndex::create_index (src/AbismalIndex.cpp:281-331) from the
 * point where the reference holds the 4-bit encoded, padded genome
 * (`genome`, `cl`, `exclude`) up to the filled `counter*` / `index*` vectors
 * that AbismalIndex::write (:1037-1072) puts on disk.  FASTA reading, the two
 * 32767-base N paddings, replacement of N runs <= 256 by the reference's LCG
 * bases and the 4-bit encoding are host work (abismal_b200/index_build.py).
 */
#ifndef ABISMAL_B200_INDEX_H
#define ABISMAL_B200_INDEX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Host arrays owned by the library; release with abg_built_index_free. */
typedef struct abg_built_index {
  uint32_t *counter;   /* counter_size + 1 start offsets into index          */
  uint32_t *counter_t; /* counter_size_three + 1                             */
  uint32_t *counter_a;
  uint32_t *index;     /* index_size genome positions, buckets sorted        */
  uint32_t *index_t;   /* index_size_three                                   */
  uint32_t *index_a;
  uint64_t counter_size, counter_size_three, index_size, index_size_three;
  uint32_t max_candidates;
  uint32_t reserved;
} abg_built_index;

const char *abg_index_build_last_error(void);

/* genome: 4-bit packed (dna_four_bit_encoding, 16 bases per word), including
 * both paddings; exclude: n_exclude pairs [first, second) of N runs longer
 * than 256 bases, ascending (AbismalIndex::exclude), paddings included. */
int abg_build_index(const uint64_t *genome, uint64_t genome_size, const uint64_t *exclude, uint32_t n_exclude,
                    int device, abg_built_index *out);
/* The same with seed::window_size given (src/AbismalIndex.hpp:73-77): 20, or 12 = the index a reference
 * configured with --enable-short builds (denser selection of positions, reads down to 36 bases). */
int abg_build_index_w(const uint64_t *genome, uint64_t genome_size, const uint64_t *exclude, uint32_t n_exclude,
                      uint32_t window_size, int device, abg_built_index *out);
void abg_built_index_free(abg_built_index *b);

#ifdef __cplusplus
}
#endif
#endif
