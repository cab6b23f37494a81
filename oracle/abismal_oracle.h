/* abismal_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's read-mapping hot path (see
 * abismal_oracle.cpp).  Same data contract as include/abismal_b200.h so that
 * tests can feed identical batches to both and compare records bit for bit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Parity status: PINNED -- oracle/_ref/abismal (the unmodified reference,
 * compiled by oracle/Makefile) reproduces all 16 golden md5s of the
 * reference's own test suite (data/md5sum.txt), and tests/test_oracle_*.py
 * check this restatement against that binary on SE/PE/PBAT/RPBAT inputs.
 */
#ifndef ABISMAL_ORACLE_H
#define ABISMAL_ORACLE_H

#include "abismal_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct abo_index abo_index;

const char *abo_last_error(void);
/* Keeps the host pointers of `view` (no copy): the caller keeps them alive. */
int abo_index_create(const abg_index_view *view, abo_index **out);
void abo_index_destroy(abo_index *idx);
/* Single-threaded; `counters` may be NULL.  Counters accumulate (+=). */
int abo_map_batch(const abo_index *idx, const abg_params *params, const abg_batch *batch,
                  abg_results *results, abg_work_counters *counters);

#ifdef __cplusplus
}
#endif
#endif
