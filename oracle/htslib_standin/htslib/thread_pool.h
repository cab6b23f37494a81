/* Minimal stand-in for <htslib/thread_pool.h>; declarations only (see sam.h). */
#ifndef ABISMAL_B200_HTS_STANDIN_TPOOL_H
#define ABISMAL_B200_HTS_STANDIN_TPOOL_H
#include "bgzf.h"
struct hts_tpool;
typedef struct htsThreadPool {
  hts_tpool *pool;
  int qsize;
} htsThreadPool;
hts_tpool *hts_tpool_init(int);
void hts_tpool_destroy(hts_tpool *);
int hts_set_thread_pool(htsFile *, htsThreadPool *);
int bgzf_thread_pool(BGZF *, hts_tpool *, int);
#endif
