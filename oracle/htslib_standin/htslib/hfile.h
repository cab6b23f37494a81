/* Minimal stand-in for <htslib/hfile.h>; test infrastructure only (see sam.h). */
#ifndef ABISMAL_B200_HTS_STANDIN_HFILE_H
#define ABISMAL_B200_HTS_STANDIN_HFILE_H
#include <sys/types.h>
#include <zlib.h>
struct hFILE {
  gzFile gz;
};
static inline off_t htell(hFILE *fp) { return static_cast<off_t>(gzoffset(fp->gz)); }
#endif
