/* Minimal stand-in for <htslib/bgzf.h> over zlib's gzFile; test infrastructure
 * only (see sam.h).  Reads plain or gzip/bgzf-compressed text. */
#ifndef ABISMAL_B200_HTS_STANDIN_BGZF_H
#define ABISMAL_B200_HTS_STANDIN_BGZF_H
#include <cstdlib>
#include <cstring>
#include <zlib.h>
#include "hfile.h"
#include "sam.h"

struct BGZF {
  hFILE *fp;
  hFILE h;
};

static inline BGZF *bgzf_open(const char *fn, const char *mode) {
  gzFile gz = gzopen(fn, mode[0] == 'r' ? "rb" : "wb");
  if (!gz) return nullptr;
  gzbuffer(gz, 1 << 20);
  BGZF *b = new BGZF();
  b->h.gz = gz;
  b->fp = &b->h;
  return b;
}
static inline int bgzf_close(BGZF *b) {
  const int r = gzclose(b->h.gz);
  delete b;
  return r == Z_OK ? 0 : -1;
}
/* returns line length (terminator stripped, also a '\r' before '\n'),
 * -1 on EOF with nothing read */
static inline int bgzf_getline(BGZF *b, int delim, kstring_t *str) {
  str->l = 0;
  bool got_any = false;
  for (;;) {
    const int c = gzgetc(b->h.gz);
    if (c < 0) break;
    got_any = true;
    if (c == delim) break;
    if (str->l + 2 > str->m) {
      str->m = str->m ? str->m * 2 : 256;
      str->s = static_cast<char *>(std::realloc(str->s, str->m));
    }
    str->s[str->l++] = static_cast<char>(c);
  }
  if (!got_any) return -1;
  if (!str->s) {
    str->m = 8;
    str->s = static_cast<char *>(std::malloc(str->m));
  }
  if (delim == '\n' && str->l > 0 && str->s[str->l - 1] == '\r') --str->l;
  str->s[str->l] = 0;
  return static_cast<int>(str->l);
}
ssize_t bgzf_write(BGZF *, const void *, size_t);
int bgzf_compression(BGZF *);
#endif
