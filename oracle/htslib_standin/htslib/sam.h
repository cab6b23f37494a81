/* Minimal stand-in for <htslib/sam.h>, TEST INFRASTRUCTURE ONLY.
 *
 * htslib is not installed in this image and cannot be fetched (no network).
 * The reference (smithlabcode/abismal) uses htslib purely for I/O: building a
 * bam1_t and printing it as one SAM text line.  This header re-implements
 * just that surface so the UNMODIFIED reference sources compile into
 * oracle/_ref/abismal (see oracle/Makefile).  It is written from the SAM
 * specification and htslib's documented behaviour, not copied from htslib.
 * BAM ("wb") output is refused.
 */
#ifndef ABISMAL_B200_HTS_STANDIN_SAM_H
#define ABISMAL_B200_HTS_STANDIN_SAM_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef int64_t hts_pos_t;

typedef struct kstring_t {
  size_t l, m;
  char *s;
} kstring_t;

enum htsFormatCategory { unknown_category, sequence_data, variant_data, index_file, region_list };
enum htsExactFormat { unknown_format, binary_format, text_format, sam, bam, bai, cram };
typedef struct htsFormat {
  enum htsFormatCategory category;
  enum htsExactFormat format;
} htsFormat;

struct sam_hdr_t {
  std::string text;
  std::vector<std::string> sq_names;
};
typedef sam_hdr_t bam_hdr_t;

struct htsFile {
  FILE *fp;
  htsFormat fmt;
};
typedef htsFile samFile;

struct bam1_core_t {
  hts_pos_t pos;
  int32_t tid;
  uint8_t mapq;
  uint16_t flag;
  int32_t mtid;
  hts_pos_t mpos;
  hts_pos_t isize;
};

struct bam1_t {
  bam1_core_t core;
  std::string qname;
  std::vector<uint32_t> cigar;
  std::string seq;
  std::string aux;  // already in SAM text form, each field preceded by '\t'
};

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CIGAR_STR "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)

static inline bam1_t *bam_init1() { return new bam1_t(); }
static inline void bam_destroy1(bam1_t *b) { delete b; }
static inline bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src) {
  if (src) *dst = *src;
  return dst;
}

/* SEQ goes through the 4-bit BAM alphabet in htslib (encode in bam_set1,
 * decode in sam_format1), which upper-cases and maps unknown symbols to N. */
static inline char hts_standin_norm_base(unsigned char c) {
  static const char dec[] = "=ACMGRSVTWYHKDBN";
  unsigned code = 15;
  switch (c) {
    case '=': code = 0; break;
    case 'A': case 'a': case '0': code = 1; break;
    case 'C': case 'c': case '1': code = 2; break;
    case 'M': case 'm': code = 3; break;
    case 'G': case 'g': case '2': code = 4; break;
    case 'R': case 'r': code = 5; break;
    case 'S': case 's': code = 6; break;
    case 'V': case 'v': code = 7; break;
    case 'T': case 't': case '3': code = 8; break;
    case 'W': case 'w': code = 9; break;
    case 'Y': case 'y': code = 10; break;
    case 'H': case 'h': code = 11; break;
    case 'K': case 'k': code = 12; break;
    case 'D': case 'd': code = 13; break;
    case 'B': case 'b': code = 14; break;
    default: code = 15;
  }
  return dec[code];
}

static inline int bam_set1(bam1_t *b, size_t l_qname, const char *qname, uint16_t flag,
                           int32_t tid, hts_pos_t pos, uint8_t mapq, size_t n_cigar,
                           const uint32_t *cigar, int32_t mtid, hts_pos_t mpos,
                           hts_pos_t isize, size_t l_seq, const char *seq,
                           const char * /*qual*/, size_t /*l_aux*/) {
  if (l_qname > 254) return -1;
  hts_pos_t qlen = 0;
  for (size_t i = 0; i < n_cigar; ++i)
    if (bam_cigar_type(bam_cigar_op(cigar[i])) & 1) qlen += bam_cigar_oplen(cigar[i]);
  if (n_cigar > 0 && l_seq > 0 && static_cast<size_t>(qlen) != l_seq) return -1;
  b->core.pos = pos;
  b->core.tid = tid;
  b->core.mapq = mapq;
  b->core.flag = flag;
  b->core.mtid = mtid;
  b->core.mpos = mpos;
  b->core.isize = isize;
  b->qname.assign(qname, l_qname);
  b->cigar.assign(cigar, cigar + n_cigar);
  b->seq.resize(l_seq);
  for (size_t i = 0; i < l_seq; ++i)
    b->seq[i] = hts_standin_norm_base(static_cast<unsigned char>(seq[i]));
  b->aux.clear();
  return 0;
}

static inline int bam_aux_update_int(bam1_t *b, const char tag[2], int64_t val) {
  b->aux += '\t';
  b->aux.append(tag, 2);
  b->aux += ":i:" + std::to_string(val);
  return 0;
}

static inline int bam_aux_append(bam1_t *b, const char tag[2], char type, int len,
                                 const uint8_t *data) {
  if (type != 'A' || len != 1) return -1;
  b->aux += '\t';
  b->aux.append(tag, 2);
  b->aux += ":A:";
  b->aux += static_cast<char>(data[0]);
  return 0;
}

static inline sam_hdr_t *sam_hdr_init() { return new sam_hdr_t(); }
static inline void bam_hdr_destroy(sam_hdr_t *h) { delete h; }
static inline sam_hdr_t *bam_hdr_dup(const sam_hdr_t *h) { return h ? new sam_hdr_t(*h) : nullptr; }

static inline int sam_hdr_add_lines(sam_hdr_t *h, const char *lines, size_t len) {
  if (len == 0) len = std::strlen(lines);
  const std::string txt(lines, len);
  h->text += txt;
  size_t st = 0;
  while (st < txt.size()) {
    size_t en = txt.find('\n', st);
    if (en == std::string::npos) en = txt.size();
    const std::string ln = txt.substr(st, en - st);
    if (ln.compare(0, 3, "@SQ") == 0) {
      size_t p = ln.find("\tSN:");
      if (p != std::string::npos) {
        p += 4;
        size_t q = ln.find('\t', p);
        h->sq_names.push_back(ln.substr(p, q == std::string::npos ? q : q - p));
      }
    }
    st = en + 1;
  }
  return 0;
}

static inline htsFile *hts_open(const char *fn, const char *mode) {
  if (std::strchr(mode, 'b')) {
    std::fprintf(stderr, "[htslib stand-in] BAM output is not supported\n");
    return nullptr;
  }
  FILE *fp = (std::strcmp(fn, "-") == 0)
               ? (mode[0] == 'r' ? stdin : stdout)
               : std::fopen(fn, mode[0] == 'r' ? "r" : "w");
  if (!fp) return nullptr;
  htsFile *f = new htsFile();
  f->fp = fp;
  f->fmt.category = sequence_data;
  f->fmt.format = sam;
  return f;
}
static inline int hts_close(htsFile *f) {
  int r = 0;
  if (f->fp && f->fp != stdout && f->fp != stdin) r = std::fclose(f->fp);
  else if (f->fp) std::fflush(f->fp);
  delete f;
  return r;
}
static inline const htsFormat *hts_get_format(htsFile *f) { return &f->fmt; }

static inline int sam_hdr_write(htsFile *f, const sam_hdr_t *h) {
  return std::fwrite(h->text.data(), 1, h->text.size(), f->fp) == h->text.size() ? 0 : -1;
}

static inline int sam_write1(htsFile *f, const sam_hdr_t *h, const bam1_t *b) {
  std::string o;
  o.reserve(512);
  const bam1_core_t &c = b->core;
  const auto ref_name = [&](int32_t tid) -> std::string {
    return (tid >= 0 && static_cast<size_t>(tid) < h->sq_names.size()) ? h->sq_names[tid] : "*";
  };
  o += b->qname;
  o += '\t';
  o += std::to_string(c.flag);
  o += '\t';
  o += ref_name(c.tid);
  o += '\t';
  o += std::to_string(c.pos + 1);
  o += '\t';
  o += std::to_string(static_cast<unsigned>(c.mapq));
  o += '\t';
  if (b->cigar.empty()) o += '*';
  else
    for (uint32_t x : b->cigar) {
      o += std::to_string(bam_cigar_oplen(x));
      o += BAM_CIGAR_STR[bam_cigar_op(x)];
    }
  o += '\t';
  if (c.mtid < 0) o += '*';
  else if (c.mtid == c.tid) o += '=';
  else o += ref_name(c.mtid);
  o += '\t';
  o += std::to_string(c.mpos + 1);
  o += '\t';
  o += std::to_string(c.isize);
  o += '\t';
  if (b->seq.empty()) o += '*';
  else o += b->seq;
  o += "\t*";
  o += b->aux;
  o += '\n';
  return std::fwrite(o.data(), 1, o.size(), f->fp) == o.size() ? static_cast<int>(o.size()) : -1;
}

/* declared only: never called by map/idx/sim */
int sam_read1(htsFile *, sam_hdr_t *, bam1_t *);
sam_hdr_t *sam_hdr_read(htsFile *);
int sam_hdr_add_line(sam_hdr_t *, const char *, ...);
const char *sam_hdr_str(sam_hdr_t *);

#endif
