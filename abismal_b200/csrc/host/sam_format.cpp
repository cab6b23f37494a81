#include "sam_format.hpp"

#include <cstdio>
#include <cstring>
#include <sstream>

namespace ab2 {

namespace {

inline void put_uint(std::string &o, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) o += tmp[--n];
}
// SEQ passes through BAM's 4-bit alphabet in htslib (bam_set1 encodes,
// sam_format1 decodes): upper-cases, keeps IUPAC codes, anything else -> N.
struct Nt16 {
  char fwd[256];  // normalised base
  char rc[256];   // normalised revcomp base (revcomp_inplace: A<->T, C<->G, else N)
  unsigned char code[256];  // seq_nt16_table
  Nt16() {
    static const char dec[] = "=ACMGRSVTWYHKDBN";
    for (int c = 0; c < 256; ++c) {
      int code = 15;
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
      fwd[c] = dec[code];
      this->code[c] = static_cast<unsigned char>(code);
      rc[c] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
  }
};
const Nt16 nt16;

// raw-pointer writers for the SAM text (one reservation per record instead of a capacity check per byte)
inline char *wr_uint(char *p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char *wr_int(char *p, int64_t v) {
  if (v < 0) {
    *p++ = '-';
    return wr_uint(p, static_cast<uint64_t>(-v));
  }
  return wr_uint(p, static_cast<uint64_t>(v));
}
inline char *wr_str(char *p, const char *s, size_t n) {
  std::memcpy(p, s, n);
  return p + n;
}
inline const std::string *rname_of(const ChromLookup &cl, int32_t tid) {
  const int64_t n_real = static_cast<int64_t>(cl.names.size()) - 2;
  return (tid >= 0 && tid < n_real) ? &cl.names[static_cast<size_t>(tid) + 1] : nullptr;
}

void put_sam_record(std::string &o, const ChromLookup &cl, const ReadView &r, bool revcomp, uint16_t flag,
                    int32_t tid, uint32_t pos, int32_t mtid, int64_t mpos, int64_t isize, int nm, char cv) {
  const std::string *rn = rname_of(cl, tid);
  const std::string *mn = (mtid >= 0 && mtid != tid) ? rname_of(cl, mtid) : nullptr;
  const size_t at = o.size();
  const size_t max_len = static_cast<size_t>(r.name_len) + r.seq_len + 12u * r.n_cigar + (rn ? rn->size() : 1) +
                         (mn ? mn->size() : 1) + 160;
  o.resize(at + max_len);
  char *p = &o[at];
  p = wr_str(p, r.name, r.name_len);
  *p++ = '\t';
  p = wr_uint(p, flag);
  *p++ = '\t';
  if (rn) p = wr_str(p, rn->data(), rn->size());
  else *p++ = '*';
  *p++ = '\t';
  p = wr_uint(p, static_cast<uint64_t>(pos) + 1);
  p = wr_str(p, "\t255\t", 5);
  if (r.n_cigar == 0) *p++ = '*';
  else {
    static const char ops[] = "MIDNSHP=XB??????";
    for (uint32_t i = 0; i < r.n_cigar; ++i) {
      p = wr_uint(p, r.cigar[i] >> 4);
      *p++ = ops[r.cigar[i] & 15u];
    }
  }
  *p++ = '\t';
  if (mtid < 0) *p++ = '*';
  else if (mtid == tid) *p++ = '=';
  else if (mn) p = wr_str(p, mn->data(), mn->size());
  else *p++ = '*';
  *p++ = '\t';
  p = wr_int(p, mpos + 1);
  *p++ = '\t';
  p = wr_int(p, isize);
  *p++ = '\t';
  if (r.seq_len == 0) *p++ = '*';
  else if (!revcomp) {
    for (uint32_t i = 0; i < r.seq_len; ++i) p[i] = nt16.fwd[static_cast<unsigned char>(r.seq[i])];
    p += r.seq_len;
  }
  else {
    const char *s = r.seq + r.seq_len - 1;
    for (uint32_t i = 0; i < r.seq_len; ++i) p[i] = nt16.rc[static_cast<unsigned char>(s[-static_cast<int64_t>(i)])];
    p += r.seq_len;
  }
  p = wr_str(p, "\t*\tNM:i:", 8);
  p = wr_int(p, nm);
  p = wr_str(p, "\tCV:A:", 6);
  *p++ = cv;
  *p++ = '\n';
  o.resize(static_cast<size_t>(p - o.data()));
}

inline void le16(std::string &o, uint32_t v) {
  o += static_cast<char>(v & 0xff);
  o += static_cast<char>((v >> 8) & 0xff);
}
inline void le32(std::string &o, uint32_t v) {
  le16(o, v & 0xffff);
  le16(o, v >> 16);
}

// UCSC binning scheme (SAM spec 5.3), the value bam_set1 stores in core.bin
inline uint32_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return static_cast<uint32_t>(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return static_cast<uint32_t>(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return static_cast<uint32_t>(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return static_cast<uint32_t>(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return static_cast<uint32_t>(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

// The record bam_set1 + bam_aux_update_int("NM") + bam_aux_append("CV", 'A') build
// (abismal.cpp:511-542, :710-770), in the on-disk BAM layout (SAM spec 4.2).
void put_bam_record(std::string &o, const ReadView &r, bool revcomp, uint16_t flag, int32_t tid, uint32_t pos,
                    int32_t mtid, int64_t mpos, int64_t isize, int nm, char cv) {
  o.clear();
  le32(o, 0);  // block_size, patched below
  le32(o, static_cast<uint32_t>(tid));
  le32(o, pos);
  uint32_t rlen = cigar_rseq_ops(r.cigar, r.n_cigar);
  if (rlen == 0) rlen = 1;
  const uint32_t l_name = r.name_len + 1;
  o += static_cast<char>(l_name & 0xff);
  o += static_cast<char>(255);  // MAPQ
  le16(o, reg2bin(pos, static_cast<int64_t>(pos) + rlen));
  le16(o, r.n_cigar);
  le16(o, flag);
  le32(o, r.seq_len);
  le32(o, static_cast<uint32_t>(mtid));
  le32(o, static_cast<uint32_t>(static_cast<int32_t>(mpos)));
  le32(o, static_cast<uint32_t>(static_cast<int32_t>(isize)));
  o.append(r.name, r.name_len);
  o += '\0';
  for (uint32_t i = 0; i < r.n_cigar; ++i) le32(o, r.cigar[i]);
  {
    const size_t at = o.size();
    o.resize(at + (r.seq_len + 1) / 2, '\0');
    unsigned char *d = reinterpret_cast<unsigned char *>(&o[at]);
    for (uint32_t i = 0; i < r.seq_len; ++i) {
      const unsigned char ch = revcomp ? static_cast<unsigned char>(nt16.rc[static_cast<unsigned char>(r.seq[r.seq_len - 1 - i])])
                                       : static_cast<unsigned char>(r.seq[i]);
      const unsigned code = nt16.code[ch];
      d[i >> 1] |= static_cast<unsigned char>((i & 1) ? code : code << 4);
    }
  }
  o.append(r.seq_len, static_cast<char>(0xff));  // no qualities
  // NM: the smallest integer type that holds the value (bam_aux_update_int)
  o += "NM";
  if (nm >= 0) {
    if (nm <= 0xff) {
      o += 'C';
      o += static_cast<char>(nm);
    }
    else if (nm <= 0xffff) {
      o += 'S';
      le16(o, static_cast<uint32_t>(nm));
    }
    else {
      o += 'I';
      le32(o, static_cast<uint32_t>(nm));
    }
  }
  else if (nm >= -128) {
    o += 'c';
    o += static_cast<char>(nm);
  }
  else if (nm >= -32768) {
    o += 's';
    le16(o, static_cast<uint32_t>(nm) & 0xffffu);
  }
  else {
    o += 'i';
    le32(o, static_cast<uint32_t>(nm));
  }
  o += "CVA";
  o += cv;
  const uint32_t block_size = static_cast<uint32_t>(o.size() - 4);
  o[0] = static_cast<char>(block_size & 0xff);
  o[1] = static_cast<char>((block_size >> 8) & 0xff);
  o[2] = static_cast<char>((block_size >> 16) & 0xff);
  o[3] = static_cast<char>((block_size >> 24) & 0xff);
}

void put_record(Emitter &em, const ChromLookup &cl, const ReadView &r, bool revcomp, uint16_t flag, int32_t tid,
                uint32_t pos, int32_t mtid, int64_t mpos, int64_t isize, int nm, char cv) {
  if (em.bam) {
    put_bam_record(em.scratch, r, revcomp, flag, tid, pos, mtid, mpos, isize, nm, cv);
    em.bam->add(em.scratch.data(), em.scratch.size());
  }
  else put_sam_record(*em.sam, cl, r, revcomp, flag, tid, pos, mtid, mpos, isize, nm, cv);
}

bool chrom_and_posn(const ChromLookup &cl, const ReadView &r, uint32_t p, uint32_t &r_p, uint32_t &r_e,
                    int32_t &r_chr) {  // abismal.cpp:464-473
  const uint32_t ref_ops = cigar_rseq_ops(r.cigar, r.n_cigar);
  if (!cl.chrom_idx_and_offset(p, ref_ops, r_chr, r_p)) return false;
  r_e = r_p + ref_ops;
  return true;
}

std::string fmt_double(double x) {  // default ostream formatting (%g, precision 6)
  std::ostringstream oss;
  oss << x;
  return oss.str();
}

}  // namespace

uint32_t cigar_rseq_ops(const uint32_t *cigar, uint32_t n) {
  uint32_t t = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t op = cigar[i] & 15u;
    if ((0x3C1A7u >> (op << 1)) & 2u) t += cigar[i] >> 4;
  }
  return t;
}

std::string make_sam_header(const ChromLookup &cl, int argc, char *const argv[], const char *version) {
  std::string out = "@HD\tVN:1.0\n";
  for (size_t i = 1; i + 1 < cl.names.size(); ++i) {
    out += "@SQ\tSN:" + cl.names[i] + "\tLN:";
    put_uint(out, cl.starts[i + 1] - cl.starts[i]);
    out += '\n';
  }
  out += "@PG\tID:ABISMAL\tVN:";
  out += version;
  out += "\tCL:\"";
  for (int i = 0; i < argc; ++i) {
    out += argv[i];
    out += ' ';
  }
  out += "\"\n";
  return out;
}

MapType format_se(bool allow_ambig, const abg_hit &res, const ChromLookup &cl, const ReadView &r,
                  Emitter &out) {
  const bool ambig = hit_ambig(res);
  const bool valid = !hit_empty(res);
  if (!allow_ambig && ambig) return map_ambig;
  uint32_t ref_s = 0, ref_e = 0;
  int32_t chrom_idx = 0;
  if (!valid || !chrom_and_posn(cl, r, res.pos, ref_s, ref_e, chrom_idx)) return map_unmapped;
  uint16_t flag = 0;
  if (hit_rc(res)) flag |= 16;
  if (allow_ambig && ambig) flag |= 256;
  put_record(out, cl, r, hit_rc(res), flag, chrom_idx - 1, ref_s, -1, -1, 0, res.diffs,
             hit_a_rich(res) ? 'A' : 'T');
  return ambig ? map_ambig : map_unique;
}

MapType format_pe(bool allow_ambig, const abg_hit &p1, const abg_hit &p2, const ChromLookup &cl,
                  const ReadView &r1, const ReadView &r2, Emitter &out) {
  if (hit_empty(p1)) return map_unmapped;
  const bool ambig = hit_ambig(p1);
  if (!allow_ambig && ambig) return map_ambig;
  int32_t chr1 = 0, chr2 = 0;
  uint32_t r_s1 = 0, r_e1 = 0, r_s2 = 0, r_e2 = 0;
  if (!chrom_and_posn(cl, r1, p1.pos, r_s1, r_e1, chr1) || !chrom_and_posn(cl, r2, p2.pos, r_s2, r_e2, chr2) ||
      chr1 != chr2)
    return map_unmapped;
  const bool rc = hit_rc(p1);
  const int isize = rc ? (static_cast<int>(r_s1) - static_cast<int>(r_e2))
                       : (static_cast<int>(r_e2) - static_cast<int>(r_s1));
  uint16_t flag1 = 1 | 2, flag2 = 1 | 2;
  if (hit_rc(p1)) {
    flag1 |= 16;
    flag2 |= 32;
  }
  if (hit_rc(p2)) {
    flag2 |= 16;
    flag1 |= 32;
  }
  if (allow_ambig && ambig) {
    flag1 |= 256;
    flag2 |= 256;
  }
  flag1 |= 64;
  flag2 |= 128;
  put_record(out, cl, r1, hit_rc(p1), flag1, chr1 - 1, r_s1, chr2 - 1, r_s2, isize, p1.diffs,
             hit_a_rich(p1) ? 'A' : 'T');
  put_record(out, cl, r2, hit_rc(p2), flag2, chr2 - 1, r_s2, chr1 - 1, r_s1, -isize, p2.diffs,
             hit_a_rich(p2) ? 'A' : 'T');
  return ambig ? map_ambig : map_unique;
}

void select_output(bool allow_ambig, const ChromLookup &cl, const ReadView &r1, const ReadView &r2,
                   abg_hit &pe1, abg_hit &pe2, abg_hit &se1, abg_hit &se2, Emitter &out) {
  const MapType pe_map_type = format_pe(allow_ambig, pe1, pe2, cl, r1, r2, out);
  const bool should_report = !hit_empty(pe1) && (allow_ambig || !hit_ambig(pe1));
  if (!should_report || pe_map_type == map_unmapped) {
    if (pe_map_type == map_unmapped) {
      hit_reset(pe1);
      hit_reset(pe2);
    }
    if (format_se(allow_ambig, se1, cl, r1, out) == map_unmapped) hit_reset(se1);
    if (format_se(allow_ambig, se2, cl, r2, out) == map_unmapped) hit_reset(se2);
  }
}

void SeStats::update(bool allow_ambig, const ReadView &r, const abg_hit &s) {
  ++total_reads;
  const bool valid = !hit_empty(s), ambig = hit_ambig(s);
  reads_mapped_unique += (valid && !ambig);
  reads_mapped_ambiguous += (valid && ambig);
  reads_skipped += (r.seq_len == 0);
  if (valid && (!ambig || allow_ambig)) {
    edit_distance += static_cast<uint64_t>(static_cast<int64_t>(s.diffs));
    total_bases += cigar_rseq_ops(r.cigar, r.n_cigar);
  }
}

void SeStats::update(const ReadView &r, const abg_hit &s) {
  ++total_reads;
  const bool valid = !hit_empty(s), ambig = hit_ambig(s);
  reads_mapped_unique += (valid && !ambig);
  reads_mapped_ambiguous += (valid && ambig);
  reads_skipped += (r.seq_len == 0);
  if (valid && !ambig) {
    edit_distance += static_cast<uint64_t>(static_cast<int64_t>(s.diffs));
    total_bases += cigar_rseq_ops(r.cigar, r.n_cigar);
  }
}

void SeStats::add(const SeStats &o) {
  total_reads += o.total_reads;
  reads_mapped_unique += o.reads_mapped_unique;
  reads_mapped_ambiguous += o.reads_mapped_ambiguous;
  reads_skipped += o.reads_skipped;
  edit_distance += o.edit_distance;
  total_bases += o.total_bases;
}

std::string SeStats::tostring(const std::string &label, size_t n_tabs) const {
  // the reference counts in 32-bit atomics (abismal.cpp:870-885)
  const uint32_t total = static_cast<uint32_t>(total_reads);
  const uint32_t uniq = static_cast<uint32_t>(reads_mapped_unique);
  const uint32_t amb = static_cast<uint32_t>(reads_mapped_ambiguous);
  const uint32_t skipped = static_cast<uint32_t>(reads_skipped);
  const uint32_t mapped = uniq + amb;
  const uint32_t unmapped = total - mapped;
  const auto frac = [&](uint32_t x) { return total > 0 ? static_cast<double>(x) / total : 0.0; };
  const double err = total_bases > 0 ? static_cast<double>(edit_distance) / total_bases : 0.0;
  static const char *tab = "    ";
  std::string t;
  for (size_t i = 0; i < n_tabs; ++i) t += tab;
  std::ostringstream oss;
  oss << t << label << ":\n";
  t += tab;
  oss << t << "total_reads: " << total << '\n'
      << t << "mapped:\n"
      << t << "    num_mapped: " << mapped << '\n'
      << t << "    num_unique: " << uniq << '\n'
      << t << "    num_ambiguous: " << amb << '\n'
      << t << "    percent_mapped: " << fmt_double(frac(mapped) * 100.0) << '\n'
      << t << "    percent_unique: " << fmt_double(frac(uniq) * 100.0) << '\n'
      << t << "    percent_ambiguous: " << fmt_double(frac(amb) * 100.0) << '\n'
      << t << "    unique_error:\n"
      << t << "        edits: " << edit_distance << '\n'
      << t << "        total_bases: " << total_bases << '\n'
      << t << "        error_rate: " << fmt_double(err) << '\n'
      << t << "num_unmapped: " << unmapped << '\n'
      << t << "num_skipped: " << skipped << '\n'
      << t << "percent_unmapped: " << fmt_double(frac(unmapped) * 100.0) << '\n'
      << t << "percent_skipped: " << fmt_double(frac(skipped) * 100.0) << '\n';
  return oss.str();
}

std::string SeStats::tojson() const {  // nlohmann compact dump, keys sorted
  std::ostringstream oss;
  oss << "{\"edit_distance\":" << edit_distance
      << ",\"reads_mapped_ambiguous\":" << static_cast<uint32_t>(reads_mapped_ambiguous)
      << ",\"reads_mapped_unique\":" << static_cast<uint32_t>(reads_mapped_unique)
      << ",\"reads_skipped\":" << static_cast<uint32_t>(reads_skipped) << ",\"total_bases\":" << total_bases
      << ",\"total_reads\":" << static_cast<uint32_t>(total_reads) << "}";
  return oss.str();
}

void PeStats::update(bool allow_ambig, const ReadView &r1, const ReadView &r2, const abg_hit &pe1,
                     const abg_hit &pe2, const abg_hit &s1, const abg_hit &s2) {
  ++read_pair_stats.total_reads;
  const bool valid = !hit_empty(pe1), ambig = hit_ambig(pe1);
  read_pair_stats.reads_mapped_unique += (valid && !ambig);
  read_pair_stats.reads_mapped_ambiguous += (valid && ambig);
  read_pair_stats.reads_skipped += (r1.seq_len == 0 || r2.seq_len == 0);
  if (valid && (allow_ambig || !ambig)) {
    read_pair_stats.edit_distance +=
      static_cast<uint64_t>(static_cast<int64_t>(static_cast<int>(pe1.diffs) + static_cast<int>(pe2.diffs)));
    read_pair_stats.total_bases +=
      cigar_rseq_ops(r1.cigar, r1.n_cigar) + cigar_rseq_ops(r2.cigar, r2.n_cigar);
  }
  else {
    end1_stats.update(r1, s1);
    end2_stats.update(r2, s2);
  }
}

void PeStats::add(const PeStats &o) {
  read_pair_stats.add(o.read_pair_stats);
  end1_stats.add(o.end1_stats);
  end2_stats.add(o.end2_stats);
}

std::string PeStats::tostring(bool allow_ambig) const {
  std::string s = read_pair_stats.tostring("pairs");
  if (!allow_ambig) {
    s += end1_stats.tostring("read1");
    s += end2_stats.tostring("read2");
  }
  return s;
}

std::string PeStats::tojson() const {
  return "{\"end1_stats\":" + end1_stats.tojson() + ",\"end2_stats\":" + end2_stats.tojson() +
         ",\"read_pair_stats\":" + read_pair_stats.tojson() + "}";
}

}  // namespace ab2
