// SAM text + mapping statistics, byte-compatible with the reference's
// format_se / format_pe / select_output (src/abismal.cpp:481-545, :648-773,
// :1073-1088), its header (:2265-2293) and stats (:865-1071), with htslib's
// sam_format1 text layout for the fields abismal sets.
#ifndef ABISMAL_B200_SAM_FORMAT_HPP
#define ABISMAL_B200_SAM_FORMAT_HPP

#include <cstdint>
#include <string>

#include "abismal_b200.h"
#include "bam_writer.hpp"
#include "index_file.hpp"

namespace ab2 {

enum MapType : uint8_t { map_unmapped, map_unique, map_ambig };

struct ReadView {
  const char *name;
  uint32_t name_len;
  const char *seq;
  uint32_t seq_len;
  const uint32_t *cigar;
  uint32_t n_cigar;
};

// Where records go: SAM text appended to *sam, or BAM records handed to *bam.
struct Emitter {
  std::string *sam = nullptr;
  BgzfRecordPacker *bam = nullptr;
  std::string scratch;  // one BAM record under construction
};

inline bool hit_empty(const abg_hit &h) { return h.pos == 0; }
inline bool hit_ambig(const abg_hit &h) { return (h.flags & ABG_FLAG_AMBIG) != 0; }
inline bool hit_rc(const abg_hit &h) { return (h.flags & ABG_FLAG_RC) != 0; }
inline bool hit_a_rich(const abg_hit &h) { return (h.flags & ABG_FLAG_A_RICH) != 0; }
inline void hit_reset(abg_hit &h) {
  h.pos = 0;
  h.diffs = 32767;
}

uint32_t cigar_rseq_ops(const uint32_t *cigar, uint32_t n);

// argv/argc are those seen by the `map` subcommand (program name included).
std::string make_sam_header(const ChromLookup &cl, int argc, char *const argv[], const char *version);

// Emits zero or one record.
MapType format_se(bool allow_ambig, const abg_hit &res, const ChromLookup &cl, const ReadView &r,
                  Emitter &out);
// Emits zero or two records.
MapType format_pe(bool allow_ambig, const abg_hit &p1, const abg_hit &p2, const ChromLookup &cl,
                  const ReadView &r1, const ReadView &r2, Emitter &out);
// select_output: may reset pe (both ends) / se1 / se2 exactly as the reference does.
void select_output(bool allow_ambig, const ChromLookup &cl, const ReadView &r1, const ReadView &r2,
                   abg_hit &pe1, abg_hit &pe2, abg_hit &se1, abg_hit &se2, Emitter &out);

struct SeStats {  // single_end_mapping_statistics
  uint64_t total_reads = 0, reads_mapped_unique = 0, reads_mapped_ambiguous = 0, reads_skipped = 0;
  uint64_t edit_distance = 0, total_bases = 0;
  void update(bool allow_ambig, const ReadView &r, const abg_hit &s);  // :974-984
  void update(const ReadView &r, const abg_hit &s);                    // :986-996
  void add(const SeStats &o);
  std::string tostring(const std::string &label, size_t n_tabs = 0) const;
  std::string tojson() const;
};

struct PeStats {  // paired_end_mapping_statistics
  SeStats read_pair_stats, end1_stats, end2_stats;
  void update(bool allow_ambig, const ReadView &r1, const ReadView &r2, const abg_hit &pe1,
              const abg_hit &pe2, const abg_hit &s1, const abg_hit &s2);  // :1039-1057
  void add(const PeStats &o);
  std::string tostring(bool allow_ambig) const;
  std::string tojson() const;
};

}  // namespace ab2
#endif
