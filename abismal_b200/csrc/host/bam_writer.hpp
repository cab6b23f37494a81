// BAM output (`-B`): BGZF blocks over zlib and the BAM header, replacing the
// htslib calls behind the reference's bamxx::bam_out (src/bamxx/bamxx.hpp:124-146,
// src/abismal.cpp:2454, :2291-2292).  Records are encoded in sam_format.cpp;
// this file frames them.
//
// Layout choices follow htslib so that a reader sees the same stream shape:
// 0xff00 bytes of payload per block at most, a record never straddles two
// blocks unless it is larger than a block (bgzf_flush_try), the header sits in
// its own block(s), and the file ends with the 28-byte EOF marker block.
#ifndef ABISMAL_B200_BAM_WRITER_HPP
#define ABISMAL_B200_BAM_WRITER_HPP

#include <cstddef>
#include <cstdint>
#include <string>

#include "index_file.hpp"

namespace ab2 {

constexpr size_t kBgzfBlockPayload = 0xff00;

// Appends one BGZF block holding data[0, n) (n <= kBgzfBlockPayload) to `out`.
void bgzf_append_block(const char *data, size_t n, int level, std::string &out);
// Appends data as a sequence of full blocks.
void bgzf_append(const char *data, size_t n, int level, std::string &out);
const std::string &bgzf_eof_marker();

// "BAM\1", header text, reference dictionary (real chromosomes only) -- uncompressed bytes.
std::string make_bam_header(const ChromLookup &cl, const std::string &sam_header_text);

// Packs BAM records of one formatter slice into BGZF blocks.
class BgzfRecordPacker {
public:
  BgzfRecordPacker(std::string &out, int level) : out_(out), level_(level) { block_.reserve(kBgzfBlockPayload); }
  // `rec` is one complete record (block_size field included)
  void add(const char *rec, size_t n);
  void finish();

private:
  std::string &out_;
  std::string block_;
  int level_;
};

}  // namespace ab2
#endif
