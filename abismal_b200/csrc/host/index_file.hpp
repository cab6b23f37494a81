// AbismalIndex file reader (host side of the drop-in boundary).
// On-disk layout follows the reference writer/reader
// (src/AbismalIndex.cpp:1037-1072 / :1082-1146, ChromLookup :1225-1258).
#ifndef ABISMAL_B200_INDEX_FILE_HPP
#define ABISMAL_B200_INDEX_FILE_HPP

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "abismal_b200.h"

namespace ab2 {

// ChromLookup (src/AbismalIndex.hpp:101-143): names include pad_start/pad_end.
struct ChromLookup {
  std::vector<std::string> names;
  std::vector<uint32_t> starts;  // names.size() + 1 entries

  // get_chrom_idx_and_offset (src/AbismalIndex.cpp:1305-1320)
  bool chrom_idx_and_offset(uint32_t pos, uint32_t ref_len, int32_t &chrom_idx, uint32_t &offset) const;
  uint32_t genome_size() const { return starts.empty() ? 0 : starts.back(); }
};

struct IndexFile {
  ChromLookup cl;
  uint32_t max_candidates = 100;
  uint32_t window_size = 20;  // seed::window_size of the file: 20, or 12 (reference configured with --enable-short)
  uint64_t counter_size = 0, counter_size_three = 0, index_size = 0, index_size_three = 0;
  std::vector<uint64_t> genome;
  std::vector<uint32_t> counter, counter_t, counter_a, index, index_t, index_a;

  IndexFile() = default;
  IndexFile(const IndexFile &) = delete;
  IndexFile &operator=(const IndexFile &) = delete;
  ~IndexFile();

  // throws std::runtime_error with the reference's messages.  map_file: the arrays stay in one buffer filled
  // by several threads (pread) instead of the vectors -- they are only read once, by the upload to HBM;
  // view() then points into that buffer.
  void read(const std::string &path, bool map_file = false);
  abg_index_view view() const;

private:
  std::unique_ptr<unsigned char[]> bulk_;
  size_t bulk_len_ = 0;
  abg_index_view mapped_{};
};

}  // namespace ab2
#endif
