// FASTQ(.gz) reader reproducing ReadLoader::load_reads
// (src/abismal.cpp:150-213) for batches of any size.
#ifndef ABISMAL_B200_READ_LOADER_HPP
#define ABISMAL_B200_READ_LOADER_HPP

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "pipeline.hpp"

namespace ab2 {

// Reads of one batch in flat arrays: read i is seq[seq_off[i] .. seq_off[i+1]),
// its name is names[name_off[i] .. name_off[i+1]).
struct ReadBatch {
  pinned_vector<char> seq;  // page-locked in the GPU build: DMA'd in place by abg_map_batch
  std::vector<uint32_t> seq_off{0};
  std::vector<char> names;
  std::vector<uint32_t> name_off{0};

  uint32_t size() const { return static_cast<uint32_t>(seq_off.size() - 1); }
  uint32_t read_len(uint32_t i) const { return seq_off[i + 1] - seq_off[i]; }
  uint32_t max_read_len() const;
  void clear();
};

class FastqReader {
public:
  // key_weight + window_size - 1 (abismal.cpp:212-213): 44, or 36 for an index with window 12
  uint32_t min_read_length = 25 + 20 - 1;
  void set_window_size(uint32_t w) { min_read_length = 25 + w - 1; }
  static constexpr size_t padding_size = 32767;              // seed::padding_size

  // n_helpers > 1: whole records already in the buffer are parsed by that many threads at a time (the trim rules
  // and the copies into the batch are independent per record; results and errors are those of the serial path)
  explicit FastqReader(const std::string &filename, unsigned n_helpers = 1);
  ~FastqReader();
  FastqReader(const FastqReader &) = delete;
  FastqReader &operator=(const FastqReader &) = delete;

  bool is_open() const { return file_ != nullptr || map_ != nullptr; }
  // true until a read attempt has hit end of file (ReadLoader::operator bool)
  bool good() const { return !eof_; }
  // Appends up to max_reads reads to `out` (cleared first).  Throws
  // std::runtime_error on an empty name line or an over-long read.
  void load_reads(ReadBatch &out, size_t max_reads);
  uint64_t current_read() const { return cur_line_ / 4; }

private:
  bool getline(const char *&line, size_t &len);  // bgzf_getline semantics
  bool fill();
  bool fill_more();
  bool fast_record(ReadBatch &out);
  size_t parse_block(ReadBatch &out, size_t max_records);
  void push_read(ReadBatch &out, const char *name, size_t name_len, const char *line, size_t len);

  std::string filename_;
  std::unique_ptr<WorkerPool> pool_;
  std::vector<std::vector<uint32_t>> nl_parts_;  // newline offsets found by each helper
  std::vector<uint32_t> nl_;
  struct Span {
    uint32_t name_at, name_len, seq_at, seq_len;
  };
  std::vector<Span> spans_;
  void *file_ = nullptr;  // gzFile (compressed input, pipes)
  const char *map_ = nullptr;  // read-only mapping of a plain file
  size_t map_len_ = 0;
  const char *data_ = nullptr;  // buf_.data() or map_
  std::vector<char> buf_;
  size_t beg_ = 0, end_ = 0;
  std::string carry_;
  bool eof_ = false, src_eof_ = false;
  uint64_t cur_line_ = 0;
};

}  // namespace ab2
#endif
