#include "read_loader.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace ab2 {

uint32_t ReadBatch::max_read_len() const {
  uint32_t m = 0;
  for (uint32_t i = 0; i + 1 < seq_off.size(); ++i) m = std::max(m, seq_off[i + 1] - seq_off[i]);
  return m;
}

void ReadBatch::clear() {
  seq.clear();
  names.clear();
  seq_off.assign(1, 0);
  name_off.assign(1, 0);
}

FastqReader::FastqReader(const std::string &filename) : filename_(filename) {
  gzFile f = gzopen(filename.c_str(), "rb");
  if (f) gzbuffer(f, 1u << 20);
  file_ = f;
  buf_.resize(4u << 20);
  if (!f) eof_ = true;
}

FastqReader::~FastqReader() {
  if (file_) gzclose(static_cast<gzFile>(file_));
}

bool FastqReader::fill() {
  if (src_eof_ || !file_) return false;
  const int n = gzread(static_cast<gzFile>(file_), buf_.data(), static_cast<unsigned>(buf_.size()));
  beg_ = 0;
  end_ = n > 0 ? static_cast<size_t>(n) : 0;
  if (n <= 0) {
    src_eof_ = true;
    return false;
  }
  return true;
}

// One line without its terminator; a '\r' before '\n' is dropped as htslib's
// bgzf_getline does.  Returns false only when nothing at all could be read.
bool FastqReader::getline(const char *&line, size_t &len) {
  carry_.clear();
  bool got_any = false;
  for (;;) {
    if (beg_ == end_ && !fill()) break;
    got_any = true;
    const char *p = buf_.data() + beg_;
    const char *nl = static_cast<const char *>(std::memchr(p, '\n', end_ - beg_));
    if (nl) {
      const size_t n = static_cast<size_t>(nl - p);
      beg_ += n + 1;
      if (carry_.empty()) {
        line = p;
        len = n;
      }
      else {
        carry_.append(p, n);
        line = carry_.data();
        len = carry_.size();
      }
      if (len > 0 && line[len - 1] == '\r') --len;
      return true;
    }
    carry_.append(p, end_ - beg_);
    beg_ = end_;
  }
  if (!got_any) return false;
  line = carry_.data();
  len = carry_.size();
  if (len > 0 && line[len - 1] == '\r') --len;
  return true;
}

void FastqReader::load_reads(ReadBatch &out, size_t max_reads) {
  out.clear();
  // one allocation up front (page-locked memory is slow to allocate); also keeps seq.data() non-null for
  // batches whose reads are all empty
  if (out.seq.capacity() == 0) out.seq.reserve(std::max<size_t>(64, std::min<size_t>(max_reads, 1u << 22) * 160));
  if (eof_) return;
  size_t line_count = 0;
  const size_t num_lines_to_read = 4 * max_reads;
  const char *line = nullptr;
  size_t len = 0;
  size_t name_beg = 0, name_len = 0;
  std::string name;
  while (line_count < num_lines_to_read) {
    if (!getline(line, len)) {
      eof_ = true;
      break;
    }
    if (line_count % 4 == 0) {
      if (len == 0)
        throw std::runtime_error("file " + filename_ + " contains an empty read name at line " +
                                 std::to_string(cur_line_));
      // name = line.substr(1, line.find_first_of(" \t") - 1)
      size_t ws = 0;
      while (ws < len && line[ws] != ' ' && line[ws] != '\t') ++ws;
      const size_t count = (ws == len) ? std::string::npos : ws - 1;  // size_t arithmetic as in the reference
      name_beg = std::min<size_t>(1, len);
      name_len = std::min(count, len - name_beg);
      name.assign(line + name_beg, name_len);
    }
    else if (line_count % 4 == 1) {
      if (len >= padding_size)
        throw std::runtime_error("found a read of size " + std::to_string(len) +
                                 ", which is too long. Maximum allowed read size = " +
                                 std::to_string(padding_size));
      size_t non_n = 0;
      for (size_t i = 0; i < len; ++i) non_n += line[i] != 'N';
      size_t b = 0, e = 0;
      if (non_n >= min_read_length) {
        e = len;
        while (e > 0 && line[e - 1] == 'N') --e;  // remove Ns from 3'
        while (b < e && line[b] != 'A' && line[b] != 'C' && line[b] != 'G' && line[b] != 'T') ++b;
        if (b == e)  // reference: substr(npos) throws std::out_of_range
          throw std::runtime_error("basic_string::substr: __pos (which is 18446744073709551615) > "
                                   "this->size() (which is " + std::to_string(e) + ")");
      }
      out.seq.insert(out.seq.end(), line + b, line + e);
      out.seq_off.push_back(static_cast<uint32_t>(out.seq.size()));
      out.names.insert(out.names.end(), name.begin(), name.end());
      out.name_off.push_back(static_cast<uint32_t>(out.names.size()));
    }
    ++line_count;
    ++cur_line_;
  }
}

}  // namespace ab2
