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

// Keeps the unread tail and tops the buffer up (the fast path wants whole records in one piece).
bool FastqReader::fill_more() {
  if (src_eof_ || !file_) return false;
  if (beg_ > 0) {
    std::memmove(buf_.data(), buf_.data() + beg_, end_ - beg_);
    end_ -= beg_;
    beg_ = 0;
  }
  if (end_ == buf_.size()) return false;  // a record larger than the buffer: the line-by-line path handles it
  const int n = gzread(static_cast<gzFile>(file_), buf_.data() + end_, static_cast<unsigned>(buf_.size() - end_));
  if (n <= 0) {
    src_eof_ = true;
    return false;
  }
  end_ += static_cast<size_t>(n);
  return true;
}

// One read (name + sequence line) under the ReadLoader rules (abismal.cpp:164-201).
void FastqReader::push_read(ReadBatch &out, const char *name, size_t name_len, const char *line, size_t len) {
  if (len >= padding_size)
    throw std::runtime_error("found a read of size " + std::to_string(len) +
                             ", which is too long. Maximum allowed read size = " + std::to_string(padding_size));
  size_t non_n = len;
  if (std::memchr(line, 'N', len) != nullptr) {
    non_n = 0;
    for (size_t i = 0; i < len; ++i) non_n += line[i] != 'N';
  }
  size_t b = 0, e = 0;
  if (non_n >= min_read_length) {
    e = len;
    while (e > 0 && line[e - 1] == 'N') --e;  // remove Ns from 3'
    while (b < e && line[b] != 'A' && line[b] != 'C' && line[b] != 'G' && line[b] != 'T') ++b;
    if (b == e)  // reference: substr(npos) throws std::out_of_range
      throw std::runtime_error("basic_string::substr: __pos (which is 18446744073709551615) > "
                               "this->size() (which is " + std::to_string(e) + ")");
    // Only reachable with characters outside ACGTN (lower case, IUPAC), on which the reference's own
    // behaviour is undefined: the 5' trim left fewer than min_read_length bases.
    // The kernels admit no such read; it is skipped like a read with too few non-N bases.
    if (e - b < min_read_length) b = e = 0;
  }
  out.seq.append(line + b, e - b);
  out.seq_off.push_back(static_cast<uint32_t>(out.seq.size()));
  out.names.insert(out.names.end(), name, name + name_len);
  out.name_off.push_back(static_cast<uint32_t>(out.names.size()));
}

// name = line.substr(1, line.find_first_of(" \t") - 1), with the reference's size_t arithmetic
static inline void name_span(const char *line, size_t len, size_t &name_beg, size_t &name_len) {
  size_t ws = 0;
  while (ws < len && line[ws] != ' ' && line[ws] != '\t') ++ws;
  const size_t count = (ws == len) ? std::string::npos : ws - 1;
  name_beg = std::min<size_t>(1, len);
  name_len = std::min(count, len - name_beg);
}

// Fast path: the next whole 4-line record lies in the buffer.  Returns false (nothing consumed) when it does
// not, or when the record needs the careful path (empty name line).
bool FastqReader::fast_record(ReadBatch &out) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    const char *p = buf_.data() + beg_, *e = buf_.data() + end_;
    const char *nl[4];
    const char *q = p;
    int k = 0;
    for (; k < 4 && q < e; ++k) {
      nl[k] = static_cast<const char *>(std::memchr(q, '\n', static_cast<size_t>(e - q)));
      if (!nl[k]) break;
      q = nl[k] + 1;
    }
    if (k == 4) {
      size_t l0 = static_cast<size_t>(nl[0] - p);
      if (l0 > 0 && p[l0 - 1] == '\r') --l0;
      if (l0 == 0) return false;
      const char *s = nl[0] + 1;
      size_t l1 = static_cast<size_t>(nl[1] - s);
      if (l1 > 0 && s[l1 - 1] == '\r') --l1;
      size_t nb, nlen;
      name_span(p, l0, nb, nlen);
      push_read(out, p + nb, nlen, s, l1);
      beg_ = static_cast<size_t>(q - buf_.data());
      return true;
    }
    if (attempt == 1 || !fill_more()) return false;
  }
  return false;
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
    if (line_count % 4 == 0 && fast_record(out)) {
      line_count += 4;
      cur_line_ += 4;
      continue;
    }
    if (!getline(line, len)) {
      eof_ = true;
      break;
    }
    if (line_count % 4 == 0) {
      if (len == 0)
        throw std::runtime_error("file " + filename_ + " contains an empty read name at line " +
                                 std::to_string(cur_line_));
      name_span(line, len, name_beg, name_len);
      name.assign(line + name_beg, name_len);
    }
    else if (line_count % 4 == 1) push_read(out, name.data(), name.size(), line, len);
    ++line_count;
    ++cur_line_;
  }
}

}  // namespace ab2
