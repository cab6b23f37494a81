#include "read_loader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
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

FastqReader::FastqReader(const std::string &filename, unsigned n_helpers) : filename_(filename) {
  if (n_helpers > 1) {
    pool_.reset(new WorkerPool(n_helpers));
    nl_parts_.resize(n_helpers);
  }
  // A plain (not gzip-compressed) regular file is parsed in place from a read-only mapping: no copies through
  // zlib and a staging buffer, and the helper threads fault its pages in parallel.
  const int fd = ::open(filename.c_str(), O_RDONLY);
  if (fd >= 0) {
    struct stat st;
    unsigned char magic[2] = {0, 0};
    if (::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && ::pread(fd, magic, 2, 0) == 2 &&
        !(magic[0] == 0x1f && magic[1] == 0x8b)) {
      void *p = ::mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
      if (p != MAP_FAILED) {
        ::madvise(p, static_cast<size_t>(st.st_size), MADV_SEQUENTIAL);
        map_ = static_cast<const char *>(p);
        map_len_ = static_cast<size_t>(st.st_size);
      }
    }
    ::close(fd);
  }
  if (map_) {
    data_ = map_;
    beg_ = 0;
    end_ = map_len_;
    src_eof_ = true;  // everything is "in the buffer"
    return;
  }
  gzFile f = gzopen(filename.c_str(), "rb");
  if (f) gzbuffer(f, 1u << 20);
  file_ = f;
  buf_.resize(pool_ ? (32u << 20) : (4u << 20));
  data_ = buf_.data();
  if (!f) eof_ = true;
}

FastqReader::~FastqReader() {
  if (file_) gzclose(static_cast<gzFile>(file_));
  if (map_) ::munmap(const_cast<char *>(map_), map_len_);
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
    const char *p = data_ + beg_;
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
    const char *p = data_ + beg_, *e = data_ + end_;
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
      beg_ = static_cast<size_t>(q - data_);
      return true;
    }
    if (attempt == 1 || !fill_more()) return false;
  }
  return false;
}

// Block mode: every whole 4-line record that is already in the buffer (topped up first), up to max_records,
// parsed in three parallel steps -- newline scan, per-record spans under the ReadLoader rules, copies into the
// batch.  Stops in front of the first record the serial path has to look at (empty name line, over-long read,
// read without A/C/G/T: it raises the reference's errors) and returns the number of records taken.
size_t FastqReader::parse_block(ReadBatch &out, size_t max_records) {
  if (!pool_ || max_records == 0) return 0;
  if (end_ - beg_ < (256u << 10)) fill_more();  // (a mapped file never refills)
  if (end_ == beg_) return 0;
  const unsigned T = pool_->size();
  // a window of at most 32 MB per call; offsets below are relative to its start
  const char *base = data_ + beg_;
  const size_t lo = 0, hi = std::min<size_t>(end_ - beg_, 32u << 20);
  // 1. newline offsets
  pool_->run(T, [&](unsigned k) {
    std::vector<uint32_t> &v = nl_parts_[k];
    v.clear();
    const size_t a = lo + (hi - lo) * k / T, b = lo + (hi - lo) * (k + 1) / T;
    const char *p = base + a, *e = base + b;
    while (p < e) {
      const char *q = static_cast<const char *>(std::memchr(p, '\n', static_cast<size_t>(e - p)));
      if (!q) break;
      v.push_back(static_cast<uint32_t>(q - base));
      p = q + 1;
    }
  });
  nl_.clear();
  for (unsigned k = 0; k < T; ++k) nl_.insert(nl_.end(), nl_parts_[k].begin(), nl_parts_[k].end());
  size_t n_rec = std::min(max_records, nl_.size() / 4);
  if (n_rec == 0) return 0;
  // 2. spans of name and (trimmed) sequence; first record that needs the serial path
  spans_.resize(n_rec);
  std::vector<size_t> stop(T, n_rec);
  const uint32_t min_len = min_read_length;
  pool_->run(T, [&](unsigned k) {
    const size_t r0 = n_rec * k / T, r1 = n_rec * (k + 1) / T;
    for (size_t r = r0; r < r1; ++r) {
      const size_t s0 = r == 0 ? lo : static_cast<size_t>(nl_[4 * r - 1]) + 1;
      size_t l0 = nl_[4 * r] - s0;
      const char *name_line = base + s0;
      if (l0 > 0 && name_line[l0 - 1] == '\r') --l0;
      const char *line = base + nl_[4 * r] + 1;
      size_t len = nl_[4 * r + 1] - (nl_[4 * r] + 1);
      if (len > 0 && line[len - 1] == '\r') --len;
      if (l0 == 0 || len >= padding_size) {
        stop[k] = r;
        return;
      }
      size_t non_n = len;
      if (std::memchr(line, 'N', len) != nullptr) {
        non_n = 0;
        for (size_t i = 0; i < len; ++i) non_n += line[i] != 'N';
      }
      size_t b = 0, e = 0;
      if (non_n >= min_len) {
        e = len;
        while (e > 0 && line[e - 1] == 'N') --e;
        while (b < e && line[b] != 'A' && line[b] != 'C' && line[b] != 'G' && line[b] != 'T') ++b;
        if (b == e) {
          stop[k] = r;
          return;
        }
        if (e - b < min_len) b = e = 0;
      }
      size_t nb, nlen;
      name_span(name_line, l0, nb, nlen);
      spans_[r] = Span{static_cast<uint32_t>(s0 + nb), static_cast<uint32_t>(nlen),
                       static_cast<uint32_t>(line + b - base), static_cast<uint32_t>(e - b)};
    }
  });
  for (unsigned k = 0; k < T; ++k) n_rec = std::min(n_rec, stop[k]);
  if (n_rec == 0) return 0;
  // 3. offsets (serial prefix sums), then the copies
  const size_t r_base = out.seq_off.size() - 1;
  out.seq_off.resize(r_base + n_rec + 1);
  out.name_off.resize(r_base + n_rec + 1);
  uint32_t so = out.seq_off[r_base], no = out.name_off[r_base];
  for (size_t r = 0; r < n_rec; ++r) {
    so += spans_[r].seq_len;
    no += spans_[r].name_len;
    out.seq_off[r_base + r + 1] = so;
    out.name_off[r_base + r + 1] = no;
  }
  out.seq.resize(so);
  out.names.resize(no);
  pool_->run(T, [&](unsigned k) {
    const size_t r0 = n_rec * k / T, r1 = n_rec * (k + 1) / T;
    for (size_t r = r0; r < r1; ++r) {
      std::memcpy(out.seq.data() + out.seq_off[r_base + r], base + spans_[r].seq_at, spans_[r].seq_len);
      std::memcpy(out.names.data() + out.name_off[r_base + r], base + spans_[r].name_at, spans_[r].name_len);
    }
  });
  beg_ += static_cast<size_t>(nl_[4 * n_rec - 1]) + 1;
  return n_rec;
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
    if (line_count % 4 == 0) {
      const size_t got = parse_block(out, (num_lines_to_read - line_count) / 4);
      if (got != 0) {
        line_count += 4 * got;
        cur_line_ += 4 * got;
        continue;
      }
    }
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
