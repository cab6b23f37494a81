#include "index_file.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace ab2 {

namespace {
struct FileCloser {
  FILE *f;
  ~FileCloser() {
    if (f) std::fclose(f);
  }
};

template <class T>
void read_pod(FILE *in, T &x, const char *msg) {
  if (std::fread(&x, sizeof(T), 1, in) != 1) throw std::runtime_error(msg);
}

template <class T>
void read_array(FILE *in, std::vector<T> &v, uint64_t n, const char *msg) {
  v.resize(n);
  // large freads in chunks keep partial-read handling simple
  uint64_t done = 0;
  while (done < n) {
    const size_t want = static_cast<size_t>(std::min<uint64_t>(n - done, 1ull << 26));
    if (std::fread(v.data() + done, sizeof(T), want, in) != want) throw std::runtime_error(msg);
    done += want;
  }
}
}  // namespace

bool ChromLookup::chrom_idx_and_offset(uint32_t pos, uint32_t ref_len, int32_t &chrom_idx,
                                       uint32_t &offset) const {
  auto it = std::upper_bound(starts.begin(), starts.end(), pos);
  if (it == starts.begin()) return false;
  --it;
  chrom_idx = static_cast<int32_t>(it - starts.begin());
  offset = pos - starts[chrom_idx];
  return pos + ref_len <= starts[chrom_idx + 1];
}

IndexFile::~IndexFile() = default;

void IndexFile::read(const std::string &path, bool map_file) {
  static const char *error_msg = "failed loading index file";
  FILE *in = std::fopen(path.c_str(), "rb");
  if (!in) throw std::runtime_error("cannot open input file " + path);
  FileCloser closer{in};

  char ident[12];
  if (std::fread(ident, 1, 12, in) != 12 || std::memcmp(ident, "AbismalIndex", 12) != 0)
    throw std::runtime_error("index file format problem: " + path);

  // seed::read (src/AbismalIndex.cpp:988-1024)
  uint32_t key_weight = 0, n_sorting = 0;
  read_pod(in, key_weight, "failed to read seed data");
  if (key_weight != 25u)
    throw std::runtime_error("inconsistent k-mer size. Expected: 25, got: " + std::to_string(key_weight));
  read_pod(in, window_size, "failed to read seed data");
  // the reference accepts the one window it was configured with (20, or 12 with --enable-short); this
  // reader takes either and the mapper follows the file
  if (window_size != 20u && window_size != 12u)
    throw std::runtime_error("inconsistent window size size. Expected: 20, got: " +
                             std::to_string(window_size));
  read_pod(in, n_sorting, "failed to read seed data");
  if (n_sorting != 256u)
    throw std::runtime_error("inconsistent sorting size size. Expected: 256, got: " +
                             std::to_string(n_sorting));

  // ChromLookup::read (:1225-1258)
  static const char *cl_msg = "failed loading chrom info from index";
  uint32_t n_chroms = 0;
  read_pod(in, n_chroms, cl_msg);
  cl.names.resize(n_chroms);
  for (uint32_t i = 0; i < n_chroms; ++i) {
    uint32_t name_size = 0;
    read_pod(in, name_size, cl_msg);
    cl.names[i].resize(name_size);
    if (name_size && std::fread(&cl.names[i][0], 1, name_size, in) != name_size) throw std::runtime_error(cl_msg);
  }
  read_array(in, cl.starts, static_cast<uint64_t>(n_chroms) + 1, cl_msg);

  const uint64_t genome_words = (static_cast<uint64_t>(cl.genome_size()) + 15) / 16;
  if (map_file) {
    // header parsed; everything from here on is fixed-size arrays.  They are read once (by the upload to HBM),
    // so they stay in one buffer that several threads fill with pread() -- a read-only mapping of the file cost
    // a page fault per 4 KB inside the single-threaded host-to-device copy (1.8 s for 2.7 GB on the GPU box).
    // The buffer starts at the first array, which also gives every array its natural alignment.
    const long at = std::ftell(in);
    struct stat st;
    if (at < 0 || fstat(fileno(in), &st) != 0 || static_cast<uint64_t>(st.st_size) < static_cast<uint64_t>(at))
      throw std::runtime_error(error_msg);
    const size_t len = static_cast<size_t>(st.st_size) - static_cast<size_t>(at);
    bulk_.reset(new unsigned char[len + 8]);
    bulk_len_ = len;
    {
      const int fd = fileno(in);
      unsigned char *dst = bulk_.get();
      const unsigned n_thr = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2u));
      const size_t per = ((len + n_thr - 1) / n_thr + 4095) & ~static_cast<size_t>(4095);
      std::atomic<bool> bad{false};
      std::vector<std::thread> th;
      for (unsigned t = 0; t < n_thr; ++t)
        th.emplace_back([&, t] {
          size_t o = std::min(len, t * per);
          const size_t e = std::min(len, o + per);
          while (o < e) {
            const ssize_t n = pread(fd, dst + o, std::min<size_t>(e - o, 16u << 20), static_cast<off_t>(at) + static_cast<off_t>(o));
            if (n <= 0) {
              bad = true;
              return;
            }
            o += static_cast<size_t>(n);
          }
        });
      for (std::thread &x : th) x.join();
      if (bad) throw std::runtime_error(error_msg);
    }
    const unsigned char *p = bulk_.get(), *end = bulk_.get() + len;
    const auto take = [&](uint64_t bytes) {
      if (static_cast<uint64_t>(end - p) < bytes) throw std::runtime_error(error_msg);
      const unsigned char *q = p;
      p += bytes;
      return q;
    };
    const auto pod32 = [&]() { uint32_t v; std::memcpy(&v, take(4), 4); return v; };
    const auto pod64 = [&]() { uint64_t v; std::memcpy(&v, take(8), 8); return v; };
    std::memset(&mapped_, 0, sizeof mapped_);
    mapped_.genome = reinterpret_cast<const uint64_t *>(take(genome_words * 8));
    mapped_.genome_words = genome_words;
    mapped_.genome_size = cl.genome_size();
    max_candidates = pod32();
    counter_size = pod64();
    counter_size_three = pod64();
    index_size = pod64();
    index_size_three = pod64();
    if (counter_size != (1ull << 25) || counter_size_three != 43046721ull) throw std::runtime_error(error_msg);
    if (index_size > (len >> 2) || index_size_three > (len >> 2)) throw std::runtime_error(error_msg);  // (before any * 4)
    mapped_.counter = reinterpret_cast<const uint32_t *>(take((counter_size + 1) * 4));
    mapped_.counter_t = reinterpret_cast<const uint32_t *>(take((counter_size_three + 1) * 4));
    mapped_.counter_a = reinterpret_cast<const uint32_t *>(take((counter_size_three + 1) * 4));
    mapped_.index = reinterpret_cast<const uint32_t *>(take(index_size * 4));
    mapped_.index_t = reinterpret_cast<const uint32_t *>(take(index_size_three * 4));
    mapped_.index_a = reinterpret_cast<const uint32_t *>(take(index_size_three * 4));
    mapped_.counter_size = counter_size;
    mapped_.counter_size_three = counter_size_three;
    mapped_.index_size = index_size;
    mapped_.index_size_three = index_size_three;
    mapped_.max_candidates = max_candidates;
    mapped_.window_size = window_size;
    return;
  }
  // one spare zero word: the look-ahead word of a compare at the very end
  read_array(in, genome, genome_words, error_msg);
  genome.push_back(0);

  read_pod(in, max_candidates, error_msg);
  read_pod(in, counter_size, error_msg);
  read_pod(in, counter_size_three, error_msg);
  read_pod(in, index_size, error_msg);
  read_pod(in, index_size_three, error_msg);
  if (counter_size != (1ull << 25) || counter_size_three != 43046721ull)
    throw std::runtime_error(error_msg);

  read_array(in, counter, counter_size + 1, error_msg);
  read_array(in, counter_t, counter_size_three + 1, error_msg);
  read_array(in, counter_a, counter_size_three + 1, error_msg);
  read_array(in, index, index_size, error_msg);
  read_array(in, index_t, index_size_three, error_msg);
  read_array(in, index_a, index_size_three, error_msg);
}

abg_index_view IndexFile::view() const {
  if (bulk_) return mapped_;
  abg_index_view v;
  std::memset(&v, 0, sizeof(v));
  v.genome = genome.data();
  v.genome_words = genome.size();
  v.genome_size = cl.genome_size();
  v.counter = counter.data();
  v.counter_size = counter_size;
  v.counter_t = counter_t.data();
  v.counter_a = counter_a.data();
  v.counter_size_three = counter_size_three;
  v.index = index.data();
  v.index_size = index_size;
  v.index_t = index_t.data();
  v.index_a = index_a.data();
  v.index_size_three = index_size_three;
  v.max_candidates = max_candidates;
  v.window_size = window_size;
  return v;
}

}  // namespace ab2
