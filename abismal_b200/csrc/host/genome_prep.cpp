#include "genome_prep.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

#ifndef ABISMAL_ENGINE_ORACLE
#include "abismal_b200_index.h"
#endif

namespace ab2 {

namespace {

constexpr uint64_t kPadding = 32767;   // seed::padding_size
constexpr uint64_t kMaxNCount = 256;   // max_n_count: longer N runs are excluded from the index

// dna_four_bit_encoding (src/dna_four_bit_bisulfite.hpp:156-165)
struct Enc4 {
  uint8_t t[256];
  Enc4() {
    std::memset(t, 0, sizeof t);
    const char *up = "ABCDGHKMRSTVWY";
    const uint8_t code[] = {1, 14, 2, 13, 4, 11, 12, 3, 5, 6, 8, 7, 9, 10};
    for (int i = 0; up[i]; ++i) {
      t[static_cast<unsigned char>(up[i])] = code[i];
      t[static_cast<unsigned char>(up[i] - 'A' + 'a')] = code[i];
    }
  }
};
const Enc4 enc4;

// random_base_generator (src/AbismalIndex.hpp:39-61): one LCG stream, x0 = 1
struct RandomBase {
  uint64_t x = 1;
  char operator()() {
    x = (1103515245ull * x + 12345ull) & 0x7fffffffull;
    return "ACGT"[x & 3u];
  }
};

}  // namespace

void prepare_genome(const std::string &path, PreparedGenome &out) {
  gzFile in = gzopen(path.c_str(), "rb");
  if (!in) throw std::runtime_error("failed to open genome file: " + path);
  gzbuffer(in, 1u << 20);
  std::vector<uint8_t> g;
  g.assign(kPadding, 'N');
  out.cl.names.assign(1, "pad_start");
  out.cl.starts.assign(1, 0);

  // load_genome_impl (:1322-1358): every non-'>' line is appended verbatim (less its line terminator)
  std::vector<char> buf(4u << 20);
  std::string header;
  bool in_header = false, at_line_start = true;
  size_t line_start = g.size();  // where the current sequence line began in g
  for (;;) {
    const int n = gzread(in, buf.data(), static_cast<unsigned>(buf.size()));
    if (n <= 0) break;
    const char *p = buf.data(), *end = p + n;
    while (p < end) {
      if (at_line_start) {
        in_header = *p == '>';
        if (in_header) header.clear();
        else line_start = g.size();
        at_line_start = false;
      }
      const char *nl = static_cast<const char *>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
      const char *stop = nl ? nl : end;
      if (in_header) header.append(p, stop);
      else g.insert(g.end(), p, stop);
      if (nl) {
        if (in_header) {
          if (!header.empty() && header.back() == '\r') header.pop_back();
          const size_t ws = header.find_first_of(" \t");
          out.cl.names.push_back(header.substr(1, ws == std::string::npos ? std::string::npos : ws - 1));
          out.cl.starts.push_back(static_cast<uint32_t>(g.size()));
        }
        else if (g.size() > line_start && g.back() == '\r') g.pop_back();
        at_line_start = true;
        p = nl + 1;
      }
      else p = end;
    }
  }
  gzclose(in);
  if (!at_line_start && in_header) {  // file ended inside a header line without a newline
    const size_t ws = header.find_first_of(" \t");
    out.cl.names.push_back(header.substr(1, ws == std::string::npos ? std::string::npos : ws - 1));
    out.cl.starts.push_back(static_cast<uint32_t>(g.size()));
  }
  if (out.cl.names.size() < 2) throw std::runtime_error("no names found in genome file");
  out.cl.names.push_back("pad_end");
  out.cl.starts.push_back(static_cast<uint32_t>(g.size()));
  g.insert(g.end(), kPadding, 'N');
  out.cl.starts.push_back(static_cast<uint32_t>(g.size()));
  if (g.size() >= (1ull << 32)) throw std::runtime_error("genome too large for 32-bit positions");
  out.genome_size = g.size();

  // contiguous_n (:125-145) keeping runs longer than max_n_count, then replace_included_n (:164-175):
  // every other N becomes the next base of the LCG stream, in genome order
  out.exclude.clear();
  RandomBase random_base;
  const size_t n = g.size();
  for (size_t i = 0; i < n;) {
    if (g[i] != 'N') {
      ++i;
      continue;
    }
    size_t j = i;
    while (j < n && g[j] == 'N') ++j;
    if (j - i > kMaxNCount) {
      out.exclude.push_back(i);
      out.exclude.push_back(j);
    }
    else
      for (size_t k = i; k < j; ++k) g[k] = static_cast<uint8_t>(random_base());
    i = j;
  }

  // encode_dna_four_bit (dna_four_bit_bisulfite.hpp:177-187)
  const size_t n_words = (n + 15) / 16;
  out.words.assign(n_words, 0);
  for (size_t w = 0; w < n_words; ++w) {
    uint64_t x = 0;
    const size_t lim = std::min<size_t>(16, n - 16 * w);
    for (size_t k = 0; k < lim; ++k) x |= static_cast<uint64_t>(enc4.t[g[16 * w + k]]) << (4 * k);
    out.words[w] = x;
  }
}

#ifndef ABISMAL_ENGINE_ORACLE
void build_index(PreparedGenome &&g, int device, IndexFile &out, uint32_t window_size) {
  abg_built_index b;
  std::memset(&b, 0, sizeof b);
  if (abg_build_index_w(g.words.data(), g.genome_size, g.exclude.data(), static_cast<uint32_t>(g.exclude.size() / 2),
                        window_size, device, &b) != 0)
    throw std::runtime_error(std::string("index construction failed: ") + abg_index_build_last_error());
  out.cl = std::move(g.cl);
  out.genome = std::move(g.words);
  out.genome.push_back(0);  // the look-ahead word IndexFile::read also appends
  out.max_candidates = b.max_candidates;
  out.window_size = window_size;
  out.counter_size = b.counter_size;
  out.counter_size_three = b.counter_size_three;
  out.index_size = b.index_size;
  out.index_size_three = b.index_size_three;
  out.counter.assign(b.counter, b.counter + b.counter_size + 1);
  out.counter_t.assign(b.counter_t, b.counter_t + b.counter_size_three + 1);
  out.counter_a.assign(b.counter_a, b.counter_a + b.counter_size_three + 1);
  out.index.assign(b.index, b.index + b.index_size);
  out.index_t.assign(b.index_t, b.index_t + b.index_size_three);
  out.index_a.assign(b.index_a, b.index_a + b.index_size_three);
  abg_built_index_free(&b);
}
#endif

void write_index_file(const IndexFile &ix, const std::string &path) {
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("failed to open output file: " + path);
  bool ok = true;
  const auto put = [&](const void *p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
  const auto put32 = [&](uint32_t v) { put(&v, 4); };
  const auto put64 = [&](uint64_t v) { put(&v, 8); };
  put("AbismalIndex", 12);
  put32(25);   // seed::key_weight
  put32(ix.window_size);  // seed::window_size
  put32(256);  // seed::n_sorting_positions
  put32(static_cast<uint32_t>(ix.cl.names.size()));
  for (const std::string &nm : ix.cl.names) {
    put32(static_cast<uint32_t>(nm.size()));
    put(nm.data(), nm.size());
  }
  put(ix.cl.starts.data(), ix.cl.starts.size() * 4);
  const uint64_t genome_words = (static_cast<uint64_t>(ix.cl.genome_size()) + 15) / 16;
  put(ix.genome.data(), genome_words * 8);
  put32(ix.max_candidates);
  put64(ix.counter_size);
  put64(ix.counter_size_three);
  put64(ix.index_size);
  put64(ix.index_size_three);
  put(ix.counter.data(), (ix.counter_size + 1) * 4);
  put(ix.counter_t.data(), (ix.counter_size_three + 1) * 4);
  put(ix.counter_a.data(), (ix.counter_size_three + 1) * 4);
  put(ix.index.data(), ix.index_size * 4);
  put(ix.index_t.data(), ix.index_size_three * 4);
  put(ix.index_a.data(), ix.index_size_three * 4);
  if (std::fclose(f) != 0) ok = false;
  if (!ok) throw std::runtime_error("failed writing index file: " + path);
}

}  // namespace ab2
