// Host side of `abismal idx` / `map -g`: FASTA -> padded, N-replaced, 4-bit
// encoded genome (the reference's load_genome, contiguous_n, replace_included_n
// and encode_dna_four_bit, src/AbismalIndex.cpp:125-175, :1322-1360), the GPU
// index builder behind include/abismal_b200_index.h, and the writer of the
// on-disk AbismalIndex format (src/AbismalIndex.cpp:1037-1072).
#ifndef ABISMAL_B200_GENOME_PREP_HPP
#define ABISMAL_B200_GENOME_PREP_HPP

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "index_file.hpp"

namespace ab2 {

struct PreparedGenome {
  ChromLookup cl;                 // names incl. pad_start / pad_end, starts
  std::vector<uint64_t> words;    // 4-bit packed, 16 bases per word
  uint64_t genome_size = 0;       // bases incl. both paddings
  std::vector<uint64_t> exclude;  // pairs [first, second) of N runs longer than 256
};

// throws std::runtime_error with the reference's messages
void prepare_genome(const std::string &fasta_path, PreparedGenome &out);

// Builds the index arrays on CUDA device `device` and fills `out` as if it had been read from a file.
// window_size: seed::window_size, 20 or 12 (--enable-short)
void build_index(PreparedGenome &&g, int device, IndexFile &out, uint32_t window_size = 20);

void write_index_file(const IndexFile &ix, const std::string &path);

}  // namespace ab2

// defined in map_main.cpp's translation unit for both engines
void build_index_from_fasta(const std::string &fasta_path, int device, ab2::IndexFile &out, uint32_t window_size = 20);

#endif
