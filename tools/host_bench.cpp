// Single-thread throughput of the host stages either side of the GPU path: FASTQ parsing
// (FastqReader, the ReadLoader rules) and SAM / BAM formatting (select_output + statistics), on real records.
// The mapping results come from the CPU oracle on the first pairs of the input (test infrastructure; nothing
// here is shipped or part of bench.py).
// Build (from the repo root):
//   g++ -std=c++17 -O2 -DABISMAL_ENGINE_ORACLE -Iinclude -Ioracle -Iabismal_b200/csrc/host -o tools/host_bench \
//     tools/host_bench.cpp abismal_b200/csrc/host/{index_file,read_loader,sam_format,pipeline,bam_writer}.cpp \
//     oracle/abismal_oracle.cpp -lz -lpthread
// Usage: host_bench <index> <reads_1.fq> <reads_2.fq> [pairs mapped by the oracle = 20000]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "abismal_oracle.h"
#include "index_file.hpp"
#include "read_loader.hpp"
#include "sam_format.hpp"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <index> <fq1> <fq2> [n_oracle_pairs]\n", argv[0]);
    return 1;
  }
  const uint32_t n_map = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 20000u;
  ab2::IndexFile index;
  index.read(argv[1]);

  // ---- parse ----
  ab2::ReadBatch b1, b2;
  double t = now();
  size_t n_reads = 0, n_bases = 0;
  {
    ab2::FastqReader r1(argv[2]);
    ab2::ReadBatch tmp;
    while (r1.good()) {
      r1.load_reads(tmp, 1u << 18);
      n_reads += tmp.size();
      n_bases += tmp.seq.size();
    }
  }
  const double parse_s = now() - t;
  std::printf("parse   : %zu reads (%zu bases) in %.3f s -> %.2f M reads/s per thread\n", n_reads, n_bases, parse_s,
              n_reads / parse_s / 1e6);
  for (unsigned helpers : {2u, 4u, 8u}) {  // block mode: whole records parsed by several threads per reader
    t = now();
    size_t n_par = 0;
    ab2::FastqReader r1(argv[2], helpers);
    ab2::ReadBatch tmp;
    while (r1.good()) {
      r1.load_reads(tmp, 1u << 18);
      n_par += tmp.size();
    }
    const double s = now() - t;
    std::printf("parse x%u: %zu reads in %.3f s -> %.2f M reads/s per reader\n", helpers, n_par, s, n_par / s / 1e6);
  }

  // ---- records to format: the oracle's results on the first n_map pairs ----
  {
    ab2::FastqReader r1(argv[2]), r2(argv[3]);
    r1.load_reads(b1, n_map);
    r2.load_reads(b2, n_map);
  }
  const uint32_t n = b1.size(), stride = 64;
  std::vector<abg_hit> pe1(n), pe2(n), se1(n), se2(n);
  std::vector<uint32_t> c1((size_t)n * stride), c2((size_t)n * stride), nc1(n), nc2(n);
  {
    const abg_index_view v = index.view();
    abo_index *ox = nullptr;
    if (abo_index_create(&v, &ox) != 0) return 2;
    abg_params p{};
    p.mode = ABG_MODE_PAIRED;
    p.min_dist = 32;
    p.max_dist = 3000;
    p.valid_frac = 0.1;
    p.cigar_stride = stride;
    abg_batch b{};
    b.n = n;
    b.seq1 = b1.seq.data();
    b.off1 = b1.seq_off.data();
    b.seq2 = b2.seq.data();
    b.off2 = b2.seq_off.data();
    abg_results r{pe1.data(), pe2.data(), se1.data(), se2.data(), c1.data(), c2.data(), nc1.data(), nc2.data()};
    t = now();
    if (abo_map_batch(ox, &p, &b, &r, nullptr) != 0) return 3;
    std::printf("oracle  : %u pairs mapped in %.2f s (single thread; checker, not product)\n", n, now() - t);
    abo_index_destroy(ox);
  }
  const auto view = [&](const ab2::ReadBatch &bb, uint32_t i, const std::vector<uint32_t> &cg, const std::vector<uint32_t> &ncg) {
    ab2::ReadView rv;
    rv.name = bb.names.data() + bb.name_off[i];
    rv.name_len = bb.name_off[i + 1] - bb.name_off[i];
    rv.seq = bb.seq.data() + bb.seq_off[i];
    rv.seq_len = bb.seq_off[i + 1] - bb.seq_off[i];
    rv.cigar = cg.data() + (size_t)i * stride;
    rv.n_cigar = ncg[i];
    return rv;
  };
  for (int bam = 0; bam < 2; ++bam) {
    const int reps = bam ? 10 : 40;
    size_t bytes_out = 0, recs = 0;
    ab2::PeStats ps;
    t = now();
    for (int rep = 0; rep < reps; ++rep) {
      std::string bytes;
      bytes.reserve((size_t)n * 700);
      ab2::Emitter em;
      ab2::BgzfRecordPacker packer(bytes, 6);
      if (bam) em.bam = &packer;
      else em.sam = &bytes;
      std::vector<abg_hit> a = pe1, b = pe2, c = se1, d = se2;  // select_output resets records in place
      for (uint32_t i = 0; i < n; ++i) {
        const ab2::ReadView r1 = view(b1, i, c1, nc1), r2 = view(b2, i, c2, nc2);
        ab2::select_output(false, index.cl, r1, r2, a[i], b[i], c[i], d[i], em);
        ps.update(false, r1, r2, a[i], b[i], c[i], d[i]);
      }
      if (bam) packer.finish();
      bytes_out += bytes.size();
      recs += 2 * (size_t)n;
    }
    const double s = now() - t;
    std::printf("format %s: %zu reads in %.3f s -> %.2f M reads/s per thread (%.0f MB/s of output)\n", bam ? "BAM" : "SAM", recs,
                s, recs / s / 1e6, bytes_out / s / 1e6);
  }
  return 0;
}
