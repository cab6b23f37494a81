// `abismal-b200 map`: drop-in front end for the reference's `abismal map`
// (src/abismal.cpp:2295-2504): same flags, same AbismalIndex file, same SAM
// and stats output.  Mapping itself runs behind the C ABI in
// include/abismal_b200.h (CUDA); with -DABISMAL_ENGINE_ORACLE this file is
// built as oracle/oracle_map, a TEST TOOL that drives the CPU restatement
// through the very same host code.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "abismal_b200.h"
#include "index_file.hpp"
#include "options.hpp"
#include "read_loader.hpp"
#include "sam_format.hpp"
#ifdef ABISMAL_ENGINE_ORACLE
#include "abismal_oracle.h"
#endif

namespace {

constexpr const char *kVersion = "3.3.0";  // reference VERSION (configure.ac:17), part of @PG

void log_msg(const std::string &s) {
  const std::time_t t = std::time(nullptr);
  std::string tf(std::ctime(&t));
  tf.pop_back();
  std::cerr << "[" << tf << "] " << s << '\n';
}

std::string fmt_secs(double s) {
  char b[64];
  std::snprintf(b, sizeof b, "%.2fs", s);
  return b;
}

struct ResultBuffers {
  std::vector<abg_hit> pe_r1, pe_r2, se1, se2;
  std::vector<uint32_t> cigar1, cigar2, n_cigar1, n_cigar2;
  uint32_t stride = 0;
  void resize(uint32_t n, uint32_t cigar_stride, bool paired) {
    stride = cigar_stride;
    se1.resize(n);
    cigar1.resize(static_cast<size_t>(n) * stride);
    n_cigar1.resize(n);
    if (paired) {
      pe_r1.resize(n);
      pe_r2.resize(n);
      se2.resize(n);
      cigar2.resize(static_cast<size_t>(n) * stride);
      n_cigar2.resize(n);
    }
  }
  abg_results view(bool paired) {
    abg_results r;
    std::memset(&r, 0, sizeof r);
    r.se1 = se1.data();
    r.cigar1 = cigar1.data();
    r.n_cigar1 = n_cigar1.data();
    if (paired) {
      r.pe_r1 = pe_r1.data();
      r.pe_r2 = pe_r2.data();
      r.se2 = se2.data();
      r.cigar2 = cigar2.data();
      r.n_cigar2 = n_cigar2.data();
    }
    return r;
  }
};

// The mapping engine behind the C ABI.
class Engine {
public:
  Engine(const abg_index_view &view, const abg_params &params, uint32_t max_batch, uint32_t max_len,
         int device)
    : params_(params) {
#ifdef ABISMAL_ENGINE_ORACLE
    (void)max_batch;
    (void)max_len;
    (void)device;
    if (abo_index_create(&view, &oidx_) != 0) throw std::runtime_error(abo_last_error());
#else
    if (abg_index_create(&view, device, &idx_) != 0) throw std::runtime_error(abg_last_error());
    if (abg_mapper_create(idx_, &params, max_batch, max_len, 0, &mapper_) != 0)
      throw std::runtime_error(abg_last_error());
#endif
  }
  ~Engine() {
#ifdef ABISMAL_ENGINE_ORACLE
    abo_index_destroy(oidx_);
#else
    abg_mapper_destroy(mapper_);
    abg_index_destroy(idx_);
#endif
  }
  void map(const abg_batch &b, abg_results &r) {
#ifdef ABISMAL_ENGINE_ORACLE
    if (abo_map_batch(oidx_, &params_, &b, &r, nullptr) != 0) throw std::runtime_error(abo_last_error());
#else
    if (abg_map_batch(mapper_, &b, &r) != 0) throw std::runtime_error(abg_last_error());
#endif
  }

private:
  abg_params params_;
#ifdef ABISMAL_ENGINE_ORACLE
  abo_index *oidx_ = nullptr;
#else
  abg_index *idx_ = nullptr;
  abg_mapper *mapper_ = nullptr;
#endif
};

ab2::ReadView make_view(const ab2::ReadBatch &b, uint32_t i, const uint32_t *cig, const uint32_t *ncig,
                        uint32_t stride) {
  ab2::ReadView v;
  v.name = b.names.data() + b.name_off[i];
  v.name_len = b.name_off[i + 1] - b.name_off[i];
  v.seq = b.seq.data() + b.seq_off[i];
  v.seq_len = b.seq_off[i + 1] - b.seq_off[i];
  v.cigar = cig + static_cast<size_t>(i) * stride;
  v.n_cigar = ncig[i];
  return v;
}

int map_main(int argc, char *argv[]) {
  try {
    bool verbose = false, g_to_a_conversion = false, allow_ambig = false, pbat_mode = false;
    bool random_pbat = false, write_bam_fmt = false, stats_as_json = false, help = false, about = false;
    uint32_t max_candidates = 0, n_threads = 1, min_dist = 32, max_dist = 3000;
    uint32_t batch_size = 1u << 16, device = 0;
    double valid_frac = 0.1;
    std::string index_file, genome_file, outfile, stats_outfile;

    ab2::Options opt;
    opt.add("help", '?', "print this help message", false, help);
    opt.add("about", '\0', "print about message", false, about);
    opt.add("index", 'i', "index file", false, index_file);
    opt.add("genome", 'g', "genome file (FASTA)", false, genome_file);
    opt.add("outfile", 'o', "output file", true, outfile);
    opt.add("bam", 'B', "output BAM format", false, write_bam_fmt);
    opt.add("stats", 's', "map statistics file (YAML)", false, stats_outfile);
    opt.add("json", 'j', "output stats as JSON", false, stats_as_json);
    opt.add("max-candidates", 'c', "max candidates per seed (0: use default)", false, max_candidates);
    opt.add("min-frag", 'l', "min fragment size (pe mode)", false, min_dist);
    opt.add("max-frag", 'L', "max fragment size (pe mode)", false, max_dist);
    opt.add("max-distance", 'm', "max fractional edit distance", false, valid_frac);
    opt.add("ambig", 'a', "report a position for ambiguous mappers", false, allow_ambig);
    opt.add("pbat", 'P', "input follows the PBAT protocol", false, pbat_mode);
    opt.add("random-pbat", 'R', "input follows random PBAT protocol", false, random_pbat);
    opt.add("a-rich", 'A', "indicates reads are a-rich (se mode)", false, g_to_a_conversion);
    opt.add("threads", 't', "number of threads", false, n_threads);
    opt.add("verbose", 'v', "print more run info", false, verbose);
    // extras of this implementation (not in the reference)
    opt.add("gpu-batch", '\0', "reads (or pairs) per GPU batch", false, batch_size);
    opt.add("device", '\0', "CUDA device ordinal", false, device);
    const std::vector<std::string> leftover = opt.parse(argc, argv);

    const std::string usage = opt.help_message(argv[0], "<reads-fq1> [<reads-fq2>]");
    if (argc == 1 || help || about) {
      std::cerr << usage << '\n';
      return EXIT_SUCCESS;
    }
    if (opt.option_missing()) {
      std::cerr << "Missing required argument\n" << opt.option_missing_message() << '\n';
      return EXIT_SUCCESS;
    }
    if (leftover.size() != 1 && leftover.size() != 2) {
      std::cerr << usage << '\n';
      return EXIT_SUCCESS;
    }
    if (n_threads == 0 || n_threads > 1024) {
      std::cerr << "Please choose a valid number of threads" << '\n';
      return EXIT_SUCCESS;
    }
    if (index_file.empty() == genome_file.empty()) {
      std::cerr << "Select one of index file (-i) or genome file (-g)\n";
      return EXIT_SUCCESS;
    }
    const std::string reads_file = leftover.front();
    std::string reads_file2;
    if (access(reads_file.c_str(), F_OK) != 0) {
      std::cerr << "cannot open read 1 FASTQ file: " << reads_file << '\n';
      return EXIT_FAILURE;
    }
    const bool paired_end = leftover.size() == 2;
    if (paired_end) {
      reads_file2 = leftover.back();
      if (access(reads_file2.c_str(), F_OK) != 0) {
        std::cerr << "cannot open read 2 FASTQ file: " << reads_file2 << '\n';
        return EXIT_FAILURE;
      }
    }
    if (!genome_file.empty())
      throw std::runtime_error("on-the-fly indexing (-g) is not part of the GPU map path; "
                               "build the index with `abismal idx` and pass it with -i");
    if (write_bam_fmt)
      throw std::runtime_error("BAM output (-B) is not implemented yet; write SAM and convert");
    if (batch_size == 0) batch_size = 1;

    if (verbose) {
      log_msg(paired_end ? "input (PE): " + reads_file + ", " + reads_file2 : "input (SE): " + reads_file);
      log_msg("output (SAM): " + outfile);
      if (!stats_outfile.empty()) log_msg("map statistics: " + stats_outfile);
    }

    ab2::IndexFile index;
    const auto t0 = std::chrono::steady_clock::now();
    if (verbose) log_msg("loading index " + index_file);
    index.read(index_file);
    if (verbose)
      log_msg("loading time: " +
              fmt_secs(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
    if (max_candidates != 0) log_msg("manually setting max_candidates to " + std::to_string(max_candidates));

    FILE *out = outfile == "-" ? stdout : std::fopen(outfile.c_str(), "w");
    if (!out) throw std::runtime_error("failed to open output file: " + outfile);
    std::vector<char> out_buf(8u << 20);  // must outlive the FILE (declared before its closer)
    std::setvbuf(out, out_buf.data(), _IOFBF, out_buf.size());
    struct Closer {
      FILE *f;
      ~Closer() {
        if (f && f != stdout) std::fclose(f);
        else if (f) std::fflush(f);
      }
    } closer{out};

    const std::string hdr = ab2::make_sam_header(index.cl, argc, argv, kVersion);
    if (std::fwrite(hdr.data(), 1, hdr.size(), out) != hdr.size()) throw std::runtime_error("error writing header");

    abg_params params;
    std::memset(&params, 0, sizeof params);
    params.mode = (paired_end ? ABG_MODE_PAIRED : 0u) | (random_pbat ? ABG_MODE_RANDOM_PBAT : 0u);
    // abismal.cpp:2468-2483: SE a-rich for -A or -P; PE a-rich for -P.  -R selects
    // the *_rand drivers inside the runner (:2215, :2246) whatever conv is.
    if (paired_end ? pbat_mode : (g_to_a_conversion || pbat_mode)) params.mode |= ABG_MODE_A_RICH;
    params.allow_ambig = allow_ambig;
    params.min_dist = min_dist;
    params.max_dist = max_dist;
    params.valid_frac = valid_frac;
    params.max_candidates = max_candidates;
    params.cigar_stride = 64;

    ab2::FastqReader rl1(reads_file);
    std::unique_ptr<ab2::FastqReader> rl2;
    if (paired_end) rl2.reset(new ab2::FastqReader(reads_file2));

    std::unique_ptr<Engine> engine;
    uint32_t engine_max_len = 0;
    const abg_index_view view = index.view();

    ab2::ReadBatch b1, b2;
    ResultBuffers rb;
    ab2::SeStats se_stats;
    ab2::PeStats pe_stats;
    std::string sam;
    const auto t_map = std::chrono::steady_clock::now();
    uint64_t n_done = 0;

    while (rl1.good() && (!paired_end || rl2->good())) {
      rl1.load_reads(b1, batch_size);
      if (paired_end) {
        rl2->load_reads(b2, batch_size);
        if (b1.size() != b2.size())
          throw std::runtime_error("paired-end batch sizes differ. Batch 1: " + std::to_string(b1.size()) +
                                   ", batch 2: " + std::to_string(b2.size()) +
                                   ". Are you sure your paired-end inputs have the same number of reads?");
      }
      const uint32_t n = b1.size();
      if (n == 0) continue;
      const uint32_t max_len = std::max(b1.max_read_len(), paired_end ? b2.max_read_len() : 0u);
      if (!engine || max_len > engine_max_len) {
        engine.reset();
        engine_max_len = std::max<uint32_t>(256, max_len);
        engine.reset(new Engine(view, params, batch_size, engine_max_len, static_cast<int>(device)));
      }
      rb.resize(n, params.cigar_stride, paired_end);
      abg_batch batch;
      std::memset(&batch, 0, sizeof batch);
      batch.n = n;
      batch.seq1 = b1.seq.data();
      batch.off1 = b1.seq_off.data();
      if (paired_end) {
        batch.seq2 = b2.seq.data();
        batch.off2 = b2.seq_off.data();
      }
      abg_results res = rb.view(paired_end);
      engine->map(batch, res);

      sam.clear();
      if (!paired_end) {
        for (uint32_t i = 0; i < n; ++i) {
          const ab2::ReadView r = make_view(b1, i, rb.cigar1.data(), rb.n_cigar1.data(), rb.stride);
          abg_hit &best = rb.se1[i];
          if (r.seq_len != 0) {
            if (ab2::format_se(allow_ambig, best, index.cl, r, sam) == ab2::map_unmapped) ab2::hit_reset(best);
          }
          se_stats.update(allow_ambig, r, best);
        }
      }
      else {
        for (uint32_t i = 0; i < n; ++i) {
          const ab2::ReadView r1 = make_view(b1, i, rb.cigar1.data(), rb.n_cigar1.data(), rb.stride);
          const ab2::ReadView r2 = make_view(b2, i, rb.cigar2.data(), rb.n_cigar2.data(), rb.stride);
          ab2::select_output(allow_ambig, index.cl, r1, r2, rb.pe_r1[i], rb.pe_r2[i], rb.se1[i], rb.se2[i], sam);
          pe_stats.update(allow_ambig, r1, r2, rb.pe_r1[i], rb.pe_r2[i], rb.se1[i], rb.se2[i]);
        }
      }
      if (std::fwrite(sam.data(), 1, sam.size(), out) != sam.size()) throw std::runtime_error("failed to write bam");
      n_done += n;
    }
    if (verbose) {
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_map).count();
      log_msg("reads mapped: " + std::to_string(n_done));
      log_msg("total mapping time: " + fmt_secs(secs));
    }

    if (!stats_outfile.empty()) {
      std::ofstream statout(stats_outfile);
      if (statout) {
        if (stats_as_json) statout << (paired_end ? pe_stats.tojson() : se_stats.tojson());
        else statout << (paired_end ? pe_stats.tostring(allow_ambig) : se_stats.tostring("read1"));
      }
      else std::cerr << "failed to open stats out file: " << stats_outfile << '\n';
    }
  }
  catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char *argv[]) {
  // same dispatch shape as the reference's abismal_main.cpp: `<prog> map ...`
  if (argc < 2 || std::strcmp(argv[1], "map") != 0) {
    std::cerr << "usage: " << argv[0] << " map [OPTIONS] <reads-fq1> [<reads-fq2>]\n"
              << "(only the `map` command is provided; use the reference's `abismal idx` to build an index)\n";
    return argc < 2 ? EXIT_SUCCESS : EXIT_FAILURE;
  }
  return map_main(argc - 1, argv + 1);
}
