// `abismal-b200 map`: drop-in front end for the reference's `abismal map`
// (src/abismal.cpp:2295-2504): same flags, same AbismalIndex file, same SAM
// and stats output.  Mapping itself runs behind the C ABI in
// include/abismal_b200.h (CUDA); with -DABISMAL_ENGINE_ORACLE this file is
// built as oracle/oracle_map, a TEST TOOL that drives the CPU restatement
// through the very same host code.
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "abismal_b200.h"
#include "bam_writer.hpp"
#include "genome_prep.hpp"
#include "index_file.hpp"
#include "options.hpp"
#include "pipeline.hpp"
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
  // hit records and CIGAR lengths are DMA'd straight into these (page-locked in the GPU build)
  ab2::pinned_vector<abg_hit> pe_r1, pe_r2, se1, se2;
  ab2::pinned_vector<uint32_t> n_cigar1, n_cigar2;
  std::vector<uint32_t> cigar1, cigar2;
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

// The index resident on one device (HBM in the product, host memory in the oracle test tool); shared by the
// mapper workers of that device.
class DeviceIndex {
public:
  DeviceIndex(const abg_index_view &view, int device) {
#ifdef ABISMAL_ENGINE_ORACLE
    (void)device;
    if (abo_index_create(&view, &oidx_) != 0) throw std::runtime_error(abo_last_error());
#else
    if (abg_index_create(&view, device, &idx_) != 0) throw std::runtime_error(abg_last_error());
#endif
  }
  ~DeviceIndex() {
#ifdef ABISMAL_ENGINE_ORACLE
    abo_index_destroy(oidx_);
#else
    abg_index_destroy(idx_);
#endif
  }
  DeviceIndex(const DeviceIndex &) = delete;
  DeviceIndex &operator=(const DeviceIndex &) = delete;
#ifdef ABISMAL_ENGINE_ORACLE
  abo_index *oidx_ = nullptr;
#else
  abg_index *idx_ = nullptr;
#endif
};

// The mapping engine behind the C ABI: streams, staging and scratch of one mapper worker.
class Engine {
public:
  Engine(std::shared_ptr<DeviceIndex> index, const abg_params &params, uint32_t max_batch, uint32_t max_len)
    : index_(std::move(index)), params_(params) {
#ifdef ABISMAL_ENGINE_ORACLE
    (void)max_batch;
    (void)max_len;
#else
    if (abg_mapper_create(index_->idx_, &params, max_batch, max_len, 0, &mapper_) != 0)
      throw std::runtime_error(abg_last_error());
#endif
  }
  ~Engine() {
#ifndef ABISMAL_ENGINE_ORACLE
    abg_mapper_destroy(mapper_);
#endif
  }
  // false: a reported CIGAR did not fit cigar_stride operations (the caller retries with a larger stride)
  bool map(const abg_batch &b, abg_results &r) {
#ifdef ABISMAL_ENGINE_ORACLE
    const int rc = abo_map_batch(index_->oidx_, &params_, &b, &r, nullptr);
    if (rc == ABG_ERR_CIGAR_OVERFLOW) return false;
    if (rc != 0) throw std::runtime_error(abo_last_error());
#else
    const int rc = abg_map_batch(mapper_, &b, &r);
    if (rc == ABG_ERR_CIGAR_OVERFLOW) return false;
    if (rc != 0) throw std::runtime_error(abg_last_error());
#endif
    return true;
  }
  uint32_t cigar_stride() const { return params_.cigar_stride; }

private:
  std::shared_ptr<DeviceIndex> index_;
  abg_params params_;
#ifndef ABISMAL_ENGINE_ORACLE
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

// One batch travelling through the pipeline.
struct WorkItem {
  uint64_t seq_no = 0;
  ab2::ReadBatch b1, b2;
  ResultBuffers rb;
};

struct StageClock {  // busy seconds of one pipeline stage (reported with -v)
  std::atomic<uint64_t> ns{0};
  struct Scope {
    StageClock &c;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Scope(StageClock &cc) : c(cc) {}
    ~Scope() {
      c.ns += static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
    }
  };
  double secs() const { return static_cast<double>(ns.load()) * 1e-9; }
};

struct MapConfig {
  bool paired_end = false, allow_ambig = false, write_bam = false, verbose = false;
  uint32_t batch_size = 0, n_threads = 1, workers_per_gpu = 2;
  int bam_level = -1;
  std::vector<int> devices;
  abg_params params{};
};

// FASTQ readers -> mapper (one per GPU) -> formatter pool -> ordered writer.
// Output order is input order, i.e. the reference's `-t 1` order.
class MapPipeline {
public:
  MapPipeline(const MapConfig &cfg, const ab2::IndexFile &index, const std::string &fq1, const std::string &fq2,
              FILE *out)
    : cfg_(cfg), index_(index), out_(out), rl1_(fq1, parse_helpers(cfg)),
      n_items_(4 + 2 * cfg.devices.size() * cfg.workers_per_gpu), free_(n_items_), q12_(n_items_), to_map_(n_items_),
      to_out_(n_items_), pool_(cfg.n_threads), free_blocks_(kOutBlocks), to_disk_(kOutBlocks) {
    if (cfg.paired_end) rl2_.reset(new ab2::FastqReader(fq2, parse_helpers(cfg)));
    rl1_.set_window_size(index.window_size);
    if (rl2_) rl2_->set_window_size(index.window_size);
    items_.resize(n_items_);
    for (WorkItem &it : items_) free_.push(&it);
    for (OutBlock &b : blocks_) free_blocks_.push(&b);
  }

  // threads each FASTQ reader parses whole records with (a quarter of -t per file, at least one)
  static unsigned parse_helpers(const MapConfig &cfg) { return std::max(1u, cfg.n_threads / 4u); }

  void run() {
    std::vector<std::thread> th;
    th.emplace_back([this] { guarded([this] { read_end1(); }); });
    if (cfg_.paired_end) th.emplace_back([this] { guarded([this] { read_end2(); }); });
    // workers_per_gpu mapper workers per device share its index: while one waits for the tail of its batch
    // (last kernel, copy back, host scatter) the other keeps the GPU busy
    n_mappers_live_ = static_cast<int>(cfg_.devices.size() * cfg_.workers_per_gpu);
    std::vector<std::shared_ptr<std::promise<std::shared_ptr<DeviceIndex>>>> promises;
    for (int dev : cfg_.devices) {
      auto pr = std::make_shared<std::promise<std::shared_ptr<DeviceIndex>>>();
      std::shared_future<std::shared_ptr<DeviceIndex>> fut = pr->get_future().share();
      for (uint32_t w = 0; w < cfg_.workers_per_gpu; ++w)
        th.emplace_back([this, dev, w, pr, fut] { guarded([&] { map_on(dev, w == 0 ? pr.get() : nullptr, fut); }); });
    }
    th.emplace_back([this] { guarded([this] { format_out(); }); });
    th.emplace_back([this] { guarded([this] { write_out(); }); });
    for (std::thread &t : th) t.join();
    if (error_) std::rethrow_exception(error_);
  }

  ab2::SeStats se_stats;
  ab2::PeStats pe_stats;
  uint64_t n_done = 0;
  StageClock t_read1, t_read2, t_map, t_format, t_write, t_upload, t_engine;

private:
  template <class F>
  void guarded(F f) {
    try {
      f();
    }
    catch (...) {
      {
        std::lock_guard<std::mutex> lk(err_mu_);
        if (!error_) error_ = std::current_exception();
      }
      failed_ = true;
      free_.close();
      q12_.close();
      to_map_.close();
      to_out_.close();
      free_blocks_.close();
      to_disk_.close();
    }
  }

  // abismal.cpp:1549-1555 / :1947-1955: `while (rl1 && rl2) { load; load; compare sizes; ... }`
  void read_end1() {
    uint64_t k = 0;
    while (!failed_ && !stop_reading_) {
      WorkItem *it = nullptr;
      if (!free_.pop(it)) break;
      bool last;
      {
        StageClock::Scope sc(t_read1);
        it->seq_no = k++;
        rl1_.load_reads(it->b1, cfg_.batch_size);
        last = !rl1_.good();
      }
      if (!(cfg_.paired_end ? q12_ : to_map_).push(it)) break;
      if (last) break;
    }
    (cfg_.paired_end ? q12_ : to_map_).close();
  }

  void read_end2() {
    WorkItem *it = nullptr;
    while (!failed_ && q12_.pop(it)) {
      {
        StageClock::Scope sc(t_read2);
        rl2_->load_reads(it->b2, cfg_.batch_size);
      }
      if (it->b1.size() != it->b2.size())
        throw std::runtime_error("paired-end batch sizes differ. Batch 1: " + std::to_string(it->b1.size()) +
                                 ", batch 2: " + std::to_string(it->b2.size()) +
                                 ". Are you sure your paired-end inputs have the same number of reads?");
      const bool last = !rl2_->good();
      if (!to_map_.push(it)) break;
      if (last) {
        stop_reading_ = true;  // the reference's loop ends as soon as either file is exhausted
        break;
      }
    }
    q12_.close();  // unblocks reader 1 if it ran ahead
    to_map_.close();
  }

  void map_on(int device, std::promise<std::shared_ptr<DeviceIndex>> *lead,
              std::shared_future<std::shared_ptr<DeviceIndex>> shared) {
    // The index upload to HBM and the engine (scratch, streams) are set up while the readers parse the first
    // batches; the lead worker of a device uploads, the others wait for it.
    if (lead) {
      try {
        StageClock::Scope sc(t_upload);
        lead->set_value(std::make_shared<DeviceIndex>(index_.view(), device));
      }
      catch (...) {
        lead->set_exception(std::current_exception());
      }
    }
    std::shared_ptr<DeviceIndex> dev_index = shared.get();
    // The reference has no CIGAR length limit (bam_cigar_t is a vector).  The engine starts with 64 slots per
    // read -- NM <= valid_frac * length keeps reported CIGARs of 150-base reads below 2 * 30 + 3 operations --
    // and is rebuilt with twice as many whenever a batch reports one that did not fit.
    uint32_t engine_max_len = 256;
    abg_params params = cfg_.params;
    std::unique_ptr<Engine> engine;
    {
      StageClock::Scope sc(t_engine);
      engine.reset(new Engine(dev_index, params, cfg_.batch_size, engine_max_len));
    }
    WorkItem *it = nullptr;
    while (!failed_ && to_map_.pop(it)) {
      const uint32_t n = it->b1.size();
      if (n != 0) {
        StageClock::Scope sc(t_map);
        const uint32_t max_len = std::max(it->b1.max_read_len(), cfg_.paired_end ? it->b2.max_read_len() : 0u);
        if (max_len > engine_max_len) {
          engine.reset();
          engine_max_len = max_len;
          engine.reset(new Engine(dev_index, params, cfg_.batch_size, engine_max_len));
        }
        abg_batch batch;
        std::memset(&batch, 0, sizeof batch);
        batch.n = n;
        batch.seq1 = it->b1.seq.data();
        batch.off1 = it->b1.seq_off.data();
        if (cfg_.paired_end) {
          batch.seq2 = it->b2.seq.data();
          batch.off2 = it->b2.seq_off.data();
        }
        for (;;) {
          it->rb.resize(n, params.cigar_stride, cfg_.paired_end);
          abg_results res = it->rb.view(cfg_.paired_end);
          if (engine->map(batch, res)) break;
          if (params.cigar_stride >= 2u * engine_max_len + 3u)  // an alignment cannot have more operations
            throw std::runtime_error("a CIGAR needed more operations than the read has bases");
          params.cigar_stride = std::min(2u * params.cigar_stride, 2u * engine_max_len + 3u);
          engine.reset();
          engine.reset(new Engine(dev_index, params, cfg_.batch_size, engine_max_len));
        }
      }
      if (!to_out_.push(it)) break;
    }
    if (--n_mappers_live_ == 0) to_out_.close();
  }

  // Formats reads [i0, i1) of the item into `bytes` (SAM text or BGZF-framed BAM records).
  void format_slice(WorkItem &w, uint32_t i0, uint32_t i1, std::string &bytes, ab2::SeStats &ss, ab2::PeStats &ps) {
    const bool allow_ambig = cfg_.allow_ambig;
    ab2::Emitter em;
    ab2::BgzfRecordPacker packer(bytes, cfg_.bam_level);
    if (cfg_.write_bam) em.bam = &packer;
    else em.sam = &bytes;
    ResultBuffers &rb = w.rb;
    if (!cfg_.paired_end) {
      for (uint32_t i = i0; i < i1; ++i) {
        const ab2::ReadView r = make_view(w.b1, i, rb.cigar1.data(), rb.n_cigar1.data(), rb.stride);
        abg_hit &best = rb.se1[i];
        if (r.seq_len != 0) {
          if (ab2::format_se(allow_ambig, best, index_.cl, r, em) == ab2::map_unmapped) ab2::hit_reset(best);
        }
        ss.update(allow_ambig, r, best);
      }
    }
    else {
      for (uint32_t i = i0; i < i1; ++i) {
        const ab2::ReadView r1 = make_view(w.b1, i, rb.cigar1.data(), rb.n_cigar1.data(), rb.stride);
        const ab2::ReadView r2 = make_view(w.b2, i, rb.cigar2.data(), rb.n_cigar2.data(), rb.stride);
        ab2::select_output(allow_ambig, index_.cl, r1, r2, rb.pe_r1[i], rb.pe_r2[i], rb.se1[i], rb.se2[i], em);
        ps.update(allow_ambig, r1, r2, rb.pe_r1[i], rb.pe_r2[i], rb.se1[i], rb.se2[i]);
      }
    }
    if (cfg_.write_bam) packer.finish();
  }

  // Formatter stage: batches in input order (multi-GPU: they finish out of order), each formatted by the pool
  // into one output block, which the writer thread puts on disk while the next batch is being formatted.
  void format_out() {
    std::map<uint64_t, WorkItem *> waiting;
    uint64_t next = 0;
    const unsigned n_slices = std::max(1u, pool_.size() * 2);
    std::vector<ab2::SeStats> ss(n_slices);
    std::vector<ab2::PeStats> ps(n_slices);
    WorkItem *in = nullptr;
    while (!failed_ && to_out_.pop(in)) {
      waiting[in->seq_no] = in;
      while (!waiting.empty() && waiting.begin()->first == next) {
        WorkItem *w = waiting.begin()->second;
        waiting.erase(waiting.begin());
        ++next;
        const uint32_t n = w->b1.size();
        if (n != 0) {
          OutBlock *blk = nullptr;
          if (!free_blocks_.pop(blk)) return;
          blk->bytes.resize(n_slices);
          {
            StageClock::Scope sc(t_format);
            pool_.run(n_slices, [&](unsigned k) {
              const uint32_t i0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * k / n_slices);
              const uint32_t i1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (k + 1) / n_slices);
              blk->bytes[k].clear();
              ss[k] = ab2::SeStats();
              ps[k] = ab2::PeStats();
              format_slice(*w, i0, i1, blk->bytes[k], ss[k], ps[k]);
            });
          }
          for (unsigned k = 0; k < n_slices; ++k) {
            se_stats.add(ss[k]);
            pe_stats.add(ps[k]);
          }
          n_done += n;
          if (!to_disk_.push(blk)) return;
        }
        free_.push(w);
      }
    }
    free_.close();
    to_disk_.close();
  }

  void write_out() {
    OutBlock *blk = nullptr;
    while (!failed_ && to_disk_.pop(blk)) {
      {
        StageClock::Scope sc(t_write);
        for (const std::string &b : blk->bytes)
          if (!b.empty() && std::fwrite(b.data(), 1, b.size(), out_) != b.size())
            throw std::runtime_error("failed to write bam");
      }
      free_blocks_.push(blk);
    }
    free_blocks_.close();
  }

  const MapConfig &cfg_;
  const ab2::IndexFile &index_;
  FILE *out_;
  ab2::FastqReader rl1_;
  std::unique_ptr<ab2::FastqReader> rl2_;
  size_t n_items_;
  std::vector<WorkItem> items_;
  ab2::BoundedQueue<WorkItem *> free_, q12_, to_map_, to_out_;
  ab2::WorkerPool pool_;
  static constexpr size_t kOutBlocks = 3;
  struct OutBlock {
    std::vector<std::string> bytes;  // one slice per formatter job, in order
  };
  OutBlock blocks_[kOutBlocks];
  ab2::BoundedQueue<OutBlock *> free_blocks_, to_disk_;
  std::atomic<bool> failed_{false}, stop_reading_{false};
  std::atomic<int> n_mappers_live_{0};
  std::mutex err_mu_;
  std::exception_ptr error_;
};

int map_main(int argc, char *argv[]) {
  try {
    bool verbose = false, g_to_a_conversion = false, allow_ambig = false, pbat_mode = false;
    bool random_pbat = false, write_bam_fmt = false, stats_as_json = false, help = false, about = false;
    uint32_t max_candidates = 0, n_threads = 0, min_dist = 32, max_dist = 3000;
    uint32_t batch_size = 1u << 18, device = 0, n_gpus = 1, gpu_workers = 2;
    double valid_frac = 0.1;
    std::string index_file, genome_file, outfile, stats_outfile;
    bool enable_short = false;

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
    opt.add("threads", 't', "number of host threads (0: all cores, at most 32)", false, n_threads);
    opt.add("verbose", 'v', "print more run info", false, verbose);
    // extras of this implementation (not in the reference)
    opt.add("gpu-batch", '\0', "reads (or pairs) per GPU batch", false, batch_size);
    opt.add("device", '\0', "first CUDA device ordinal", false, device);
    opt.add("gpus", '\0', "number of GPUs to shard batches over (index replicated)", false, n_gpus);
    opt.add("gpu-workers", '\0', "mapper workers (streams) per GPU", false, gpu_workers);
    opt.add("enable-short", '\0', "with -g: index with window 12 (the reference's --enable-short build)", false, enable_short);
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
    if (n_threads > 1024) {
      std::cerr << "Please choose a valid number of threads" << '\n';
      return EXIT_SUCCESS;
    }
    if (n_threads == 0) n_threads = std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    n_threads = std::min(n_threads, std::max(1u, std::thread::hardware_concurrency()));  // abismal.cpp:2398-2400
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
    if (batch_size == 0) batch_size = 1;
    if (n_gpus == 0) n_gpus = 1;
    if (gpu_workers == 0) gpu_workers = 1;

    if (verbose) {
      log_msg(paired_end ? "input (PE): " + reads_file + ", " + reads_file2 : "input (SE): " + reads_file);
      log_msg(std::string("output (") + (write_bam_fmt ? "BAM" : "SAM") + "): " + outfile);
      if (!stats_outfile.empty()) log_msg("map statistics: " + stats_outfile);
    }

#ifndef ABISMAL_ENGINE_ORACLE
    ab2::host_mem_hooks().alloc = abg_host_alloc;
    ab2::host_mem_hooks().release = abg_host_free;
#endif

    ab2::IndexFile index;
    const auto t0 = std::chrono::steady_clock::now();
    if (!index_file.empty()) {
      if (verbose) log_msg("loading index " + index_file);
#ifdef ABISMAL_ENGINE_ORACLE
      // (the test tool reads into vectors; ABISMAL_B200_INDEX_BULK=1 makes it take the product's loader)
      index.read(index_file, std::getenv("ABISMAL_B200_INDEX_BULK") != nullptr);
#else
      index.read(index_file, true);  // arrays stay in one buffer filled by several threads: they are read once, by the upload to HBM
#endif
      if (verbose)
        log_msg("loading time: " +
                fmt_secs(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
    }
    else {
      // abismal.cpp:2439-2446: index the genome on the fly
      if (verbose) log_msg("indexing genome " + genome_file);
      build_index_from_fasta(genome_file, static_cast<int>(device), index, enable_short ? 12u : 20u);
      if (verbose)
        log_msg("indexing time: " +
                fmt_secs(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
    }
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

    MapConfig cfg;
    cfg.paired_end = paired_end;
    cfg.allow_ambig = allow_ambig;
    cfg.write_bam = write_bam_fmt;
    cfg.verbose = verbose;
    cfg.batch_size = batch_size;
    cfg.n_threads = n_threads;
    cfg.workers_per_gpu = gpu_workers;
    if (const char *e = std::getenv("ABISMAL_B200_BAM_LEVEL")) cfg.bam_level = std::atoi(e);
    for (uint32_t g = 0; g < n_gpus; ++g) cfg.devices.push_back(static_cast<int>(device + g));

    const std::string hdr = ab2::make_sam_header(index.cl, argc, argv, kVersion);
    {
      std::string bytes;
      if (write_bam_fmt) {
        const std::string bh = ab2::make_bam_header(index.cl, hdr);
        ab2::bgzf_append(bh.data(), bh.size(), cfg.bam_level, bytes);
      }
      else bytes = hdr;
      if (std::fwrite(bytes.data(), 1, bytes.size(), out) != bytes.size())
        throw std::runtime_error("error writing header");
    }

    abg_params &params = cfg.params;
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

    const auto t_map = std::chrono::steady_clock::now();
    MapPipeline pipe(cfg, index, reads_file, reads_file2, out);
    pipe.run();
    if (write_bam_fmt) {
      const std::string &eof = ab2::bgzf_eof_marker();
      if (std::fwrite(eof.data(), 1, eof.size(), out) != eof.size()) throw std::runtime_error("failed to write bam");
    }
    if (verbose) {
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_map).count();
      log_msg("reads mapped: " + std::to_string(pipe.n_done));
      log_msg("total mapping time: " + fmt_secs(secs));
      // summed over the devices (they upload concurrently); the readers parse the first batches meanwhile
      log_msg("index upload to HBM: " + fmt_secs(pipe.t_upload.secs() / static_cast<double>(cfg.devices.size())) + " per device");
      log_msg("engine set-up (device scratch, streams; summed over the mapper workers): " + fmt_secs(pipe.t_engine.secs()));
      log_msg("stage busy time: read1 " + fmt_secs(pipe.t_read1.secs()) + ", read2 " + fmt_secs(pipe.t_read2.secs()) +
              ", map (" + std::to_string(cfg.devices.size()) + " GPU) " + fmt_secs(pipe.t_map.secs()) + ", format (" +
              std::to_string(n_threads) + " threads) " + fmt_secs(pipe.t_format.secs()) + ", write " +
              fmt_secs(pipe.t_write.secs()));
    }

    if (!stats_outfile.empty()) {
      std::ofstream statout(stats_outfile);
      if (statout) {
        if (stats_as_json) statout << (paired_end ? pipe.pe_stats.tojson() : pipe.se_stats.tojson());
        else statout << (paired_end ? pipe.pe_stats.tostring(allow_ambig) : pipe.se_stats.tostring("read1"));
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

// `abismal-b200 idx`: the reference's abismalidx (src/abismalidx.cpp:30-120) with the index built on the GPU.
int idx_main(int argc, char *argv[]) {
  try {
    bool verbose = false, help = false, about = false, enable_short = false;
    uint32_t n_threads = 1, device = 0;
    std::string target_regions_file;
    ab2::Options opt;
    opt.add("help", '?', "print this help message", false, help);
    opt.add("about", '\0', "print about message", false, about);
    opt.add("targets", 'A', "target regions", false, target_regions_file);
    opt.add("threads", 't', "number of threads", false, n_threads);
    opt.add("verbose", 'v', "print more run info", false, verbose);
    opt.add("device", '\0', "CUDA device ordinal", false, device);
    opt.add("enable-short", '\0', "window 12 instead of 20 (the reference's --enable-short build)", false, enable_short);
    const std::vector<std::string> leftover = opt.parse(argc, argv);
    const std::string usage = opt.help_message(argv[0], "<genome-fasta> <abismal-index-file>");
    if (argc == 1 || help || about) {
      std::cerr << usage << '\n';
      return EXIT_SUCCESS;
    }
    if (leftover.size() != 2) {
      std::cerr << usage << '\n';
      return EXIT_SUCCESS;
    }
    if (!target_regions_file.empty())
      throw std::runtime_error("target regions (-A) are not supported by the GPU index builder");
    const auto t0 = std::chrono::steady_clock::now();
    if (verbose) log_msg("indexing genome " + leftover.front());
    ab2::IndexFile index;
    build_index_from_fasta(leftover.front(), static_cast<int>(device), index, enable_short ? 12u : 20u);
    if (verbose) log_msg("writing index file " + leftover.back());
    ab2::write_index_file(index, leftover.back());
    if (verbose)
      log_msg("total indexing time: " +
              fmt_secs(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
  }
  catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

}  // namespace

void build_index_from_fasta(const std::string &fasta_path, int device, ab2::IndexFile &out, uint32_t window_size) {
#ifdef ABISMAL_ENGINE_ORACLE
  (void)fasta_path;
  (void)device;
  (void)out;
  (void)window_size;
  throw std::runtime_error("index construction is not part of the oracle test tool; pass an index with -i");
#else
  ab2::PreparedGenome g;
  ab2::prepare_genome(fasta_path, g);
  ab2::build_index(std::move(g), device, out, window_size);
#endif
}

int main(int argc, char *argv[]) {
  // same dispatch shape as the reference's abismal_main.cpp: `<prog> <command> ...`
  if (argc >= 2 && std::strcmp(argv[1], "map") == 0) return map_main(argc - 1, argv + 1);
  if (argc >= 2 && std::strcmp(argv[1], "idx") == 0) return idx_main(argc - 1, argv + 1);
  std::cerr << "usage: " << argv[0] << " <command> [OPTIONS]\n"
            << "commands:\n  idx   make an index for a reference genome (on the GPU)\n"
            << "  map   map bisulfite converted reads (on the GPU)\n"
            << "(`sim` is not provided; use the reference's `abismal sim`)\n";
  return argc < 2 ? EXIT_SUCCESS : EXIT_FAILURE;
}
