// Host-side plumbing of the `map` front end: a bounded queue between pipeline
// stages, a fork-join worker pool for the per-batch host work (SAM/BAM
// formatting, statistics) and an allocator that puts batch and result buffers
// in page-locked memory so that the mapper DMAs them in place.
//
// The reference runs N symmetric OpenMP threads that each load, map and write
// a 1000-read batch under two mutexes (src/abismal.cpp:1541-1596); here the
// GPU does the mapping, so `-t N` host threads are spent on the stages either
// side of it: FASTQ readers -> mapper (one per GPU) -> formatters -> writer.
#ifndef ABISMAL_B200_PIPELINE_HPP
#define ABISMAL_B200_PIPELINE_HPP

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <deque>
#include <exception>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <utility>
#include <vector>

namespace ab2 {

// ---- page-locked memory hook -------------------------------------------------
// Set once by main() before any buffer exists (abg_host_alloc/abg_host_free in
// the GPU build, left null in the oracle test tool => plain malloc).
struct HostMemHooks {
  int (*alloc)(size_t, void **) = nullptr;
  void (*release)(void *) = nullptr;
};
HostMemHooks &host_mem_hooks();

void *host_buf_alloc(size_t bytes);  // page-locked when the hooks are set, malloc otherwise; throws std::bad_alloc
void host_buf_free(void *p);

// Growable array of trivially-copyable T in (possibly page-locked) host memory; elements are never
// value-initialised: these are plain buffers the parser or the GPU fills.
template <class T>
class pinned_vector {
public:
  pinned_vector() = default;
  pinned_vector(const pinned_vector &) = delete;
  pinned_vector &operator=(const pinned_vector &) = delete;
  pinned_vector(pinned_vector &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_) { o.p_ = nullptr; o.n_ = o.cap_ = 0; }
  ~pinned_vector() { host_buf_free(p_); }
  T *data() { return p_; }
  const T *data() const { return p_; }
  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  bool empty() const { return n_ == 0; }
  T &operator[](size_t i) { return p_[i]; }
  const T &operator[](size_t i) const { return p_[i]; }
  void clear() { n_ = 0; }
  void reserve(size_t cap) {
    if (cap <= cap_) return;
    T *q = static_cast<T *>(host_buf_alloc(cap * sizeof(T)));
    if (n_) std::memcpy(q, p_, n_ * sizeof(T));
    host_buf_free(p_);
    p_ = q;
    cap_ = cap;
  }
  void resize(size_t n) {
    if (n > cap_) reserve(std::max(n, cap_ + cap_ / 2));
    n_ = n;
  }
  void append(const T *src, size_t n) {
    if (n_ + n > cap_) reserve(std::max(n_ + n, cap_ + cap_ / 2));
    std::memcpy(p_ + n_, src, n * sizeof(T));
    n_ += n;
  }

private:
  T *p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
};

// ---- bounded multi-producer multi-consumer queue ---------------------------------
template <class T>
class BoundedQueue {
public:
  explicit BoundedQueue(size_t cap) : cap_(cap) {}
  // false when the queue was closed before the item could be queued
  bool push(T v) {
    std::unique_lock<std::mutex> lk(mu_);
    not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
    if (closed_) return false;
    q_.push_back(std::move(v));
    not_empty_.notify_one();
    return true;
  }
  // false when the queue is closed and drained
  bool pop(T &out) {
    std::unique_lock<std::mutex> lk(mu_);
    not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    not_full_.notify_one();
    return true;
  }
  void close() {
    std::lock_guard<std::mutex> lk(mu_);
    closed_ = true;
    not_full_.notify_all();
    not_empty_.notify_all();
  }

private:
  std::mutex mu_;
  std::condition_variable not_full_, not_empty_;
  std::deque<T> q_;
  size_t cap_;
  bool closed_ = false;
};

// ---- fork-join pool -----------------------------------------------------------
// run(n, f) calls f(k) for k in [0, n) on the pool's threads plus the caller
// and returns when all are done; the first exception is rethrown.
class WorkerPool {
public:
  explicit WorkerPool(unsigned n_threads);
  ~WorkerPool();
  WorkerPool(const WorkerPool &) = delete;
  WorkerPool &operator=(const WorkerPool &) = delete;
  unsigned size() const { return static_cast<unsigned>(threads_.size()) + 1; }
  void run(unsigned n, const std::function<void(unsigned)> &f);

private:
  void worker();
  void drain(std::unique_lock<std::mutex> &lk);

  std::vector<std::thread> threads_;
  std::mutex mu_, run_mu_;
  std::condition_variable wake_, done_;
  const std::function<void(unsigned)> *job_ = nullptr;
  unsigned next_ = 0, total_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  std::exception_ptr error_;
  bool stop_ = false;
};

}  // namespace ab2
#endif
