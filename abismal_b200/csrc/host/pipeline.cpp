#include "pipeline.hpp"

namespace ab2 {

HostMemHooks &host_mem_hooks() {
  static HostMemHooks h;
  return h;
}

void *host_buf_alloc(size_t bytes) {
  void *p = nullptr;
  const HostMemHooks &h = host_mem_hooks();
  if (h.alloc) {
    if (h.alloc(bytes ? bytes : 1, &p) != 0 || !p) throw std::bad_alloc();
  }
  else {
    p = std::malloc(bytes ? bytes : 1);
    if (!p) throw std::bad_alloc();
  }
  return p;
}

void host_buf_free(void *p) {
  if (!p) return;
  const HostMemHooks &h = host_mem_hooks();
  if (h.release) h.release(p);
  else std::free(p);
}

WorkerPool::WorkerPool(unsigned n_threads) {
  const unsigned extra = n_threads > 1 ? n_threads - 1 : 0;  // the caller of run() works too
  threads_.reserve(extra);
  for (unsigned i = 0; i < extra; ++i) threads_.emplace_back([this] { worker(); });
}

WorkerPool::~WorkerPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  wake_.notify_all();
  for (std::thread &t : threads_) t.join();
}

// Take indices of the current job until none are left.  Called with mu_ held.
void WorkerPool::drain(std::unique_lock<std::mutex> &lk) {
  while (job_ != nullptr && next_ < total_) {
    const unsigned k = next_++;
    const std::function<void(unsigned)> *f = job_;
    lk.unlock();
    std::exception_ptr err;
    try {
      (*f)(k);
    }
    catch (...) {
      err = std::current_exception();
    }
    lk.lock();
    if (err && !error_) error_ = err;
    if (--pending_ == 0) done_.notify_all();
  }
}

void WorkerPool::worker() {
  std::unique_lock<std::mutex> lk(mu_);
  uint64_t seen = 0;
  for (;;) {
    wake_.wait(lk, [&] { return stop_ || (generation_ != seen && job_ != nullptr && next_ < total_); });
    if (stop_) return;
    seen = generation_;
    drain(lk);
  }
}

void WorkerPool::run(unsigned n, const std::function<void(unsigned)> &f) {
  if (n == 0) return;
  std::lock_guard<std::mutex> one_job(run_mu_);
  std::unique_lock<std::mutex> lk(mu_);
  job_ = &f;
  next_ = 0;
  total_ = n;
  pending_ = n;
  error_ = nullptr;
  ++generation_;
  wake_.notify_all();
  drain(lk);
  done_.wait(lk, [&] { return pending_ == 0; });
  job_ = nullptr;
  const std::exception_ptr err = error_;
  error_ = nullptr;
  lk.unlock();
  if (err) std::rethrow_exception(err);
}

}  // namespace ab2
