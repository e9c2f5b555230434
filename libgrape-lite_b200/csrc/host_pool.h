// host_pool.h — a small persistent pool of host threads for result post-processing
// (widening the one-byte BFS depths that crossed PCIe into the reference's int64
// depth array).  Independent of OpenMP on purpose: launchers such as torchrun
// export OMP_NUM_THREADS=1, which would serialise an `omp parallel for`.
//
// One job = a sequence of chunks that become available over time (the main
// thread publishes `ready` after each chunk's D2H copy completed); every worker
// converts its slice of each published chunk.  Workers are woken once per job
// and then poll the atomics, so a job costs one condition-variable round trip.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include <emmintrin.h>

namespace gl {

class WidenPool {
 public:
  static WidenPool& instance() {
    static WidenPool p;
    return p;
  }
  int threads() const { return (int) th_.size() + 1; }

  // chunks are [bounds[c], bounds[c+1]); returns after everything was converted.
  // wait_chunk(c) blocks until chunk c of `in` is valid (called by the calling thread).
  template <class WaitFn>
  void widen_u8_to_i64(const uint8_t* in, int64_t* out, const uint32_t* bounds, uint32_t nchunks, WaitFn wait_chunk) {
    {
      std::lock_guard<std::mutex> g(mu_);
      in_ = in;
      out_ = out;
      bounds_ = bounds;
      nchunks_ = nchunks;
      ready_.store(0, std::memory_order_relaxed);
      done_.store(0, std::memory_order_relaxed);
      ++job_;
    }
    cv_.notify_all();
    const int K = threads();
    for (uint32_t c = 0; c < nchunks; ++c) {
      wait_chunk(c);
      ready_.store(c + 1, std::memory_order_release);
      slice(c, K - 1, K);   // the caller is worker K-1
    }
    // wait for the pool threads
    const uint32_t want = (uint32_t) th_.size();
    while (done_.load(std::memory_order_acquire) < want) _mm_pause();
  }

 private:
  WidenPool() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 8;
    unsigned local = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = (unsigned) std::max(1, atoi(e));
    // measured on the 128-thread host of the GPU box: a streaming-store thread sustains only
    // ~1.7 GB/s into the pinned result buffer, so the conversion scales with the thread count
    unsigned k = hw * 3 / (4 * local);
    if (const char* e = getenv("GL_HOST_THREADS")) k = (unsigned) std::max(1, atoi(e));
    k = std::max(1u, std::min(96u, k));
    for (unsigned i = 0; i + 1 < k; ++i) th_.emplace_back([this, i] { loop((int) i); });
  }
  ~WidenPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void slice(uint32_t c, int me, int K) {
    const uint64_t b = bounds_[c], e = bounds_[c + 1], n = e - b;
    const uint64_t lo = b + n * (uint64_t) me / (uint64_t) K, hi = b + n * (uint64_t) (me + 1) / (uint64_t) K;
    const uint8_t* in = in_;
    int64_t* out = out_;
    for (uint64_t i = lo; i < hi; ++i)
      _mm_stream_si64((long long*) (out + i), in[i] == 0xFFu ? (long long) INT64_MAX : (long long) in[i]);
  }
  void loop(int me) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || job_ != seen; });
        if (stop_) return;
        seen = job_;
      }
      const int K = threads();
      const uint32_t nch = nchunks_;
      for (uint32_t c = 0; c < nch; ++c) {
        while (ready_.load(std::memory_order_acquire) <= c) _mm_pause();
        slice(c, me, K);
      }
      _mm_sfence();
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
  uint64_t job_ = 0;
  const uint8_t* in_ = nullptr;
  int64_t* out_ = nullptr;
  const uint32_t* bounds_ = nullptr;
  uint32_t nchunks_ = 0;
  std::atomic<uint32_t> ready_{0}, done_{0};
};

}  // namespace gl
