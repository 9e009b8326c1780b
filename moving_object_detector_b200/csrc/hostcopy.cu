// hostcopy.cu -- see hostcopy.cuh.  Host code only (no kernels).
#include "hostcopy.cuh"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define MOD_HAVE_SSE2_STREAM 1
#endif

namespace mod {
namespace {

inline void cpu_relax() {
#ifdef MOD_HAVE_SSE2_STREAM
  _mm_pause();  // spin-wait hint: leaves the core's resources to the sibling hyper-thread
#endif
}

// memcpy whose stores bypass the cache (SSE2 non-temporal, baseline x86-64): the destination - the pinned window on
// the way in, the message buffer on the way out - is not read again by this thread, so it should neither be fetched
// for ownership nor displace the source lines.  One third less memory traffic than a cached copy.
void stream_copy(void* dst, const void* src, size_t n) {
#ifdef MOD_HAVE_SSE2_STREAM
  uint8_t* d = static_cast<uint8_t*>(dst);
  const uint8_t* s = static_cast<const uint8_t*>(src);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (n < 256 + head) {
    std::memcpy(d, s, n);
    return;
  }
  std::memcpy(d, s, head);
  d += head;
  s += head;
  n -= head;
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; i++, d += 64, s += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
  }
  _mm_sfence();
  std::memcpy(d, s, n - blocks * 64);
#else
  std::memcpy(dst, src, n);
#endif
}

// A fork-join pool: run(n, fn) executes fn(0..n-1) on the workers and the calling thread and returns when all are
// done.  Workers spin for a short while after a job, then sleep on a condition variable.
class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool pool;
    return pool;
  }
  ~CopyPool() { resize(0); }

  int threads() {
    std::lock_guard<std::mutex> l(cfg_mu_);
    if (!configured_) configure_locked(default_threads());
    return threads_;
  }
  void set_threads(int n) {
    std::lock_guard<std::mutex> l(cfg_mu_);
    configure_locked(n);
  }

  void run(int n, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mu_);  // one fork-join at a time (contexts may live on several threads)
    if (workers_.empty() || n <= 1) {
      for (int i = 0; i < n; i++) fn(i);
      return;
    }
    Job job;
    job.fn = &fn;
    job.n = n;
    job.pending.store(n, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> l(mu_);
      job_ = &job;
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_job_.notify_all();
    work(job);
    while (job.pending.load(std::memory_order_acquire) > 0) cpu_relax();  // stragglers are inside one chunk's memcpy
    {
      std::lock_guard<std::mutex> l(mu_);  // from here on no worker can pick the job up ...
      job_ = nullptr;
    }
    while (job.active.load(std::memory_order_acquire) > 0) cpu_relax();  // ... and those that did have left it
  }

 private:
  static int default_threads() {
    if (const char* e = std::getenv("MOD_B200_COPY_THREADS")) return std::atoi(e);
    const unsigned hw = std::thread::hardware_concurrency();
    int n = static_cast<int>(hw / 2);
    if (n > 4) n = 4;
    if (n < 1) n = 1;
    return n;
  }
  void configure_locked(int n) {
    if (n < 0) n = 0;
    if (n > 16) n = 16;
    resize(n > 0 ? n - 1 : 0);  // the calling thread is one of the n
    threads_ = n;
    configured_ = true;
  }
  void resize(int n_workers) {
    std::lock_guard<std::mutex> serial(run_mu_);
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_job_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
    stop_ = false;
    for (int i = 0; i < n_workers; i++) workers_.emplace_back([this] { worker(); });
  }
  struct Job {
    const std::function<void(int)>* fn = nullptr;
    int n = 0;
    std::atomic<int> next{0}, pending{0}, active{0};
  };
  static void work(Job& job) {
    for (;;) {
      const int i = job.next.fetch_add(1, std::memory_order_acq_rel);
      if (i >= job.n) return;
      (*job.fn)(i);
      job.pending.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  void worker() {
    uint64_t seen = epoch_.load(std::memory_order_acquire);
    for (;;) {
      // spin briefly for the next wave (the stager issues one every ~100 us), then sleep
      const auto t0 = std::chrono::steady_clock::now();
      bool woke = false;
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
        if (epoch_.load(std::memory_order_acquire) != seen) { woke = true; break; }
        cpu_relax();
      }
      Job* job = nullptr;
      {
        std::unique_lock<std::mutex> l(mu_);
        // stop_ is part of the condition: a worker that starts after the epoch of a resize was published would
        // otherwise record that epoch as seen and sleep through its own stop request
        if (!woke) cv_job_.wait(l, [&] { return stop_ || epoch_.load(std::memory_order_acquire) != seen; });
        seen = epoch_.load(std::memory_order_acquire);
        if (stop_) return;
        job = job_;  // registering under mu_ keeps the job alive until active drops back (see run)
        if (job) job->active.fetch_add(1, std::memory_order_acq_rel);
      }
      if (job) {
        work(*job);
        job->active.fetch_sub(1, std::memory_order_acq_rel);
      }
    }
  }

  std::mutex cfg_mu_, run_mu_, mu_;
  std::condition_variable cv_job_;
  std::vector<std::thread> workers_;
  Job* job_ = nullptr;  // guarded by mu_
  std::atomic<uint64_t> epoch_{0};
  bool stop_ = false, configured_ = false;
  int threads_ = 0;
};

}  // namespace

int host_copy_threads() { return CopyPool::get().threads(); }
void set_host_copy_threads(int n) { CopyPool::get().set_threads(n); }

HostStager::~HostStager() {
  for (cudaEvent_t& e : done_)
    if (e) {
      cudaEventSynchronize(e);
      cudaEventDestroy(e);
    }
  if (window_) cudaFreeHost(window_);
}

bool HostStager::is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

cudaError_t HostStager::ensure(int threads) {
  if (threads == threads_ && window_) return cudaSuccess;
  cudaError_t e;
  for (cudaEvent_t& ev : done_) {
    if (!ev) {
      if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    } else if ((e = cudaEventSynchronize(ev)) != cudaSuccess) {
      return e;
    }
  }
  if (window_) {
    cudaFreeHost(window_);
    window_ = nullptr;
  }
  if ((e = cudaMallocHost(reinterpret_cast<void**>(&window_), 2 * static_cast<size_t>(threads) * kChunk)) != cudaSuccess)
    return e;
  threads_ = threads;
  return cudaSuccess;
}

cudaError_t HostStager::h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  // (the pool is only started when a pageable buffer actually shows up)
  const int T = (bytes < kMinBytes || is_pinned(h_src)) ? 0 : host_copy_threads();
  if (T <= 0) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream);
  cudaError_t e = ensure(T);
  if (e != cudaSuccess) return e;
  const size_t wave = static_cast<size_t>(T) * kChunk;
  const uint8_t* src = static_cast<const uint8_t*>(h_src);
  uint8_t* dst = static_cast<uint8_t*>(d_dst);
  int half = 0;
  for (size_t off = 0; off < bytes; off += wave, half ^= 1) {
    const size_t len = bytes - off < wave ? bytes - off : wave;
    uint8_t* stage = window_ + half * wave;
    if ((e = cudaEventSynchronize(done_[half])) != cudaSuccess) return e;  // the DMA that last read this half
    const int parts = static_cast<int>((len + kChunk - 1) / kChunk);
    CopyPool::get().run(parts, [&](int i) {
      const size_t o = static_cast<size_t>(i) * kChunk;
      stream_copy(stage + o, src + off + o, len - o < kChunk ? len - o : kChunk);
    });
    if ((e = cudaMemcpyAsync(dst + off, stage, len, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(done_[half], stream)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t HostStager::d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  const int T = (bytes < kMinBytes || is_pinned(h_dst)) ? 0 : host_copy_threads();
  if (T <= 0) {
    const cudaError_t e = cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(stream);
  }
  cudaError_t e = ensure(T);
  if (e != cudaSuccess) return e;
  const size_t wave = static_cast<size_t>(T) * kChunk;
  const size_t n_waves = (bytes + wave - 1) / wave;
  const uint8_t* src = static_cast<const uint8_t*>(d_src);
  uint8_t* dst = static_cast<uint8_t*>(h_dst);
  auto issue = [&](size_t k) -> cudaError_t {  // DMA of wave k into half k & 1
    const size_t off = k * wave;
    const size_t len = bytes - off < wave ? bytes - off : wave;
    const int half = static_cast<int>(k & 1);
    cudaError_t r = cudaMemcpyAsync(window_ + half * wave, src + off, len, cudaMemcpyDeviceToHost, stream);
    return r != cudaSuccess ? r : cudaEventRecord(done_[half], stream);
  };
  // both halves may still be the source of an earlier h2d's DMA
  for (int half = 0; half < 2; half++)
    if ((e = cudaEventSynchronize(done_[half])) != cudaSuccess) return e;
  for (size_t k = 0; k < n_waves && k < 2; k++)
    if ((e = issue(k)) != cudaSuccess) return e;
  for (size_t k = 0; k < n_waves; k++) {
    const int half = static_cast<int>(k & 1);
    if ((e = cudaEventSynchronize(done_[half])) != cudaSuccess) return e;
    const size_t off = k * wave;
    const size_t len = bytes - off < wave ? bytes - off : wave;
    const uint8_t* stage = window_ + half * wave;
    const int parts = static_cast<int>((len + kChunk - 1) / kChunk);
    CopyPool::get().run(parts, [&](int i) {
      const size_t o = static_cast<size_t>(i) * kChunk;
      stream_copy(dst + off + o, stage + o, len - o < kChunk ? len - o : kChunk);
    });
    if (k + 2 < n_waves && (e = issue(k + 2)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace mod
