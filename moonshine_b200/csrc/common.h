// Shared host-side helpers for the B200 Moonshine runtime.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <algorithm>
#include <thread>
#include <mutex>
#include <functional>
#include <condition_variable>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

namespace msb {

inline std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return std::string(buf);
}

#define MSB_LOGF(...)                                   \
  do {                                                  \
    fprintf(stderr, "[moonshine-b200] " __VA_ARGS__);   \
    fprintf(stderr, "\n");                              \
  } while (0)

#define CUDA_CHECK(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      throw std::runtime_error(msb::format("CUDA error %s at %s:%d: %s",        \
                                           cudaGetErrorName(_e), __FILE__,      \
                                           __LINE__, cudaGetErrorString(_e)));  \
    }                                                                           \
  } while (0)

// Model dimensions.  TINY/BASE follow core/moonshine-model.cpp:41-79 of the
// reference (6/8/36 and 8/8/52) plus the HF config the graphs were exported
// from; 100/101 are reduced-size unit-test configs.
struct Dims {
  int arch = 0;
  int dim = 0;         // D
  int enc_layers = 0;
  int dec_layers = 0;
  int heads = 0;
  int head_dim = 0;
  int ffn = 0;         // I
  int vocab = 32768;
  int rot_dim = 0;     // rotated dims per head (interleaved pairs)
  int rope_den = 0;    // int(head_dim * partial_rotary_factor)
  float rope_theta = 10000.0f;
  int bos = 1;
  int eos = 2;
  // --- streaming family (arch 2..5 and unit-test ids 102/103): dimensions are data, read from the
  // weight container's "streaming.config" record (the reference reads streaming_config.json,
  // core/moonshine-streaming-model.cpp:75-116) ---
  bool streaming = false;
  int enc_dim = 0;       // encoder hidden size E (adapter projects E -> D when different)
  int enc_ffn = 0;
  bool tied = true;      // logits off embed_tokens, else off proj_out
  int max_seq_len = 448;
  int max_pos_emb = 4096;
  int n_windows = 0;
  int win_past[16] = {0}, win_future[16] = {0};  // per encoder layer, inclusive
  int lookahead() const {
    int s = 0;
    for (int i = 0; i < n_windows; i++) s += win_future[i];
    return s;
  }
};

inline bool is_streaming_arch(uint32_t arch) { return (arch >= 2 && arch <= 5) || arch == 102 || arch == 103; }

// streaming.config: [version, E, D, enc_layers, dec_layers, H, hd, enc_ffn, ffn, vocab, rope_den, rot_dim,
//                    rope_theta, tied, max_seq_len, max_pos_emb, bos, eos, n_windows, (past, future)...]
inline Dims dims_from_streaming_config(uint32_t arch, const float* c, size_t n) {
  if (n < 19 || c[0] != 1.0f) throw std::runtime_error("streaming.config: unsupported record");
  Dims d;
  d.arch = (int)arch;
  d.streaming = true;
  d.enc_dim = (int)c[1]; d.dim = (int)c[2]; d.enc_layers = (int)c[3]; d.dec_layers = (int)c[4];
  d.heads = (int)c[5]; d.head_dim = (int)c[6]; d.enc_ffn = (int)c[7]; d.ffn = (int)c[8];
  d.vocab = (int)c[9]; d.rope_den = (int)c[10]; d.rot_dim = (int)c[11]; d.rope_theta = c[12];
  d.tied = c[13] != 0.0f; d.max_seq_len = (int)c[14]; d.max_pos_emb = (int)c[15];
  d.bos = (int)c[16]; d.eos = (int)c[17]; d.n_windows = (int)c[18];
  // every field is data from the weight file: range-check before anything divides or allocates by it
  auto in_range = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
  if (!in_range(d.enc_dim, 4, 8192) || !in_range(d.dim, 4, 8192) || !in_range(d.enc_layers, 1, 16) ||
      !in_range(d.dec_layers, 1, 64) || !in_range(d.heads, 1, 256) || !in_range(d.head_dim, 4, 1024) ||
      !in_range(d.enc_ffn, 4, 65536) || !in_range(d.ffn, 4, 65536) || !in_range(d.vocab, 3, 1 << 24) ||
      !in_range(d.rope_den, 1, d.head_dim) || !in_range(d.rot_dim, 2, d.head_dim) || (d.rot_dim & 1) ||
      !(d.rope_theta > 0.0f) || !in_range(d.max_seq_len, 1, 1 << 20) || !in_range(d.max_pos_emb, 1, 1 << 24) ||
      !in_range(d.bos, 0, d.vocab - 1) || !in_range(d.eos, 0, d.vocab - 1) ||
      (int64_t)d.heads * d.head_dim != d.dim || d.enc_dim % d.heads != 0)
    throw std::runtime_error("streaming.config: field out of range (heads * head_dim must equal the decoder size)");
  if (d.n_windows != d.enc_layers || d.n_windows > 16 || n < (size_t)19 + 2 * d.n_windows)
    throw std::runtime_error("streaming.config: one (past, future) window per encoder layer expected");
  for (int i = 0; i < d.n_windows; i++) {
    d.win_past[i] = (int)c[19 + 2 * i];
    d.win_future[i] = (int)c[20 + 2 * i];
    if (d.win_past[i] < 0 || d.win_future[i] < 0 || d.win_past[i] > 1 << 20 || d.win_future[i] > 1 << 20)
      throw std::runtime_error("streaming.config: negative attention window");
  }
  return d;
}

inline Dims dims_for_arch(uint32_t arch) {
  Dims d;
  d.arch = (int)arch;
  auto set = [&](int D, int el, int dl, int H, int hd, int I, int V) {
    d.dim = D; d.enc_layers = el; d.dec_layers = dl; d.heads = H;
    d.head_dim = hd; d.ffn = I; d.vocab = V;
    d.rope_den = (int)((double)hd * 0.9);   // HF: int(head_dim * partial_rotary_factor), in double like Python
    d.rot_dim = 2 * ((d.rope_den + 1) / 2); // arange(0, dim, 2) frequencies, one pair each
  };
  switch (arch) {
    case 0: set(288, 6, 6, 8, 36, 1152, 32768); break;   // MOONSHINE_MODEL_ARCH_TINY
    case 1: set(416, 8, 8, 8, 52, 1664, 32768); break;   // MOONSHINE_MODEL_ARCH_BASE
    case 100: set(64, 2, 2, 4, 16, 96, 512); break;      // unit-test config "test"
    case 101: set(72, 2, 3, 2, 36, 80, 300); break;      // unit-test config "test2"
    default:
      if (is_streaming_arch(arch))
        throw std::runtime_error(format("architecture %u takes its dimensions from the weight file", arch));
      throw std::runtime_error(format("Unsupported model architecture %u", arch));
  }
  return d;
}

// Frame counts of the conv frontend (no padding): k127/s64, k7/s3, k3/s2.
inline void frontend_lengths(int64_t n, int& t1, int& t2, int& t3) {
  t1 = n >= 127 ? (int)((n - 127) / 64 + 1) : 0;
  t2 = t1 >= 7 ? (t1 - 7) / 3 + 1 : 0;
  t3 = t2 >= 3 ? (t2 - 3) / 2 + 1 : 0;
}

// Per-device cache for cudaFuncSetAttribute(MaxDynamicSharedMemorySize): the attribute belongs to the (function,
// device) pair, and one process may hold transcribers on several devices (option `device`).
struct SmemAttrCache {
  size_t configured[64] = {0};
  // true when `bytes` exceeds what this device has been configured for (and records it)
  bool needs(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (bytes <= configured[dev]) return false;
    configured[dev] = bytes;
    return true;
  }
};

template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  // Grow-only allocation.
  void reserve(size_t n) {
    if (n <= count) return;
    release();
    CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
    count = n;
  }
  size_t bytes() const { return count * sizeof(T); }
};

template <typename T>
struct PinnedBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  ~PinnedBuffer() { if (ptr) cudaFreeHost(ptr); }
  void reserve(size_t n) {
    if (n <= count) return;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    CUDA_CHECK(cudaMallocHost(&ptr, n * sizeof(T)));
    count = n;
  }
};

// A small persistent pool for the host-side fan-out of one call (segmentation of a batch, staging copies): spawning
// std::threads per call cost 20-50 us each, i.e. ~0.5 ms of a 24 ms batch.  parallel_for(n, fn) runs fn(i) for i in [0, n)
// on the workers plus the calling thread and returns when all are done; the first exception is rethrown in the caller.
class WorkerPool {
 public:
  static WorkerPool& instance() {
    static WorkerPool pool(std::max(1u, std::min(16u, std::thread::hardware_concurrency())) - 1);
    return pool;
  }
  template <typename F>
  void parallel_for(int n, F&& fn) {
    if (n <= 0) return;
    if (n == 1 || workers_.empty()) {
      for (int i = 0; i < n; i++) fn(i);
      return;
    }
    Job job;
    job.n = n;
    job.fn = [&fn](int i) { fn(i); };
    {
      std::lock_guard<std::mutex> lock(mu_);
      jobs_.push_back(&job);
    }
    cv_.notify_all();
    run(job);  // the caller works too
    {
      std::unique_lock<std::mutex> lock(mu_);
      done_cv_.wait(lock, [&] { return job.finished == job.n && job.pinned == 0; });
      jobs_.erase(std::find(jobs_.begin(), jobs_.end(), &job));
    }
    if (job.error) std::rethrow_exception(job.error);
  }

 private:
  struct Job {
    int n = 0;
    std::atomic<int> next{0};
    int finished = 0;  // guarded by mu_
    int pinned = 0;    // workers currently holding a pointer to this job (guarded by mu_): the caller's frame outlives them
    std::function<void(int)> fn;
    std::exception_ptr error;
  };
  explicit WorkerPool(unsigned n) {
    for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void run(Job& job) {
    int mine = 0;
    std::exception_ptr err;
    for (;;) {
      const int i = job.next.fetch_add(1);
      if (i >= job.n) break;
      try {
        job.fn(i);
      } catch (...) {
        if (!err) err = std::current_exception();
      }
      mine++;
    }
    if (mine || err) {
      std::lock_guard<std::mutex> lock(mu_);
      job.finished += mine;
      if (err && !job.error) job.error = err;
      if (job.finished == job.n) done_cv_.notify_all();
    }
  }
  void loop() {
    for (;;) {
      Job* job = nullptr;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] {
          if (stop_) return true;
          for (Job* j : jobs_)
            if (j->next.load() < j->n) return true;
          return false;
        });
        if (stop_) return;
        for (Job* j : jobs_)
          if (j->next.load() < j->n) { job = j; break; }
        if (job) job->pinned++;
      }
      if (job) {
        run(*job);
        std::lock_guard<std::mutex> lock(mu_);
        job->pinned--;
        done_cv_.notify_all();
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<Job*> jobs_;
  std::vector<std::thread> workers_;
  bool stop_ = false;
};

}  // namespace msb
