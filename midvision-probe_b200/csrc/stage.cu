// stage.cu -- host -> device upload of PAGEABLE host memory through a pinned staging ring filled by worker threads.
//
// The reference's callers hand the matching helpers plain CPU tensors (`.detach().cpu()`,
// evaluate_navi_correspondence.py:149-150, render_scannet_correspondence.py:201): pageable memory.  A
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at ~10 GB/s -- a fifth of the
// PCIe 5 link -- which made the helper's upload (19.6 MB per NAVI-shaped pair) three times slower than from pinned
// memory.  mv_h2d_staged cuts the source into chunks; a small pool of worker threads copies chunk i into slot
// i mod SLOTS of a pinned ring and issues the chunk's own cudaMemcpyAsync on the caller's stream, so the host-side
// memcpy of several chunks and the DMA of earlier ones run at the same time.  A slot is reused once the event recorded
// after its DMA has fired.  The call returns when every chunk has been ISSUED (the source may then be modified);
// the stream completes the transfers.  No batched-memcpy API is involved: plain cudaMemcpyAsync calls, one per run of
// adjacent staged chunks.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr size_t STAGE_CHUNK = 1u << 18;  // 256 KiB: ~5 us of PCIe time, ~30 us of one core's memcpy; a 9.6 MB map = 37 chunks
constexpr int STAGE_SLOTS = 64;           // 16 MiB pinned ring
constexpr int STAGE_MAX_THREADS = 16;
constexpr int STAGE_MAX_BATCH = 8;        // ready chunks that are adjacent in the ring go out as ONE cudaMemcpyAsync (<= 2 MiB)
constexpr int STAGE_EVENTS = 64;

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

struct Job {
  const char* src = nullptr;
  size_t bytes = 0;
  size_t chunks = 0;
  unsigned long long base = 0;            // ring position (global chunk id) of chunk 0
  std::atomic<size_t> next{0};            // next chunk index to claim
  std::atomic<unsigned char>* ready = nullptr;  // per chunk: staged into its ring slot
};

// Round 2: the workers ONLY copy into the ring; every CUDA call (one cudaMemcpyAsync per run of adjacent ready chunks, one event
// per run, event queries that free ring slots) comes from the calling thread.  With eight threads each issuing their own
// chunk's cudaMemcpyAsync + cudaEventRecord + cudaEventSynchronize the driver calls serialised on the context lock and a 9.6 MB
// map uploaded at 14 GB/s (64 MB: 25 GB/s); the ring position now runs on across calls, so consecutive tensors need no reset.
class Stager {
 public:
  static Stager& get() {
    static Stager s;
    return s;
  }

  int upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return 0;
    std::lock_guard<std::mutex> serial(call_mutex_);  // one upload at a time per process: the ring is shared
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int rc = ensure(dev);
    if (rc) return rc;
    Job job;
    job.src = static_cast<const char*>(src);
    job.bytes = bytes;
    job.chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    job.base = next_id_;
    std::unique_ptr<std::atomic<unsigned char>[]> ready(new std::atomic<unsigned char>[job.chunks]);
    for (size_t i = 0; i < job.chunks; ++i) ready[i].store(0, std::memory_order_relaxed);
    job.ready = ready.get();
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      ++generation_;
      gen_atomic_.store(generation_, std::memory_order_release);
    }
    cv_.notify_all();

    // ---- the issuer: DMAs go out in chunk order as soon as their chunks are staged
    char* d = static_cast<char*>(dst);
    int err = 0;
    size_t i = 0;
    while (i < job.chunks) {
      poll_drained();
      if (!job.ready[i].load(std::memory_order_acquire)) {
        if (workers_.empty()) stage_one(job);  // single-threaded configuration: the caller stages too
        else cpu_relax();
        continue;
      }
      size_t j = i + 1;
      while (j < job.chunks && j - i < (size_t)STAGE_MAX_BATCH && ((job.base + j) % STAGE_SLOTS) != 0 &&
             job.ready[j].load(std::memory_order_acquire))
        ++j;
      const size_t off = i * STAGE_CHUNK;
      const size_t len = (j * STAGE_CHUNK <= bytes ? j * STAGE_CHUNK : bytes) - off;
      const char* pin = ring_ + (size_t)((job.base + i) % STAGE_SLOTS) * STAGE_CHUNK;
      e = cudaMemcpyAsync(d + off, pin, len, cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) {
        if (pending_.size() >= (size_t)STAGE_EVENTS) {  // every event is in flight: wait for the oldest
          cudaEventSynchronize(events_[pending_.front().ev]);
          poll_drained();
        }
        const int ev = ev_next_++ % STAGE_EVENTS;
        e = cudaEventRecord(events_[ev], stream);
        pending_.push_back({ev, job.base + j});
      }
      if (e != cudaSuccess && !err) err = (int)e;
      i = j;
    }
    next_id_ = job.base + job.chunks;
    {
      std::unique_lock<std::mutex> lk(m_);
      done_cv_.wait(lk, [&] { return active_ == 0; });  // no worker still looks at this job
      job_ = nullptr;
    }
    return err;
  }

  int threads() const { return (int)workers_.size() + 1; }

 private:
  Stager() {
    // measured on the B200 box (16 cores, tools/h2d_staged_probe.py): the copy threads saturate around 8 (torch's pageable
    // copy: 12-20 GB/s, pinned: 55 GB/s) -> half of the cores, at most 8, one of them the issuing caller
    const int hw = (int)std::thread::hardware_concurrency();
    int n = hw > 0 ? (hw / 2 < 8 ? hw / 2 : 8) : 4;
    if (const char* env = getenv("MVMATCH_STAGE_THREADS")) n = atoi(env);
    if (hw > 0 && n > hw) n = hw;
    if (n < 1) n = 1;
    if (n > STAGE_MAX_THREADS) n = STAGE_MAX_THREADS;
    for (int i = 0; i < n - 1; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~Stager() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      stop_flag_.store(true);
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
    // the pinned ring and the events are left to process teardown (the CUDA context may already be gone here)
  }

  int ensure(int dev) {
    if (ring_ && ring_dev_ == dev) return 0;
    if (ring_) {  // another device became current: the events belong to the old one
      for (auto& pb : pending_) cudaEventSynchronize(events_[pb.ev]);
      pending_.clear();
      drained_.store(next_id_, std::memory_order_release);
      cudaFreeHost(ring_);
      ring_ = nullptr;
      for (auto& ev : events_) cudaEventDestroy(ev);
      events_.clear();
    }
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ring_), (size_t)STAGE_SLOTS * STAGE_CHUNK, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      ring_ = nullptr;
      return (int)e;
    }
    events_.resize(STAGE_EVENTS);
    for (auto& ev : events_) {
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return (int)e;
    }
    ring_dev_ = dev;
    return 0;
  }

  // DMAs whose event has fired free their ring slots: every chunk id below drained_ may be overwritten
  void poll_drained() {
    while (!pending_.empty()) {
      if (cudaEventQuery(events_[pending_.front().ev]) != cudaSuccess) {
        (void)cudaGetLastError();  // cudaErrorNotReady is not an error: do not leave it for the next launch check
        break;
      }
      drained_.store(pending_.front().end_id, std::memory_order_release);
      pending_.pop_front();
    }
  }

  // claim one chunk and copy it into its ring slot (strictly after the DMA of the chunk that used the slot before)
  bool stage_one(Job& job) {
    const size_t i = job.next.fetch_add(1);
    if (i >= job.chunks) return false;
    const unsigned long long g = job.base + i;
    while (g >= drained_.load(std::memory_order_acquire) + STAGE_SLOTS) {
      if (workers_.empty()) poll_drained();  // single-threaded: nobody else polls
      else cpu_relax();
      if (stop_flag_.load(std::memory_order_relaxed)) return false;
    }
    const size_t off = i * STAGE_CHUNK;
    const size_t len = (off + STAGE_CHUNK <= job.bytes) ? STAGE_CHUNK : job.bytes - off;
    std::memcpy(ring_ + (size_t)(g % STAGE_SLOTS) * STAGE_CHUNK, job.src + off, len);
    job.ready[i].store(1, std::memory_order_release);
    return true;
  }

  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      Job* job = nullptr;
      {
        // uploads come in bursts (four tensors per image pair, a pair every millisecond): spin for ~300 us before going to
        // sleep, a condition-variable wake-up costs more than staging several chunks
        const auto t0 = std::chrono::steady_clock::now();
        for (int spin = 0;; ++spin) {
          if (stop_flag_.load(std::memory_order_relaxed) || gen_atomic_.load(std::memory_order_acquire) != seen) break;
          cpu_relax();
          if ((spin & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
        }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (job_ && generation_ != seen); });
        if (stop_) return;
        seen = generation_;
        job = job_;
        ++active_;
      }
      while (stage_one(*job)) {
      }
      {
        std::lock_guard<std::mutex> lk(m_);
        --active_;
      }
      done_cv_.notify_all();
    }
  }

  struct PendingBatch {
    int ev;
    unsigned long long end_id;  // chunk ids below this are drained once the event has fired
  };

  std::mutex call_mutex_, m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  Job* job_ = nullptr;
  unsigned long long generation_ = 0;
  int active_ = 0;
  bool stop_ = false;
  std::atomic<bool> stop_flag_{false};
  std::atomic<unsigned long long> gen_atomic_{0};
  char* ring_ = nullptr;
  int ring_dev_ = -1;
  std::vector<cudaEvent_t> events_;
  std::deque<PendingBatch> pending_;       // issued DMA runs, oldest first (caller thread only)
  unsigned long long next_id_ = 0;         // ring position of the next call's chunk 0 (caller thread only)
  unsigned int ev_next_ = 0;
  std::atomic<unsigned long long> drained_{0};
};

}  // namespace

extern "C" {

int mv_h2d_staged(void* dst_device, const void* src_host, size_t bytes, mv_stream_t stream) {
  MV_REQUIRE(bytes == 0 || (dst_device && src_host), MV_E_ARG, "mv_h2d_staged: null pointer");
  const int rc = Stager::get().upload(dst_device, src_host, bytes, mv_cuda_stream(stream));
  if (rc != 0) mv_set_error("mv_h2d_staged: CUDA error %d (%s)", rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}

int mv_h2d_staged_threads(void) { return Stager::get().threads(); }

}  // extern "C"
