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
// the stream completes the transfers.  No batched-memcpy API is involved: one plain cudaMemcpyAsync per chunk.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr size_t STAGE_CHUNK = 1u << 18;  // 256 KiB: ~5 us of PCIe time, ~30 us of one core's memcpy; a 9.6 MB map = 37 chunks
constexpr int STAGE_SLOTS = 32;
constexpr int STAGE_MAX_THREADS = 16;

struct Job {
  const char* src = nullptr;
  char* dst = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  int device = 0;
  std::atomic<size_t> next{0};   // next chunk index to claim
  size_t chunks = 0;
  std::atomic<size_t> done{0};
  std::atomic<int> error{0};
};

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
    job.dst = static_cast<char*>(dst);
    job.bytes = bytes;
    job.stream = stream;
    job.device = dev;
    job.chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      ++generation_;
      gen_atomic_.store(generation_, std::memory_order_release);
    }
    cv_.notify_all();
    work(job);  // the calling thread stages chunks too
    {
      std::unique_lock<std::mutex> lk(m_);
      done_cv_.wait(lk, [&] { return job.done.load() == job.chunks && active_ == 0; });
      job_ = nullptr;
    }
    return job.error.load();
  }

  int threads() const { return (int)workers_.size() + 1; }

 private:
  Stager() {
    // measured on the B200 box (16 cores, tools/h2d_staged_probe.py, 64 MB): 1 thread 8 GB/s, 2: 18, 4: 36, 8: 42, 16: 28
    // (torch's pageable copy: 14 GB/s, pinned: 55 GB/s) -> half of the cores, at most 8
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
      cudaFreeHost(ring_);
      ring_ = nullptr;
      for (auto& ev : events_) cudaEventDestroy(ev);
      events_.clear();
    }
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ring_), STAGE_SLOTS * STAGE_CHUNK, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      ring_ = nullptr;
      return (int)e;
    }
    events_.resize(STAGE_SLOTS);
    for (auto& ev : events_) {
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return (int)e;
    }
    used_.assign(STAGE_SLOTS, false);
    ring_dev_ = dev;
    return 0;
  }

  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      Job* job = nullptr;
      {
        // uploads come in bursts (four tensors per image pair): spin briefly before going to sleep, a condition-variable
        // wake-up costs more than staging a chunk
        for (int spin = 0; spin < 20000; ++spin) {
            if (stop_flag_.load(std::memory_order_relaxed) || gen_atomic_.load(std::memory_order_acquire) != seen) break;
        }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (job_ && generation_ != seen); });
        if (stop_) return;
        seen = generation_;
        job = job_;
        ++active_;
      }
      cudaSetDevice(job->device);
      work(*job);
      {
        std::lock_guard<std::mutex> lk(m_);
        --active_;
      }
      done_cv_.notify_all();
    }
  }

  // claim chunks until none is left: chunk i uses slot i mod SLOTS, strictly after the DMA of chunk i - SLOTS
  void work(Job& job) {
    for (;;) {
      const size_t i = job.next.fetch_add(1);
      if (i >= job.chunks) break;
      const int slot = (int)(i % STAGE_SLOTS);
      const size_t off = i * STAGE_CHUNK;
      const size_t len = (off + STAGE_CHUNK <= job.bytes) ? STAGE_CHUNK : job.bytes - off;
      char* pin = ring_ + (size_t)slot * STAGE_CHUNK;
      cudaError_t e = cudaSuccess;
      {
        // slot ownership: chunk i may touch the slot only after chunk i - SLOTS has recorded its event
        std::unique_lock<std::mutex> lk(slot_m_);
        slot_cv_.wait(lk, [&] { return slot_turn_[slot] == i / STAGE_SLOTS; });
      }
      if (used_[slot]) e = cudaEventSynchronize(events_[slot]);  // its previous DMA has drained
      if (e == cudaSuccess) {
        std::memcpy(pin, job.src + off, len);
        e = cudaMemcpyAsync(job.dst + off, pin, len, cudaMemcpyHostToDevice, job.stream);
      }
      if (e == cudaSuccess) e = cudaEventRecord(events_[slot], job.stream);
      used_[slot] = true;
      {
        std::lock_guard<std::mutex> lk(slot_m_);
        slot_turn_[slot] = i / STAGE_SLOTS + 1;
      }
      slot_cv_.notify_all();
      if (e != cudaSuccess) job.error.store((int)e);
      if (job.done.fetch_add(1) + 1 == job.chunks) {
        // last chunk of the call: reset the slot turns for the next call
        std::lock_guard<std::mutex> lk(slot_m_);
        for (auto& t : slot_turn_) t = 0;
      }
    }
    done_cv_.notify_all();
  }

  std::mutex call_mutex_, m_, slot_m_;
  std::condition_variable cv_, done_cv_, slot_cv_;
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
  std::vector<char> used_;
  size_t slot_turn_[STAGE_SLOTS] = {};
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
