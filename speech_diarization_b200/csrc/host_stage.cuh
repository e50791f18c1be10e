// host_stage.cuh — upload of PAGEABLE host audio through a page-locked staging ring.
//
// The reference's callers hand `ecapa_encode_batch` ordinary numpy memory (np.stack-ed window batches,
// /root/reference/anti_stick_diarize.py:164-166 and :423-424).  cudaMemcpyAsync from pageable memory is a
// synchronous, driver-staged copy on the calling thread: nothing overlaps it.  Here the span is cut into
// pieces; a small pool of worker threads copies piece i into slot i % NSLOT of a pinned ring (several
// memcpy streams in parallel: one core does not saturate PCIe 5), the calling thread queues the slot's
// asynchronous H2D copy as soon as the piece is staged, and the caller's kernels for a chunk of windows start
// when that chunk's samples have landed.  Page-locked callers' buffers bypass all this (direct DMA).
#pragma once
#include <cuda_runtime.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <stdlib.h>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace sd {

class StagingRing {
 public:
  static constexpr size_t kPiece = size_t(2) << 20;   // bytes per piece / slot
  static constexpr int kSlots = 12;

  ~StagingRing() { shutdown(); }

  // false (and a cleared CUDA error) when the pinned ring cannot be allocated: the caller falls back to the
  // driver's own staged copy
  bool init(int n_threads) {
    if (ring_) return true;
    if (cudaHostAlloc(reinterpret_cast<void**>(&ring_), kPiece * kSlots, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      ring_ = nullptr;
      return false;
    }
    for (int i = 0; i < kSlots; ++i) {
      if (cudaEventCreateWithFlags(&slot_ev_[i], cudaEventDisableTiming) != cudaSuccess) return false;
      slot_used_[i] = false;
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 16) n_threads = 16;
    stop_ = false;
    for (int t = 0; t < n_threads; ++t) workers_.emplace_back([this] { work(); });
    return true;
  }

  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
    if (ring_) {
      for (int i = 0; i < kSlots; ++i) cudaEventDestroy(slot_ev_[i]);
      cudaFreeHost(ring_);
      ring_ = nullptr;
    }
  }

  // Copies src[0, bytes) (pageable) to dst_dev on `copy_stream`.  `on_bytes(done)` is called on the calling
  // thread each time the H2D copies queued so far cover `done` bytes (monotonic), so the caller can record an
  // event / launch the consumers of a finished chunk.  Returns a CUDA error code.
  template <typename F>
  cudaError_t upload(const char* src, char* dst_dev, size_t bytes, cudaStream_t copy_stream, F&& on_bytes) {
    const size_t n_pieces = (bytes + kPiece - 1) / kPiece;
    size_t issued = 0, queued = 0;   // pieces handed to the workers / pieces whose H2D copy is queued
    while (queued < n_pieces) {
      // keep the workers fed: a slot may be refilled once the H2D copy that last read it has finished
      while (issued < n_pieces && issued < queued + kSlots) {
        const int slot = static_cast<int>(issued % kSlots);
        if (slot_used_[slot]) {
          cudaError_t e = cudaEventSynchronize(slot_ev_[slot]);
          if (e != cudaSuccess) return e;
        }
        const size_t off = issued * kPiece;
        const size_t len = bytes - off < kPiece ? bytes - off : kPiece;
        done_[slot].store(0, std::memory_order_relaxed);
        {
          std::lock_guard<std::mutex> lk(mu_);
          jobs_.push_back(Job{src + off, ring_ + slot * kPiece, len, &done_[slot]});
        }
        cv_.notify_one();
        ++issued;
      }
      const int slot = static_cast<int>(queued % kSlots);
      while (done_[slot].load(std::memory_order_acquire) == 0) std::this_thread::yield();
      const size_t off = queued * kPiece;
      const size_t len = bytes - off < kPiece ? bytes - off : kPiece;
      cudaError_t e = cudaMemcpyAsync(dst_dev + off, ring_ + slot * kPiece, len, cudaMemcpyHostToDevice, copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(slot_ev_[slot], copy_stream);
      if (e != cudaSuccess) return e;
      slot_used_[slot] = true;
      ++queued;
      on_bytes(off + len);
    }
    return cudaSuccess;
  }

 private:
  struct Job {
    const char* src;
    char* dst;
    size_t len;
    std::atomic<int>* done;
  };
  // Copy into the pinned ring with NON-TEMPORAL stores: a plain memcpy leaves the 2 MB piece dirty in the copying
  // core's cache, and the DMA engine then reads it through cache snoops at ~15 GB/s (measured: the H2D copies of a
  // 24.6 MB batch took 1.4 ms from the ring against 0.43 ms from cold page-locked memory); streamed past the cache
  // it is read from DRAM at the PCIe rate.  dst is 64-byte aligned (ring slots), src is not.
  static void stream_copy(char* dst, const char* src, size_t len) {
#if defined(__x86_64__) || defined(_M_X64)
    if (!stream_stores()) {
      std::memcpy(dst, src, len);
      return;
    }
    size_t i = 0;
    for (; i + 64 <= len; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    if (i < len) std::memcpy(dst + i, src + i, len - i);
    _mm_sfence();
#else
    std::memcpy(dst, src, len);
#endif
  }
  static bool stream_stores() {   // SD_ECAPA_STAGING_NT=0: plain memcpy (A/B)
    static const bool on = [] { const char* e = getenv("SD_ECAPA_STAGING_NT"); return !e || atoi(e) != 0; }();
    return on;
  }

  void work() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
        if (stop_ && jobs_.empty()) return;
        j = jobs_.front();
        jobs_.erase(jobs_.begin());
      }
      stream_copy(j.dst, j.src, j.len);
      j.done->store(1, std::memory_order_release);
    }
  }

  char* ring_ = nullptr;
  cudaEvent_t slot_ev_[kSlots] = {};
  bool slot_used_[kSlots] = {};
  std::atomic<int> done_[kSlots];
  std::vector<std::thread> workers_;
  std::vector<Job> jobs_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
};

}  // namespace sd
