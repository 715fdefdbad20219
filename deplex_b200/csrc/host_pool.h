// host_pool.h -- a small persistent thread pool for the host side of the narrow label transport.
//
// The host-pointer entry points can bring the labels back over PCIe as uint16 (2 B/pixel instead of 4) and widen them
// into the caller's int32 buffer here, overlapped with the copies of the following chunks.  Pure host plumbing: no
// arithmetic on labels other than the zero extension.
#pragma once
#include <immintrin.h>

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace dpx {

// uint16 -> int32 with 256-bit loads / streaming stores (chosen at run time when the CPU has AVX2)
__attribute__((target("avx2"))) inline size_t widen_u16_to_i32_avx2(const uint16_t* src, int32_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_cvtepu16_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(src + i)));
    const __m256i b = _mm256_cvtepu16_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(src + i + 8)));
    const __m256i c = _mm256_cvtepu16_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(src + i + 16)));
    const __m256i d = _mm256_cvtepu16_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(src + i + 24)));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 16), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 24), d);
  }
  _mm_sfence();
  return i;
}

// uint16 -> int32, streaming stores when the destination allows (the widened labels are not read again here)
inline void widen_u16_to_i32(const uint16_t* src, int32_t* dst, size_t n) {
  size_t i = 0;
  static const bool has_avx2 = __builtin_cpu_supports("avx2");
  if (has_avx2 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0)
    i = widen_u16_to_i32_avx2(src, dst, n);
  if ((reinterpret_cast<uintptr_t>(dst + i) & 15u) == 0 && (reinterpret_cast<uintptr_t>(src + i) & 15u) == 0) {
    const __m128i zero = _mm_setzero_si128();
    for (; i + 8 <= n; i += 8) {
      const __m128i v = _mm_load_si128(reinterpret_cast<const __m128i*>(src + i));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi16(v, zero));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), _mm_unpackhi_epi16(v, zero));
    }
    _mm_sfence();
  }
  for (; i < n; ++i) dst[i] = src[i];
}

class HostPool {
 public:
  explicit HostPool(int n_threads) {
    for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { run(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;

  int threads() const { return static_cast<int>(workers_.size()); }

  // Queue the widening of n labels, split into one slice per worker; returns a ticket for wait().
  uint64_t widen(const uint16_t* src, int32_t* dst, size_t n) {
    const size_t parts = workers_.empty() ? 1 : workers_.size();
    const size_t step = ((n + parts - 1) / parts + 63) & ~static_cast<size_t>(63);
    std::lock_guard<std::mutex> lk(m_);
    const uint64_t ticket = ++issued_;
    for (size_t b = 0; b < n; b += step) {
      jobs_.push_back({src + b, dst + b, (n - b < step ? n - b : step), ticket});
      ++pending_[ticket % kTickets];
    }
    cv_.notify_all();
    return ticket;
  }

  // Block until every slice of `ticket` has been written.  (At most kTickets tickets may be outstanding.)
  void wait(uint64_t ticket) {
    if (ticket == 0) return;
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_[ticket % kTickets] == 0; });
  }

 private:
  struct Job {
    const uint16_t* src;
    int32_t* dst;
    size_t n;
    uint64_t ticket;
  };
  static constexpr int kTickets = 16;

  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
        if (jobs_.empty()) return;
        j = jobs_.front();
        jobs_.pop_front();
      }
      widen_u16_to_i32(j.src, j.dst, j.n);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_[j.ticket % kTickets] == 0) done_cv_.notify_all();
      }
    }
  }

  std::vector<std::thread> workers_;
  std::deque<Job> jobs_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  int pending_[kTickets] = {};
  uint64_t issued_ = 0;
  bool stop_ = false;
};

}  // namespace dpx
