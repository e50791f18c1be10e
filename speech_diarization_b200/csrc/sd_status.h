// sd_status.h — status codes (shared with include/sd_b200.h) and the thread-local
// error string behind sd_last_error().
#pragma once
#include "../../include/sd_b200.h"
#include <cstdarg>
#include <cstdio>
#include <atomic>

namespace sd {
// kernels launched by this library since load (bench.py reports it as gpu_launches)
inline std::atomic<long>& launch_counter() {
  static std::atomic<long> c{0};
  return c;
}
inline void count_launch(int n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace sd

#define SD_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return sd::fail(SD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                                \
  } while (0)
#define SD_TRY(expr)            \
  do {                          \
    int s__ = (expr);           \
    if (s__ != SD_OK) return s__; \
  } while (0)
