// Shared host/device helpers for libb2of (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <string>

#include "../../include/b2of.h"

namespace b2of {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

#define B2OF_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return b2of::fail(B2OF_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// cv2-style assertion: same text shape the reference's user would have seen from cv2.error
#define B2OF_ASSERT(cond, fn)                                                                    \
  do {                                                                                           \
    if (!(cond)) return b2of::fail(B2OF_E_BADARG, "(-215:Assertion failed) %s in function '%s'", #cond, fn); \
  } while (0)

#define B2OF_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    b2of::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      return b2of::fail(B2OF_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__host__ __device__ inline int reflect101(int i, int n) {
  // BORDER_REFLECT_101, valid for any offset.  In range and single reflections (every access of the kernels except
  // on images smaller than a filter) take no division.
  if ((unsigned)i < (unsigned)n) return i;
  if (n == 1) return 0;
  if (i < 0 && -i < n) return -i;
  if (i >= n && i < 2 * n - 1) return 2 * n - 2 - i;
  int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - i : i;
}

__host__ __device__ inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// round-half-to-even, cv::cvRound for doubles
inline int cv_round(double v) { return (int)nearbyint(v); }

// ---- optional per-kernel timing (CUDA events on the launching stream); off by default ----
enum ProfTag {
  PT_GRAY = 0, PT_PYRDOWN, PT_FB_LEVEL_H, PT_FB_LEVEL_V, PT_FB_POLYEXP, PT_FB_ITER_FINEST, PT_FB_ITER_COARSE,
  PT_LK_SCHARR, PT_LK_TRACK, PT_GFTT_MINEIG, PT_GFTT_NMS, PT_GFTT_SELECT, PT_FILTER, PT_STATS, PT_FB_UPSAMPLE, PT_COUNT
};
extern std::atomic<int> g_prof_on;
void prof_mark(int tag, cudaStream_t st, bool end, double bytes);
struct ProfScope {
  int tag; cudaStream_t st; bool on;
  ProfScope(int t, cudaStream_t s, double bytes = 0) : tag(t), st(s), on(g_prof_on.load(std::memory_order_relaxed) != 0) {
    if (on) prof_mark(tag, st, false, bytes);
  }
  ~ProfScope() { if (on) prof_mark(tag, st, true, 0); }
};

// Function attributes (the dynamic shared-memory opt-in) apply to the device that is current when they are set, so
// every "set it once" site keeps one record per device ordinal.
constexpr int B2OF_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 || dev >= B2OF_MAX_DEVICES ? 0 : dev;
}
struct PerDeviceOnce {            // first(): true exactly once per device (benign race: the guarded calls are idempotent)
  std::atomic<unsigned long long> mask{0};
  bool first() {
    const unsigned long long bit = 1ull << current_device();
    return (mask.fetch_or(bit) & bit) == 0;
  }
};
struct PerDeviceMax {             // raise(n): true when n exceeds what was recorded for the current device
  std::atomic<size_t> v[B2OF_MAX_DEVICES];
  PerDeviceMax() { for (auto& x : v) x.store(0); }
  bool raise(size_t n) {
    auto& x = v[current_device()];
    if (n <= x.load()) return false;
    x.store(n);
    return true;
  }
};

// linear bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

}  // namespace b2of
