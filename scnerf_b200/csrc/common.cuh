// Shared helpers for the scnerf_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/scnerf_b200.h"

namespace scnerf {

// ---- error state (thread-local message, C-ABI returns a code) -------------------------------
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
#define SCNERF_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return scnerf::fail(SCNERF_ERR_ARG, __VA_ARGS__); } while (0)

inline std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}
// Per-launch timing with CUDA events on the launching stream (scnerf_kernel_timing / _report): off by default, so
// the timed region of a benchmark carries no events; bench.py switches it on for a separate short pass to split a
// step's time by kernel (roofline object).  One event pair per launch, recycled from a pool.
struct KernelTimes {
  struct Rec { const char* name; unsigned grid; cudaEvent_t e0, e1; };
  std::atomic<bool> on{false};
  std::mutex mu;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(const char* name, unsigned grid, cudaStream_t st) {
    std::lock_guard<std::mutex> g(mu);
    Rec r{name, grid, get(), get()};
    cudaEventRecord(r.e0, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) {
    std::lock_guard<std::mutex> g(mu);
    if (!recs.empty()) cudaEventRecord(recs.back().e1, st);
  }
};
inline KernelTimes& kernel_times() {
  static KernelTimes k;
  return k;
}
inline unsigned grid_x(unsigned g) { return g; }
inline unsigned grid_x(int g) { return (unsigned)g; }
inline unsigned grid_x(const dim3& g) { return g.x * g.y * g.z; }
// Every kernel launch in the library goes through this macro: counts it and checks the launch.
#define SCNERF_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                          \
    (void)cudaGetLastError(); /* drop stale non-sticky errors left by other libraries */        \
    const bool t__ = scnerf::kernel_times().on.load(std::memory_order_relaxed);                 \
    if (t__) scnerf::kernel_times().begin(#kernel, scnerf::grid_x(grid), (cudaStream_t)(stream)); \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                   \
    if (t__) scnerf::kernel_times().end((cudaStream_t)(stream));                                \
    scnerf::launch_counter().fetch_add(1, std::memory_order_relaxed);                           \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return scnerf::fail(SCNERF_ERR_CUDA, "%s launch failed: %s", #kernel,                     \
                          cudaGetErrorString(e__));                                             \
  } while (0)
#define SCNERF_CUDA(call)                                                                       \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return scnerf::fail(SCNERF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));           \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap, off;
  bool dry;  // size query only
  Arena(void* p, size_t c) : base((char*)p), cap(c), off(0), dry(p == nullptr) {}
  template <typename T>
  T* get(size_t n) {
    off = align_up(off, 256);
    T* r = dry ? nullptr : (T*)(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return dry || off <= cap; }
};

// ---- Philox4x32-10 counter RNG (Salmon et al. 2011) -------------------------------------------
struct Philox {
  __device__ static inline void round_(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // 4 x 32 random bits for (seed, stream, counter)
  __device__ static inline uint4 draw(uint64_t seed, uint32_t stream, uint64_t counter) {
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), stream, 0x5C4E4B46u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round_(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c[0], c[1], c[2], c[3]);
  }
  // U[0,1): 24 random mantissa bits (same support as torch.rand for fp32)
  __device__ static inline float u01(uint32_t bits) { return (bits >> 8) * (1.0f / 16777216.0f); }
  __device__ static inline float uniform(uint64_t seed, uint32_t stream, uint64_t i) {
    uint4 r = draw(seed, stream, i >> 2);
    uint32_t w = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
    return u01(w);
  }
  __device__ static inline float normal(uint64_t seed, uint32_t stream, uint64_t i) {
    uint4 r = draw(seed, stream, i >> 1);
    uint32_t a = (i & 1) ? r.z : r.x, b = (i & 1) ? r.w : r.y;
    float u1 = ((a >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
    float u2 = u01(b);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
};
enum : uint32_t { RNG_T_RAND = 1, RNG_U = 2, RNG_NOISE0 = 3, RNG_NOISE1 = 4 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace scnerf
