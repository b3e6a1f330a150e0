// Shared helpers for the mocha_b200 CUDA library (sm_100a only).
#pragma once
#include "../../include/mocha_b200.h"  // MOCHA_OK / MOCHA_ERR_* codes
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

namespace mocha {

// ---------------------------------------------------------------------------------------------
// Error channel of the C ABI: every entry point returns 0 on success or a negative code and
// leaves a message readable through mocha_last_error(). No exceptions cross the boundary.
// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
const char* last_error();

#define MOCHA_CHECK_ARG(cond, ...)                                     \
  do {                                                                 \
    if (!(cond)) return ::mocha::set_error(MOCHA_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define MOCHA_CUDA(call)                                                               \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return ::mocha::set_error(MOCHA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                                cudaGetErrorString(e__), __FILE__, __LINE__);          \
  } while (0)

#define MOCHA_LAUNCH_CHECK(name)                                                        \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return ::mocha::set_error(MOCHA_ERR_CUDA, "launch of %s failed: %s", name, \
                                cudaGetErrorString(e__));                              \
  } while (0)

#define MOCHA_TRY(expr)          \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

// launch counter (bench.py reports "gpu_launches" from it)
void count_launch(int n = 1);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace (the library never cudaMallocs on the
// data path; see include/mocha_b200.h "Ownership").
struct Workspace {
  char* base;
  size_t cap;
  size_t off;
  bool overflow;
  Workspace(void* p, size_t n) : base((char*)p), cap(n), off(0), overflow(false) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) { overflow = true; off += bytes; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float lrelu02(float x) { return x > 0.f ? x : 0.2f * x; }
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mocha
