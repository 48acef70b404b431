// Shared helpers for the vocalie_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/vocalie_b200.h"

namespace vt {

// ---- thread-local error slot + launch counter ------------------------------------------
void set_error(const char* fmt, ...);
int& launch_counter();

#define VT_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      vt::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                   \
                    cudaGetErrorString(_e));                                               \
      return VT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VT_REQUIRE(cond, ...)                                                              \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      vt::set_error(__VA_ARGS__);                                                          \
      return VT_ERR_INVALID;                                                               \
    }                                                                                      \
  } while (0)

// Counts a launch of one of OUR kernels and checks the launch error.
#define VT_LAUNCHED()                                                                      \
  do {                                                                                     \
    vt::launch_counter()++;                                                                \
    VT_CUDA_OK(cudaGetLastError());                                                        \
  } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// 128-bit streaming load that does not allocate in L1 (data is touched once per pass).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace vt
