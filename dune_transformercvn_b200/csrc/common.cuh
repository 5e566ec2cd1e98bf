// Shared helpers for libtcvn (sm_100a).  Host + device.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tcvn.h"

namespace tcvn {

// thread-local error text behind tcvn_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define TCVN_CHECK_ARG(cond, ...)                         \
  do {                                                    \
    if (!(cond)) return ::tcvn::fail(TCVN_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define TCVN_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return ::tcvn::fail(TCVN_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

// device pointer whose value every seeded kernel launched from this host thread adds to its seed (tcvn_set_seed_offset)
const unsigned long long* seed_offset_ptr();
__device__ __forceinline__ unsigned long long seed_with_offset(unsigned long long seed, const unsigned long long* off) {
  return off ? seed + __ldg(off) : seed;
}

// every kernel launch of the library goes through this: counts launches for tcvn_launch_count()
void count_launch();
#define TCVN_LAUNCH_CHECK()        \
  do {                             \
    ::tcvn::count_launch();        \
    TCVN_CUDA(cudaGetLastError()); \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float prelu(float y, float a) { return y >= 0.f ? y : a * y; }

}  // namespace tcvn
