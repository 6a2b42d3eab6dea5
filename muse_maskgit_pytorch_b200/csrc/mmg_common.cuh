// Shared host/device helpers for libmmg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/mmg.h"

namespace mmg {

typedef __nv_bfloat16 bf16;

// ---- thread-local error string -------------------------------------------------------------------
inline char* err_buf() { static thread_local char buf[512] = {0}; return buf; }
inline int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
  return code;
}
inline std::atomic<int64_t>& launch_counter() { static std::atomic<int64_t> c{0}; return c; }

#define MMG_CHECK_ARG(cond, ...) do { if (!(cond)) return ::mmg::fail(MMG_EINVAL, __VA_ARGS__); } while (0)
#define MMG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return ::mmg::fail(MMG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
// after a kernel launch: count it, surface launch-configuration errors (no sync)
#define MMG_LAUNCHED() do { ::mmg::launch_counter()++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    return ::mmg::fail(MMG_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

inline int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

// ---- device helpers --------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (F.gelu default), ref muse_maskgit_pytorch.py:77
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float leaky01(float x) { return x > 0.f ? x : 0.1f * x; }

// load/store 64 consecutive elements of T from/to a 16B-aligned row chunk
template <typename T> struct Vec64;
template <> struct Vec64<float> {
  static __device__ __forceinline__ void store(float* p, const float (&v)[64]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[64]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { float4 t = reinterpret_cast<const float4*>(p)[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
  }
};
template <> struct Vec64<bf16> {
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[64]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 t;
      __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), b = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), d = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
      t.z = *reinterpret_cast<uint32_t*>(&c); t.w = *reinterpret_cast<uint32_t*>(&d);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[64]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 t = reinterpret_cast<const uint4*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y; }
    }
  }
};

}  // namespace mmg
