// Shared host/device helpers for libmagcache_b200.so
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/magcache_b200.h"

namespace mc {

void set_error(const char* fmt, ...);  // defined in controller.cu

inline int32_t cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return MC_ERR_CUDA;
}

#define MC_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::mc::set_error(__VA_ARGS__);    \
      return MC_ERR_INVALID;           \
    }                                  \
  } while (0)

#define MC_CHECK_LAUNCH(what)                                      \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::mc::cuda_fail(e__, what);     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();  // cudaDevAttrMultiProcessorCount of the current device, cached per device (controller.cu)

constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per (kernel, device), not per process.
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
};
template <typename Kernel>
inline int32_t set_max_smem_once(Kernel kernel, int bytes, PerDeviceOnce& once, const char* what) {
  const int d = current_device();
  if (once.done[d]) return MC_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cuda_fail(e, what);
  once.done[d] = true;
  return MC_OK;
}

// ---- device-side dtype helpers -----------------------------------------------------------------
template <typename T>
struct Vec8;  // 8 elements of T as one or two 128-bit words

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

// unpack 8 bf16 (one uint4) to 8 floats
__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xFFFF0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xFFFF0000u);
  f[4] = __uint_as_float(v.z << 16);
  f[5] = __uint_as_float(v.z & 0xFFFF0000u);
  f[6] = __uint_as_float(v.w << 16);
  f[7] = __uint_as_float(v.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {  // round-to-nearest-even, like torch
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }  // nn.SiLU in fp32
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// torch.nn.GELU() (exact, erf form) in fp32: 0.5*x*(1+erf(x/sqrt(2))) — `img_emb` of the i2v models
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// torch.nn.GELU(approximate='tanh') in fp32: 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float kBeta = 0.7978845608028654f;  // sqrt(2/pi)
  const float kKappa = 0.044715f;
  const float u = kBeta * (x + kKappa * x * x * x);
  // one MUFU.TANH (tanh.approx.f32, |rel err| ~ 2^-11) instead of tanhf's ~20 instructions: the result is rounded to
  // bf16 (2^-9 relative) by every caller, so the approximation is invisible after rounding except on exact ties
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mc
