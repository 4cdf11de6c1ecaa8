// Shared helpers for the ifcb_classifier_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

namespace ifcb {

// Thread-local last-error string, returned through the C ABI by ifcb_last_error().
void set_error(const char* fmt, ...);
const char* get_error();

// Status convention of include/ifcb_b200.h: 0 ok, <0 argument error, >0 cudaError_t.
#define IFCB_ARG_CHECK(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::ifcb::set_error(__VA_ARGS__);                              \
      return -1;                                                   \
    }                                                              \
  } while (0)

#define IFCB_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::ifcb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                        __FILE__, __LINE__);                                         \
      return (int)_e;                                                                \
    }                                                                                \
  } while (0)

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 16-bit activation storage: bf16 (dtype 0) or fp16 (dtype 1, saturating).
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi, int fp16) {
  if (fp16) {
    lo = fminf(fmaxf(lo, -65504.f), 65504.f);
    hi = fminf(fmaxf(hi, -65504.f), 65504.f);
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return pack_bf16x2(lo, hi);
}

__device__ __forceinline__ float2 unpack_act2(uint32_t u, int fp16) {
  if (fp16) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
  }
  return unpack_bf16x2(u);
}

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

// floor(n / d) for 0 <= n < 2^31 through the host-computed magic = floor((2^64 - 1) / d) + 1
// (exact while n * d < 2^64; magic == 0 encodes d == 1).  ~5 integer instructions
// instead of the ~40 of an emulated 32-bit division -- map_row runs per accumulator.
__device__ __forceinline__ uint32_t fast_div(uint32_t n, unsigned long long magic) {
  if (magic == 0ull) return n;
  const uint32_t m_lo = (uint32_t)magic, m_hi = (uint32_t)(magic >> 32);
  const unsigned long long t = (unsigned long long)n * m_hi + (((unsigned long long)n * m_lo) >> 32);
  return (uint32_t)(t >> 32);
}

// host side of fast_div
inline unsigned long long div_magic(int d) { return d <= 1 ? 0ull : (~0ull) / (unsigned long long)d + 1ull; }

int sm_count();

// deterministic TRAIN mode (ifcb_train_deterministic): the workspace if one of at least `need` bytes is set, else NULL
void* det_workspace(long long need);
bool det_enabled();

}  // namespace ifcb
