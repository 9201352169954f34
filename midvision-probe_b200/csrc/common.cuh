// common.cuh -- shared host/device helpers for libmvmatch (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvmatch.h"

// ---- host-side error plumbing -----------------------------------------------------------
void mv_set_error(const char* fmt, ...);

#define MV_REQUIRE(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      mv_set_error(__VA_ARGS__);      \
      return (code);                  \
    }                                 \
  } while (0)

#define MV_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      mv_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return (int)e__;                                                                      \
    }                                                                                       \
  } while (0)

#define MV_LAUNCH_CHECK()                                                                   \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      mv_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return (int)e__;                                                                      \
    }                                                                                       \
  } while (0)

static inline cudaStream_t mv_cuda_stream(mv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int mv_sm_count();  // cached multiprocessor count of the current device
int mv_device_slot();  // index of the current device in [0, 64) for per-device caches (function attributes are per device)
constexpr int MV_MAX_DEVICES = 64;

// ---- device helpers ---------------------------------------------------------------------
#define MV_MASKED_F (-3.0e38f)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// monotone map float -> uint32 (a < b  <=>  ord(a) < ord(b) for non-NaN)
__device__ __forceinline__ uint32_t f32_orderable(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float orderable_f32(uint32_t o) {
  uint32_t b = o ^ ((o & 0x80000000u) ? 0x80000000u : 0xffffffffu);
  return __uint_as_float(b);
}
