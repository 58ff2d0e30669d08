// common.cuh -- shared helpers for libvbnn.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/vbnn.h"

typedef __nv_bfloat16 bf16;

namespace vbnn {

void set_error(const char* fmt, ...);

#define VB_CUDA(expr)                                                                 \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      vbnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                      __FILE__, __LINE__);                                            \
      return VBNN_E_CUDA;                                                             \
    }                                                                                 \
  } while (0)

#define VB_CHECK(cond, code, ...)                                                     \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      vbnn::set_error(__VA_ARGS__);                                                   \
      return (code);                                                                  \
    }                                                                                 \
  } while (0)

#define VB_TRY(expr)                                                                  \
  do {                                                                                \
    int _r = (expr);                                                                  \
    if (_r != VBNN_OK) return _r;                                                     \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline long long round_up_ll(long long x, long long m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int x, int m) { return (x + m - 1) / m; }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- activation element type helpers (float in FP32 mode, bf16 in BF16 mode) -----------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// warp / block reductions ----------------------------------------------------------------
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
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of a double; result valid in thread 0.  `sh` needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace vbnn
