// philox.cuh -- counter-based Gaussian noise for libvbnn.so.
//
// Replaces the reference's host-side randomkit.normal fill + H2D copy (VBLinear.lua:55-57):
// epsilon is a pure function of (seed, step, stream, sample, row, col), so the forward-time
// weight sample and the backward-time regeneration inside the dW epilogue agree bit for bit,
// on every rank, with no storage and no communication ("hard parts").
//
// Layout (mirrored by oracle/vbnn_oracle.py: philox_normal_matrix):
//   counter = (uint32(row * ceil(cols/4) + col/4), stream, sample, step)
//   key     = (seed_lo, seed_hi)
//   the 4 outputs of one Philox4x32-10 call are the normals of columns 4q .. 4q+3
//   (Box-Muller on (x0,x1) -> (n0,n1) and (x2,x3) -> (n2,n3)).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace vbnn {

struct PhiloxStream {
  uint32_t key0, key1;   // seed
  uint32_t stream;       // layer id | kind << 16
  uint32_t sample;       // MC sample index
  uint32_t step;         // minibatch counter
};

// stream ids: kind in the upper half so weight-eps, activation-zeta and init never collide
constexpr uint32_t kStreamEps = 0u << 16;
constexpr uint32_t kStreamZeta = 1u << 16;
constexpr uint32_t kStreamInit = 2u << 16;

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  // one 32 x 32 -> 64 multiply per product (IMAD.WIDE.U32 yields both halves)
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k0;
  uint32_t n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// MUFU.RSQ without the denormal / IEEE fix-up sequence of rsqrtf()
__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {
  // (n + 0.5) * 2^-24 in (0,1), 24 bits; as one FMA (a single rounding of the same real number, so
  // bit-identical to ((float)n + 0.5f) * 2^-24)
  return fmaf((float)(x >> 8), 1.0f / 16777216.0f, 1.0f / 33554432.0f);
}

// 4 standard normals for quad index `idx4` of the stream.
__device__ __forceinline__ void philox_normal4(const PhiloxStream& ps, uint32_t idx4, float (&n)[4]) {
  uint32_t c[4] = {idx4, ps.stream, ps.sample, ps.step};
  philox4x32_10(c, ps.key0, ps.key1);
#pragma unroll
  for (int a = 0; a < 4; a += 2) {
    float u0 = u32_to_unit(c[a]), u1 = u32_to_unit(c[a + 1]);
    float t = -2.0f * __logf(u0);          // >= 0; exactly 0 when u0 rounds to 1.0 (p ~ 2^-25)
    float r = t * rsqrt_fast(fmaxf(t, 1e-30f)); // sqrt via MUFU.RSQ (no IEEE slow path), 0 -> 0
    float s, co;
    __sincosf(6.283185307179586f * u1, &s, &co);
    n[a] = r * co;
    n[a + 1] = r * s;
  }
}

}  // namespace vbnn
