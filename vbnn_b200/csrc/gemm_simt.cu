// gemm_simt.cu -- fp32-operand / fp32-accumulate CUDA-core GEMM with the fused VB epilogues.
//
// This is the VBNN_PREC_FP32 "exact parity" mode: the reference computes everything in fp32
// (main.lua:10), so this path is directly comparable with the oracle to ~1e-6.  It is also what
// serves the tiny launch-bound C1 config.  The throughput path is gemm_tc.cu (tcgen05).
//
// D[m,n] = sum_k A(m,k) * B(k,n) with arbitrary element strides, so one kernel covers
//   forward   Y  = X W^T      (A = X [N x I], B(k,n) = W[n*ldw + k])          nn.Linear:updateOutput
//   backward  dX = G W        (A = G [N x O], B(k,n) = W[k*ldw + n])          nn.Linear:updateGradInput
//   wgrad     dW = G^T X      (A(m,k) = G[k*ldg + m], B(k,n) = X[k*ldx + n])  VBLinear.lua:113-115
#include "gemm.h"

namespace vbnn {

constexpr int TM = 64, TN = 64, TK = 16;

template <bool DUAL>
struct SimtSmem {
  float a[DUAL ? 2 : 1][TK][TM + 4];
  float b[DUAL ? 2 : 1][TK][TN + 4];
};

__device__ __forceinline__ void load_tile(float (*dst)[TM + 4], const float* __restrict__ src,
                                          long long s_mn, long long s_k, int mn0, int k0, int MN,
                                          int K) {
  // 64 (mn) x 16 (k) elements, 256 threads -> 4 each; make the unit-stride dim the fastest.
  const int tid = threadIdx.x;
  if (s_k == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int k = e & 15, mn = e >> 4;
      int gm = mn0 + mn, gk = k0 + k;
      dst[k][mn] = (gm < MN && gk < K) ? src[gm * s_mn + gk] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int mn = e & 63, k = e >> 6;
      int gm = mn0 + mn, gk = k0 + k;
      dst[k][mn] = (gm < MN && gk < K) ? src[gm * s_mn + gk * s_k] : 0.f;
    }
  }
}

template <int MODE, bool DUAL>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtGemmArgs g, EpiParams p) {
  __shared__ SimtSmem<DUAL> sm;
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* A1 = g.A1 + z * g.zsA1;
  const float* B1 = g.B1 + z * g.zsB1;
  const float* A2 = DUAL ? g.A2 + z * g.zsA2 : nullptr;
  const float* B2 = DUAL ? g.B2 + z * g.zsB2 : nullptr;

  float acc1[4][4] = {}, acc2[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    load_tile(sm.a[0], A1, g.sA1m, g.sA1k, m0, k0, g.M, g.K);
    load_tile(sm.b[0], B1, g.sB1n, g.sB1k, n0, k0, g.N, g.K);
    if (DUAL) {
      load_tile(sm.a[DUAL ? 1 : 0], A2, g.sA2m, g.sA2k, m0, k0, g.M, g.K);
      load_tile(sm.b[DUAL ? 1 : 0], B2, g.sB2n, g.sB2k, n0, k0, g.N, g.K);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&sm.a[0][k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&sm.b[0][k][tx * 4]);
      float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc1[i][j] = fmaf(a[i], b[j], acc1[i][j]);
      if (DUAL) {
        float4 av2 = *reinterpret_cast<const float4*>(&sm.a[DUAL ? 1 : 0][k][ty * 4]);
        float4 bv2 = *reinterpret_cast<const float4*>(&sm.b[DUAL ? 1 : 0][k][tx * 4]);
        float a2[4] = {av2.x, av2.y, av2.z, av2.w}, b2[4] = {bv2.x, bv2.y, bv2.z, bv2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = fmaf(a2[i], b2[j], acc2[i][j]);
      }
    }
    __syncthreads();
  }
  const PhiloxStream ps = epi_stream(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) epi_quad<MODE, float>(p, ps, z, m0 + ty * 4 + i, n0 + tx * 4, acc1[i], acc2[i]);
}

template <int MODE>
static int launch_mode(const SimtGemmArgs& g, const EpiParams& p, int batch, cudaStream_t st) {
  constexpr bool DUAL = epi_is_dual(MODE);
  if (g.M <= 0 || g.N <= 0) return VBNN_OK;
  if (epi_z_accumulates(MODE)) {
    // z = MC sample accumulates into one output: serialise over z (deterministic order)
    for (int z = 0; z < batch; ++z) {
      SimtGemmArgs gz = g;
      gz.A1 += z * g.zsA1; gz.B1 += z * g.zsB1;
      if (DUAL) { gz.A2 += z * g.zsA2; gz.B2 += z * g.zsB2; }
      EpiParams pz = p;
      pz.ps.sample += z;
      if (pz.noise) pz.noise += z * p.zs_noise;
      pz.accumulate = p.accumulate || z > 0;
      dim3 grid(ceil_div(g.N, TN), ceil_div(g.M, TM), 1);
      gemm_simt_kernel<MODE, DUAL><<<grid, 256, 0, st>>>(gz, pz);
    }
  } else {
    dim3 grid(ceil_div(g.N, TN), ceil_div(g.M, TM), batch);
    gemm_simt_kernel<MODE, DUAL><<<grid, 256, 0, st>>>(g, p);
  }
  VB_CUDA(cudaGetLastError());
  return VBNN_OK;
}

int gemm_simt_launch(int mode, const SimtGemmArgs& g, const EpiParams& p, int batch,
                     cudaStream_t st, long long* launches) {
  if (launches) *launches += epi_z_accumulates(mode) ? batch : 1;
  switch (mode) {
    case EPI_STORE: return launch_mode<EPI_STORE>(g, p, batch, st);
    case EPI_FWD: return launch_mode<EPI_FWD>(g, p, batch, st);
    case EPI_FWD_LRT: return launch_mode<EPI_FWD_LRT>(g, p, batch, st);
    case EPI_DX: return launch_mode<EPI_DX>(g, p, batch, st);
    case EPI_DX_LRT: return launch_mode<EPI_DX_LRT>(g, p, batch, st);
    case EPI_DW: return launch_mode<EPI_DW>(g, p, batch, st);
    case EPI_DW_LRT: return launch_mode<EPI_DW_LRT>(g, p, batch, st);
  }
  set_error("gemm_simt_launch: bad mode %d", mode);
  return VBNN_E_INVALID;
}

}  // namespace vbnn
