// CUDA-core implicit-GEMM conv (fp32 FMA).  Two jobs:
//   1. the source_downs strided convs (K = 18*k is tiny and strided), on every path;
//   2. the GNV_DTYPE_FP32 / GNV_FLAG_SIMT_CONV validation path for all conv layers, with the same
//      operands and the same fused epilogue as the tcgen05 kernel, so the tensor-core kernel can be
//      checked layer by layer on the GPU.
// 64x64 output tile, K step 16, 256 threads, 4x4 register micro-tile.
#pragma once
#include "common.cuh"

namespace gnv {

template <typename Ein, typename Ew, typename Eout>
__global__ void __launch_bounds__(256) conv_simt_kernel(const Ein* __restrict__ A, const Ew* __restrict__ W,
                                                         ConvGeom g, EpiParams ep) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN, b = blockIdx.z;
  const int tx = tid % 16, ty = tid / 16;
  const int lrow = tid / 4, lk = (tid % 4) * 4;     // loader mapping: 64 rows x 4 groups of 4 k
  const int Kw = g.n_taps * g.C_in_w;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long row_stride = g.a_row_stride ? g.a_row_stride : g.C_in_ld;
  const Ein* Ab = A + (size_t)b * (g.a_batch_stride ? g.a_batch_stride : (long long)g.L_in * g.C_in_ld);
  for (int tap = 0; tap < g.n_taps; ++tap) {
    const int in_row = (m0 + lrow) * g.in_stride + g.off0 + tap * g.tap_step;
    const bool row_ok = (m0 + lrow) < g.M_rows && in_row >= 0 && in_row < g.L_in;
    const bool n_ok = (n0 + lrow) < g.N_total;
    for (int c0 = 0; c0 < g.C_in; c0 += TK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ci = c0 + lk + i;
        float a = 0.f, w = 0.f;
        if (ci < g.C_in) {
          if (row_ok) a = ElemIO<Ein>::load(Ab + (size_t)in_row * row_stride + ci);
          if (n_ok) w = ElemIO<Ew>::load(W + (size_t)(n0 + lrow) * Kw + tap * g.C_in_w + ci);
        }
        As[lk + i][lrow] = a;
        Ws[lk + i][lrow] = w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float av[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Ws[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < g.M_rows) epi_apply<4, Eout>(ep, b, m, n0 + tx * 4, acc[i]);
  }
}

template <typename Ein, typename Ew, typename Eout>
inline cudaError_t launch_conv_simt(const void* A, const void* W, const ConvGeom& g, const EpiParams& ep,
                                    cudaStream_t st) {
  dim3 grid((g.M_rows + 63) / 64, (g.N_total + 63) / 64, g.B);
  conv_simt_kernel<Ein, Ew, Eout><<<grid, 256, 0, st>>>(reinterpret_cast<const Ein*>(A),
                                                        reinterpret_cast<const Ew*>(W), g, ep);
  return cudaGetLastError();
}

}  // namespace gnv
