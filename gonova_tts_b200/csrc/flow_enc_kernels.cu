// Small kernels of the flow front (oracle/flow_enc_ref.py): token embedding, LayerNorm(512), the relative
// position table, relative-position self-attention, the mu transpose and the speaker projection.  Every Linear / Conv1d of
// the Conformer encoder is a conv_tc2_kernel launch (tcgen05 / TMEM / TMA) with bias / residual / SiLU / LeakyReLU in its
// epilogue; these kernels are what sits between them.
// Activations are time-major [B, T, 512]; rows at or beyond an utterance's length are kept at zero (an utterance in a
// ragged batch is encoded exactly as if it were alone: upstream's flow.inference asserts a batch of one).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "flow_enc_kernels.h"

namespace gnv { void preload_kernel(const void* kernel); }   // conv_tc.cu

namespace gnv {

namespace {

__device__ __forceinline__ int enc_len(const int32_t* lengths, int b, int len_mul, int T) {
  if (!lengths) return T;
  const int n = lengths[b] * len_mul;
  return n < 0 ? 0 : (n > T ? T : n);
}

template <typename E>
__device__ __forceinline__ void enc_store4(E* dst, const float* v, int round_tf32v) {
  float y[4] = {v[0], v[1], v[2], v[3]};
  if constexpr (sizeof(E) == 4) {
    if (round_tf32v) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = round_tf32(y[k]);
    }
  }
  ElemIO<E>::template store_vec<4>(dst, y);
}

// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void enc_embed_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ token_len,
                                 const float* __restrict__ table, int vocab, int B, int L, E* __restrict__ out, int round_tf32v) {
  const size_t n = (size_t)B * L * (kEncC / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / (kEncC / 4);
    const int c = (int)(i % (kEncC / 4)) * 4;
    const int b = (int)(row / L), l = (int)(row - (size_t)b * L);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (l < enc_len(token_len, b, 1, L)) {
      int tok = tokens[row];
      tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
      const float4 a = *reinterpret_cast<const float4*>(table + (size_t)tok * kEncC + c);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    enc_store4<E>(out + row * kEncC + c, v, round_tf32v);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm(512): one warp per row, a lane holds 4 x float4 (channels lane*4 + 128 j): coalesced 512-byte warp accesses.
template <typename E>
__global__ void __launch_bounds__(256) enc_ln_kernel(const float* __restrict__ in, int B, int T, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, const int32_t* __restrict__ lengths,
                                                     int len_mul, E* __restrict__ out_e, float* __restrict__ out_f, int round_tf32v) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int rows = B * T;
  float g[16], bt[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 a = *reinterpret_cast<const float4*>(gamma + 128 * j + lane * 4);
    const float4 c = *reinterpret_cast<const float4*>(beta + 128 * j + lane * 4);
    g[4 * j] = a.x; g[4 * j + 1] = a.y; g[4 * j + 2] = a.z; g[4 * j + 3] = a.w;
    bt[4 * j] = c.x; bt[4 * j + 1] = c.y; bt[4 * j + 2] = c.z; bt[4 * j + 3] = c.w;
  }
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const int b = row / T, t = row - b * T;
    const bool live = t < enc_len(lengths, b, len_mul, T);
    float v[16];
    if (live) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(in + (size_t)row * kEncC + 128 * j + lane * 4);
        v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.f / kEncC);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * (1.f / kEncC) + eps);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf((v[i] - mean) * rstd, g[i], bt[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t o = (size_t)row * kEncC + 128 * j + lane * 4;
      if (out_f) *reinterpret_cast<float4*>(out_f + o) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      if (out_e) enc_store4<E>(out_e + o, v + 4 * j, round_tf32v);
    }
  }
}

template <typename E>
__global__ void enc_cast_kernel(const float* __restrict__ in, int B, int T, const int32_t* __restrict__ lengths, int len_mul,
                                E* __restrict__ out, int round_tf32v) {
  const size_t n = (size_t)B * T * (kEncC / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / (kEncC / 4);
    const int c = (int)(i % (kEncC / 4)) * 4;
    const int b = (int)(row / T), t = (int)(row - (size_t)b * T);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < enc_len(lengths, b, len_mul, T)) {
      const float4 a = *reinterpret_cast<const float4*>(in + row * kEncC + c);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    enc_store4<E>(out + row * kEncC + c, v, round_tf32v);
  }
}

// ------------------------------------------------------------------------------------------------
// The relative-position table of one layer at one length (built once per plan): a block takes kPosRows rows r of pos_emb
// (sinusoids of relative position (T-1) - r, the fp32 arithmetic of upstream's EspnetRelPositionalEncoding) and thread o
// computes output channel o of linear_pos for each of them, reading its weight row once.
constexpr int kPosRows = 8;

__global__ void __launch_bounds__(kEncC) enc_pos_kernel(const float* __restrict__ w_pos, int T, float* __restrict__ P,
                                                        __nv_bfloat16* __restrict__ Pb) {
  __shared__ float pe[kPosRows][kEncC];
  const int R = 2 * T - 1;
  const int r0 = blockIdx.x * kPosRows;
  const int c = threadIdx.x;
  {
    const float div = expf((float)(c & ~1) * (float)(-(9.210340371976184 / (double)kEncC)));   // ln(10000) / d
    for (int k = 0; k < kPosRows; ++k) {
      const int r = r0 + k;
      float val = 0.f;
      if (r < R) {
        const int rel = (T - 1) - r;
        const float arg = (float)(rel < 0 ? -rel : rel) * div;
        val = (c & 1) ? cosf(arg) : (rel < 0 ? -sinf(arg) : sinf(arg));
      }
      pe[k][c] = val;
    }
  }
  __syncthreads();
  float acc[kPosRows];
#pragma unroll
  for (int k = 0; k < kPosRows; ++k) acc[k] = 0.f;
  const float4* wrow = reinterpret_cast<const float4*>(w_pos + (size_t)c * kEncC);
  for (int j = 0; j < kEncC / 4; ++j) {
    const float4 w4 = wrow[j];
#pragma unroll
    for (int k = 0; k < kPosRows; ++k) {
      acc[k] = fmaf(w4.x, pe[k][4 * j], acc[k]);
      acc[k] = fmaf(w4.y, pe[k][4 * j + 1], acc[k]);
      acc[k] = fmaf(w4.z, pe[k][4 * j + 2], acc[k]);
      acc[k] = fmaf(w4.w, pe[k][4 * j + 3], acc[k]);
    }
  }
  const int h = c >> 6, d = c & 63;
  for (int k = 0; k < kPosRows; ++k) {
    const int r = r0 + k;
    if (r < R) {
      P[((size_t)h * R + r) * 64 + d] = acc[k];
      Pb[((size_t)h * R + r) * 64 + d] = __float2bfloat16_rn(acc[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Relative-position attention, fp32 arithmetic on the CUDA cores (the encoder runs once per utterance against twenty
// estimator evaluations: ~1 % of the flow's FLOPs).  A block = kAeQ consecutive query rows of one (utterance, head), one
// warp per row; per tile of 32 keys lane j owns key j for the scores (K and the window of P rows the 16 x 32 (i, j) pairs
// touch sit in shared memory, read as float4), then owns channels 2 lane, 2 lane + 1 for P V.
constexpr int kAeQ = 16;
constexpr int kAeK = 32;
constexpr int kAeWin = kAeQ + kAeK - 1;

template <typename E>
__global__ void __launch_bounds__(kAeQ * 32) enc_attn_kernel(const E* __restrict__ qkv, const float* __restrict__ P, int T,
                                                             const int32_t* __restrict__ lengths, int len_mul,
                                                             E* __restrict__ out, int round_tf32v) {
  // K and the P window with a 68-word pitch: rows are 16-byte aligned and a quarter-warp's float4 loads (lanes 68 words apart)
  // cover all 32 banks
  __shared__ __align__(16) float s_qu[kAeQ][64], s_qv[kAeQ][64];
  __shared__ __align__(16) float s_k[kAeK][68], s_v[kAeK][64], s_p[kAeWin][68];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAeQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = enc_len(lengths, b, len_mul, T);
  const int i = i0 + warp;
  E* orow = out + ((size_t)b * T + i) * kEncC + h * 64 + 2 * lane;
  if (i0 >= len) {                                           // the whole block is padding
    if (i < T) {
      if constexpr (sizeof(E) == 2) *reinterpret_cast<uint32_t*>(orow) = 0u;
      else *reinterpret_cast<float2*>(orow) = make_float2(0.f, 0.f);
    }
    return;
  }
  const int R = 2 * T - 1;
  const E* base = qkv + (size_t)b * T * kEncQkv + h * 64;
  {
    const int ii = i < T ? i : T - 1;
    const E* q = base + (size_t)ii * kEncQkv;
    s_qu[warp][lane] = ElemIO<E>::load(q + lane);       s_qu[warp][lane + 32] = ElemIO<E>::load(q + lane + 32);
    s_qv[warp][lane] = ElemIO<E>::load(q + 512 + lane); s_qv[warp][lane + 32] = ElemIO<E>::load(q + 512 + lane + 32);
  }
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  const float sc = 0.125f * 1.4426950408889634f;             // 1 / sqrt(64), in the exp2 domain
  // the next tile's K / V / P-window elements travel in registers while this tile is computed (one global round trip per
  // tile would otherwise sit between the two barriers)
  constexpr int kKvPer = kAeK * 64 / (kAeQ * 32), kPPer = (kAeWin * 64 + kAeQ * 32 - 1) / (kAeQ * 32);
  float rk[kKvPer], rv[kKvPer], rp[kPPer];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int u = 0; u < kKvPer; ++u) {
      const int e = threadIdx.x + u * kAeQ * 32;
      const int j = e >> 6, d = e & 63;
      const int jj = j0 + j;
      rk[u] = rv[u] = 0.f;
      if (jj < len) {
        rk[u] = ElemIO<E>::load(base + (size_t)jj * kEncQkv + 1024 + d);
        rv[u] = ElemIO<E>::load(base + (size_t)jj * kEncQkv + 1536 + d);
      }
    }
    // window of P: local row w holds r = (T-1) - (i0 + kAeQ - 1) + j0 + w; pair (query i0 + q, key j0 + j) reads w = j + kAeQ-1 - q
    const int rbase = (T - 1) - (i0 + kAeQ - 1) + j0;
#pragma unroll
    for (int u = 0; u < kPPer; ++u) {
      const int e = threadIdx.x + u * kAeQ * 32;
      const int w = e >> 6, d = e & 63;
      const int r = rbase + w;
      rp[u] = (e < kAeWin * 64 && r >= 0 && r < R) ? P[((size_t)h * R + r) * 64 + d] : 0.f;
    }
  };
  fetch(0);
  for (int j0 = 0; j0 < len; j0 += kAeK) {
    __syncthreads();                                         // every warp is done with the previous tile
#pragma unroll
    for (int u = 0; u < kKvPer; ++u) {
      const int e = threadIdx.x + u * kAeQ * 32;
      s_k[e >> 6][e & 63] = rk[u];
      s_v[e >> 6][e & 63] = rv[u];
    }
#pragma unroll
    for (int u = 0; u < kPPer; ++u) {
      const int e = threadIdx.x + u * kAeQ * 32;
      if (e < kAeWin * 64) s_p[e >> 6][e & 63] = rp[u];
    }
    __syncthreads();
    if (j0 + kAeK < len) fetch(j0 + kAeK);
    float ac = 0.f, bd = 0.f;
    const float4* kr = reinterpret_cast<const float4*>(s_k[lane]);
    const float4* pr = reinterpret_cast<const float4*>(s_p[lane + kAeQ - 1 - warp]);
    const float4* qu4 = reinterpret_cast<const float4*>(s_qu[warp]);
    const float4* qv4 = reinterpret_cast<const float4*>(s_qv[warp]);
#pragma unroll 8
    for (int d = 0; d < 16; ++d) {
      const float4 k4 = kr[d], p4 = pr[d], u4 = qu4[d], v4 = qv4[d];
      ac = fmaf(u4.x, k4.x, ac); ac = fmaf(u4.y, k4.y, ac); ac = fmaf(u4.z, k4.z, ac); ac = fmaf(u4.w, k4.w, ac);
      bd = fmaf(v4.x, p4.x, bd); bd = fmaf(v4.y, p4.y, bd); bd = fmaf(v4.z, p4.z, bd); bd = fmaf(v4.w, p4.w, bd);
    }
    const bool valid = j0 + lane < len;
    const float s = valid ? (ac + bd) * sc : -INFINITY;
    float mt = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    const float mn = fmaxf(m, mt);                           // finite: key j0 is valid
    const float corr = exp2f(m - mn);
    const float p = valid ? exp2f(s - mn) : 0.f;
    float ps = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    l = fmaf(l, corr, ps);
    o0 *= corr; o1 *= corr;
#pragma unroll 8
    for (int j = 0; j < kAeK; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const float2 v2 = *reinterpret_cast<const float2*>(&s_v[j][2 * lane]);
      o0 = fmaf(pj, v2.x, o0);
      o1 = fmaf(pj, v2.y, o1);
    }
    m = mn;
  }
  if (i < T) {
    const bool live = i < len;
    const float inv = live ? 1.0f / l : 0.f;
    float y0 = live ? o0 * inv : 0.f, y1 = live ? o1 * inv : 0.f;
    if constexpr (sizeof(E) == 2) {
      *reinterpret_cast<uint32_t*>(orow) = ElemIO<E>::pack2(y0, y1);
    } else {
      if (round_tf32v) { y0 = round_tf32(y0); y1 = round_tf32(y1); }
      *reinterpret_cast<float2*>(orow) = make_float2(y0, y1);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// The bf16 path: the same attention as a flash loop on mma.sync m16n8k16 (bf16 operands, fp32 accumulate / softmax).  Block
// = 4 warps = 64 queries of one (utterance, head); per tile of 64 keys a two-stage cp.async ring brings K, V and the 127
// rows of the position table the 64 x 64 (i, j) pairs touch (row (T-1) - i + j).  A warp (16 queries) computes
//   AC = (q + u) K^T                                   16 x 64   (32 MMAs)
//   BD' = (q + v) P_win^T over its 79-row sub-window   16 x 80   (40 MMAs), BD'[ii][c] with c = 15 - ii + jj
// and "rel_shift" is the skewed read-back of BD' through 5.6 KB of per-warp shared memory: S[ii][jj] = AC + BD'[ii][15 - ii + jj].
// Then the usual online softmax and P V (32 MMAs, V through ldmatrix.trans).  (The CUDA-core kernel above took 2.5 ms per
// layer at B x T = 32 x 500: 87 % of an encode.)
// ------------------------------------------------------------------------------------------------
constexpr int kEmQ = 64, kEmK = 64, kEmPitch = 72, kEmWin = 128, kEmBdPitch = 88;
constexpr uint32_t kEmKvStage = kEmK * kEmPitch * 2, kEmPStage = kEmWin * kEmPitch * 2;
constexpr int kEmSmem = 2 * (2 * kEmKvStage + kEmPStage) + 4 * 16 * kEmBdPitch * 4;

__device__ __forceinline__ void enc_mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t enc_pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float enc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void enc_cp16(uint32_t dst, const void* src, bool live) {
  const int n = live ? 16 : 0;                               // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(128, 2) enc_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             const __nv_bfloat16* __restrict__ Pb, int T,
                                                             const int32_t* __restrict__ lengths, int len_mul,
                                                             __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char enc_smem[];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kEmQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int len = enc_len(lengths, b, len_mul, T);
  const int R = 2 * T - 1;
  const __nv_bfloat16* base = qkv + (size_t)b * T * kEncQkv + h * 64;
  const __nv_bfloat16* pbase = Pb + (size_t)h * R * 64;
  const int r0 = q0 + 16 * warp + g, r1 = r0 + 8;
  const int n_tiles = q0 < len ? (len + kEmK - 1) / kEmK : 0;
  const uint32_t ks_u = (uint32_t)__cvta_generic_to_shared(enc_smem);
  const uint32_t vs_u = ks_u + 2 * kEmKvStage;
  const uint32_t ps_u = vs_u + 2 * kEmKvStage;
  float* bd_w = reinterpret_cast<float*>(enc_smem + 4 * kEmKvStage + 2 * kEmPStage) + warp * 16 * kEmBdPitch;
  auto load_tile = [&](int tile, int stage) {
    const int k0 = tile * kEmK;
#pragma unroll
    for (int i = threadIdx.x; i < kEmK * 8; i += 128) {     // 64 rows x 8 sixteen-byte words, K and V
      const int kk = i >> 3, w = i & 7;
      const int tk = k0 + kk;
      const bool live = tk < len;
      const __nv_bfloat16* kp = base + (size_t)(live ? tk : 0) * kEncQkv + 1024 + 8 * w;
      const uint32_t off = stage * kEmKvStage + (uint32_t)(kk * kEmPitch + 8 * w) * 2u;
      enc_cp16(ks_u + off, kp, live);
      enc_cp16(vs_u + off, kp + 512, live);
    }
    const int rb = (T - 1) - (q0 + kEmQ - 1) + k0;           // window row wl holds table row rb + wl
#pragma unroll
    for (int i = threadIdx.x; i < kEmWin * 8; i += 128) {
      const int wl = i >> 3, w = i & 7;
      const int r = rb + wl;
      const bool live = r >= 0 && r < R;
      enc_cp16(ps_u + stage * kEmPStage + (uint32_t)(wl * kEmPitch + 8 * w) * 2u, pbase + (size_t)(live ? r : 0) * 64 + 8 * w, live);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_tiles > 0) load_tile(0, 0);
  uint32_t qu[4][4], qv[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = 16 * ks + 2 * t;
    const __nv_bfloat16* p0 = base + (size_t)r0 * kEncQkv + c;
    const __nv_bfloat16* p1 = base + (size_t)r1 * kEncQkv + c;
    qu[ks][0] = r0 < len ? *reinterpret_cast<const uint32_t*>(p0) : 0u;
    qu[ks][1] = r1 < len ? *reinterpret_cast<const uint32_t*>(p1) : 0u;
    qu[ks][2] = r0 < len ? *reinterpret_cast<const uint32_t*>(p0 + 8) : 0u;
    qu[ks][3] = r1 < len ? *reinterpret_cast<const uint32_t*>(p1 + 8) : 0u;
    qv[ks][0] = r0 < len ? *reinterpret_cast<const uint32_t*>(p0 + 512) : 0u;
    qv[ks][1] = r1 < len ? *reinterpret_cast<const uint32_t*>(p1 + 512) : 0u;
    qv[ks][2] = r0 < len ? *reinterpret_cast<const uint32_t*>(p0 + 520) : 0u;
    qv[ks][3] = r1 < len ? *reinterpret_cast<const uint32_t*>(p1 + 520) : 0u;
  }
  const float sc2 = 0.125f * 1.4426950408889634f;            // 1 / sqrt(64), softmax in base 2
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t k_lane = (uint32_t)((lane & 7) * kEmPitch + 8 * (lane >> 3)) * 2u;
  const uint32_t v_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kEmPitch + 8 * (lane >> 4)) * 2u;
  const int wb = 48 - 16 * warp;                             // first window row of this warp's 80-row sub-window
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1, k0 = tile * kEmK;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                         // tile `tile` has landed; every warp is done with tile - 1
    if (tile + 1 < n_tiles) load_tile(tile + 1, stage ^ 1);
    const uint32_t kb = ks_u + stage * kEmKvStage + k_lane, vb = vs_u + stage * kEmKvStage + v_lane;
    const uint32_t pb = ps_u + stage * kEmPStage + (uint32_t)(wb * kEmPitch) * 2u + k_lane;
    // BD' first: through shared memory while the AC MMAs run
#pragma unroll
    for (int n = 0; n < 10; ++n) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(pb + (uint32_t)(8 * n * kEmPitch + 32 * kp) * 2u));
        enc_mma_bf16(d, qv[2 * kp], b0, b1);
        enc_mma_bf16(d, qv[2 * kp + 1], b2, b3);
      }
      *reinterpret_cast<float2*>(bd_w + g * kEmBdPitch + 8 * n + 2 * t) = make_float2(d[0], d[1]);
      *reinterpret_cast<float2*>(bd_w + (g + 8) * kEmBdPitch + 8 * n + 2 * t) = make_float2(d[2], d[3]);
    }
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(kb + (uint32_t)(8 * j * kEmPitch + 32 * kp) * 2u));
        enc_mma_bf16(s[j], qu[2 * kp], b0, b1);
        enc_mma_bf16(s[j], qu[2 * kp + 1], b2, b3);
      }
    }
    __syncwarp();
    {
      const float* ra = bd_w + g * kEmBdPitch + (15 - g) + 2 * t;          // row g:     column 15 - g + jj
      const float* rc = bd_w + (g + 8) * kEmBdPitch + (7 - g) + 2 * t;     // row g + 8: column 15 - (g + 8) + jj
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] += ra[8 * j]; s[j][1] += ra[8 * j + 1];
        s[j][2] += rc[8 * j]; s[j][3] += rc[8 * j + 1];
      }
    }
    if (k0 + kEmK > len) {                                   // the utterance's last tile: keys past its length
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + 8 * j + 2 * t;
        if (key >= len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0 * sc2), mn1 = fmaxf(m1, mx1 * sc2);      // finite: the tile holds at least one valid key
    const float c0 = enc_ex2(m0 - mn0), c1 = enc_ex2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
      s[j][0] = enc_ex2(fmaf(s[j][0], sc2, -mn0)); s[j][1] = enc_ex2(fmaf(s[j][1], sc2, -mn0));
      s[j][2] = enc_ex2(fmaf(s[j][2], sc2, -mn1)); s[j][3] = enc_ex2(fmaf(s[j][3], sc2, -mn1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                        // 16 keys per k-step: S tiles 2 kk and 2 kk + 1
      uint32_t pa[4];
      pa[0] = enc_pack2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = enc_pack2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = enc_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = enc_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {                      // d tiles 2 dp, 2 dp + 1
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vb + (uint32_t)(16 * kk * kEmPitch + 16 * dp) * 2u));
        enc_mma_bf16(o[2 * dp], pa, b0, b1);
        enc_mma_bf16(o[2 * dp + 1], pa, b2, b3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = (r0 < len && l0 > 0.f) ? 1.f / l0 : 0.f, i1 = (r1 < len && l1 > 0.f) ? 1.f / l1 : 0.f;
#pragma unroll
  for (int dj = 0; dj < 8; ++dj) {
    const int c = h * 64 + 8 * dj + 2 * t;
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r0) * kEncC + c) = enc_pack2(o[dj][0] * i0, o[dj][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r1) * kEncC + c) = enc_pack2(o[dj][2] * i1, o[dj][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------
// The tf32 path: the same attention on mma.sync m16n8k8 (tf32 operands, fp32 accumulate / softmax) over the fp32 buffer.
// Block = 4 warps = 64 queries; per tile of 64 keys K, V and the 127-row window of the position table are staged by
// cp.async (rows padded to 68 floats: every fragment load below is conflict-free); two blocks per SM overlap each other's
// loads.  AC and BD' as in the bf16 kernel (64 + 80 MMAs per warp), the same skewed read-back; for P V the accumulator
// layout of S (a thread's columns 2t, 2t + 1) serves as the A fragment (k indices t, t + 4) by reading V's rows in that
// permuted key order.  (The CUDA-core kernel took 2.5 ms per layer at B x T = 32 x 500.)
// ------------------------------------------------------------------------------------------------
constexpr int kEtQ = 64, kEtK = 64, kEtPitch = 68, kEtWin = 128, kEtBdPitch = 88;
constexpr int kEtSmem = ((2 * kEtK + kEtWin) * kEtPitch + 4 * 16 * kEtBdPitch) * 4;

__device__ __forceinline__ void enc_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t enc_tf32_bits(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

__global__ void __launch_bounds__(128, 2) enc_attn_tf32_kernel(const float* __restrict__ qkv, const float* __restrict__ P, int T,
                                                              const int32_t* __restrict__ lengths, int len_mul,
                                                              float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char enc_smem[];
  float* Ks = reinterpret_cast<float*>(enc_smem);
  float* Vs = Ks + kEtK * kEtPitch;
  float* Ps = Vs + kEtK * kEtPitch;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kEtQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* bd_w = Ps + kEtWin * kEtPitch + warp * 16 * kEtBdPitch;
  const int len = enc_len(lengths, b, len_mul, T);
  const int R = 2 * T - 1;
  const float* base = qkv + (size_t)b * T * kEncQkv + h * 64;
  const float* pbase = P + (size_t)h * R * 64;
  const int r0 = q0 + 16 * warp + g, r1 = r0 + 8;
  const int n_tiles = q0 < len ? (len + kEtK - 1) / kEtK : 0;
  const uint32_t ks_u = (uint32_t)__cvta_generic_to_shared(Ks), vs_u = (uint32_t)__cvta_generic_to_shared(Vs);
  const uint32_t ps_u = (uint32_t)__cvta_generic_to_shared(Ps);
  // A fragments of (q + u) and (q + v): a0 = (g, t), a1 = (g + 8, t), a2 = (g, t + 4), a3 = (g + 8, t + 4) per k-step of 8
  uint32_t qu[8][4], qv[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const float* p0 = base + (size_t)r0 * kEncQkv + 8 * ks + t;
    const float* p1 = base + (size_t)r1 * kEncQkv + 8 * ks + t;
    qu[ks][0] = r0 < len ? __float_as_uint(p0[0]) : 0u;   qu[ks][1] = r1 < len ? __float_as_uint(p1[0]) : 0u;
    qu[ks][2] = r0 < len ? __float_as_uint(p0[4]) : 0u;   qu[ks][3] = r1 < len ? __float_as_uint(p1[4]) : 0u;
    qv[ks][0] = r0 < len ? __float_as_uint(p0[512]) : 0u; qv[ks][1] = r1 < len ? __float_as_uint(p1[512]) : 0u;
    qv[ks][2] = r0 < len ? __float_as_uint(p0[516]) : 0u; qv[ks][3] = r1 < len ? __float_as_uint(p1[516]) : 0u;
  }
  const float sc2 = 0.125f * 1.4426950408889634f;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int wb = 48 - 16 * warp;                             // first window row of this warp's 80-row sub-window
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int k0 = tile * kEtK;
    __syncthreads();                                         // every warp is done with the previous tile
    for (int i = threadIdx.x; i < kEtK * 16; i += 128) {     // 64 rows x 16 sixteen-byte words, K and V
      const int kk = i >> 4, w = i & 15;
      const int tk = k0 + kk;
      const bool live = tk < len;
      const float* kp = base + (size_t)(live ? tk : 0) * kEncQkv + 1024 + 4 * w;
      const uint32_t off = (uint32_t)(kk * kEtPitch + 4 * w) * 4u;
      enc_cp16(ks_u + off, kp, live);
      enc_cp16(vs_u + off, kp + 512, live);
    }
    const int rb = (T - 1) - (q0 + kEtQ - 1) + k0;           // window row wl holds table row rb + wl
    for (int i = threadIdx.x; i < kEtWin * 16; i += 128) {
      const int wl = i >> 4, w = i & 15;
      const int r = rb + wl;
      const bool live = r >= 0 && r < R;
      enc_cp16(ps_u + (uint32_t)(wl * kEtPitch + 4 * w) * 4u, pbase + (size_t)(live ? r : 0) * 64 + 4 * w, live);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // BD' = (q + v) P_win^T over this warp's 80 window rows, through shared memory
    // (B fragments of two k-steps per ldmatrix.x4: an 8 x 8 b16 matrix is 8 rows x 4 floats and thread i receives the 32-bit
    //  word (row i / 4, word i % 4) — b0 = (k = t, n = g), b1 = (k = t + 4, n = g) of m16n8k8)
    const uint32_t ld_lane = (uint32_t)(((lane & 7) * kEtPitch + 4 * (lane >> 3)) * 4);
    {
      const uint32_t pw = ps_u + (uint32_t)(wb * kEtPitch * 4) + ld_lane;
#pragma unroll
      for (int n = 0; n < 10; ++n) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          uint32_t b0, b1, b2, b3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(pw + (uint32_t)((8 * n * kEtPitch + 16 * kp) * 4)));
          enc_mma_tf32(d, qv[2 * kp], b0, b1);
          enc_mma_tf32(d, qv[2 * kp + 1], b2, b3);
        }
        *reinterpret_cast<float2*>(bd_w + g * kEtBdPitch + 8 * n + 2 * t) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2*>(bd_w + (g + 8) * kEtBdPitch + 8 * n + 2 * t) = make_float2(d[2], d[3]);
      }
    }
    float s[8][4];
    {
      const uint32_t kw = ks_u + ld_lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          uint32_t b0, b1, b2, b3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(kw + (uint32_t)((8 * j * kEtPitch + 16 * kp) * 4)));
          enc_mma_tf32(s[j], qu[2 * kp], b0, b1);
          enc_mma_tf32(s[j], qu[2 * kp + 1], b2, b3);
        }
      }
    }
    __syncwarp();
    {
      const float* ra = bd_w + g * kEtBdPitch + (15 - g) + 2 * t;          // row g:     column 15 - g + jj
      const float* rc = bd_w + (g + 8) * kEtBdPitch + (7 - g) + 2 * t;     // row g + 8: column 15 - (g + 8) + jj
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] += ra[8 * j]; s[j][1] += ra[8 * j + 1];
        s[j][2] += rc[8 * j]; s[j][3] += rc[8 * j + 1];
      }
    }
    if (k0 + kEtK > len) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + 8 * j + 2 * t;
        if (key >= len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0 * sc2), mn1 = fmaxf(m1, mx1 * sc2);      // finite: the tile holds at least one valid key
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
      s[j][0] = exp2f(fmaf(s[j][0], sc2, -mn0)); s[j][1] = exp2f(fmaf(s[j][1], sc2, -mn0));
      s[j][2] = exp2f(fmaf(s[j][2], sc2, -mn1)); s[j][3] = exp2f(fmaf(s[j][3], sc2, -mn1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    {
      const float* vw = Vs + (2 * t) * kEtPitch + g;         // B fragment in the permuted key order: b0 = row 2t, b1 = row 2t + 1
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {                       // 8 keys per k-step: S tile kk
        uint32_t pa[4] = {enc_tf32_bits(s[kk][0]), enc_tf32_bits(s[kk][2]), enc_tf32_bits(s[kk][1]), enc_tf32_bits(s[kk][3])};
#pragma unroll
        for (int dj = 0; dj < 8; ++dj)
          enc_mma_tf32(o[dj], pa, __float_as_uint(vw[(8 * kk) * kEtPitch + 8 * dj]), __float_as_uint(vw[(8 * kk + 1) * kEtPitch + 8 * dj]));
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = (r0 < len && l0 > 0.f) ? 1.f / l0 : 0.f, i1 = (r1 < len && l1 > 0.f) ? 1.f / l1 : 0.f;
#pragma unroll
  for (int dj = 0; dj < 8; ++dj) {
    const int c = h * 64 + 8 * dj + 2 * t;
    if (r0 < T) *reinterpret_cast<float2*>(out + ((size_t)b * T + r0) * kEncC + c) = make_float2(round_tf32(o[dj][0] * i0), round_tf32(o[dj][1] * i0));
    if (r1 < T) *reinterpret_cast<float2*>(out + ((size_t)b * T + r1) * kEncC + c) = make_float2(round_tf32(o[dj][2] * i1), round_tf32(o[dj][3] * i1));
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void enc_mu_kernel(const float* __restrict__ in, int B, int T, const int32_t* __restrict__ lengths, int len_mul,
                              float* __restrict__ mu) {
  __shared__ float tile[32][81];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int len = enc_len(lengths, b, len_mul, T);
  for (int e = threadIdx.x; e < 32 * 80; e += blockDim.x) {
    const int t = e / 80, c = e - t * 80;
    const int tt = t0 + t;
    tile[t][c] = (tt < len) ? in[((size_t)b * T + tt) * 80 + c] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * 80; e += blockDim.x) {
    const int c = e >> 5, t = e & 31;
    const int tt = t0 + t;
    if (tt < T) mu[((size_t)b * 80 + c) * T + tt] = tile[t][c];
  }
}

__global__ void __launch_bounds__(192) enc_spk_kernel(const float* __restrict__ embedding, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ spks) {
  __shared__ float xn[192];
  __shared__ float part[6];
  const int b = blockIdx.x, c = threadIdx.x;
  const float x = embedding[(size_t)b * 192 + c];
  float s = x * x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((c & 31) == 0) part[c >> 5] = s;
  __syncthreads();
  const float norm = sqrtf(part[0] + part[1] + part[2] + part[3] + part[4] + part[5]);
  xn[c] = x / fmaxf(norm, 1e-12f);                           // F.normalize(dim=1), eps 1e-12
  __syncthreads();
  if (c < 80) {
    float acc = bias[c];
    for (int k = 0; k < 192; ++k) acc = fmaf(w[(size_t)c * 192 + k], xn[k], acc);
    spks[(size_t)b * 80 + c] = acc;
  }
}

inline int enc_blocks(size_t n, int per_block) {
  const size_t need = (n + per_block - 1) / per_block;
  return (int)(need < (size_t)148 * 8 ? (need ? need : 1) : (size_t)148 * 8);
}

}  // namespace

// Per device (function attributes are per device): called by gnv_flow_enc_create under its device guard.
cudaError_t flow_enc_init() {
  preload_kernel((const void*)enc_attn_mma_kernel);
  preload_kernel((const void*)enc_attn_tf32_kernel);
  cudaError_t e = cudaFuncSetAttribute(enc_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(enc_attn_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEtSmem);
}

cudaError_t launch_enc_embed(const int32_t* tokens, const int32_t* token_len, const float* table, int vocab, int B, int L,
                             void* out_e, int elem_bytes, int round_tf32v, cudaStream_t st) {
  const int blocks = enc_blocks((size_t)B * L * (kEncC / 4), 256);
  if (elem_bytes == 2)
    enc_embed_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(tokens, token_len, table, vocab, B, L, (__nv_bfloat16*)out_e, 0);
  else
    enc_embed_kernel<float><<<blocks, 256, 0, st>>>(tokens, token_len, table, vocab, B, L, (float*)out_e, round_tf32v);
  return cudaGetLastError();
}

cudaError_t launch_enc_ln(const float* in, int B, int T, const float* gamma, const float* beta, float eps,
                          const int32_t* lengths, int len_mul, void* out_e, int elem_bytes, int round_tf32v, float* out_f,
                          cudaStream_t st) {
  const int blocks = enc_blocks((size_t)B * T, 8);
  if (elem_bytes == 2)
    enc_ln_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, B, T, gamma, beta, eps, lengths, len_mul, (__nv_bfloat16*)out_e,
                                                         out_f, 0);
  else
    enc_ln_kernel<float><<<blocks, 256, 0, st>>>(in, B, T, gamma, beta, eps, lengths, len_mul, (float*)out_e, out_f, round_tf32v);
  return cudaGetLastError();
}

cudaError_t launch_enc_cast(const float* in, int B, int T, const int32_t* lengths, int len_mul, void* out_e, int elem_bytes,
                            int round_tf32v, cudaStream_t st) {
  const int blocks = enc_blocks((size_t)B * T * (kEncC / 4), 256);
  if (elem_bytes == 2)
    enc_cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, B, T, lengths, len_mul, (__nv_bfloat16*)out_e, 0);
  else
    enc_cast_kernel<float><<<blocks, 256, 0, st>>>(in, B, T, lengths, len_mul, (float*)out_e, round_tf32v);
  return cudaGetLastError();
}

cudaError_t launch_enc_pos(const float* w_pos, int T, float* P, void* P_bf16, cudaStream_t st) {
  const int R = 2 * T - 1;
  enc_pos_kernel<<<(R + kPosRows - 1) / kPosRows, kEncC, 0, st>>>(w_pos, T, P, (__nv_bfloat16*)P_bf16);
  return cudaGetLastError();
}

cudaError_t launch_enc_attn(const void* qkv, const float* P, const void* P_bf16, int B, int T, const int32_t* lengths,
                            int len_mul, void* out_e, int elem_bytes, int round_tf32v, cudaStream_t st) {
  // tensor cores (mma.sync bf16 / tf32); GONOVA_ENC_ATTN_MMA=0 keeps the fp32 CUDA-core kernel (tests run both)
  static const bool mma_env = [] { const char* v = getenv("GONOVA_ENC_ATTN_MMA"); return !(v && atoi(v) == 0); }();
  if (elem_bytes == 2 && mma_env && P_bf16) {
    dim3 grid((T + kEmQ - 1) / kEmQ, kEncH, B);
    enc_attn_mma_kernel<<<grid, 128, kEmSmem, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)P_bf16, T, lengths, len_mul,
                                                    (__nv_bfloat16*)out_e);
    return cudaGetLastError();
  }
  if (elem_bytes == 4 && mma_env && round_tf32v) {
    dim3 grid((T + kEtQ - 1) / kEtQ, kEncH, B);
    enc_attn_tf32_kernel<<<grid, 128, kEtSmem, st>>>((const float*)qkv, P, T, lengths, len_mul, (float*)out_e);
    return cudaGetLastError();
  }
  dim3 grid((T + kAeQ - 1) / kAeQ, kEncH, B);
  if (elem_bytes == 2)
    enc_attn_kernel<__nv_bfloat16><<<grid, kAeQ * 32, 0, st>>>((const __nv_bfloat16*)qkv, P, T, lengths, len_mul,
                                                               (__nv_bfloat16*)out_e, 0);
  else
    enc_attn_kernel<float><<<grid, kAeQ * 32, 0, st>>>((const float*)qkv, P, T, lengths, len_mul, (float*)out_e, round_tf32v);
  return cudaGetLastError();
}

cudaError_t launch_enc_mu(const float* in, int B, int T, const int32_t* lengths, int len_mul, float* mu, cudaStream_t st) {
  dim3 grid((T + 31) / 32, B);
  enc_mu_kernel<<<grid, 256, 0, st>>>(in, B, T, lengths, len_mul, mu);
  return cudaGetLastError();
}

cudaError_t launch_enc_spk(const float* embedding, const float* w, const float* bias, int B, float* spks, cudaStream_t st) {
  enc_spk_kernel<<<B, 192, 0, st>>>(embedding, w, bias, spks);
  return cudaGetLastError();
}

}  // namespace gnv
