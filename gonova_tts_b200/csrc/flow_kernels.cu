// Byte-moving and small kernels of the CFM flow decoder (oracle/flow_ref.py): time embedding, input pack,
// LayerNorm (+ Mish, + time bias, + mask), self-attention, CFG + Euler step.  The GEMMs of the estimator (causal 3-tap
// convs, 1x1 convs, q/k/v / output / feed-forward projections) run on conv_tc2_kernel (tcgen05 / TMEM / TMA) with fused
// bias / residual / GELU epilogues; these kernels are what sits between them.
// Activations are time-major [B2, T, C] (B2 = 2 B: the conditioned rows, then the zero-condition rows of classifier-free
// guidance); rows at or beyond an utterance's length are kept at zero.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "flow_kernels.h"

namespace gnv { void preload_kernel(const void* kernel); }   // conv_tc.cu

namespace gnv {

__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)); softplus with torch's threshold 20
  // tanh(log(1 + e^x)) = ((1 + e^x)^2 - 1) / ((1 + e^x)^2 + 1) = n / (n + 2) with n = e^x (e^x + 2): one MUFU.EX2 and one
  // MUFU.RCP instead of expf + log1pf + tanhf (the LayerNorm + Mish kernel was bound by them); x > 20 returns x like torch
  if (x > 20.f) return x;
  const float e = __expf(x);
  const float n = e * (e + 2.f);
  return x * __fdividef(n, n + 2.f);
}

// ------------------------------------------------------------------------------------------------
// Time embedding for every Euler step at once, for t = t_steps[s]:
//   emb = [sin(1000 t w_i), cos(1000 t w_i)] (320)  ->  Linear 320 -> 1024, SiLU = h1  ->  Linear 1024 -> 1024, Mish = m
//   and for each ResNet block r:  tb[s][r][:] = Linear_r(m)   (1024 -> 256)
// (the same for every batch row: t is shared).  fp32 weights, PyTorch [out, in] layout.  Three launches, one warp per output
// channel (lanes stride the input: coalesced weight rows, shuffle reduction), 128 / 128 / 14 x 32 blocks per step: as ONE
// block per step with a thread per output it took 2.6 ms per decode (6 % of a single sentence's ten Euler steps).
// ------------------------------------------------------------------------------------------------
constexpr int kTmWarps = 8;

template <int kIn>
__device__ __forceinline__ float time_dot(const float* __restrict__ w, const float* x, int lane) {
  float acc = 0.f;
#pragma unroll 4
  for (int i = lane; i < kIn; i += 32) acc = fmaf(w[i], x[i], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(kTmWarps * 32) flow_time_h1_kernel(const FlowTimes t_steps, const float* __restrict__ w1,
                                                                    const float* __restrict__ b1, float* __restrict__ h1) {
  __shared__ float emb[320];
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float t = t_steps.t[s];
  const int half = 160;
  const float kf = logf(10000.f) / (float)(half - 1);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = 1000.f * t * expf(-kf * (float)i);
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
  }
  __syncthreads();
  const int o = blockIdx.y * kTmWarps + warp;
  const float acc = b1[o] + time_dot<320>(w1 + (size_t)o * 320, emb, lane);
  if (lane == 0) h1[(size_t)s * 1024 + o] = acc / (1.f + expf(-acc));       // SiLU
}

__global__ void __launch_bounds__(kTmWarps * 32) flow_time_h2_kernel(const float* __restrict__ h1, const float* __restrict__ w2,
                                                                    const float* __restrict__ b2, float* __restrict__ m) {
  __shared__ float x[1024];
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) x[i] = h1[(size_t)s * 1024 + i];
  __syncthreads();
  const int o = blockIdx.y * kTmWarps + warp;
  const float acc = b2[o] + time_dot<1024>(w2 + (size_t)o * 1024, x, lane);
  if (lane == 0) m[(size_t)s * 1024 + o] = mish_f(acc);
}

__global__ void __launch_bounds__(kTmWarps * 32) flow_time_res_kernel(const float* __restrict__ m, const float* const* __restrict__ wr,
                                                                     const float* const* __restrict__ br, int n_res,
                                                                     float* __restrict__ tb) {
  __shared__ float x[1024];
  const int s = blockIdx.x, r = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) x[i] = m[(size_t)s * 1024 + i];
  __syncthreads();
  const int o = blockIdx.z * kTmWarps + warp;
  const float acc = br[r][o] + time_dot<1024>(wr[r] + (size_t)o * 1024, x, lane);
  if (lane == 0) tb[((size_t)s * n_res + r) * 256 + o] = acc;
}

cudaError_t launch_flow_time(const FlowTimes& t_steps, int n_steps, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* const* wr, const float* const* br, int n_res, float* scratch,
                             float* tb, cudaStream_t st) {
  float* h1 = scratch;                                       // [n_steps, 1024] each
  float* m = scratch + (size_t)n_steps * 1024;
  flow_time_h1_kernel<<<dim3(n_steps, 1024 / kTmWarps), kTmWarps * 32, 0, st>>>(t_steps, w1, b1, h1);
  flow_time_h2_kernel<<<dim3(n_steps, 1024 / kTmWarps), kTmWarps * 32, 0, st>>>(h1, w2, b2, m);
  flow_time_res_kernel<<<dim3(n_steps, n_res, 256 / kTmWarps), kTmWarps * 32, 0, st>>>(m, wr, br, n_res, tb);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Input pack (once per decode): X0 [2B, T, 320] (E) = [x | mu | spks | cond] for the conditioned rows, [x | 0 | 0 | 0] for the
// guidance rows; x state [B, T, 80] fp32 = z.  Inputs are upstream's layouts: z, mu, cond [B, 80, T] fp32, spks [B, 80].
// Rows >= length are zero.
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void flow_pack_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ spks,
                                 const float* __restrict__ cond, const int* __restrict__ lengths, int B, int T,
                                 float* __restrict__ x_state, E* __restrict__ X0) {
  __shared__ float tile[3][80][33];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int len = lengths ? min(T, max(0, lengths[b])) : T;
  for (int i = threadIdx.x; i < 3 * 80 * 32; i += blockDim.x) {
    const int which = i / (80 * 32), c = (i / 32) % 80, tt = i % 32;
    const int t = t0 + tt;
    const float* src = which == 0 ? z : (which == 1 ? mu : cond);
    tile[which][c][tt] = (t < len) ? src[((size_t)b * 80 + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 320; i += blockDim.x) {
    const int tt = i / 320, c = i % 320;
    const int t = t0 + tt;
    if (t >= T) continue;
    const bool live = t < len;
    float v;
    if (c < 80) v = tile[0][c][tt];
    else if (c < 160) v = tile[1][c - 80][tt];
    else if (c < 240) v = live ? spks[b * 80 + (c - 160)] : 0.f;
    else v = tile[2][c - 240][tt];
    ElemIO<E>::store(X0 + ((size_t)b * T + t) * 320 + c, v);
    ElemIO<E>::store(X0 + ((size_t)(B + b) * T + t) * 320 + c, c < 80 ? v : 0.f);
    if (c < 80) x_state[((size_t)b * T + t) * 80 + c] = v;
  }
}

cudaError_t launch_flow_pack(const float* z, const float* mu, const float* spks, const float* cond, const int* lengths,
                             int B, int T, float* x_state, void* X0, int elem_bytes, cudaStream_t st) {
  dim3 grid((T + 31) / 32, B);
  if (elem_bytes == 2)
    flow_pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(z, mu, spks, cond, lengths, B, T, x_state, (__nv_bfloat16*)X0);
  else
    flow_pack_kernel<float><<<grid, 256, 0, st>>>(z, mu, spks, cond, lengths, B, T, x_state, (float*)X0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the 256 channels of a row (eps 1e-5), optional Mish, optional per-channel bias (the ResNet block's time
// embedding), mask; writes the conv / GEMM operand (E) and / or fp32.  One warp per row, 8 channels per lane.
// ------------------------------------------------------------------------------------------------
// A warp takes kLnRows rows per trip and has all their loads in flight before the first reduction (one row per warp and
// one block per eight rows ran at 1.8 TB/s: 4000 short blocks whose load -> reduce -> reduce -> store chains did not overlap).
constexpr int kLnRows = 4;

template <typename E>
__global__ void __launch_bounds__(256) flow_ln_kernel(const float* __restrict__ in, int rows, int T,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ tb, const int* __restrict__ lengths,
                                                      int mish, int round_tf32v, E* __restrict__ out_e, float* __restrict__ out_f) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float g[8], bt[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + lane * 8), b1 = *reinterpret_cast<const float4*>(beta + lane * 8 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
  }
  float tbv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tbv[i] = tb ? tb[lane * 8 + i] : 0.f;
  for (int row0 = gw * kLnRows; row0 < rows; row0 += warps * kLnRows) {
    float v[kLnRows][8];
    bool live[kLnRows];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      const int row = row0 + r;
      live[r] = row < rows;
      if (live[r] && lengths) { const int b = row / T; live[r] = row - b * T < lengths[b]; }
      if (live[r]) {
        const float4* p = reinterpret_cast<const float4*>(in + (size_t)row * 256 + lane * 8);
        const float4 a = p[0], c = p[1];
        v[r][0] = a.x; v[r][1] = a.y; v[r][2] = a.z; v[r][3] = a.w; v[r][4] = c.x; v[r][5] = c.y; v[r][6] = c.z; v[r][7] = c.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
      }
    }
    float s[kLnRows], q[kLnRows];
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      s[r] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[r] += v[r][i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRows; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      s[r] *= (1.f / 256.f);
      q[r] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[r][i] - s[r]; q[r] = fmaf(d, d, q[r]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kLnRows; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
    }
#pragma unroll
    for (int r = 0; r < kLnRows; ++r) {
      const int row = row0 + r;
      if (row >= rows) break;
      if (live[r]) {
        const float rstd = rsqrtf(q[r] * (1.f / 256.f) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float y = (v[r][i] - s[r]) * rstd * g[i] + bt[i];
          if (mish) y = mish_f(y);
          v[r][i] = y + tbv[i];
        }
      }
      if (out_f) {
        float4* o = reinterpret_cast<float4*>(out_f + (size_t)row * 256 + lane * 8);
        o[0] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        o[1] = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
      }
      if (out_e) {
        if constexpr (sizeof(E) == 4) {
          if (round_tf32v) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] = round_tf32(v[r][i]);
          }
        }
        ElemIO<E>::template store_vec<8>(out_e + (size_t)row * 256 + lane * 8, v[r]);
      }
    }
  }
}

cudaError_t launch_flow_ln(const float* in, int rows, int T, const float* gamma, const float* beta, const float* tb,
                           const int* lengths, int mish, int round_tf32v, void* out_e, int elem_bytes, float* out_f,
                           cudaStream_t st) {
  const int wpb = 8;
  const int need = (rows + wpb * kLnRows - 1) / (wpb * kLnRows);
  dim3 grid(need < 148 * 8 ? need : 148 * 8);
  if (elem_bytes == 2)
    flow_ln_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, st>>>(in, rows, T, gamma, beta, tb, lengths, mish, 0,
                                                             (__nv_bfloat16*)out_e, out_f);
  else
    flow_ln_kernel<float><<<grid, wpb * 32, 0, st>>>(in, rows, T, gamma, beta, tb, lengths, mish, round_tf32v, (float*)out_e,
                                                     out_f);
  return cudaGetLastError();
}

template <typename E>
__global__ void flow_cast_kernel(const float* __restrict__ in, int rows, int T, const int* __restrict__ lengths,
                                 E* __restrict__ dst, int dst_pitch, int dst_off, int round_tf32v) {
  const size_t n = (size_t)rows * 64;                      // four channels per thread
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i >> 6;
    const int c = (int)(i & 63) * 4;
    const int b = (int)(row / T), t = (int)(row - (size_t)b * T);
    const bool live = !lengths || t < lengths[b];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float4 a = *reinterpret_cast<const float4*>(in + row * 256 + c);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    if constexpr (sizeof(E) == 4) {
      if (round_tf32v) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = round_tf32(v[k]);
      }
    }
    ElemIO<E>::template store_vec<4>(dst + row * dst_pitch + dst_off + c, v);
  }
}

cudaError_t launch_flow_cast(const float* in, int rows, int T, const int* lengths, void* dst, int dst_pitch, int dst_off,
                             int elem_bytes, int round_tf32v, cudaStream_t st) {
  const size_t n = (size_t)rows * 64;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (elem_bytes == 2)
    flow_cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, rows, T, lengths, (__nv_bfloat16*)dst, dst_pitch, dst_off, 0);
  else
    flow_cast_kernel<float><<<blocks, 256, 0, st>>>(in, rows, T, lengths, (float*)dst, dst_pitch, dst_off, round_tf32v);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Self-attention, 8 heads x 64, full (non-causal) over the valid keys of the utterance.
// QKV [B2, T, 1536] (E): q | k | v, each 8 heads x 64.  O [B2, T, 512] (E).
// First correct version on CUDA cores: one warp per query row (a lane owns two of the 64 dimensions), K / V tiles of 64 keys
// staged in shared memory, online softmax in fp32.  (The tcgen05 version — S = Q K^T into TMEM, softmax in the epilogue
// warps, P back through shared memory as the A operand of P V, the structure of conv_pair_kernel — is the next step.)
// ------------------------------------------------------------------------------------------------
constexpr int kAttQ = 8;       // queries (warps) per block
constexpr int kAttK = 64;      // keys per shared-memory tile

template <typename E>
__global__ void __launch_bounds__(kAttQ * 32) flow_attn_kernel(const E* __restrict__ qkv, int T,
                                                               const int* __restrict__ lengths, float scale, int round_tf32v,
                                                               E* __restrict__ out) {
  __shared__ float2 ks[kAttK][32], vs[kAttK][32];
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tq = blockIdx.x * kAttQ + warp;
  const int len = lengths ? min(T, max(0, lengths[b])) : T;
  const E* base = qkv + (size_t)b * T * 1536 + h * 64;
  float2 q = make_float2(0.f, 0.f);
  const bool qlive = tq < len;
  if (qlive) {
    const E* qp = base + (size_t)tq * 1536 + 2 * lane;
    q = make_float2(ElemIO<E>::load(qp) * scale, ElemIO<E>::load(qp + 1) * scale);
  }
  float m = -INFINITY, l = 0.f;
  float2 o = make_float2(0.f, 0.f);
  for (int k0 = 0; k0 < len; k0 += kAttK) {
    __syncthreads();
    for (int i = threadIdx.x; i < kAttK * 32; i += blockDim.x) {
      const int kk = i >> 5, d2 = i & 31;
      const int tk = k0 + kk;
      float2 kv = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
      if (tk < len) {
        const E* kp = base + (size_t)tk * 1536 + 512 + 2 * d2;
        kv = make_float2(ElemIO<E>::load(kp), ElemIO<E>::load(kp + 1));
        vv = make_float2(ElemIO<E>::load(kp + 512), ElemIO<E>::load(kp + 513));
      }
      ks[kk][d2] = kv;
      vs[kk][d2] = vv;
    }
    __syncthreads();
    const int nk = min(kAttK, len - k0);
    if (qlive) {
      for (int kk = 0; kk < nk; ++kk) {
        const float2 kv = ks[kk][lane];
        float s = fmaf(q.x, kv.x, q.y * kv.y);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), p = __expf(s - mn);
        const float2 vv = vs[kk][lane];
        l = fmaf(l, corr, p);
        o.x = fmaf(o.x, corr, p * vv.x);
        o.y = fmaf(o.y, corr, p * vv.y);
        m = mn;
      }
    }
  }
  if (tq < T) {
    float ox = 0.f, oy = 0.f;
    if (qlive && l > 0.f) { ox = o.x / l; oy = o.y / l; }
    if constexpr (sizeof(E) == 4) {
      if (round_tf32v) { ox = round_tf32(ox); oy = round_tf32(oy); }
    }
    E* op = out + ((size_t)b * T + tq) * 512 + h * 64 + 2 * lane;
    ElemIO<E>::store(op, ox);
    ElemIO<E>::store(op + 1, oy);
  }
}

// ------------------------------------------------------------------------------------------------
// The bf16 path: flash-attention on the warp-level tensor-core instruction (mma.sync m16n8k16, fp32 accumulate).
// Block = 8 warps = 128 queries of one (utterance, head); K / V tiles of 64 keys arrive by cp.async into a two-stage ring
// (rows padded to 144 B: every ldmatrix phase is conflict-free), so tile i + 1 loads under tile i's math.  S = Q K^T stays in
// registers — its accumulator layout IS the A-fragment layout of P V (two adjacent 8-key tiles make one 16-key k-step); K
// fragments come through ldmatrix.x4, V fragments through ldmatrix.x4.trans; online softmax in fp32 with ex2.approx, the
// key mask only on an utterance's last tile.
// History (B2 = 64, T = 500, one call): CUDA cores 5.6 ms -> 64-query blocks, synchronous tiles, 32-bit fragment loads
// 0.200 ms -> this version (see profiles/).
// ------------------------------------------------------------------------------------------------
constexpr int kAmQ = 128, kAmK = 64, kAmPitch = 72, kAmWarps = kAmQ / 16;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool live) {
  const int n = live ? 16 : 0;                               // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kAmWarps * 32, 2) flow_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int T,
                                                                      const int* __restrict__ lengths, float scale,
                                                                      __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) __nv_bfloat16 Ks[2][kAmK][kAmPitch];
  __shared__ __align__(16) __nv_bfloat16 Vs[2][kAmK][kAmPitch];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAmQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int len = lengths ? min(T, max(0, lengths[b])) : T;
  const __nv_bfloat16* base = qkv + (size_t)b * T * 1536 + h * 64;
  const int r0 = q0 + 16 * warp + g, r1 = r0 + 8;          // this thread's two query rows
  const int n_tiles = (len + kAmK - 1) / kAmK;
  const uint32_t ks_u = (uint32_t)__cvta_generic_to_shared(&Ks[0][0][0]);
  const uint32_t vs_u = (uint32_t)__cvta_generic_to_shared(&Vs[0][0][0]);
  constexpr uint32_t kStage = kAmK * kAmPitch * 2;
  auto load_tile = [&](int tile, int stage) {
    const int k0 = tile * kAmK;
#pragma unroll
    for (int i = threadIdx.x; i < kAmK * 8; i += kAmWarps * 32) {     // 64 rows x 8 sixteen-byte words, K and V
      const int kk = i >> 3, w = i & 7;
      const int tk = k0 + kk;
      const bool live = tk < len;
      const __nv_bfloat16* kp = base + (size_t)(live ? tk : 0) * 1536 + 512 + 8 * w;
      const uint32_t off = stage * kStage + (uint32_t)(kk * kAmPitch + 8 * w) * 2u;
      cp_async16(ks_u + off, kp, live);
      cp_async16(vs_u + off, kp + 512, live);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_tiles > 0) load_tile(0, 0);
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = 16 * ks + 2 * t;
    qa[ks][0] = r0 < len ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * 1536 + c) : 0u;
    qa[ks][1] = r1 < len ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * 1536 + c) : 0u;
    qa[ks][2] = r0 < len ? *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * 1536 + c + 8) : 0u;
    qa[ks][3] = r1 < len ? *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * 1536 + c + 8) : 0u;
  }
  const float sc2 = scale * 1.4426950408889634f;           // softmax in base 2
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  // ldmatrix lane addresses inside a stage: K (non-transposed, two k-steps per x4) and V (transposed, two d-tiles per x4)
  const uint32_t k_lane = (uint32_t)((lane & 7) * kAmPitch + 8 * (lane >> 3)) * 2u;
  const uint32_t v_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kAmPitch + 8 * (lane >> 4)) * 2u;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1, k0 = tile * kAmK;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                         // tile `tile` has landed; every warp is done with tile - 1
    if (tile + 1 < n_tiles) load_tile(tile + 1, stage ^ 1);
    const uint32_t kb = ks_u + stage * kStage + k_lane, vb = vs_u + stage * kStage + v_lane;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {                      // k-steps 2 kp, 2 kp + 1
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(kb + (uint32_t)(8 * j * kAmPitch + 32 * kp) * 2u));
        mma_bf16_16816(s[j], qa[2 * kp], b0, b1);
        mma_bf16_16816(s[j], qa[2 * kp + 1], b2, b3);
      }
    }
    if (k0 + kAmK > len) {                                   // the utterance's last tile: keys past its length
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
    const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
      s[j][0] = ex2_approx(fmaf(s[j][0], sc2, -mn0)); s[j][1] = ex2_approx(fmaf(s[j][1], sc2, -mn0));
      s[j][2] = ex2_approx(fmaf(s[j][2], sc2, -mn1)); s[j][3] = ex2_approx(fmaf(s[j][3], sc2, -mn1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                        // 16 keys per k-step: S tiles 2 kk and 2 kk + 1
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {                      // d tiles 2 dp, 2 dp + 1
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vb + (uint32_t)(16 * kk * kAmPitch + 16 * dp) * 2u));
        mma_bf16_16816(o[2 * dp], pa, b0, b1);
        mma_bf16_16816(o[2 * dp + 1], pa, b2, b3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = (r0 < len && l0 > 0.f) ? 1.f / l0 : 0.f, i1 = (r1 < len && l1 > 0.f) ? 1.f / l1 : 0.f;
#pragma unroll
  for (int dj = 0; dj < 8; ++dj) {
    const int c = h * 64 + 8 * dj + 2 * t;
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r0) * 512 + c) = pack_bf16x2(o[dj][0] * i0, o[dj][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + r1) * 512 + c) = pack_bf16x2(o[dj][2] * i1, o[dj][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------
// The tf32 path: the same flash-attention on mma.sync m16n8k8 (tf32 operands, fp32 accumulate) over the fp32 q | k | v
// buffer.  Block = 4 warps = 64 queries, K / V tiles of 64 keys by cp.async into a two-stage ring (rows padded to 68
// floats: every fragment load is conflict-free).  S's accumulator layout has a thread's two columns at 2t, 2t + 1 while
// the A fragment of the next MMA wants k-indices t, t + 4: the sum over keys does not care about their order, so the P V
// step simply reads V's rows in that permuted order (b0 = row 2t, b1 = row 2t + 1) — no shuffles.
// (The CUDA-core kernel above took 5.0 ms per call at 2B x T = 64 x 500: 93 % of a tf32 estimator evaluation.)
// ------------------------------------------------------------------------------------------------
constexpr int kAtQ = 64, kAtK = 64, kAtPitch = 68;
constexpr int kAtSmem = 2 * 2 * kAtK * kAtPitch * 4;

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

__global__ void __launch_bounds__(128) flow_attn_tf32_kernel(const float* __restrict__ qkv, int T, const int* __restrict__ lengths,
                                                             float scale, int round_tf32v, float* __restrict__ out) {
  extern __shared__ __align__(16) float at_smem[];
  float* Ks = at_smem;                                   // [2][64][68]
  float* Vs = at_smem + 2 * kAtK * kAtPitch;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAtQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int len = lengths ? min(T, max(0, lengths[b])) : T;
  const float* base = qkv + (size_t)b * T * 1536 + h * 64;
  const int r0 = q0 + 16 * warp + g, r1 = r0 + 8;
  const int n_tiles = (len + kAtK - 1) / kAtK;
  const uint32_t ks_u = (uint32_t)__cvta_generic_to_shared(Ks), vs_u = (uint32_t)__cvta_generic_to_shared(Vs);
  constexpr uint32_t kStage = kAtK * kAtPitch * 4;
  auto load_tile = [&](int tile, int stage) {
    const int k0 = tile * kAtK;
    for (int i = threadIdx.x; i < kAtK * 16; i += 128) {   // 64 rows x 16 sixteen-byte words, K and V
      const int kk = i >> 4, w = i & 15;
      const int tk = k0 + kk;
      const bool live = tk < len;
      const float* kp = base + (size_t)(live ? tk : 0) * 1536 + 512 + 4 * w;
      const uint32_t off = stage * kStage + (uint32_t)(kk * kAtPitch + 4 * w) * 4u;
      cp_async16(ks_u + off, kp, live);
      cp_async16(vs_u + off, kp + 512, live);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_tiles > 0) load_tile(0, 0);
  uint32_t qa[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int c = 8 * ks + t;
    qa[ks][0] = r0 < len ? __float_as_uint(base[(size_t)r0 * 1536 + c]) : 0u;
    qa[ks][1] = r1 < len ? __float_as_uint(base[(size_t)r1 * 1536 + c]) : 0u;
    qa[ks][2] = r0 < len ? __float_as_uint(base[(size_t)r0 * 1536 + c + 4]) : 0u;
    qa[ks][3] = r1 < len ? __float_as_uint(base[(size_t)r1 * 1536 + c + 4]) : 0u;
  }
  const float sc2 = scale * 1.4426950408889634f;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1, k0 = tile * kAtK;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tile + 1 < n_tiles) load_tile(tile + 1, stage ^ 1);
    const float* Kt = Ks + stage * kAtK * kAtPitch;
    const float* Vt = Vs + stage * kAtK * kAtPitch;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    // B fragments of two k-steps per ldmatrix.x4: an 8 x 8 b16 matrix is 8 keys x 16 bytes = 4 floats, and thread i receives
    // the 32-bit word (row i / 4, word i % 4) = K[key g][4 floats' column t]: exactly b0 / b1 of m16n8k8 (four scalar loads
    // per MMA pair otherwise: the kernel was bound by shared-memory instructions)
    const uint32_t k_lane = (uint32_t)__cvta_generic_to_shared(Kt) + (uint32_t)(((lane & 7) * kAtPitch + 4 * (lane >> 3)) * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(k_lane + (uint32_t)((8 * j * kAtPitch + 16 * kp) * 4)));
        mma_tf32_1688(s[j], qa[2 * kp], b0, b1);
        mma_tf32_1688(s[j], qa[2 * kp + 1], b2, b3);
      }
    }
    if (k0 + kAtK > len) {
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
    const float mn0 = fmaxf(m0, mx0 * sc2), mn1 = fmaxf(m1, mx1 * sc2);
    const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
      s[j][0] = ex2_approx(fmaf(s[j][0], sc2, -mn0)); s[j][1] = ex2_approx(fmaf(s[j][1], sc2, -mn0));
      s[j][2] = ex2_approx(fmaf(s[j][2], sc2, -mn1)); s[j][3] = ex2_approx(fmaf(s[j][3], sc2, -mn1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {                           // eight keys per k-step: S tile j, keys in the order 2t | 2t + 1
      uint32_t pa[4];
      pa[0] = tf32_bits(s[j][0]); pa[1] = tf32_bits(s[j][2]); pa[2] = tf32_bits(s[j][1]); pa[3] = tf32_bits(s[j][3]);
#pragma unroll
      for (int dj = 0; dj < 8; ++dj) {
        const uint32_t b0 = __float_as_uint(Vt[(8 * j + 2 * t) * kAtPitch + 8 * dj + g]);
        const uint32_t b1 = __float_as_uint(Vt[(8 * j + 2 * t + 1) * kAtPitch + 8 * dj + g]);
        mma_tf32_1688(o[dj], pa, b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = (r0 < len && l0 > 0.f) ? 1.f / l0 : 0.f, i1 = (r1 < len && l1 > 0.f) ? 1.f / l1 : 0.f;
#pragma unroll
  for (int dj = 0; dj < 8; ++dj) {
    const int c = h * 64 + 8 * dj + 2 * t;
    float a0 = o[dj][0] * i0, a1 = o[dj][1] * i0, b0 = o[dj][2] * i1, b1 = o[dj][3] * i1;
    if (round_tf32v) { a0 = round_tf32(a0); a1 = round_tf32(a1); b0 = round_tf32(b0); b1 = round_tf32(b1); }
    if (r0 < T) *reinterpret_cast<float2*>(out + ((size_t)b * T + r0) * 512 + c) = make_float2(a0, a1);
    if (r1 < T) *reinterpret_cast<float2*>(out + ((size_t)b * T + r1) * 512 + c) = make_float2(b0, b1);
  }
}

// Per device (function attributes are per device): called by gnv_flow_create under its device guard.
cudaError_t flow_kernels_init() {
  preload_kernel((const void*)flow_attn_tf32_kernel);
  preload_kernel((const void*)flow_attn_mma_kernel);
  return cudaFuncSetAttribute(flow_attn_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem);
}

cudaError_t launch_flow_attn(const void* qkv, int B2, int T, const int* lengths, float scale, int round_tf32v, void* out,
                             int elem_bytes, cudaStream_t st) {
  dim3 grid((T + kAttQ - 1) / kAttQ, 8, B2);
  static const bool use_mma = [] { const char* v = getenv("GONOVA_FLOW_ATTN_MMA"); return !(v && atoi(v) == 0); }();
  if (elem_bytes == 2 && use_mma) {
    dim3 gm((T + kAmQ - 1) / kAmQ, 8, B2);
    flow_attn_mma_kernel<<<gm, kAmWarps * 32, 0, st>>>((const __nv_bfloat16*)qkv, T, lengths, scale, (__nv_bfloat16*)out);
    return cudaGetLastError();
  }
  if (elem_bytes == 2) {
    flow_attn_kernel<__nv_bfloat16><<<grid, kAttQ * 32, 0, st>>>((const __nv_bfloat16*)qkv, T, lengths, scale, 0,
                                                                 (__nv_bfloat16*)out);
    return cudaGetLastError();
  }
  if (use_mma) {
    dim3 gt((T + kAtQ - 1) / kAtQ, 8, B2);
    flow_attn_tf32_kernel<<<gt, 128, kAtSmem, st>>>((const float*)qkv, T, lengths, scale, round_tf32v, (float*)out);
    return cudaGetLastError();
  }
  flow_attn_kernel<float><<<grid, kAttQ * 32, 0, st>>>((const float*)qkv, T, lengths, scale, round_tf32v, (float*)out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Classifier-free guidance + Euler step:  v = (1 + cfg) v[b] - cfg v[B + b];  x += dt v;  the x channels of X0 (both
// halves) are rewritten for the next estimator call.  v [2B, T, v_pitch] fp32 (80 channels used).
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void flow_euler_kernel(const float* __restrict__ v, int v_pitch, int B, int T, const int* __restrict__ lengths,
                                  float dt, float cfg, float* __restrict__ x_state, E* __restrict__ X0, int round_tf32v) {
  const size_t n = (size_t)B * T * 80;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 80);
    const size_t row = i / 80;
    const int b = (int)(row / T), t = (int)(row - (size_t)b * T);
    const bool live = !lengths || t < lengths[b];
    const float vc = v[row * v_pitch + c], vu = v[(row + (size_t)B * T) * v_pitch + c];
    float x = x_state[i] + dt * ((1.f + cfg) * vc - cfg * vu);
    if (!live) x = 0.f;
    x_state[i] = x;
    float xe = x;
    if constexpr (sizeof(E) == 4) { if (round_tf32v) xe = round_tf32(xe); }
    ElemIO<E>::store(X0 + row * 320 + c, xe);
    ElemIO<E>::store(X0 + (row + (size_t)B * T) * 320 + c, xe);
  }
}

cudaError_t launch_flow_euler(const float* v, int v_pitch, int B, int T, const int* lengths, float dt, float cfg,
                              float* x_state, void* X0, int elem_bytes, int round_tf32v, cudaStream_t st) {
  const size_t n = (size_t)B * T * 80;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (elem_bytes == 2)
    flow_euler_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(v, v_pitch, B, T, lengths, dt, cfg, x_state, (__nv_bfloat16*)X0, 0);
  else
    flow_euler_kernel<float><<<blocks, 256, 0, st>>>(v, v_pitch, B, T, lengths, dt, cfg, x_state, (float*)X0, round_tf32v);
  return cudaGetLastError();
}

// x_state [B, T, 80] fp32 -> mel [B, 80, T] fp32 (upstream's layout)
__global__ void flow_unpack_kernel(const float* __restrict__ x_state, int T, float* __restrict__ mel) {
  __shared__ float tile[32][81];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * 80; i += blockDim.x) {
    const int tt = i / 80, c = i % 80;
    tile[tt][c] = (t0 + tt < T) ? x_state[((size_t)b * T + t0 + tt) * 80 + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) {
    const int c = i / 32, tt = i % 32;
    if (t0 + tt < T) mel[((size_t)b * 80 + c) * T + t0 + tt] = tile[tt][c];
  }
}

cudaError_t launch_flow_unpack(const float* x_state, int B, int T, float* mel, cudaStream_t st) {
  dim3 grid((T + 31) / 32, B);
  flow_unpack_kernel<<<grid, 256, 0, st>>>(x_state, T, mel);
  return cudaGetLastError();
}

}  // namespace gnv
