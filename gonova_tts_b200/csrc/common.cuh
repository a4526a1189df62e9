// Shared device/host definitions for libgonova_hift.so (sm_100a only).
//
// Data layout in HBM: every intermediate of the decoder is "time-major /
// channels-last"  [B, L, C]  so that one output row of a conv is one GEMM row and the channel
// contraction is contiguous.  Two copies of a tensor may exist:
//   raw  : fp32, the residual stream (what the next residual add reads)
//   act  : element type E (bf16, or fp32 holding tf32-rounded values), already passed through the
//          activation of the conv that will consume it — the A operand of that conv's GEMM.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnv {

enum { ACT_NONE = 0, ACT_SNAKE = 1, ACT_LRELU = 2, ACT_ELU = 3, ACT_SNAKE_FAST = 4, ACT_ELU_FAST = 5, ACT_GELU = 6, ACT_SILU = 7 };

constexpr int kMaxAct = 3;

// Fused epilogue of every conv GEMM.  A GEMM element (row m, column n) of batch b maps to
//   r = n / C_out, c = n % C_out            (r > 0 only for the polyphase ConvTranspose1d GEMM)
//   p0 = m*up + r - pad_out                 (unshifted output row; must be in [0, L_store))
//   row = p0 + shift                        (ReflectionPad1d((1,0)) => shift 1, and p0 == dup_row
//                                            is stored a second time at row 0)
//   w = raw_scale * (acc + bias[c] + res[b,row,c])  (+ raw[b,row,c] if raw_accum)
//   raw[b,row,c] = w;   act_out[i][b,row,c] = act_i(w)  for i < n_act
// Rows at or beyond the utterance's valid length (lengths[b]*len_mul + len_add) are written as 0
// so the next conv sees exactly the zero padding it would see at the end of that utterance.
struct EpiParams {
  int C_out, N_valid;
  int C_pitch;            // channel pitch (elements) of the output tensors; == C_out except conv_post (18 -> 20)
  int up, pad_out, shift, dup_row;
  int L_store, L_out;
  const int* lengths;
  int len_mul, len_add;
  const float* bias;
  const float* res;
  float* raw;
  float raw_scale;
  int raw_accum;
  int n_act;
  void* act_out[kMaxAct];
  const float* act_alpha[kMaxAct];
  int act_kind[kMaxAct];
  float act_slope[kMaxAct];
  int round_tf32;
};

// Geometry of one conv layer seen as an implicit GEMM
//   D[m, n] = sum_{tap, ci} A[b, m*in_stride + off0 + tap*tap_step, ci] * W[n, tap*C_in_w + ci]
// A rows outside [0, L_in) contribute zero (conv zero padding; TMA out-of-bounds fill on the
// tensor-core path, a predicate on the CUDA-core path).
struct ConvGeom {
  int B, L_in;
  int C_in;       // channels actually contracted per tap
  int C_in_ld;    // channel stride of the A tensor (>= C_in; padded to the 128-byte K block)
  int C_in_w;     // channel stride per tap in the packed weights (== C_in_ld on the TC path)
  int M_rows;     // GEMM rows per batch element
  int N_total;    // GEMM columns (weights rows)
  int n_taps, off0, tap_step, in_stride;
  // A as a strided view (elements): row r of utterance b starts at A + b*a_batch_stride + r*a_row_stride.
  // 0 = dense [B, L_in, C_in_ld].  Rows may overlap (a_row_stride < C_in_ld): that is how the strided
  // source_downs convs become plain GEMMs (one GEMM row = k consecutive STFT frames).
  long long a_row_stride, a_batch_stride;
};

#ifdef __CUDACC__

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// GELU (erf form) for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 /
// tf32 rounding of the stored operand), e^{-z^2} through MUFU.EX2 and 1 / (1 + p z) through MUFU.RCP: ~16 instructions
// against ~85 of erff (which made the feed-forward GEMM of the flow estimator epilogue-bound).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float h = 0.5f * p * t * exp2f(-1.4426950408889634f * z * z);   // (1 - erf(z)) / 2
  return x >= 0.f ? fmaf(-x, h, x) : x * h;
}

__device__ __forceinline__ float act_apply(int kind, float x, float alpha, float slope) {
  switch (kind) {
    case ACT_SNAKE: {
      float s = sinf(x * alpha);
      return x + (1.0f / (alpha + 1e-9f)) * (s * s);
    }
    case ACT_SNAKE_FAST: {
      // MUFU.SIN after the hardware range reduction: |abs err| ~ 2^-21 for moderate |alpha*x|, far
      // below the bf16 / tf32 rounding the value gets when it is stored as the next conv's operand.
      float s = __sinf(x * alpha);
      return fmaf(__fdividef(1.0f, alpha + 1e-9f), s * s, x);
    }
    case ACT_LRELU: return x > 0.f ? x : x * slope;
    case ACT_ELU:   return x > 0.f ? x : expm1f(x);
    // exp(x) - 1 through MUFU.EX2: absolute error ~6e-8 near 0, far below the bf16 / tf32 rounding of the stored operand
    case ACT_ELU_FAST: return x > 0.f ? x : __expf(x) - 1.0f;
    case ACT_GELU:  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));     // exact (erf) GELU, torch's default
    case ACT_SILU:  return x / (1.0f + expf(-x));                                  // x sigmoid(x) (the Conformer's "swish")
    default:        return x;
  }
}

template <typename E> struct ElemIO;
template <> struct ElemIO<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
  template <int NC>
  static __device__ __forceinline__ void store_vec(float* p, const float (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC; i += 4)
      *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
};
template <> struct ElemIO<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  template <int NC>
  static __device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&v)[NC]) {
    if constexpr (NC % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NC; i += 8)
        *reinterpret_cast<uint4*>(p + i) = make_uint4(pack2(v[i], v[i + 1]), pack2(v[i + 2], v[i + 3]),
                                                      pack2(v[i + 4], v[i + 5]), pack2(v[i + 6], v[i + 7]));
    } else {
#pragma unroll
      for (int i = 0; i < NC; i += 4)
        *reinterpret_cast<uint2*>(p + i) = make_uint2(pack2(v[i], v[i + 1]), pack2(v[i + 2], v[i + 3]));
    }
  }
};

template <int NC, typename Eout>
__device__ __forceinline__ void epi_emit(const EpiParams& p, int b, int row, int c, int nvalid,
                                         const float (&v)[NC]) {
  int valid_rows = p.L_out;
  if (p.lengths) {
    int lv = p.lengths[b] * p.len_mul + p.len_add;
    valid_rows = lv < valid_rows ? lv : valid_rows;
  }
  const bool live = row < valid_rows;
  const size_t base = ((size_t)b * p.L_out + row) * p.C_pitch + c;
  const bool vec = (nvalid == NC) && ((base % NC) == 0);
  float w[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) w[i] = 0.f;
  if (live) {
    if (vec) {
#pragma unroll
      for (int i = 0; i < NC; i += 4) {
        float4 bb = p.bias ? *reinterpret_cast<const float4*>(p.bias + c + i) : make_float4(0, 0, 0, 0);
        w[i] = v[i] + bb.x; w[i + 1] = v[i + 1] + bb.y; w[i + 2] = v[i + 2] + bb.z; w[i + 3] = v[i + 3] + bb.w;
      }
      if (p.res) {
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
          float4 rr = *reinterpret_cast<const float4*>(p.res + base + i);
          w[i] += rr.x; w[i + 1] += rr.y; w[i + 2] += rr.z; w[i + 3] += rr.w;
        }
      }
#pragma unroll
      for (int i = 0; i < NC; ++i) w[i] *= p.raw_scale;
      if (p.raw_accum) {
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
          float4 rr = *reinterpret_cast<const float4*>(p.raw + base + i);
          w[i] += rr.x; w[i + 1] += rr.y; w[i + 2] += rr.z; w[i + 3] += rr.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (i < nvalid) {
          float t = v[i] + (p.bias ? p.bias[c + i] : 0.f);
          if (p.res) t += p.res[base + i];
          t *= p.raw_scale;
          if (p.raw_accum) t += p.raw[base + i];
          w[i] = t;
        }
      }
    }
  }
  if (p.raw) {
    if (vec) {
      ElemIO<float>::store_vec<NC>(p.raw + base, w);
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (i < nvalid) p.raw[base + i] = w[i];
    }
  }
#pragma unroll
  for (int a = 0; a < kMaxAct; ++a) {
    if (a < p.n_act) {
      float y[NC];
      const int kind = p.act_kind[a];
      const float slope = p.act_slope[a];
      const float* al = p.act_alpha[a];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        float alpha = (al && i < nvalid) ? al[c + i] : 1.f;
        float t = live ? act_apply(kind, w[i], alpha, slope) : 0.f;
        if constexpr (sizeof(Eout) == 4) { if (p.round_tf32) t = round_tf32(t); }
        y[i] = t;
      }
      Eout* o = reinterpret_cast<Eout*>(p.act_out[a]) + base;
      if (vec) {
        ElemIO<Eout>::template store_vec<NC>(o, y);
      } else {
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i < nvalid) ElemIO<Eout>::store(o + i, y[i]);
      }
    }
  }
}

// v[i] is GEMM element (m, n0 + i) of batch b.  Requires n0 % NC == 0.
template <int NC, typename Eout>
__device__ __forceinline__ void epi_apply(const EpiParams& p, int b, int m, int n0, const float (&v)[NC]) {
  if (n0 >= p.N_valid) return;
  const int r = n0 / p.C_out;
  const int c = n0 - r * p.C_out;
  const int p0 = m * p.up + r - p.pad_out;
  if (p0 < 0 || p0 >= p.L_store) return;
  int nvalid = NC;
  if (p.C_out - c < nvalid) nvalid = p.C_out - c;
  if (p.N_valid - n0 < nvalid) nvalid = p.N_valid - n0;
  epi_emit<NC, Eout>(p, b, p0 + p.shift, c, nvalid, v);
  if (p0 == p.dup_row) epi_emit<NC, Eout>(p, b, 0, c, nvalid, v);
}

#endif  // __CUDACC__

}  // namespace gnv
