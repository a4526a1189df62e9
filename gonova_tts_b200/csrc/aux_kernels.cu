// Bandwidth-bound kernels of the decoder: layout packers, f0 head, NSF source, STFT, iSTFT +
// overlap-add, and the streaming PCM tail.  All are HBM-bound byte movers: coalesced, vectorised,
// staged through shared memory where a thread's natural access would be strided.
#include <cmath>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace gnv {

// ------------------------------------------------------------------------------------------------
// layout packers
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void nct_to_nlc_kernel(const float* __restrict__ in, int C, int L, const int* __restrict__ lengths,
                                  E* __restrict__ out, int C_ld, int round) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int Lb = lengths ? min(L, lengths[b]) : L;       // rows past the utterance's end become zeros
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    tile[i][tx] = (c < C && l < Lb) ? in[((size_t)b * C + c) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    if (l < L && c < C_ld) {
      float v = tile[tx][i];
      if constexpr (sizeof(E) == 4) { if (round) v = round_tf32(v); }
      ElemIO<E>::store(out + ((size_t)b * L + l) * C_ld + c, v);
    }
  }
}

template <typename E>
__global__ void nlc_to_nct_kernel(const E* __restrict__ in, int L, int C, int C_ld, long long in_batch_stride,
                                  float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, c = c0 + tx;
    tile[i][tx] = (l < L && c < C) ? ElemIO<E>::load(in + (size_t)b * in_batch_stride + (size_t)l * C_ld + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + tx;
    if (c < C && l < L) out[((size_t)b * C + c) * L + l] = tile[tx][i];
  }
}

cudaError_t launch_nct_to_nlc(const float* in, int B, int C, int L, const int* lengths, void* out, int C_ld,
                              int elem_bytes, int round, cudaStream_t st) {
  dim3 grid((L + 31) / 32, (C_ld + 31) / 32, B), block(32, 8);
  if (elem_bytes == 2)
    nct_to_nlc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(in, C, L, lengths, (__nv_bfloat16*)out, C_ld, 0);
  else
    nct_to_nlc_kernel<float><<<grid, block, 0, st>>>(in, C, L, lengths, (float*)out, C_ld, round);
  return cudaGetLastError();
}

cudaError_t launch_nlc_to_nct(const void* in, int B, int L, int C, int C_ld, int elem_bytes, float* out,
                              cudaStream_t st, long long in_batch_stride) {
  dim3 grid((L + 31) / 32, (C + 31) / 32, B), block(32, 8);
  if (in_batch_stride <= 0) in_batch_stride = (long long)L * C_ld;
  if (elem_bytes == 2)
    nlc_to_nct_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, L, C, C_ld, in_batch_stride, out);
  else
    nlc_to_nct_kernel<float><<<grid, block, 0, st>>>((const float*)in, L, C, C_ld, in_batch_stride, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// f0 head: |Linear(512 -> 1)|, one warp per (b, t) row
// ------------------------------------------------------------------------------------------------
template <typename E>
__global__ void f0_head_kernel(const E* __restrict__ h, int rows, int T, const int* __restrict__ lengths, int C,
                               const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ f0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const E* p = h + (size_t)row * C;
  float acc = 0.f;
  // frames past the utterance's length hold zero rows by definition (a ragged batch may leave them unwritten)
  const bool live = !lengths || (row % T) < lengths[row / T];
  if (live)
    for (int c = lane; c < C; c += 32) acc = fmaf(ElemIO<E>::load(p + c), w[c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) f0[row] = fabsf(acc + bias[0]);
}

cudaError_t launch_f0_head(const void* h, int elem_bytes, int rows, int T, const int* lengths, int C, const float* w,
                           const float* bias, float* f0, cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
  if (elem_bytes == 2)
    f0_head_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)h, rows, T, lengths, C, w, bias, f0);
  else
    f0_head_kernel<float><<<grid, block, 0, st>>>((const float*)h, rows, T, lengths, C, w, bias, f0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// NSF harmonic source (SineGen + SourceModuleHnNSF).  f0 is piecewise constant over a mel frame
// (nearest x480 upsampling), so the running phase cumsum(f0*h/24000) at sample j of frame t is
//   h * ( sum_{t'<t} f0[t']/50  +  (j+1) * f0[t]/24000 )
// The frame prefix is reduced in fp64 by every block for its own frames (T <= a few thousand), so
// the phase never loses precision however long the utterance is; the [9, 480T] harmonic bank
// lives in registers only.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) {      // (0, 1]
  return (__uint2float_rn(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

constexpr int kSrcFramesPerBlock = 4;
constexpr int kSPF = 480;

__global__ void __launch_bounds__(256) source_kernel(const float* __restrict__ f0, int T, uint64_t seed,
                                                      const float* __restrict__ phase_vec,
                                                      const float* __restrict__ noise,
                                                      const float* __restrict__ lin_w,
                                                      const float* __restrict__ lin_b, float* __restrict__ s,
                                                      const double* __restrict__ f0_sum0, long long sample0,
                                                      const uint64_t* __restrict__ seed_dev, int seed_per_row) {
  // a device-resident seed (gnv_inference_dseed) is read at run time: a captured CUDA graph then draws fresh noise at
  // every replay (seed_bump_kernel advances it behind this kernel).  Per-row seeds: row b draws what an utterance
  // decoded ALONE (as row 0) with seed_dev[b] draws — a micro-batched request gets its own noise stream.
  if (seed_dev) seed = seed_per_row ? seed_dev[blockIdx.y] : *seed_dev;
  const uint32_t bkey = (seed_dev && seed_per_row) ? 0u : (uint32_t)blockIdx.y;
  __shared__ double red[8];
  __shared__ double base_s[kSrcFramesPerBlock];
  __shared__ float f0_s[kSrcFramesPerBlock];
  __shared__ float ca_s[9], sa_s[9], lw_s[9];            // 0.1 * lw_h * cos(phi_h), 0.1 * lw_h * sin(phi_h), lw_h
  const int b = blockIdx.y, t0 = blockIdx.x * kSrcFramesPerBlock;
  const float* f0b = f0 + (size_t)b * T;
  double part = 0.0;
  for (int t = threadIdx.x; t < t0; t += blockDim.x) part += (double)f0b[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  if (threadIdx.x < 9) {
    float ph;
    if (phase_vec) {
      ph = phase_vec[b * 9 + threadIdx.x];
    } else if (threadIdx.x == 0) {
      ph = 0.f;                                          // the fundamental keeps phase 0
    } else {
      uint32_t r[4];
      philox4x32_10(bkey, 0xFFFFFFFFu, threadIdx.x, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      ph = (2.f * u01(r[0]) - 1.f) * 3.14159265358979f;
    }
    const float lw = lin_w[threadIdx.x];
    float sp, cp;
    sincosf(ph, &sp, &cp);                               // once per block and harmonic: the precise version
    ca_s[threadIdx.x] = 0.1f * lw * cp;
    sa_s[threadIdx.x] = 0.1f * lw * sp;
    lw_s[threadIdx.x] = lw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = f0_sum0 ? f0_sum0[b] : 0.0;             // streaming: sum of f0 over the frames already emitted
    for (int i = 0; i < 8; ++i) acc += red[i];
    acc /= 50.0;                                         // 480 / 24000 per frame
    for (int i = 0; i < kSrcFramesPerBlock; ++i) {
      const float f = (t0 + i < T) ? f0b[t0 + i] : 0.f;
      base_s[i] = acc - floor(acc);
      f0_s[i] = f;
      acc += (double)f / 50.0;
    }
  }
  __syncthreads();
  const float lb = lin_b[0];
  const size_t L = (size_t)T * kSPF;
  // The nine per-harmonic noise terms enter the output only through the 9 -> 1 linear: sum_h lw_h * namp * N_h(0,1) is ONE
  // normal of standard deviation namp * |lw|, so (unless a test injects `noise`) one normal per sample is drawn, four
  // samples per Philox call, keyed by the absolute sample index.
  float lw_norm = 0.f;
#pragma unroll
  for (int h = 0; h < 9; ++h) lw_norm = fmaf(lw_s[h], lw_s[h], lw_norm);
  lw_norm = sqrtf(lw_norm);
  constexpr int kQuads = kSPF / 4;
  for (int q = threadIdx.x; q < kSrcFramesPerBlock * kQuads; q += blockDim.x) {
    const int fi = q / kQuads, j0 = (q - fi * kQuads) * 4;
    const int t = t0 + fi;
    if (t >= T) break;
    const size_t n0 = (size_t)t * kSPF + j0;
    const float f = f0_s[fi];
    const float uv = f > 10.f ? 1.f : 0.f;
    const float namp = uv * 0.003f + (1.f - uv) * (0.1f / 3.f);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (!noise) {
      const unsigned long long nq = ((unsigned long long)n0 + (unsigned long long)sample0) >> 2;   // absolute quad index
      uint32_t r[4];
      philox4x32_10((uint32_t)nq, (uint32_t)(nq >> 32) ^ (bkey << 8), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      const float r0 = sqrtf(-2.f * __logf(u01(r[0]))), r1 = sqrtf(-2.f * __logf(u01(r[2])));
      float s0, c0, s1, c1;
      __sincosf(6.2831853f * u01(r[1]) - 3.14159265f, &s0, &c0);       // argument in (-pi, pi]: MUFU accuracy range
      __sincosf(6.2831853f * u01(r[3]) - 3.14159265f, &s1, &c1);
      z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
    }
    // Running phase of the fundamental in cycles: reduced mod 1 in fp64 ONCE per four samples; the next three samples
    // add the per-sample increment (<= 0.03 cycles) in fp32.
    const double based = base_s[fi] + (double)(j0 + 1) * ((double)f / 24000.0);
    const float bfrac0 = (float)(based - floor(based));
    const float dcyc = f * (1.0f / 24000.0f);
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // One MUFU sincos of the fundamental (argument reduced to (-pi, pi]); harmonics n = 2..9 by the Chebyshev
      // recurrence  sin((n+1)x) = 2 cos(x) sin(nx) - sin((n-1)x)  (same for cos), two FMAs each instead of a range
      // reduction and a MUFU.SIN per harmonic.  Each harmonic's random phase is a rotation folded into its two
      // coefficients: 0.1 lw_h sin(n x + phi_h) = ca_h sin(n x) + sa_h cos(n x).  Error growth of the recurrence over
      // 9 steps stays below 1e-5 of the 0.1 amplitude (tests/test_gpu_decode.py: <= 2e-5 against the fp64 oracle).
      float tt = bfrac0 + (float)e * dcyc;
      tt -= rintf(tt);                                                  // (-0.5, 0.5]
      float s1v, c1v;
      __sincosf(6.283185307179586f * tt, &s1v, &c1v);
      const float k2 = 2.f * c1v;
      float sp = 0.f, cp = 1.f, sc = s1v, cc = c1v;                      // (n-1) and n
      float acc = 0.f;
      if (noise) {
#pragma unroll
        for (int h = 0; h < 9; ++h) {
          const float sw = ca_s[h] * sc + sa_s[h] * cc;                  // 0.1 lw_h sin(n x + phi_h)
          acc += sw * uv + lw_s[h] * namp * noise[((size_t)b * 9 + h) * L + n0 + e];
          const float sn = fmaf(k2, sc, -sp), cn = fmaf(k2, cc, -cp);
          sp = sc; cp = cc; sc = sn; cc = cn;
        }
        acc += lb;
      } else {
#pragma unroll
        for (int h = 0; h < 9; ++h) {
          acc = fmaf(ca_s[h], sc, acc);
          acc = fmaf(sa_s[h], cc, acc);
          const float sn = fmaf(k2, sc, -sp), cn = fmaf(k2, cc, -cp);
          sp = sc; cp = cc; sc = sn; cc = cn;
        }
        acc = fmaf(acc, uv, lb);
        acc = fmaf(namp * lw_norm, z[e], acc);
      }
      // tanh(a) = 1 - 2 / (exp(2a) + 1): two MUFU ops, absolute error ~1e-7 for the |a| < 2 this layer produces
      const float ex = __expf(2.f * acc);
      o[e] = 1.f - __fdividef(2.f, ex + 1.f);
    }
    *reinterpret_cast<float4*>(s + (size_t)b * L + n0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// f0_sum_out[b] = (f0_sum0 ? f0_sum0[b] : 0) + sum_t f0[b, t]   (fp64: the running phase a stream carries between pushes)
__global__ void __launch_bounds__(256) f0_sum_kernel(const float* __restrict__ f0, int T, const double* __restrict__ f0_sum0,
                                                     double* __restrict__ f0_sum_out) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double part = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) part += (double)f0[(size_t)b * T + t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = f0_sum0 ? f0_sum0[b] : 0.0;
    for (int i = 0; i < 8; ++i) acc += red[i];
    f0_sum_out[b] = acc;
  }
}

__global__ void seed_bump_kernel(uint64_t* seed_dev) { *seed_dev += 1ull; }

cudaError_t launch_source(const float* f0, int B, int T, uint64_t seed, const float* phase_vec, const float* noise,
                          const float* lin_w, const float* lin_b, float* s, cudaStream_t st, const double* f0_sum0,
                          long long sample0, double* f0_sum_out, uint64_t* seed_dev, int seed_per_row) {
  dim3 grid((T + kSrcFramesPerBlock - 1) / kSrcFramesPerBlock, B);
  source_kernel<<<grid, 256, 0, st>>>(f0, T, seed, phase_vec, noise, lin_w, lin_b, s, f0_sum0, sample0, seed_dev,
                                      seed_per_row);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && seed_dev && !seed_per_row) {
    seed_bump_kernel<<<1, 1, 0, st>>>(seed_dev);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess || !f0_sum_out) return e;
  f0_sum_kernel<<<B, 256, 0, st>>>(f0, T, f0_sum0, f0_sum_out);      // after the source kernel: out may alias f0_sum0
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// STFT (n_fft 16, hop 4, periodic Hann, center=True with reflect padding) -> [B, F, 18]
// (9 real parts then 9 imaginary parts per frame, the channel order of torch.cat([real, imag], 1)).
// ------------------------------------------------------------------------------------------------
constexpr int kStftFrames = 256;

// Windowed DFT bases of the 16-point transforms, filled once per device by init_dft_tables():
//   STFT : cw[k][n] = w[n] cos(2 pi k n / 16),  sw[k][n] = -w[n] sin(2 pi k n / 16)          (periodic Hann w)
//   iSTFT: cb[k][n] = w[n] c_k cos(2 pi k n / 16) / 16,  sb[k][n] = -w[n] c_k sin(...) / 16 (c_k = 1 for DC/Nyquist, else 2;
//          the imaginary parts of DC and Nyquist are ignored, as cuFFT C2R does),  w2[n] = w[n]^2
__constant__ float c_stft_cw[9][16], c_stft_sw[9][16];
__constant__ float c_istft_cb[9][16], c_istft_sb[9][16], c_istft_w2[16];

static cudaError_t init_dft_tables() {
  static std::mutex mu;
  static unsigned long long done_mask = 0ull;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 64 && (done_mask >> dev) & 1ull) return cudaSuccess;
  float cw[9][16], sw[9][16], cb[9][16], sb[9][16], w2[16];
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < 9; ++k)
    for (int n = 0; n < 16; ++n) {
      const double win = 0.5 - 0.5 * cos(2.0 * pi * n / 16.0);
      const double ang = 2.0 * pi * ((k * n) % 16) / 16.0;
      const double ck = (k == 0 || k == 8) ? 1.0 : 2.0;
      cw[k][n] = (float)(win * cos(ang));
      sw[k][n] = (float)(-win * sin(ang));
      cb[k][n] = (float)(win * ck * cos(ang) / 16.0);
      sb[k][n] = (k == 0 || k == 8) ? 0.f : (float)(-win * ck * sin(ang) / 16.0);
      if (k == 0) w2[n] = (float)(win * win);
    }
  if ((e = cudaMemcpyToSymbol(c_stft_cw, cw, sizeof(cw))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_stft_sw, sw, sizeof(sw))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_istft_cb, cb, sizeof(cb))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_istft_sb, sb, sizeof(sb))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_istft_w2, w2, sizeof(w2))) != cudaSuccess) return e;
  if (dev < 64) done_mask |= 1ull << dev;
  return cudaSuccess;
}

// Output rows: [front zero rows | F frames | back zero rows] = total_rows rows of C_ld elements of E
// (channels >= 18 are zero).  The padding rows are what lets the strided source_downs convs read their
// conv padding (and the K padding of the GEMM view) as plain zeros.
template <typename E>
__global__ void __launch_bounds__(kStftFrames) stft_kernel(const float* __restrict__ s, int L,
                                                           const int* __restrict__ lengths, E* __restrict__ spec,
                                                           int C_ld, int front, int total_rows, int round) {
  __shared__ __align__(16) float x_s[kStftFrames * 4 + 12];
  const int b = blockIdx.y, rb = blockIdx.x * kStftFrames;     // first output row of this block
  const int fb = rb - front;                                   // its frame index (may be negative)
  const int Lb = lengths ? min(L, lengths[b] * kSPF) : L;      // this utterance's samples
  const int Fb = Lb / 4 + 1;
  const float* sb = s + (size_t)b * L;
  {
    constexpr int kN = kStftFrames * 4 + 12, kIt = (kN + kStftFrames - 1) / kStftFrames;
    float v[kIt];
#pragma unroll
    for (int u = 0; u < kIt; ++u) {                            // all of the thread's loads in flight together
      const int i = threadIdx.x + u * kStftFrames;
      int idx = fb * 4 - 8 + i;
      if (idx < 0 && idx >= -8) idx = -idx;
      if (idx >= Lb && idx < Lb + 8) idx = 2 * (Lb - 1) - idx;
      v[u] = (i < kN && idx >= 0 && idx < Lb) ? sb[idx] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kIt; ++u) {
      const int i = threadIdx.x + u * kStftFrames;
      if (i < kN) x_s[i] = v[u];
    }
  }
  __syncthreads();
  const int f = fb + threadIdx.x;
  float x[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v4 = *reinterpret_cast<const float4*>(&x_s[threadIdx.x * 4 + 4 * q]);
    x[4 * q] = v4.x; x[4 * q + 1] = v4.y; x[4 * q + 2] = v4.z; x[4 * q + 3] = v4.w;
  }
  const bool live = f >= 0 && f < Fb;
  float o[24];                                   // the frame's output row: 9 re, 9 im, zero padding channels
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      re = fmaf(x[n], c_stft_cw[k][n], re);
      im = fmaf(x[n], c_stft_sw[k][n], im);
    }
    o[k] = live ? re : 0.f;
    o[9 + k] = (live && k != 0 && k != 8) ? im : 0.f;
  }
#pragma unroll
  for (int c = 18; c < 24; ++c) o[c] = 0.f;
  const int row = rb + threadIdx.x;
  if (row >= total_rows) return;
  E* orow = spec + ((size_t)b * total_rows + row) * C_ld;
  // one thread = one row: 16-byte stores straight from registers when the row pitch allows it
  if constexpr (sizeof(E) == 2) {
    if (C_ld == 24) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        reinterpret_cast<uint4*>(orow)[q] = make_uint4(ElemIO<E>::pack2(o[8 * q], o[8 * q + 1]), ElemIO<E>::pack2(o[8 * q + 2], o[8 * q + 3]),
                                                      ElemIO<E>::pack2(o[8 * q + 4], o[8 * q + 5]), ElemIO<E>::pack2(o[8 * q + 6], o[8 * q + 7]));
      return;
    }
  } else {
    if (round) {
#pragma unroll
      for (int c = 0; c < 18; ++c) o[c] = round_tf32(o[c]);
    }
    if (C_ld == 20) {
#pragma unroll
      for (int q = 0; q < 5; ++q)
        reinterpret_cast<float4*>(orow)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      return;
    }
  }
#pragma unroll
  for (int c = 0; c < 24; ++c)
    if (c < C_ld) ElemIO<E>::store(orow + c, o[c]);
}

cudaError_t launch_stft(const float* s, int B, int L, const int* lengths, void* spec_nlc, int elem_bytes, int round,
                        int C_ld, int front_rows, int total_rows, cudaStream_t st) {
  if (C_ld < 18 || C_ld > 24 || front_rows < 0 || total_rows < front_rows + L / 4 + 1) return cudaErrorInvalidValue;
  if (cudaError_t e = init_dft_tables()) return e;
  dim3 grid((total_rows + kStftFrames - 1) / kStftFrames, B);
  if (elem_bytes == 2)
    stft_kernel<__nv_bfloat16><<<grid, kStftFrames, 0, st>>>(s, L, lengths, (__nv_bfloat16*)spec_nlc, C_ld, front_rows,
                                                            total_rows, 0);
  else
    stft_kernel<float><<<grid, kStftFrames, 0, st>>>(s, L, lengths, (float*)spec_nlc, C_ld, front_rows, total_rows, round);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// iSTFT head: mag = min(exp(x[:9]), 100), ph = sin(x[9:]), X = mag*exp(i*ph); inverse 16-point real
// DFT per frame, synthesis window, 4-frame overlap-add, division by the window envelope, trim of the
// 8-sample centre padding, clamp(+-limit).  A block owns 253 frames' worth of output (1012 samples)
// and recomputes the 3 halo frames it overlaps with, so there is no inter-block exchange.
// ------------------------------------------------------------------------------------------------
constexpr int kIstftThreads = 256;
constexpr int kIstftNew = kIstftThreads - 3;

__global__ void __launch_bounds__(kIstftThreads) istft_kernel(const float* __restrict__ x, int F, int C_ld,
                                                              const int* __restrict__ lengths, float limit,
                                                              float* __restrict__ wav) {
  __shared__ float xin[kIstftThreads * 19];             // 18 used floats per frame, odd pitch: no bank conflicts
  __shared__ float fr[kIstftThreads][17];
  const int b = blockIdx.y;
  const int L = 4 * (F - 1);
  const int Fb = lengths ? min(F, lengths[b] * (kSPF / 4) + 1) : F;
  const int Lb = 4 * (Fb - 1);
  const int m0 = blockIdx.x * (kIstftNew * 4);
  const int fbase = m0 / 4 - 1;
  {
    const long long lo = (long long)fbase * C_ld, total = (long long)F * C_ld;
    const float* xb = x + (size_t)b * F * C_ld;
    if (C_ld == 20) {
      // 256 frames x 80 B = 1280 16-byte words, five per thread, all five loads in flight together (rows of 20 floats keep
      // every word inside one frame: word w = frame w / 5, floats 4 (w % 5) ...).  One round trip to HBM per block instead
      // of five dependent ones of scalar loads: the kernel is a 22 B/sample byte mover and was latency-bound.
      const float4* xb4 = reinterpret_cast<const float4*>(xb);
      const long long lo4 = lo / 4, total4 = total / 4;
      float4 v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const long long g = lo4 + threadIdx.x + (long long)u * kIstftThreads;
        v[u] = (g >= 0 && g < total4) ? __ldg(xb4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int w = threadIdx.x + u * kIstftThreads;
        const int fr_i = w / 5, k = (w - fr_i * 5) * 4;
        float* dst = xin + fr_i * 19 + k;
        dst[0] = v[u].x; dst[1] = v[u].y;
        if (k < 16) { dst[2] = v[u].z; dst[3] = v[u].w; }      // floats 18, 19 of a row are the pitch padding
      }
    } else {
    // flat, coalesced copy of 256 frames x C_ld floats, four loads in flight per thread (one at a time the block spent
    // its life waiting for ~20 dependent global loads: 0.30 ms for a kernel whose bytes take 0.05 ms)
    const int n_it = C_ld;                                   // kIstftThreads * C_ld elements / kIstftThreads threads
    for (int it0 = 0; it0 < n_it; it0 += 4) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long g = lo + threadIdx.x + (long long)(it0 + u) * kIstftThreads;
        v[u] = (it0 + u < n_it && g >= 0 && g < total) ? xb[g] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (it0 + u < n_it) {
          const int i = threadIdx.x + (it0 + u) * kIstftThreads;
          const int fr_i = i / C_ld, k = i - fr_i * C_ld;
          if (k < 18) xin[fr_i * 19 + k] = v[u];
        }
      }
    }
    }
  }
  __syncthreads();
  {
    const int f = fbase + threadIdx.x;
    float re[9], im[9];
    const bool live = f >= 0 && f < Fb;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float mag = fminf(expf(xin[threadIdx.x * 19 + k]), 100.f);
      // sin of the raw phase channel: one explicit 2*pi reduction, then MUFU.SIN on |r| <= pi (abs error ~ 2^-21)
      const float xr = xin[threadIdx.x * 19 + 9 + k];
      const float ph = __sinf(fmaf(-6.283185307179586f, rintf(xr * 0.15915494309189535f), xr));
      float sn, cs;
      __sincosf(ph, &sn, &cs);                               // |ph| <= 1: MUFU abs error ~ 2^-21
      re[k] = live ? mag * cs : 0.f;
      im[k] = live ? mag * sn : 0.f;
    }
    // k outer, n inner: the 16 coefficients of a bin are contiguous in constant memory, so they arrive as wide uniform
    // loads (n outer made ptxas issue one LDCU per FMA, 298 of them: the kernel's whole run time).  The order of
    // the additions into each acc[n] is unchanged (k ascending, cos term before sin term).
    float acc[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        acc[n] = fmaf(re[k], c_istft_cb[k][n], acc[n]);
        acc[n] = fmaf(im[k], c_istft_sb[k][n], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) fr[threadIdx.x][n] = acc[n];
  }
  __syncthreads();
  if (threadIdx.x < kIstftNew) {
    const int m = m0 + 4 * threadIdx.x;
    if (m < L) {
      const int fhi = m0 / 4 + 2 + threadIdx.x;      // newest frame covering padded sample m + 8
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float acc = 0.f, env = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = fhi - j;
          if (f >= 0 && f < Fb) {
            acc += fr[threadIdx.x + 3 - j][e + 4 * j];
            env += c_istft_w2[e + 4 * j];
          }
        }
        float v = (m + e < Lb) ? acc / env : 0.f;
        o[e] = fminf(fmaxf(v, -limit), limit);
      }
      *reinterpret_cast<float4*>(wav + (size_t)b * L + m) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

cudaError_t launch_istft(const float* x_nlc, int B, int F, int C_ld, const int* lengths, float limit, float* wav,
                         cudaStream_t st) {
  const int L = 4 * (F - 1);
  if (L <= 0) return cudaSuccess;
  if (C_ld < 18 || C_ld > 20) return cudaErrorInvalidValue;
  if (cudaError_t e = init_dft_tables()) return e;
  dim3 grid((L + kIstftNew * 4 - 1) / (kIstftNew * 4), B);
  istft_kernel<<<grid, kIstftThreads, 0, st>>>(x_nlc, F, C_ld, lengths, limit, wav);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// streaming tail: head fade / crossfade, clamp, float -> int16 pack.  6 B/sample (4 in, 2 out).
// Every fp32 operation is an explicitly rounded intrinsic (no FMA contraction) so the result is
// bit-identical to the numpy definition in oracle/tail_ref.py.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tail_one(float c, int i, int fade, const float* __restrict__ prev,
                                          const float* __restrict__ fw, float limit) {
  float v = c;
  if (i < fade) {
    const float w = fw[i];
    if (prev) v = __fadd_rn(__fmul_rn(prev[i], __fsub_rn(1.0f, w)), __fmul_rn(c, w));
    else v = __fmul_rn(c, w);
  }
  return fminf(fmaxf(v, -limit), limit);
}
__device__ __forceinline__ int16_t pack_i16(float v) {
  int q = __float2int_rn(__fmul_rn(v, 32767.0f));
  q = max(-32768, min(32767, q));
  return (int16_t)q;
}

template <bool VEC>
__global__ void __launch_bounds__(256) pcm_tail_kernel(const float* __restrict__ cur, int64_t cur_stride,
                                                        const float* __restrict__ prev,
                                                        const float* __restrict__ fw, int n, int fade, float limit,
                                                        int16_t* __restrict__ o16, float* __restrict__ o32,
                                                        int64_t out_stride) {
  const int row = blockIdx.y;
  const float* c = cur + (size_t)row * cur_stride;
  const float* p = prev ? prev + (size_t)row * fade : nullptr;
  if constexpr (VEC) {
    const int n4 = n >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += gridDim.x * blockDim.x) {
      const int i = q << 2;
      const float4 x = __ldg(reinterpret_cast<const float4*>(c + i));
      float v0 = tail_one(x.x, i, fade, p, fw, limit), v1 = tail_one(x.y, i + 1, fade, p, fw, limit);
      float v2 = tail_one(x.z, i + 2, fade, p, fw, limit), v3 = tail_one(x.w, i + 3, fade, p, fw, limit);
      if (o32) *reinterpret_cast<float4*>(o32 + (size_t)row * out_stride + i) = make_float4(v0, v1, v2, v3);
      if (o16) {
        short4 s4 = make_short4(pack_i16(v0), pack_i16(v1), pack_i16(v2), pack_i16(v3));
        *reinterpret_cast<short4*>(o16 + (size_t)row * out_stride + i) = s4;
      }
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const float v = tail_one(c[i], i, fade, p, fw, limit);
      if (o32) o32[(size_t)row * out_stride + i] = v;
      if (o16) o16[(size_t)row * out_stride + i] = pack_i16(v);
    }
  }
}

cudaError_t launch_pcm_tail(const float* cur, int64_t cur_stride, const float* prev_tail, const float* fade_w,
                            int rows, int n, int fade, float limit, int16_t* out_i16, float* out_f32,
                            int64_t out_stride, cudaStream_t st) {
  if (rows <= 0 || n <= 0) return cudaSuccess;
  if (!fade_w) fade = 0;
  const bool vec = (n % 4 == 0) && (((uintptr_t)cur & 15) == 0) && (cur_stride % 4 == 0) &&
                   (out_stride % 4 == 0) && (!out_f32 || ((uintptr_t)out_f32 & 15) == 0) &&
                   (!out_i16 || ((uintptr_t)out_i16 & 7) == 0);
  const int per = vec ? n / 4 : n;
  int bx = (per + 255) / 256;
  const int cap = 148 * 8;                           // a few waves of the 148 SMs; grid-stride beyond
  if (bx > cap) bx = cap;
  dim3 grid(bx, rows);
  if (vec)
    pcm_tail_kernel<true><<<grid, 256, 0, st>>>(cur, cur_stride, prev_tail, fade_w, n, fade, limit, out_i16, out_f32,
                                                 out_stride);
  else
    pcm_tail_kernel<false><<<grid, 256, 0, st>>>(cur, cur_stride, prev_tail, fade_w, n, fade, limit, out_i16,
                                                  out_f32, out_stride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// G.711 mu-law companding of int16 PCM (wire format of telephony clients): 2 B in, 1 B out per sample.
// Same arithmetic as CPython's audioop.lin2ulaw (14-bit magnitude, bias 33, clip 8159); bit-exact.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulaw_one(int v16) {
  int x = v16 >> 2;
  const uint32_t mask = x < 0 ? 0x7Fu : 0xFFu;
  x = x < 0 ? -x : x;
  x = min(x, 8159) + 33;
  // segment = index of the first end in {0x3F,0x7F,...,0x1FFF} that is >= x  ==  max(0, position of the MSB - 5)
  const int seg = max(0, 26 - __clz(x));          // x in [33, 8192]: MSB position 5..13 -> seg 0..8
  const uint32_t u = seg >= 8 ? 0x7Fu : (uint32_t)((seg << 4) | ((x >> (seg + 1)) & 0xF));
  return (u ^ mask) & 0xFFu;
}

__global__ void __launch_bounds__(256) mulaw_kernel(const int16_t* __restrict__ in, long long n, uint8_t* __restrict__ out) {
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n8; q += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + q);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = mulaw_one((int)(short)(w[i] & 0xFFFFu)), b = mulaw_one((int)(short)(w[i] >> 16));
      const uint32_t pr = a | (b << 8);
      if (i < 2) lo |= pr << (16 * i); else hi |= pr << (16 * (i - 2));
    }
    reinterpret_cast<uint2*>(out)[q] = make_uint2(lo, hi);
  }
  // tail (n not a multiple of 8)
  const long long t0 = n8 << 3;
  for (long long i = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint8_t)mulaw_one(in[i]);
}

cudaError_t launch_mulaw(const int16_t* in, long long n, uint8_t* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (((uintptr_t)in & 15) || ((uintptr_t)out & 7)) return cudaErrorInvalidValue;
  long long blocks = ((n >> 3) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mulaw_kernel<<<(int)blocks, 256, 0, st>>>(in, n, out);
  return cudaGetLastError();
}

}  // namespace gnv
