// Host-callable launchers of the bandwidth-bound kernels (aux_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnv {

// [B, C, L] fp32 (NCT) -> [B, L, C_ld] in E (elem_bytes 2 = bf16, 4 = fp32), channels >= C zero.
cudaError_t launch_nct_to_nlc(const float* in, int B, int C, int L, const int* lengths, void* out, int C_ld,
                              int elem_bytes, int round_tf32, cudaStream_t st);
// [B, L, C_ld] (elem_bytes 2/4) -> [B, C, L] fp32.
cudaError_t launch_nlc_to_nct(const void* in, int B, int L, int C, int C_ld, int elem_bytes, float* out,
                              cudaStream_t st, long long in_batch_stride = 0);
// f0[b,t] = | dot(h[b,t,:C], w) + bias |   (ConvRNNF0Predictor.classifier + abs)
cudaError_t launch_f0_head(const void* h, int elem_bytes, int rows, int T, const int* lengths, int C, const float* w,
                           const float* bias, float* f0, cudaStream_t st);
// SineGen + SourceModuleHnNSF: f0 [B,T] -> s [B, 480T]
cudaError_t launch_source(const float* f0, int B, int T, uint64_t seed, const float* phase_vec, const float* noise,
                          const float* lin_w, const float* lin_b, float* s, cudaStream_t st,
                          const double* f0_sum0 = nullptr, long long sample0 = 0, double* f0_sum_out = nullptr,
                          uint64_t* seed_dev = nullptr, int seed_per_row = 0);
// STFT n_fft 16 hop 4, periodic Hann, center/reflect: s [B, L] (row stride L) ->
// spec [B, total_rows, C_ld] (elem_bytes 2 = bf16 / 4 = fp32): front_rows zero rows, F = L/4+1 frames, zero rows after
cudaError_t launch_stft(const float* s, int B, int L, const int* lengths, void* spec_nlc, int elem_bytes, int round_tf32,
                        int C_ld, int front_rows, int total_rows, cudaStream_t st);
// exp/min/sin + iSTFT + clamp: x [B, F, C_ld >= 18] fp32 (first 18 channels used) -> wav [B, 4(F-1)]
cudaError_t launch_istft(const float* x_nlc, int B, int F, int C_ld, const int* lengths, float limit, float* wav,
                         cudaStream_t st);
// streaming tail
cudaError_t launch_pcm_tail(const float* cur, int64_t cur_stride, const float* prev_tail, const float* fade_w,
                            int rows, int n, int fade, float limit, int16_t* out_i16, float* out_f32,
                            int64_t out_stride, cudaStream_t st);

// G.711 mu-law: int16 PCM [n] (16-byte aligned) -> uint8 [n] (8-byte aligned)
cudaError_t launch_mulaw(const int16_t* in, long long n, uint8_t* out, cudaStream_t st);

}  // namespace gnv
