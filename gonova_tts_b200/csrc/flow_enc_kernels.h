// Host-callable launchers of the flow front's small kernels (flow_enc_kernels.cu): everything of the token embedding +
// upsampling Conformer encoder that is not a GEMM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnv {

constexpr int kEncC = 512;       // model width
constexpr int kEncH = 8;         // heads (x 64)
constexpr int kEncQkv = 2048;    // [q + pos_bias_u | q + pos_bias_v | k | v] per row

// opt-in to > 48 KB of dynamic shared memory for the tensor-core attention kernel, on the CURRENT device
cudaError_t flow_enc_init();
// E[b, l, :] = table[clamp(tokens[b, l], 0, vocab - 1)] for l < token_len[b], else 0
cudaError_t launch_enc_embed(const int32_t* tokens, const int32_t* token_len, const float* table, int vocab, int B, int L,
                             void* out_e, int elem_bytes, int round_tf32v, cudaStream_t st);
// LayerNorm over 512 channels of fp32 rows [B*T, 512] -> E and / or fp32; rows at or past lengths[b] * len_mul are zero
cudaError_t launch_enc_ln(const float* in, int B, int T, const float* gamma, const float* beta, float eps,
                          const int32_t* lengths, int len_mul, void* out_e, int elem_bytes, int round_tf32v, float* out_f,
                          cudaStream_t st);
// fp32 [B*T, 512] -> E, rows past the length zero
cudaError_t launch_enc_cast(const float* in, int B, int T, const int32_t* lengths, int len_mul, void* out_e, int elem_bytes,
                            int round_tf32v, cudaStream_t st);
// P[h][r][64] = linear_pos(pos_emb)[r, h*64 ...], r in [0, 2T-1): row r is relative position (T-1) - r (ESPnet layout);
// fp32 and a bf16 copy
cudaError_t launch_enc_pos(const float* w_pos, int T, float* P, void* P_bf16, cudaStream_t st);
// relative-position self-attention of one Conformer layer: QKV [B*T, 2048] (E) -> O [B*T, 512] (E)
//   score(i, j) = ((q_i + u) k_j + (q_i + v) p_{(T-1) - i + j}) / 8, softmax over the keys j < len, rows i >= len are zero
//   (bf16 with P_bf16: mma.sync tensor-core kernel; otherwise fp32 on the CUDA cores)
cudaError_t launch_enc_attn(const void* qkv, const float* P, const void* P_bf16, int B, int T, const int32_t* lengths,
                            int len_mul, void* out_e, int elem_bytes, int round_tf32v, cudaStream_t st);
// [B*T, 80] fp32 -> mu [B, 80, T], frames past the length zero
cudaError_t launch_enc_mu(const float* in, int B, int T, const int32_t* lengths, int len_mul, float* mu, cudaStream_t st);
// spks[b] = W normalize(embedding[b]) + bias   (192 -> 80)
cudaError_t launch_enc_spk(const float* embedding, const float* w, const float* bias, int B, float* spks, cudaStream_t st);

}  // namespace gnv
