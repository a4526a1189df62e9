// Host-callable launchers of the flow decoder's small kernels (flow_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnv {

struct FlowTimes { float t[32]; };      // the Euler grid, passed by value (no copy, no synchronisation)

// tb[s][r][256] = Linear_r(Mish(time_mlp(sinusoidal(t_steps[s]))))  for every Euler step s and ResNet block r
// (scratch: 2 * n_steps * 1024 floats)
cudaError_t launch_flow_time(const FlowTimes& t_steps, int n_steps, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* const* wr, const float* const* br, int n_res, float* scratch,
                             float* tb, cudaStream_t st);
// z / mu / cond [B,80,T], spks [B,80] -> x_state [B,T,80] fp32 and X0 [2B,T,320] (E)
cudaError_t launch_flow_pack(const float* z, const float* mu, const float* spks, const float* cond, const int* lengths,
                             int B, int T, float* x_state, void* X0, int elem_bytes, cudaStream_t st);
// LayerNorm(256) [+ Mish] [+ tb] [* mask] of fp32 rows -> E and / or fp32
cudaError_t launch_flow_ln(const float* in, int rows, int T, const float* gamma, const float* beta, const float* tb,
                           const int* lengths, int mish, int round_tf32v, void* out_e, int elem_bytes, float* out_f,
                           cudaStream_t st);
// fp32 [rows, 256] -> E [rows, dst_pitch] at channel offset dst_off, masked
cudaError_t launch_flow_cast(const float* in, int rows, int T, const int* lengths, void* dst, int dst_pitch, int dst_off,
                             int elem_bytes, int round_tf32v, cudaStream_t st);
// opt-in to > 48 KB of dynamic shared memory for the tf32 attention kernel, on the CURRENT device
cudaError_t flow_kernels_init();
// QKV [B2,T,1536] -> O [B2,T,512]: 8 heads x 64, softmax over the valid keys
cudaError_t launch_flow_attn(const void* qkv, int B2, int T, const int* lengths, float scale, int round_tf32v, void* out,
                             int elem_bytes, cudaStream_t st);
// x += dt ((1 + cfg) v[b] - cfg v[B + b]); rewrites the x channels of X0
cudaError_t launch_flow_euler(const float* v, int v_pitch, int B, int T, const int* lengths, float dt, float cfg,
                              float* x_state, void* X0, int elem_bytes, int round_tf32v, cudaStream_t st);
cudaError_t launch_flow_unpack(const float* x_state, int B, int T, float* mel, cudaStream_t st);

}  // namespace gnv
