// Cycles per MUFU.EX2 warp instruction: fp32 against the packed f16x2 / bf16x2 forms (two exponentials per lane per instruction).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench tools/mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = MODE == 0 ? __float_as_uint(-0.001f * (threadIdx.x + i)) : (0xB000B000u + threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) {   // what a softmax needs around a packed exp: two FFMAs, one pack, one MUFU
        float a = __uint_as_float(x[i]), b = a + 1.f;
        a = fmaf(a, 0.5f, -1.f); b = fmaf(b, 0.5f, -1.f);
        uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        x[i] = h;
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "2 FFMA + cvt.f16x2 + ex2.f16x2"};
  for (int warps : {4, 8, 16, 32}) {
    for (int m = 0; m < 4; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
        if (m == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
        if (m == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
        if (m == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_sm_instr = (double)warps * iters * 8;
      printf("%2d warps/SM  %-34s %.2f cycles per warp instruction per SM -> %.1f exps per clock per SM\n", warps, names[m],
             h[0] / per_sm_instr, per_sm_instr * 32 * (m == 0 ? 1 : 2) / h[0]);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
