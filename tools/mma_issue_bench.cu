// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS operands) as a function of N, of the issue style
// and of how many accumulators the stream alternates between.  One CTA per SM, operands resident in shared memory
// (contents irrelevant), no TMA, no epilogue: this is the tensor pipe + operand fetch + issue path floor that the
// conv kernels' main loops are measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_issue_bench tools/mma_issue_bench.cu && ./mma_issue_bench
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: start address >> 4, LBO 1, SBO 1024 B (8 rows x 128 B), version 1, swizzle 2
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// style 0: whole role under `if (lane == 0)`;  style 1: warp-uniform loop, elect_one around the MMAs
template <int STYLE, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int N, int groups, long long* cycles, int flags = 0) {
  constexpr int per_group = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t full[4], empty[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const uint32_t sA = smem_u32(base), sB = sA + 4 * 16384;          // 4 A tiles (128 x 64 bf16), then B (N x 64 bf16)
  for (int i = threadIdx.x; i < (4 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // instruction descriptor: D = f32, A = B = bf16, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t a0 = desc_sw128(sA), b0 = desc_sw128(sB);
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    if (STYLE == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int g = 0; g < groups; ++g) {
          const uint64_t ag = a0 + (uint64_t)((g & 3) * 1024);
#pragma unroll
          for (int i = 0; i < per_group; ++i)
            umma(tmem + (uint32_t)((i % NACC) * N), ag + 2 * (i & 3), b0 + 2 * (i & 3), idesc, (g | (i >= NACC)) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
      }
    } else if (STYLE == 2) {
      // the conv kernels' protocol: per group wait on a "full" barrier (armed by a producer warp that itself waits for
      // the MMAs' commit on "empty"), fence, elected lane issues 8 MMAs and commits to "empty"
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        const int sl = g & 3;
        mbar_wait(smem_u32(&full[sl]), (uint32_t)((g >> 2) & 1));
        if (!(flags & 1)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t ag = a0 + (uint64_t)((g & 3) * 1024);
          const int reps = (flags & 2) ? 4 : 1;           // 8 or 32 MMAs per barrier round
          for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < per_group; ++i)
              umma(tmem + (uint32_t)((i % NACC) * N), ag + 2 * (i & 3), b0 + 2 * (i & 3), idesc, (g | r | (i >= NACC)) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty[sl]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
      t1 = clock64();
      if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    } else {
      // flags & 4: a tcgen05.commit (to a barrier nobody waits on) after every group of 8 — isolates the cost of the
      // commit itself from the cost of the full/empty handshake
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        if (elect_one()) {
          const uint64_t ag = a0 + (uint64_t)((g & 3) * 1024);
#pragma unroll
          for (int i = 0; i < per_group; ++i)
            umma(tmem + (uint32_t)((i % NACC) * N), ag + 2 * (i & 3), b0 + 2 * (i & 3), idesc, (g | (i >= NACC)) ? 1u : 0u);
          if (flags & 4) umma_commit(smem_u32(&empty[g & 3]));
        }
        __syncwarp();
        if (flags & 8) {      // ... plus a wait on an already-completed barrier phase (the cost of the wait instruction alone)
          mbar_wait(smem_u32(&full[0]), 1u);
        }
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
      t1 = clock64();
      if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
  }
  if (STYLE == 2 && warp == 2) {
    for (int g = 0; g < groups; ++g) {
      const int sl = g & 3;
      mbar_wait(smem_u32(&empty[sl]), (uint32_t)((g >> 2) & 1) ^ 1u);
      if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[sl])) : "memory");
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 4 * 16384 + 32768 + 1024;
  cudaFuncSetAttribute(bench<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d = nullptr;
  cudaMalloc(&d, sms * sizeof(long long));
  std::vector<long long> h(sms);
  const int groups = 512, per_group = 8;
  (void)per_group;
  printf("tcgen05.mma kind::f16 M=128 K=16 SS, %d CTAs, %d MMAs each; cycles per MMA (median over CTAs)\n", sms, groups * per_group);
  printf("%6s %6s %22s %22s %22s %10s\n", "N", "accs", "lane0-if (ELECT loops)", "elected lane, uniform", "elected + ring protocol", "floor");
  auto run = [&](int style, int n_acc, int N, int flags) -> double {
    for (int rep = 0; rep < 2; ++rep) {
      if (style == 0 && n_acc == 1) bench<0, 1><<<sms, 128, smem>>>(N, groups, d);
      if (style == 0 && n_acc == 2) bench<0, 2><<<sms, 128, smem>>>(N, groups, d);
      if (style == 1 && n_acc == 1) bench<1, 1><<<sms, 128, smem>>>(N, groups, d);
      if (style == 1 && n_acc == 2) bench<1, 2><<<sms, 128, smem>>>(N, groups, d, flags);
      if (style == 2 && n_acc == 1) bench<2, 1><<<sms, 128, smem>>>(N, groups, d, flags);
      if (style == 2 && n_acc == 2) bench<2, 2><<<sms, 128, smem>>>(N, groups, d, flags);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<long long> v(h);
    std::sort(v.begin(), v.end());
    return (double)v[sms / 2] / (groups * per_group * ((style == 2 && (flags & 2)) ? 4 : 1));
  };
  for (int N : {32, 64, 128, 256})
    for (int n_acc : {1, 2}) {
      if (n_acc * N > 512) continue;
      printf("%6d %6d %22.1f %22.1f %22.1f %10d\n", N, n_acc, run(0, n_acc, N, 0), run(1, n_acc, N, 0), run(2, n_acc, N, 0), 128 * N / 256);
    }
  printf("ring protocol variants, 2 accumulators (cycles per MMA):\n%6s %22s %22s %22s %22s\n", "N", "8 per round", "8, no tcgen05.fence", "32 per round", "32, no fence");
  for (int N : {64, 128})
    printf("%6d %22.1f %22.1f %22.1f %22.1f\n", N, run(2, 2, N, 0), run(2, 2, N, 1), run(2, 2, N, 2), run(2, 2, N, 3));
  printf("no handshake, 2 accumulators (cycles per MMA):\n%6s %22s %22s %22s\n", "N", "plain", "commit per 8", "commit + passed wait per 8");
  for (int N : {64, 128})
    printf("%6d %22.1f %22.1f %22.1f\n", N, run(1, 2, N, 0), run(1, 2, N, 4), run(1, 2, N, 12));
  cudaFree(d);
  return 0;
}
