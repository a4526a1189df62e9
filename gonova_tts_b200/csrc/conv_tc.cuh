// tcgen05 implicit-GEMM Conv1d / polyphase ConvTranspose1d for sm_100a.
//
//   D[m, n] = sum_{tap} sum_{ci} A[b, m + off0 + tap*tap_step, ci] * W[n, tap*C_in + ci]
//
// GEMM view: M = time rows (128 per CTA = the 128 TMEM lanes), N = output channels
// (BLOCK_N <= 256 fp32 TMEM columns), K = taps x input channels, walked in 128-byte K blocks
// (64 bf16 / 32 tf32 channels).  Per K block one TMA box brings the A tile — 128 consecutive time
// rows of one channel block, *shifted by the tap offset*; rows before 0 or past L_in are filled
// with zeros by TMA, which is exactly the conv's zero padding at each utterance's ends (the tensor
// map is 3-D [B, L, C] so the fill is per utterance) — and one TMA box brings the W tile.  Both land
// in the canonical K-major SWIZZLE_128B layout the UMMA shared-memory descriptors expect.
//
// Warp roles (256 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane,
// tcgen05.mma.cta_group::1, fp32 accumulate in TMEM), warp 2 = TMEM alloc/dealloc,
// warps 4..7 = epilogue (tcgen05.ld 32 lanes x 16 columns -> fused bias/residual/activation ->
// global).  Pipelines: smem full/empty mbarrier ring between producer and MMA issuer, one
// TMEM-full mbarrier between MMA issuer and epilogue.  Up to two CTAs are resident per SM (each
// owns <= 256 TMEM columns) so one CTA's epilogue overlaps the other's main loop.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gnv {

struct ConvTcParams {
  ConvGeom g;
  EpiParams ep;
  int block_n;      // UMMA N (multiple of 16, <= 256)
  int stages;       // smem ring depth
  int n_chunks;     // K blocks per tap = C_in_ld / (128 / sizeof(E))
  int tmem_cols;    // power of two >= max(32, block_n)
  uint32_t idesc;   // UMMA instruction descriptor
};

#ifdef __CUDACC__
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int KBLK_BYTES = 128;                        // one swizzle-128B row
constexpr int A_TILE_BYTES = BLOCK_M * KBLK_BYTES;     // 16 KiB
constexpr uint32_t kSpinLimit = 1u << 22;

// Host-mapped debug words (set by tc_debug_init): a barrier wait that times out records
// {0xDEAD0000 | role, blockIdx.x, barrier smem address, parity} before trapping, so the host can
// say WHICH wait hung even though the context is lost.
static __device__ uint32_t* g_tc_debug = nullptr;   // one copy per translation unit; each TU's init sets its own

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) {
      if (g_tc_debug) {
        g_tc_debug[1] = blockIdx.x; g_tc_debug[2] = bar; g_tc_debug[3] = parity;
        g_tc_debug[0] = 0xDEAD0000u | tag;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows are 128 B apart, 8-row groups 1024 B
// apart (SBO), LBO unused for swizzled K-major (encoded 1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
template <typename E>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (sizeof(E) == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc

// E = __nv_bfloat16 (kind::f16) or float (kind::tf32).
template <typename E>
__global__ void __launch_bounds__(256, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_tile_bytes = p.block_n * KBLK_BYTES;
  const int stage_bytes = A_TILE_BYTES + b_tile_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;    // 8-byte aligned (multiple of 1024)
  // barriers: full[stages], empty[stages], tmem_full; then the TMEM base-address slot
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * p.stages;
  const uint32_t bar_tmem_full = bar_base + 16u * p.stages;
  const uint32_t tmem_slot = bar_tmem_full + 8u;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * p.block_n;
  const int b = blockIdx.z;
  const int total_k = p.g.n_taps * p.n_chunks;
  constexpr int KBE = KBLK_BYTES / (int)sizeof(E);    // elements per K block

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < total_k; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        const int tap = it / p.n_chunks, ch = it - tap * p.n_chunks;
        const uint32_t sa = smem_base + s * stage_bytes, sb = sa + A_TILE_BYTES;
        mbar_expect_tx(bar_full + 8u * s, (uint32_t)stage_bytes);
        tma_load_3d(&tmA, bar_full + 8u * s, sa, ch * KBE, m0 + p.g.off0 + tap * p.g.tap_step, b);
        tma_load_2d(&tmW, bar_full + 8u * s, sb, it * KBE, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (one thread issues for the CTA; the same thread commits) =====
      for (int it = 0; it < total_k; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(bar_full + 8u * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * stage_bytes, sb = sa + A_TILE_BYTES;
        const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < KBLK_BYTES / 32; ++k) {
          // advance 32 bytes (one UMMA_K) inside the 128-byte swizzle row: +2 in 16-byte units
          umma<E>(tmem_acc, ad + 2u * k, bd + 2u * k, p.idesc, (it | k) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8u * s);          // frees the smem slot when those MMAs retire
      }
      umma_commit(bar_tmem_full);                 // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> fused bias / residual / activation -> global =====
    mbar_wait(bar_tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int m = m0 + q * 32 + lane;
    const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c0, v);          // warp-collective: all lanes participate
      if (m < p.g.M_rows) epi_apply<16, E>(p.ep, b, m, n0 + c0, v);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

#endif  // __CUDACC__

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();   // resolved through cudaGetDriverEntryPoint (no libcuda link)
// Lazy module loading (the CUDA 12 default) defers a kernel's load to its FIRST LAUNCH: 0.1-0.5 s for the large persistent
// kernels, in the middle of whichever request first needs that variant (measured: tools/t2w_plan_cost.py).  The *_init()
// functions call this for every kernel they configure, so a handle's kernels are resident when its create call returns.
// cuFuncLoad through cudaGetDriverEntryPoint; a no-op on drivers without it.
void preload_kernel(const void* kernel);

struct ConvTcLaunch {
  CUtensorMap tmA, tmW;
  ConvTcParams p;
  dim3 grid;
  size_t smem_bytes;
  int elem_bytes;
};

// Fills tensor maps + launch geometry for one layer.  `act` = A tensor [B, L_in, C_in_ld] (E),
// `w` = packed weights [N_rows_alloc, n_taps*C_in_ld] (E).  Returns "" or an error text.
const char* make_conv_tc_launch(ConvTcLaunch* out, int elem_bytes, const void* act, const void* w, int w_rows_alloc,
                                const ConvGeom& g, const EpiParams& ep, int block_n);
cudaError_t launch_conv_tc(const ConvTcLaunch& L, cudaStream_t st);
cudaError_t conv_tc_init();           // sets the max-dynamic-smem attribute once per process/device
const char* tc_debug_string();        // "" or a description of the barrier wait that timed out
cudaError_t tc_debug_device_ptr(uint32_t** out);   // device address of the host-mapped debug words

}  // namespace gnv
