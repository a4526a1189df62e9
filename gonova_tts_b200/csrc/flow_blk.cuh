// The transformer block of the CFM flow estimator on tcgen05, bf16 operands, CTA pairs.
//
// One kernel template, three modes, all over the flat row space M = 2 B T of the residual stream R [M, 256] fp32:
//
//   FB_FF   : R' = R + W2 gelu(W1 x + b1) + b2;   N' = LayerNorm(R') gamma + beta   (or bf16(R') when ln = 0)
//             the WHOLE feed-forward in one launch: x tile [128, 256] resident in shared memory, the 1024 hidden
//             channels in eight chunks of 128:  G1(c): acc1[c & 1] = x W1_c^T  ->  E1(c): gelu(acc1 + b1_c) -> bf16 ->
//             shared-memory operand H[c & 1]  ->  G2(c): acc2 += H W2_c^T.   The hidden activation [M, 1024] never
//             exists in HBM (it was written and read back: 131 MB per block), and there is one launch instead of three.
//   FB_OUT  : R' = R + A W^T + b  (the attention output projection, K = 512 streamed), same epilogue
//   FB_WIDE : Y = A W^T (+ b) as bf16, N in tiles of 256 (the fused q / k / v projection, N = 1536)
//   FB_OUTFF: FB_OUT and FB_FF of one transformer block chained in ONE launch: the out-projection's rows (x = R + O W^T + b)
//             stay in TMEM as the accumulator the feed-forward adds onto, LayerNorm(x) goes straight into shared memory
//             as the feed-forward's A operand; R is read once and written once per block, N never touches HBM between them
//   FB_CONV : Y = Mish(LayerNorm(causal 3-tap conv(A) + b)) (+ time bias), the two convs of a ResNet block with the
//             LayerNorm / Mish kernel that followed each folded into the epilogue; tiles per utterance (the taps are
//             row-shifted TMA loads of a [C, T, 2B] map: rows before an utterance's first frame read as zero)
//
// The residual + LayerNorm epilogue (E2) runs on the same sixteen warps once a tile's last MMA has retired: a warp owns 32
// rows x 64 columns, x = acc + b2 + R goes back into TMEM with the thread's mean / M2, the four warps of a lane quarter
// combine them through shared memory, and the second pass writes LayerNorm(x).
//
// CTA pair (cluster of 2, tcgen05.mma.cta_group::2, M = 256): each CTA owns 128 rows and stages HALF of every weight
// tile; the leader's MMA warp issues for both.  TMEM (512 columns): acc1 x 2 at 0 / 128, acc2 at 256 (OUT / WIDE: two
// 256-column buffers).  Shared memory: x 64 KB (FF / WIDE: resident A tile; OUT: A ring) | H 64 KB (WIDE: output staging) |
// weight ring sw x 16 KB.  640 threads: 16 epilogue warps, TMEM allocator, barrier init, TMA producer, MMA issuer.
#pragma once
#include <type_traits>

#include "conv_tc2.cuh"

namespace gnv {

enum { FB_FF = 0, FB_OUT = 1, FB_WIDE = 2, FB_CONV = 3, FB_OUTFF = 4 };
constexpr int kFbSlot = 16384;     // one K block (64 bf16) of 128 rows: A tile of a CTA / weight ring slot
constexpr int kFbMaxSw = 6;
constexpr int kFbTab = 1536 + 3 * 256 + 16 * 32 * 2 + 256;     // floats: b1 (| b3 | g3) | b2 | gamma | beta | LayerNorm exchange | be3

struct FlowBlkParams {
  int M, T, tiles;          // rows, rows per utterance, pair tiles of 256 rows
  int kb_a;                 // OUT: K blocks streamed (WIDE / FF: 4, resident)
  int n_tiles, npu;         // WIDE: N / 256; n tiles per work unit (the unit's A tile stays resident)
  int n_bias1;              // entries of b1 (FF: 1024, WIDE: N or 0)
  int sw;                   // weight ring depth
  int ln;                   // FF / OUT: 1 LayerNorm, 0 plain cast
  int dbg;                  // GONOVA_FB_DBG: 8 = record CTA 0's timeline (tools/flow_blk_trace.py)
  const int* lengths;       // [M / T] valid rows per utterance, or NULL
  const float *b1, *b2, *gamma, *beta;
  const float *b3, *g3, *be3;   // OUTFF: the out-projection's bias and the LayerNorm between it and the feed-forward
  int qkv;                      // OUTFF: 1 = the NEXT block's q/k/v projection runs on the tile's LayerNorm output before it leaves
  int noff;                     // OUTFF: 1 = no feed-forward: projection + residual + LayerNorm, then the q/k/v tail (a ResNet
                                // block's res_conv in front of a level's first transformer block)
  float* r;                 // FF / OUT: the residual stream [M, 256] fp32 (output); CONV with out_f32: the fp32 output
  const float* r_in;        // FF / OUT: the residual input (== r when updated in place)
  int conv_nch, conv_tpb;   // CONV: K blocks per tap (C_in / 64), pair tiles per utterance
  int out_f32;              // CONV: 1 = fp32 output rows at r, 0 = bf16 at n_out
  __nv_bfloat16* n_out;     // FF / OUT: bf16 output rows, pitch n_pitch elements
  int n_pitch;
  // WIDE: columns at or beyond vt_col0 (the V third of q | k | v) are written TRANSPOSED into vt [M / T * 8, 64, vt_tp]
  // (keys contiguous: the K-major B operand of P V in flow_attn_tc_kernel) instead of the row-major output
  __nv_bfloat16* vt;
  int vt_col0, vt_tp;
  uint32_t idesc128, idesc256;
  uint32_t off_x, off_h, off_w, off_sc, off_tab, off_bar;   // off_sc: the row warps' scratch (4 x 4 KB)
};

struct FlowBlkMaps { CUtensorMap A, W1, W2, Nout, W3, W4; };

constexpr int kFbThreads = 640;
// warps 0..15 epilogue warps (TMEM lane quarter w % 4, column block w / 4): the GELU epilogue of FF, the residual +
// LayerNorm epilogue of FF / OUT, the output of WIDE; the single-lane control warps last (highest ids: favoured by the
// warp arbiter)
constexpr int kFbWarpTmem = 16, kFbWarpInit = 17, kFbWarpProducer = 18, kFbWarpMma = 19;

#ifdef __CUDACC__
namespace tc2 {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// Arrive on the barrier at the same offset in the pair's LEADER CTA with the default (.release.cta) semantics.  The
// `.release.cluster` form (mbar_arrive_cluster) compiles to MEMBAR.ALL.CTA + ERRBAR and cost 2.6 k cycles per use on the
// timeline of the q/k/v GEMM; what these arrivals publish is either a drained TMEM buffer (ordered by
// tcgen05.fence::before_thread_sync) or shared memory of the arriving CTA itself that its own tensor core will read
// (ordered by fence.proxy.async): nothing another CTA's threads read through generic loads.
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// tcgen05.ld of 16 columns in two halves: issue now, consume after tmem_ld_wait16 (which names the registers as in-out
// operands, so that nothing reads them before the wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}

// 256-bit global accesses (sm_100): a thread moves one whole 32-byte sector.  The row warps own one row each, so a 128-bit
// store is half a sector per thread — every one a partial-sector write that L2 has to merge (measured: the residual +
// LayerNorm epilogue took 29 k cycles per tile with 128-bit accesses).
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void stg256_bf16x16(__nv_bfloat16* p, const float (&v)[16]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(ElemIO<__nv_bfloat16>::pack2(v[0], v[1])), "r"(ElemIO<__nv_bfloat16>::pack2(v[2], v[3])),
                 "r"(ElemIO<__nv_bfloat16>::pack2(v[4], v[5])), "r"(ElemIO<__nv_bfloat16>::pack2(v[6], v[7])),
                 "r"(ElemIO<__nv_bfloat16>::pack2(v[8], v[9])), "r"(ElemIO<__nv_bfloat16>::pack2(v[10], v[11])),
                 "r"(ElemIO<__nv_bfloat16>::pack2(v[12], v[13])), "r"(ElemIO<__nv_bfloat16>::pack2(v[14], v[15]))
               : "memory");
}

// GELU (erf form): gelu(x) = relu(x) - |x| Phi(-|x|), Phi(-a) = 2^g(a) with g a degree-5 fit of log2(erfc(a / sqrt 2) / 2) on
// [0, 4 sqrt 2] (max |error| of the GELU 8e-7 in fp32; beyond the clamp Phi(-|x|) < 4e-9).  Nine instructions, one of them
// MUFU.EX2: the feed-forward epilogue handles 16 K elements per 2 K MMA cycles, every instruction is on the critical path.
__device__ __forceinline__ float gelu_poly(float x) {
  const float a = fminf(fabsf(x), 5.656854249f);
  float g = fmaf(-4.991072352e-04f, a, 7.270108298e-03f);
  g = fmaf(g, a, -5.230520455e-02f);
  g = fmaf(g, a, -4.594566783e-01f);
  g = fmaf(g, a, -1.151037651e+00f);
  g = fmaf(g, a, -1.000002375e+00f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(g));
  return fmaf(-a, h, fmaxf(x, 0.f));
}
// Mish = x tanh(softplus x) = x n / (n + 2) with n = e^x (e^x + 2): one MUFU.EX2 and one MUFU.RCP; x > 20 returns x like torch
__device__ __forceinline__ float mish_fast(float x) {
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  return x > 20.f ? x : x * __fdividef(n, n + 2.f);
}
}  // namespace tc2

// Timeline of CTA 0 for tuning (GONOVA_FB_DBG = 8): (role + 1) << 56 | a << 48 | b << 40 | event << 32 | clock32
constexpr int kFbTraceCap = 4 * 2048;
static __device__ unsigned long long g_fb_trace[kFbTraceCap];
__device__ __forceinline__ void fb_trace(int on, int role, int a, int b, int ev, unsigned int& idx) {
  if (!on) return;
  unsigned int c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  if (idx < 2048u)
    g_fb_trace[role * 2048 + idx] = ((unsigned long long)(role + 1) << 56) | ((unsigned long long)(a & 255) << 48) |
                                    ((unsigned long long)(b & 255) << 40) | ((unsigned long long)ev << 32) | c;
  ++idx;
}

template <int MODE>
__global__ void __launch_bounds__(kFbThreads, 1)
flow_blk_kernel(const FlowBlkMaps* __restrict__ maps_g, const __grid_constant__ FlowBlkParams p) {
  using namespace tc2;
  typedef __nv_bfloat16 E;
  const FlowBlkMaps& maps = *maps_g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base + p.off_x, sH = smem_base + p.off_h, sW = smem_base + p.off_w;
  float* tab = reinterpret_cast<float*>(smem_gen + p.off_tab);
  float* tb1 = tab;
  float* tb2 = tab + 1536;
  float* tg = tb2 + 256;
  float* tbt = tg + 256;
  const uint32_t bar0 = smem_base + p.off_bar;
  const uint32_t b_a_full = bar0, b_a_empty = bar0 + 32u;
  const uint32_t b_w_full = bar0 + 64u, b_w_empty = b_w_full + 8u * kFbMaxSw;
  const uint32_t b_acc1_full = b_w_empty + 8u * kFbMaxSw;
  const uint32_t b_e1_done = b_acc1_full + 16u, b_h_empty = b_e1_done + 16u;
  const uint32_t b_acc2_full = b_h_empty + 16u, b_acc2_free = b_acc2_full + 16u;
  const uint32_t b_x_ready = b_acc2_free + 16u, b_x_free = b_x_ready + 8u;   // OUTFF
  const uint32_t b_q_full = b_x_free + 8u, b_q_free = b_q_full + 16u;          // OUTFF with the q/k/v tail: two 256-column buffers
  const uint32_t tmem_slot = b_q_free + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int pair0 = (int)(blockIdx.x >> 1), G = (int)(gridDim.x >> 1);
  const int n_groups = MODE == FB_WIDE ? p.n_tiles / p.npu : 1;      // work units per m tile
  const int n_units = p.tiles * n_groups;
  const int tr = ((p.dbg & 8) && blockIdx.x == 0 && lane == 0) ? 1 : 0;
  unsigned int tri = 0;

  if (warp == kFbWarpProducer && lane == 0) {
    prefetch_tmap(&maps.A);
    prefetch_tmap(&maps.W1);
    prefetch_tmap(&maps.W2);
    if constexpr (MODE == FB_OUTFF) { prefetch_tmap(&maps.W3); prefetch_tmap(&maps.W4); }
  }
  if (warp == kFbWarpInit && lane == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(b_a_full + 8u * s, 1); mbar_init(b_a_empty + 8u * s, 1); }
    for (int s = 0; s < kFbMaxSw; ++s) { mbar_init(b_w_full + 8u * s, 1); mbar_init(b_w_empty + 8u * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_acc1_full + 8u * s, 1);
      mbar_init(b_e1_done + 8u * s, 32);                               // 16 column warps of both CTAs arrive on the leader
      mbar_init(b_h_empty + 8u * s, 1);
      mbar_init(b_acc2_full + 8u * s, 1);
      mbar_init(b_acc2_free + 8u * s, 32);
    }
    mbar_init(b_x_ready, 32);
    mbar_init(b_x_free, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(b_q_full + 8u * s, 1); mbar_init(b_q_free + 8u * s, 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFbWarpTmem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int c = threadIdx.x; c < (MODE == FB_OUTFF ? 1024 : 1536); c += blockDim.x) tb1[c] = (p.b1 && c < p.n_bias1) ? p.b1[c] : 0.f;
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    tb2[c] = p.b2 ? p.b2[c] : 0.f;
    tg[c] = p.gamma ? p.gamma[c] : 1.f;
    tbt[c] = p.beta ? p.beta[c] : 0.f;
    if constexpr (MODE == FB_OUTFF) {
      tb1[1024 + c] = p.b3 ? p.b3[c] : 0.f;
      tb1[1280 + c] = p.g3 ? p.g3[c] : 1.f;
      tab[1536 + 768 + 1024 + c] = p.be3 ? p.be3[c] : 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait_then_release();

  if (warp == kFbWarpProducer) {
    // ===== TMA producer (both CTAs; loads complete on the LEADER's barriers) =====
    Ring ra, rw;
    auto load_a = [&](int kb, int row0) {
      mbar_wait(b_a_empty + 8u * ra.slot, ra.phase ^ 1u, 1);
      if (elect_one()) {
        if (crank == 0) mbar_expect_tx(b_a_full + 8u * ra.slot, 2u * kFbSlot);
        tma_load_2d_2sm(&maps.A, (b_a_full + 8u * ra.slot) & kPeerBitMask, sX + ra.slot * kFbSlot, kb * 64, row0);
      }
      __syncwarp();
      ra.advance(4);
    };
    auto w_slot = [&]() -> uint32_t {                   // waits for the ring slot and arms its barrier
      mbar_wait(b_w_empty + 8u * rw.slot, rw.phase ^ 1u, 1);
      if (elect_one() && crank == 0) mbar_expect_tx(b_w_full + 8u * rw.slot, 2u * kFbSlot);
      __syncwarp();
      return (uint32_t)rw.slot;
    };
    if constexpr (MODE == FB_FF || MODE == FB_OUTFF) {
      auto load_w1 = [&](int c) {                       // W1 rows [128 c, 128 c + 128): this CTA's 64, K blocks 2 s, 2 s + 1
        for (int s = 0; s < 2; ++s) {
          const uint32_t sl = w_slot();
          if (elect_one()) {
            for (int j = 0; j < 2; ++j)
              tma_load_2d_2sm(&maps.W1, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot + j * (kFbSlot / 2),
                              (2 * s + j) * 64, 128 * c + 64 * crank);
          }
          __syncwarp();
          rw.advance(p.sw);
        }
      };
      auto load_w2 = [&](int c) {                       // W2[:, 128 c + 64 s ...): this CTA's 128 of the 256 rows
        for (int s = 0; s < 2; ++s) {
          const uint32_t sl = w_slot();
          if (elect_one())
            tma_load_2d_2sm(&maps.W2, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, 128 * c + 64 * s, 128 * crank);
          __syncwarp();
          rw.advance(p.sw);
        }
      };
      int itp = 0;
      for (int t = pair0; t < p.tiles; t += G, ++itp) {
        const int row0 = t * 256 + crank * 128;
        fb_trace(tr, 2, t, 0, 0, tri);
        if constexpr (MODE == FB_OUTFF) {
          // the x region is the out-projection's A ring now and the feed-forward's operand afterwards: the next tile's
          // attention-output blocks may land only when this CTA's last G1 of the previous tile has read it
          if (itp > 0) mbar_wait(b_x_free, (uint32_t)((itp - 1) & 1), 1);
          for (int kb = 0; kb < p.kb_a; ++kb) {
            load_a(kb, row0);
            const uint32_t sl = w_slot();
            if (elect_one())
              tma_load_2d_2sm(&maps.W3, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, kb * 64, 128 * crank);
            __syncwarp();
            rw.advance(p.sw);
          }
        } else {
          for (int kb = 0; kb < 4; ++kb) load_a(kb, row0);
        }
        if (!(MODE == FB_OUTFF && p.noff)) {
          load_w1(0);
          load_w1(1);
          for (int c = 0; c < 8; ++c) {
            load_w2(c);
            if (c + 2 < 8) load_w1(c + 2);
            fb_trace(tr, 2, t, c, 3, tri);
          }
        }
        if constexpr (MODE == FB_OUTFF) {
          if (p.qkv) {
            for (int n = 0; n < 6; ++n)
              for (int kb = 0; kb < 4; ++kb) {
                const uint32_t sl = w_slot();
                if (elect_one())
                  tma_load_2d_2sm(&maps.W4, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, kb * 64, n * 256 + 128 * crank);
                __syncwarp();
                rw.advance(p.sw);
              }
          }
        }
      }
    } else if constexpr (MODE == FB_CONV) {
      for (int t = pair0; t < p.tiles; t += G) {
        const int b = t / p.conv_tpb, t0 = (t - b * p.conv_tpb) * 256 + crank * 128;
        for (int kb = 0; kb < p.kb_a; ++kb) {
          const int tap = kb / p.conv_nch, ch = kb - tap * p.conv_nch;
          mbar_wait(b_a_empty + 8u * ra.slot, ra.phase ^ 1u, 1);
          if (elect_one()) {
            if (crank == 0) mbar_expect_tx(b_a_full + 8u * ra.slot, 2u * kFbSlot);
            // causal: tap j reads frame t - 2 + j; frames before 0 (and past T) arrive as zeros
            tma_load_3d_2sm(&maps.A, (b_a_full + 8u * ra.slot) & kPeerBitMask, sX + ra.slot * kFbSlot, ch * 64, t0 - 2 + tap, b);
          }
          __syncwarp();
          ra.advance(4);
          const uint32_t sl = w_slot();
          if (elect_one())
            tma_load_2d_2sm(&maps.W2, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, kb * 64, 128 * crank);
          __syncwarp();
          rw.advance(p.sw);
        }
      }
    } else if constexpr (MODE == FB_OUT) {
      for (int t = pair0; t < p.tiles; t += G) {
        const int row0 = t * 256 + crank * 128;
        for (int kb = 0; kb < p.kb_a; ++kb) {
          load_a(kb, row0);
          const uint32_t sl = w_slot();
          if (elect_one())
            tma_load_2d_2sm(&maps.W2, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, kb * 64, 128 * crank);
          __syncwarp();
          rw.advance(p.sw);
          fb_trace(tr, 2, t, kb, 2, tri);
        }
      }
    } else {
      for (int u = pair0; u < n_units; u += G) {
        const int m = u / n_groups, n0 = (u - m * n_groups) * p.npu;
        const int row0 = m * 256 + crank * 128;
        for (int kb = 0; kb < 4; ++kb) load_a(kb, row0);
        for (int n = n0; n < n0 + p.npu; ++n) {
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t sl = w_slot();
            if (elect_one())
              tma_load_2d_2sm(&maps.W1, (b_w_full + 8u * sl) & kPeerBitMask, sW + sl * kFbSlot, kb * 64, n * 256 + 128 * crank);
            __syncwarp();
            rw.advance(p.sw);
          }
          fb_trace(tr, 2, u, n, 2, tri);
        }
      }
    }
  } else if (warp == kFbWarpMma) {
    if (crank == 0) {
      // ===== MMA issuer (leader; whole warp walks the loops, one elected lane issues) =====
      const uint64_t x_desc0 = umma_desc_sw128(sX), w_desc0 = umma_desc_sw128(sW), h_desc0 = umma_desc_sw128(sH);
      constexpr uint64_t kSlotUnits = kFbSlot >> 4;
      Ring rw;
      if constexpr (MODE == FB_FF || MODE == FB_OUTFF) {
        int it = 0;
        Ring ra;
        for (int t = pair0; t < p.tiles; t += G, ++it) {
          auto g1 = [&](int c, bool first) {
            const uint32_t acc = tmem_base + (uint32_t)(((it * 8 + c) & 1) * 128);
            for (int s = 0; s < 2; ++s) {
              if (first) {
                mbar_wait(b_a_full + 8u * (2 * s), (uint32_t)(it & 1), 2);
                mbar_wait(b_a_full + 8u * (2 * s + 1), (uint32_t)(it & 1), 2);
              }
              mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint64_t ad = x_desc0 + (uint64_t)(2 * s + j) * kSlotUnits;
                  const uint64_t bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits + (uint64_t)j * (kSlotUnits / 2);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_2sm<E>(acc, ad + 2u * kk, bd + 2u * kk, p.idesc128, (s | j | kk) ? 1u : 0u);
                }
                umma_commit_2sm(b_w_empty + 8u * rw.slot);
              }
              __syncwarp();
              rw.advance(p.sw);
            }
            if (elect_one()) umma_commit_2sm(b_acc1_full + 8u * ((it * 8 + c) & 1));
            __syncwarp();
          };
          auto g2 = [&](int c) {
            const uint32_t acc = tmem_base + 256u;
            const uint32_t buf = (uint32_t)((it * 8 + c) & 1);
            for (int s = 0; s < 2; ++s) {
              mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint64_t ad = h_desc0 + (uint64_t)(buf * 2 + s) * kSlotUnits;
                const uint64_t bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_2sm<E>(acc, ad + 2u * kk, bd + 2u * kk, p.idesc256, (MODE == FB_OUTFF || (c | s | kk)) ? 1u : 0u);
                umma_commit_2sm(b_w_empty + 8u * rw.slot);
              }
              __syncwarp();
              rw.advance(p.sw);
            }
            if (elect_one()) umma_commit_2sm(b_h_empty + 8u * buf);
            __syncwarp();
          };
          fb_trace(tr, 1, t, 0, 0, tri);
          if constexpr (MODE == FB_OUTFF) {
            // ---- the out-projection: acc2 = O W3^T (K streamed through the x ring) ----
            mbar_wait(b_acc2_free, (uint32_t)(it & 1) ^ 1u, 2);                     // the previous tile's rows have left acc2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < p.kb_a; ++kb) {
              mbar_wait(b_a_full + 8u * ra.slot, ra.phase, 2);
              mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint64_t ad = x_desc0 + (uint64_t)ra.slot * kSlotUnits, bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_2sm<E>(tmem_base + 256u, ad + 2u * kk, bd + 2u * kk, p.idesc256, (kb | kk) ? 1u : 0u);
                umma_commit_2sm(b_a_empty + 8u * ra.slot);
                umma_commit_2sm(b_w_empty + 8u * rw.slot);
              }
              __syncwarp();
              ra.advance(4);
              rw.advance(p.sw);
            }
            if (elect_one()) umma_commit_2sm(b_acc2_full);                          // (phase 0 of the tile: the projection is done)
            __syncwarp();
            // x is back in acc2, LayerNorm(x) in the x region (two x_ready / acc2_full events per tile with feed-forward AND tail)
            mbar_wait(b_x_ready, (p.qkv && !p.noff) ? 0u : (uint32_t)(it & 1), 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const bool ff = !(MODE == FB_OUTFF && p.noff);
          if (ff) { g1(0, MODE == FB_FF); g1(1, false); }
          for (int c = 0; ff && c < 8; ++c) {
            const int nn = it * 8 + c;
            mbar_wait(b_e1_done + 8u * (nn & 1), (uint32_t)((nn >> 1) & 1), 2);     // H[nn & 1] valid, acc1[nn & 1] free
            if (MODE == FB_FF && c == 0) mbar_wait(b_acc2_free, (uint32_t)(it & 1) ^ 1u, 2);   // the previous tile's rows have left acc2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            fb_trace(tr, 1, t, c, 3, tri);
            g2(c);
            fb_trace(tr, 1, t, c, 4, tri);
            if (c + 2 < 8) g1(c + 2, false);
            fb_trace(tr, 1, t, c, 5, tri);
            if (c + 2 == 7) {                                                       // the tile's last read of x
              if (elect_one()) {
                if constexpr (MODE == FB_OUTFF) { if (!p.qkv) umma_commit_2sm(b_x_free); }
                else for (int s = 0; s < 4; ++s) umma_commit_2sm(b_a_empty + 8u * s);
              }
              __syncwarp();
            }
          }
          if (ff) {
            if (elect_one()) umma_commit_2sm(b_acc2_full);
            __syncwarp();
          }
          if constexpr (MODE == FB_OUTFF) {
            if (p.qkv) {
              // ---- the next block's q/k/v projection on the LayerNorm output the second epilogue left in the x region ----
              if (ff) {
                mbar_wait(b_x_ready, 1u, 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              for (int n = 0; n < 6; ++n) {
                const uint32_t buf = (uint32_t)(n & 1);
                const uint32_t idx = (uint32_t)(3 * it + (n >> 1));
                mbar_wait(b_q_free + 8u * buf, (idx & 1u) ^ 1u, 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < 4; ++kb) {
                  mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
                  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                  if (elect_one()) {
                    const uint64_t ad = x_desc0 + (uint64_t)kb * kSlotUnits, bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                      umma_2sm<E>(tmem_base + buf * 256u, ad + 2u * kk, bd + 2u * kk, p.idesc256, (kb | kk) ? 1u : 0u);
                    umma_commit_2sm(b_w_empty + 8u * rw.slot);
                  }
                  __syncwarp();
                  rw.advance(p.sw);
                }
                if (elect_one()) umma_commit_2sm(b_q_full + 8u * buf);
                __syncwarp();
              }
              if (elect_one()) umma_commit_2sm(b_x_free);
              __syncwarp();
            }
          }
        }
      } else if constexpr (MODE == FB_OUT || MODE == FB_CONV) {
        Ring ra;
        int it = 0;
        for (int t = pair0; t < p.tiles; t += G, ++it) {
          const uint32_t buf = (uint32_t)(it & 1);
          const uint32_t acc = tmem_base + buf * 256u;
          mbar_wait(b_acc2_free + 8u * buf, (uint32_t)((it >> 1) & 1) ^ 1u, 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          fb_trace(tr, 1, t, 0, 0, tri);
          for (int kb = 0; kb < p.kb_a; ++kb) {
            mbar_wait(b_a_full + 8u * ra.slot, ra.phase, 2);
            mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const uint64_t ad = x_desc0 + (uint64_t)ra.slot * kSlotUnits, bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_2sm<E>(acc, ad + 2u * kk, bd + 2u * kk, p.idesc256, (kb | kk) ? 1u : 0u);
              umma_commit_2sm(b_a_empty + 8u * ra.slot);
              umma_commit_2sm(b_w_empty + 8u * rw.slot);
            }
            __syncwarp();
            ra.advance(4);
            rw.advance(p.sw);
          }
          if (elect_one()) umma_commit_2sm(b_acc2_full + 8u * buf);
          __syncwarp();
          fb_trace(tr, 1, t, 0, 5, tri);
        }
      } else {
        int it = 0, cnt = 0;
        for (int u = pair0; u < n_units; u += G, ++it) {
          for (int s = 0; s < 4; ++s) mbar_wait(b_a_full + 8u * s, (uint32_t)(it & 1), 2);
          for (int j = 0; j < p.npu; ++j, ++cnt) {
            const uint32_t buf = (uint32_t)(cnt & 1);
            const uint32_t acc = tmem_base + buf * 256u;
            mbar_wait(b_acc2_free + 8u * buf, (uint32_t)((cnt >> 1) & 1) ^ 1u, 2);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            fb_trace(tr, 1, u, j, 0, tri);
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint64_t ad = x_desc0 + (uint64_t)kb * kSlotUnits, bd = w_desc0 + (uint64_t)rw.slot * kSlotUnits;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_2sm<E>(acc, ad + 2u * kk, bd + 2u * kk, p.idesc256, (kb | kk) ? 1u : 0u);
                umma_commit_2sm(b_w_empty + 8u * rw.slot);
              }
              __syncwarp();
              rw.advance(p.sw);
            }
            if (elect_one()) umma_commit_2sm(b_acc2_full + 8u * buf);
            __syncwarp();
            fb_trace(tr, 1, u, j, 5, tri);
          }
          if (elect_one()) {
            for (int s = 0; s < 4; ++s) umma_commit_2sm(b_a_empty + 8u * s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 16) {
    // ===== the sixteen epilogue warps: TMEM lane quarter q = warp % 4 (32 rows), column block cb = warp / 4 =====
    const int q = warp & 3, cb = warp >> 2;
    const int erow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int tre = tr && warp == 0;

    // ---- E2 (FF / OUT): x = acc + b2 + R  ->  R' (fp32, in place)  and  N' = LayerNorm(x) gamma + beta  (or bf16(x)) ----
    // A warp owns 32 rows x 64 columns (four 16-column steps).  TMEM gives a thread one ROW; global memory wants a warp
    // instruction to cover whole 64-byte row segments: every step passes through a 2 KB scratch of the warp (in the H region,
    // idle between a tile's last G2 and the next tile's first E1): the R segment arrives there by cp.async (lane -> row
    // 8 i + lane / 4, 16-byte piece lane % 4), the thread reads its row, writes x back in place, and the warp stores the
    // scratch with the same coalesced mapping.  LayerNorm: each thread's mean / M2 over its 64 columns, combined across the
    // four warps of the lane quarter through shared memory (Chan's formula), x parked in TMEM for the second pass.
    // (History: four dedicated "row" warps with a whole row per thread ran this at one warp per scheduler — every
    // instruction's latency exposed, 25 k cycles per tile; these sixteen warps do it in ~5 k.)
    const uint32_t sc0 = sH + (uint32_t)warp * 4096u;
    const uint32_t my_row = (uint32_t)lane * 64u, my_sw = ((uint32_t)lane >> 1) & 3u;
    uint32_t co_off[4];                                 // coalesced mapping, fp32: instruction i -> row 8 i + lane / 4
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t r = 8u * i + ((uint32_t)lane >> 2), pc = (uint32_t)lane & 3u;
      co_off[i] = r * 64u + ((pc ^ ((r >> 1) & 3u)) << 4);
    }
    uint32_t cbo[2];                                    // coalesced mapping, bf16 (rows of 32 B): i -> row 16 i + lane / 2
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t r = 16u * i + ((uint32_t)lane >> 1), pc = (uint32_t)lane & 1u;
      cbo[i] = r * 64u + ((pc ^ ((r >> 1) & 3u)) << 4);
    }
    float* red = tab + 1536 + 768;                      // [16 warps][32 lanes][2]: mean, M2 of a thread's 64 columns
    // resin: x includes the residual rows (cp.async through the scratch); store_r: x is written to r (R'); to_x: the bf16
    // output goes into the x region in the operand layout (OUTFF's first epilogue) instead of global memory; do_ln:
    // LayerNorm with (gtab, betab) or the plain cast; btab: the GEMM's bias; free_bar = 0: no arrival (OUTFF's first epilogue)
    auto e2_tile = [&](int t, uint32_t acc_col, uint32_t full_bar, uint32_t full_par, uint32_t free_bar, bool resin, bool store_r,
                       bool to_x, bool do_ln, bool early_loads, const float* btab, const float* gtab, const float* betab) {
      const int wrow0 = t * 256 + crank * 128 + q * 32;   // first row of this warp
      const int row = wrow0 + lane;
      bool live = row < p.M;
      if (live && p.lengths) { const int b = row / p.T; live = row - b * p.T < p.lengths[b]; }
      const uint32_t acc = lane_base + acc_col + (uint32_t)(cb * 64);
      const float* rsrc[4];
      bool rok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = wrow0 + 8 * i + (lane >> 2);
        rok[i] = r < p.M;
        rsrc[i] = p.r_in + (size_t)(rok[i] ? r : 0) * 256 + cb * 64 + (lane & 3) * 4;
      }
      __nv_bfloat16* ndst[2];
      bool nok[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = wrow0 + 16 * i + (lane >> 1);
        nok[i] = r < p.M;
        ndst[i] = p.n_out + (size_t)(nok[i] ? r : 0) * p.n_pitch + cb * 64 + (lane & 1) * 8;
      }
      auto issue_load = [&](int st) {
        const uint32_t sc = sc0 + (uint32_t)(st & 1) * 2048u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nb = rok[i] ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sc + co_off[i]), "l"(rsrc[i] + st * 16), "r"(nb) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      if (resin && early_loads) {                          // (the H region is idle: the R segments load under the MMAs)
        issue_load(0);
        issue_load(1);
      }
      mbar_wait(full_bar, full_par, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      fb_trace(tre, 3, t, 0, 4, tri);
      if (resin && !early_loads) {
        issue_load(0);
        issue_load(1);
      }
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      uint32_t tv0[16], tv1[16];
      tmem_ld16_issue(acc, tv0);
      const ptrdiff_t rdelta = p.r - p.r_in;               // R' goes to r (== r_in when the stream is updated in place)
      // (instruction count matters here: sixteen warps run this pass issue-bound.  Everything that moves by a constant per
      // step is a pointer bumped once per loop trip plus a compile-time offset, and the row-in-range predicates exist
      // only in the instantiation for a tensor's last, partial tile.)
      const bool full = wrow0 + 32 <= p.M;                 // warp-uniform
      auto pass1 = [&](auto FULLC) {
        constexpr bool FULL = decltype(FULLC)::value;
        const float* rp[4] = {rsrc[0], rsrc[1], rsrc[2], rsrc[3]};
        __nv_bfloat16* np[2] = {ndst[0], ndst[1]};
        const float* btp = btab + cb * 64;
        uint32_t accp = acc;
        auto step1 = [&](auto ODD, bool more, uint32_t (&tcur)[16], uint32_t (&tnext)[16]) {
          constexpr int odd = decltype(ODD)::value;
          const uint32_t sc = sc0 + (uint32_t)odd * 2048u;
          if (resin) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
          }
          tmem_ld_wait16(tcur);
          if (odd == 0 || more) tmem_ld16_issue(accp + (uint32_t)(odd * 16 + 16), tnext);
          float v[16];
          const float4* bt = reinterpret_cast<const float4*>(btp + odd * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = bt[j];
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (resin) r4 = lds128(sc + my_row + (((uint32_t)j ^ my_sw) << 4));
            v[4 * j] = __uint_as_float(tcur[4 * j]) + b4.x + r4.x;
            v[4 * j + 1] = __uint_as_float(tcur[4 * j + 1]) + b4.y + r4.y;
            v[4 * j + 2] = __uint_as_float(tcur[4 * j + 2]) + b4.z + r4.z;
            v[4 * j + 3] = __uint_as_float(tcur[4 * j + 3]) + b4.w + r4.w;
          }
          if (!live) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          if (store_r) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(sc + my_row + (((uint32_t)j ^ my_sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 o = lds128(sc + co_off[i]);
              if ((FULL || rok[i]) && !(p.dbg & 16)) *reinterpret_cast<float4*>(const_cast<float*>(rp[i]) + rdelta + odd * 16) = o;
            }
          }
          if (do_ln) {
            if (odd == 0 && more) shift = v[0];           // (first step: sums relative to a value of the row itself)
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float d = v[j] - shift; s1 += d; s2 = fmaf(d, d, s2); }
            tmem_st16(accp + (uint32_t)(odd * 16), v);
          } else {
            __syncwarp();                                 // the fp32 rows have been read: the scratch takes the bf16 rows
            sts128u(sc + my_row + ((0u ^ my_sw) << 4), ElemIO<E>::pack2(v[0], v[1]), ElemIO<E>::pack2(v[2], v[3]),
                    ElemIO<E>::pack2(v[4], v[5]), ElemIO<E>::pack2(v[6], v[7]));
            sts128u(sc + my_row + ((1u ^ my_sw) << 4), ElemIO<E>::pack2(v[8], v[9]), ElemIO<E>::pack2(v[10], v[11]),
                    ElemIO<E>::pack2(v[12], v[13]), ElemIO<E>::pack2(v[14], v[15]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4 o = lds128(sc + cbo[i]);
              if ((FULL || nok[i]) && !(p.dbg & 32)) *reinterpret_cast<float4*>(np[i] + odd * 16) = o;
            }
          }
          __syncwarp();                                   // this scratch buffer is free again
          if (resin) {
            if (more) {                                   // the segment two steps on, into the buffer just freed
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int nb = (FULL || rok[i]) ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sc + co_off[i]), "l"(rp[i] + odd * 16 + 32), "r"(nb) : "memory");
              }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty: keeps the wait count uniform)
          }
        };
#pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {
          const bool more = sp == 0;
          step1(std::integral_constant<int, 0>{}, more, tv0, tv1);
          step1(std::integral_constant<int, 1>{}, more, tv1, tv0);
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] += 32;
          np[0] += 32; np[1] += 32;
          btp += 32;
          accp += 32u;
        }
      };
      if (full) pass1(std::true_type{}); else pass1(std::false_type{});
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (do_ln) {
        tmem_wait_st();
        // this thread: 64 columns -> (mean, M2); the row's other three quarters sit in warps q + 4 k
        const float m1 = s1 * (1.f / 64.f);
        red[(warp * 32 + lane) * 2] = shift + m1;
        red[(warp * 32 + lane) * 2 + 1] = fmaxf(s2 - s1 * m1, 0.f);
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        float mean = 0.f, m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) mean += red[((q + 4 * k) * 32 + lane) * 2];
        mean *= 0.25f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = red[((q + 4 * k) * 32 + lane) * 2] - mean;
          m2 += red[((q + 4 * k) * 32 + lane) * 2 + 1] + 64.f * d * d;
        }
        const float rstd = rsqrtf(m2 * (1.f / 256.f) + 1e-5f);
        const float nmr = -mean * rstd;
        fb_trace(tre, 3, t, 0, 5, tri);
        tmem_ld16_issue(acc, tv0);
        auto pass2 = [&](auto FULLC) {
          constexpr bool FULL = decltype(FULLC)::value;
          __nv_bfloat16* np[2] = {ndst[0], ndst[1]};
          const float* gp = gtab + cb * 64;
          const float* bp = betab + cb * 64;
          uint32_t accp = acc;
          uint32_t xrow = sX + (uint32_t)cb * kFbSlot + (uint32_t)erow * 128u;   // to_x: this row of K block cb of the x operand
          uint32_t xch = 0;                                                      // 16-byte chunk of the step inside the row
          auto step2 = [&](auto ODD, bool more, uint32_t (&tcur)[16], uint32_t (&tnext)[16]) {
            constexpr int odd = decltype(ODD)::value;
            const uint32_t sc = sc0 + (uint32_t)odd * 2048u;
            tmem_ld_wait16(tcur);
            if (odd == 0 || more) tmem_ld16_issue(accp + (uint32_t)(odd * 16 + 16), tnext);
            float v[16];
            const float4* g4 = reinterpret_cast<const float4*>(gp + odd * 16);
            const float4* b4p = reinterpret_cast<const float4*>(bp + odd * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 gg = g4[j], bb = b4p[j];
              v[4 * j] = fmaf(fmaf(__uint_as_float(tcur[4 * j]), rstd, nmr), gg.x, bb.x);
              v[4 * j + 1] = fmaf(fmaf(__uint_as_float(tcur[4 * j + 1]), rstd, nmr), gg.y, bb.y);
              v[4 * j + 2] = fmaf(fmaf(__uint_as_float(tcur[4 * j + 2]), rstd, nmr), gg.z, bb.z);
              v[4 * j + 3] = fmaf(fmaf(__uint_as_float(tcur[4 * j + 3]), rstd, nmr), gg.w, bb.w);
            }
            if (!live) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (to_x) {
              const uint32_t c0 = xch + (uint32_t)odd * 2u;
              sts128u(xrow + ((c0 ^ ((uint32_t)erow & 7u)) << 4), ElemIO<E>::pack2(v[0], v[1]), ElemIO<E>::pack2(v[2], v[3]),
                      ElemIO<E>::pack2(v[4], v[5]), ElemIO<E>::pack2(v[6], v[7]));
              sts128u(xrow + (((c0 + 1u) ^ ((uint32_t)erow & 7u)) << 4), ElemIO<E>::pack2(v[8], v[9]), ElemIO<E>::pack2(v[10], v[11]),
                      ElemIO<E>::pack2(v[12], v[13]), ElemIO<E>::pack2(v[14], v[15]));
            } else {
              sts128u(sc + my_row + ((0u ^ my_sw) << 4), ElemIO<E>::pack2(v[0], v[1]), ElemIO<E>::pack2(v[2], v[3]),
                      ElemIO<E>::pack2(v[4], v[5]), ElemIO<E>::pack2(v[6], v[7]));
              sts128u(sc + my_row + ((1u ^ my_sw) << 4), ElemIO<E>::pack2(v[8], v[9]), ElemIO<E>::pack2(v[10], v[11]),
                      ElemIO<E>::pack2(v[12], v[13]), ElemIO<E>::pack2(v[14], v[15]));
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float4 o = lds128(sc + cbo[i]);
                if (FULL || nok[i]) *reinterpret_cast<float4*>(np[i] + odd * 16) = o;
              }
              // (the next step writes the OTHER scratch buffer; this one is rewritten two steps on, after two more __syncwarp)
            }
          };
#pragma unroll 1
          for (int sp = 0; sp < 2; ++sp) {
            const bool more = sp == 0;
            step2(std::integral_constant<int, 0>{}, more, tv0, tv1);
            step2(std::integral_constant<int, 1>{}, more, tv1, tv0);
            np[0] += 32; np[1] += 32;
            gp += 32; bp += 32;
            accp += 32u;
            xch += 4u;
          }
        };
        if (full) pass2(std::true_type{}); else pass2(std::false_type{});
        // the next tile rewrites the exchange words and the scratch: every warp of the quarter is done reading both
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      }
      if (to_x) fence_async_smem();                        // the x operand was written through the generic proxy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) mbar_arrive_leader(free_bar);
      fb_trace(tre, 3, t, 0, 6, tri);
    };

    // one 256-column tile of a plain bf16 output (q | k | v): this warp's 32 rows x 64 columns, staged and stored by TMA; the
    // V third leaves transposed (WIDE, and the q/k/v tail of OUTFF)
    auto wide_item = [&](int n, uint32_t acc_col, int row0, int row, bool live) {
      // both 32-column halves of this warp's block in flight behind one wait (a tcgen05.ld round trip is ~400 cycles)
      uint32_t tw[2][32];
      tmem_ld32_issue(lane_base + acc_col + (uint32_t)(cb * 64), tw[0]);
      tmem_ld32_issue(lane_base + acc_col + (uint32_t)(cb * 64 + 32), tw[1]);
      tmem_wait_ld();
      tmem_ld_pin32(tw[0]);
      tmem_ld_pin32(tw[1]);
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        const int col = cb * 64 + i2 * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(tw[i2][j]);
        if (MODE == FB_WIDE && p.n_bias1) {
          const float4* bt = reinterpret_cast<const float4*>(tb1 + n * 256 + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bt[j];
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        if (!live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (p.vt && n * 256 + col >= p.vt_col0) {
          // a warp = 32 consecutive rows = (mostly) 32 consecutive keys of one utterance: one 64-byte run per channel
          if (row < p.M) {
            const int b = row / p.T, t = row - b * p.T;
            const int c = n * 256 + col - p.vt_col0;              // channel of v[0] inside the V third: head c / 64
            __nv_bfloat16* dst = p.vt + ((size_t)(b * 8 + (c >> 6)) * 64 + (c & 63)) * p.vt_tp + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[(size_t)j * p.vt_tp] = __float2bfloat16_rn(v[j]);
          }
          continue;
        }
        const uint32_t sbuf = sH + (uint32_t)warp * 4096u + (uint32_t)i2 * 2048u;
        if (elect_one()) bulk_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(sbuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), ElemIO<E>::pack2(v[8 * j], v[8 * j + 1]),
                  ElemIO<E>::pack2(v[8 * j + 2], v[8 * j + 3]), ElemIO<E>::pack2(v[8 * j + 4], v[8 * j + 5]),
                  ElemIO<E>::pack2(v[8 * j + 6], v[8 * j + 7]));
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&maps.Nout, sbuf, n * 256 + col, row0 + q * 32);
          bulk_commit();
        }
      }
    };

    // ---- CONV epilogue: y = Mish(LayerNorm(acc + b)) (+ time bias) -> bf16 rows (the next conv's operand) or fp32 rows ----
    auto e2_conv = [&](int b, int t0w, uint32_t acc_col, uint32_t full_bar, uint32_t full_par, uint32_t free_bar) {
      const int nvalid = min(32, max(0, p.T - t0w));       // rows of this warp inside the utterance
      bool live = lane < nvalid;
      if (live && p.lengths) live = t0w + lane < p.lengths[b];
      const size_t grow0 = (size_t)b * p.T + t0w;
      const uint32_t acc = lane_base + acc_col + (uint32_t)(cb * 64);
      mbar_wait(full_bar, full_par, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      uint32_t tv0[16], tv1[16];
      tmem_ld16_issue(acc, tv0);
      auto c1 = [&](int st, uint32_t (&tcur)[16], uint32_t (&tnext)[16]) {
        tmem_ld_wait16(tcur);
        if (st + 1 < 4) tmem_ld16_issue(acc + (uint32_t)((st + 1) * 16), tnext);
        float v[16];
        const float4* bt = reinterpret_cast<const float4*>(tb2 + cb * 64 + st * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = bt[j];
          v[4 * j] = __uint_as_float(tcur[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(tcur[4 * j + 1]) + b4.y;
          v[4 * j + 2] = __uint_as_float(tcur[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(tcur[4 * j + 3]) + b4.w;
        }
        if (st == 0) shift = v[0];
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float d = v[j] - shift; s1 += d; s2 = fmaf(d, d, s2); }
        tmem_st16(acc + (uint32_t)(st * 16), v);
      };
#pragma unroll 1
      for (int sp = 0; sp < 2; ++sp) { c1(2 * sp, tv0, tv1); c1(2 * sp + 1, tv1, tv0); }
      tmem_wait_st();
      const float m1 = s1 * (1.f / 64.f);
      red[(warp * 32 + lane) * 2] = shift + m1;
      red[(warp * 32 + lane) * 2 + 1] = fmaxf(s2 - s1 * m1, 0.f);
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      float mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) mean += red[((q + 4 * k) * 32 + lane) * 2];
      mean *= 0.25f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = red[((q + 4 * k) * 32 + lane) * 2] - mean;
        m2 += red[((q + 4 * k) * 32 + lane) * 2 + 1] + 64.f * d * d;
      }
      const float rstd = rsqrtf(m2 * (1.f / 256.f) + 1e-5f);
      const float nmr = -mean * rstd;
      tmem_ld16_issue(acc, tv0);
      auto c2 = [&](int st, uint32_t (&tcur)[16], uint32_t (&tnext)[16]) {
        const uint32_t sc = sc0 + (uint32_t)(st & 1) * 2048u;
        tmem_ld_wait16(tcur);
        if (st + 1 < 4) tmem_ld16_issue(acc + (uint32_t)((st + 1) * 16), tnext);
        float v[16];
        const int c0 = cb * 64 + st * 16;
        const float4* g4 = reinterpret_cast<const float4*>(tg + c0);
        const float4* b4p = reinterpret_cast<const float4*>(tbt + c0);
        const float4* t4 = reinterpret_cast<const float4*>(tb1 + c0);     // the ResNet block's time bias (zeros when absent)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 gg = g4[j], bb = b4p[j], tt = t4[j];
          v[4 * j] = mish_fast(fmaf(fmaf(__uint_as_float(tcur[4 * j]), rstd, nmr), gg.x, bb.x)) + tt.x;
          v[4 * j + 1] = mish_fast(fmaf(fmaf(__uint_as_float(tcur[4 * j + 1]), rstd, nmr), gg.y, bb.y)) + tt.y;
          v[4 * j + 2] = mish_fast(fmaf(fmaf(__uint_as_float(tcur[4 * j + 2]), rstd, nmr), gg.z, bb.z)) + tt.z;
          v[4 * j + 3] = mish_fast(fmaf(fmaf(__uint_as_float(tcur[4 * j + 3]), rstd, nmr), gg.w, bb.w)) + tt.w;
        }
        if (!live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (p.out_f32) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(sc + my_row + (((uint32_t)j ^ my_sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2);
            const float4 o = lds128(sc + co_off[i]);
            if (r < nvalid) *reinterpret_cast<float4*>(p.r + (grow0 + r) * 256 + c0 + (lane & 3) * 4) = o;
          }
        } else {
          sts128u(sc + my_row + ((0u ^ my_sw) << 4), ElemIO<E>::pack2(v[0], v[1]), ElemIO<E>::pack2(v[2], v[3]),
                  ElemIO<E>::pack2(v[4], v[5]), ElemIO<E>::pack2(v[6], v[7]));
          sts128u(sc + my_row + ((1u ^ my_sw) << 4), ElemIO<E>::pack2(v[8], v[9]), ElemIO<E>::pack2(v[10], v[11]),
                  ElemIO<E>::pack2(v[12], v[13]), ElemIO<E>::pack2(v[14], v[15]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = 16 * i + (lane >> 1);
            const float4 o = lds128(sc + cbo[i]);
            if (r < nvalid) *reinterpret_cast<float4*>(p.n_out + (grow0 + r) * p.n_pitch + c0 + (lane & 1) * 8) = o;
          }
        }
      };
#pragma unroll 1
      for (int sp = 0; sp < 2; ++sp) { c2(2 * sp, tv0, tv1); c2(2 * sp + 1, tv1, tv0); }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");    // exchange words and scratch are free for the next tile
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) mbar_arrive_leader(free_bar);
    };

    if constexpr (MODE == FB_CONV) {
      int it = 0;
      for (int t = pair0; t < p.tiles; t += G, ++it) {
        const uint32_t buf = (uint32_t)(it & 1);
        const int b = t / p.conv_tpb, t0w = (t - b * p.conv_tpb) * 256 + crank * 128 + q * 32;
        e2_conv(b, t0w, buf * 256u, b_acc2_full + 8u * buf, (uint32_t)((it >> 1) & 1), b_acc2_free + 8u * buf);
      }
    } else if constexpr (MODE == FB_FF || MODE == FB_OUTFF) {
      // E1(c): gelu(acc1 + b1_c) -> bf16 -> H[c & 1] in the K-major SWIZZLE_128B operand layout; this warp: 32 of the 128 columns
      int it = 0;
      const uint32_t row_off = (uint32_t)(cb >> 1) * kFbSlot + (uint32_t)erow * 128u;
      for (int t = pair0; t < p.tiles; t += G, ++it) {
        if constexpr (MODE == FB_OUTFF) {
          // the out-projection's epilogue: x = acc + b3 + R stays in TMEM (the feed-forward accumulates onto it),
          // LayerNorm(x) becomes the feed-forward's A operand in the x region; nothing goes to global memory
          // (without a feed-forward the same epilogue also writes R': it is the block's only one)
          // (acc2_full completes twice per tile — projection, feed-forward — unless there is no feed-forward)
          e2_tile(t, 256u, b_acc2_full, p.noff ? (uint32_t)(it & 1) : 0u, b_x_ready, true, p.noff != 0, true, true, true,
                  tb1 + 1024, tb1 + 1280, tab + 1536 + 768 + 1024);
        }
        const bool ff = !(MODE == FB_OUTFF && p.noff);
        for (int c = 0; ff && c < 8; ++c) {
          const int nn = it * 8 + c;
          const uint32_t buf = (uint32_t)(nn & 1);
          mbar_wait(b_acc1_full + 8u * buf, (uint32_t)((nn >> 1) & 1), 4);
          if (nn >= 2) mbar_wait(b_h_empty + 8u * buf, (uint32_t)(((nn >> 1) - 1) & 1), 4);   // G2(nn - 2) has read H[buf]
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          fb_trace(tre, 0, t, c, 2, tri);
          float v[32];
          tmem_ld32(lane_base + buf * 128u + (uint32_t)(cb * 32), v);
          const float4* bt = reinterpret_cast<const float4*>(tb1 + c * 128 + cb * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bt[j];
            v[4 * j] = gelu_poly(v[4 * j] + b4.x);
            v[4 * j + 1] = gelu_poly(v[4 * j + 1] + b4.y);
            v[4 * j + 2] = gelu_poly(v[4 * j + 2] + b4.z);
            v[4 * j + 3] = gelu_poly(v[4 * j + 3] + b4.w);
          }
          const uint32_t rowa = sH + buf * 2u * kFbSlot + row_off;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128u(rowa + ((((uint32_t)((cb & 1) * 4 + j)) ^ ((uint32_t)erow & 7u)) << 4), ElemIO<E>::pack2(v[8 * j], v[8 * j + 1]),
                    ElemIO<E>::pack2(v[8 * j + 2], v[8 * j + 3]), ElemIO<E>::pack2(v[8 * j + 4], v[8 * j + 5]),
                    ElemIO<E>::pack2(v[8 * j + 6], v[8 * j + 7]));
          fence_async_smem();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) mbar_arrive_leader(b_e1_done + 8u * buf);
          fb_trace(tre, 0, t, c, 3, tri);
        }
        // the tile's rows are complete once the last G2 has retired (which also frees the H region for the scratch)
        if constexpr (MODE == FB_OUTFF) {
          if (ff)
            e2_tile(t, 256u, b_acc2_full, 1u, p.qkv ? b_x_ready : b_acc2_free, false, true, p.qkv != 0, p.ln != 0, false, tb2, tg, tbt);
        } else
          e2_tile(t, 256u, b_acc2_full, (uint32_t)(it & 1), b_acc2_free, true, true, false, p.ln != 0, false, tb2, tg, tbt);
        if constexpr (MODE == FB_OUTFF) {
          if (p.qkv) {
            const int row0 = t * 256 + crank * 128;
            const int row = row0 + erow;
            bool live = row < p.M;
            if (live && p.lengths) { const int b = row / p.T; live = row - b * p.T < p.lengths[b]; }
            asm volatile("bar.sync 1, 512;" ::: "memory");   // every warp is done with the E2 scratch: the H region stages the q/k/v stores
            for (int n = 0; n < 6; ++n) {
              const uint32_t buf = (uint32_t)(n & 1);
              const uint32_t idx = (uint32_t)(3 * it + (n >> 1));
              mbar_wait(b_q_full + 8u * buf, idx & 1u, 4);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              wide_item(n, buf * 256u, row0, row, live);
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (elect_one()) mbar_arrive_leader(b_q_free + 8u * buf);
            }
            if (elect_one()) bulk_wait_read<0>();            // the staged boxes have left the H region
            __syncwarp();
            if (elect_one()) mbar_arrive_leader(b_acc2_free);
          }
        }
        // every warp's generic-proxy writes to the H region are done before the next tile's E1 hands it to the tensor core
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
    } else if constexpr (MODE == FB_OUT) {
      int it = 0;
      for (int t = pair0; t < p.tiles; t += G, ++it) {
        const uint32_t buf = (uint32_t)(it & 1);
        e2_tile(t, buf * 256u, b_acc2_full + 8u * buf, (uint32_t)((it >> 1) & 1), b_acc2_free + 8u * buf, true, true, false,
                p.ln != 0, true, tb2, tg, tbt);
      }
    } else {
      // plain output: this warp's 32 rows x 64 columns of every 256-column tile, staged and stored by TMA
      int cnt = 0;
      for (int u = pair0; u < n_units; u += G) {
        const int m = u / n_groups, n0 = (u - m * n_groups) * p.npu;
        const int row0 = m * 256 + crank * 128;
        const int row = row0 + erow;
        bool live = row < p.M;
        if (live && p.lengths) { const int b = row / p.T; live = row - b * p.T < p.lengths[b]; }
        for (int n = n0; n < n0 + p.npu; ++n, ++cnt) {
          const uint32_t buf = (uint32_t)(cnt & 1);
          mbar_wait(b_acc2_full + 8u * buf, (uint32_t)((cnt >> 1) & 1), 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          fb_trace(tre, 0, u, n, 4, tri);
          wide_item(n, buf * 256u, row0, row, live);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) mbar_arrive_leader(b_acc2_free + 8u * buf);
          fb_trace(tre, 0, u, n, 6, tri);
        }
      }
      if (elect_one()) bulk_wait_read<0>();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == kFbWarpTmem)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#endif  // __CUDACC__

struct FlowBlkLaunch {
  FlowBlkMaps maps;
  const FlowBlkMaps* d_maps = nullptr;
  FlowBlkParams p;
  int mode = 0, grid = 0;
  size_t smem_bytes = 0;
};

// a: A operand [M, K] bf16 (FF: K = 256; OUT: any multiple of 64; WIDE: 256 .. 512); w1 / w2: packed bf16 weights
// [N, K] (FF: w1 [1024, 256], w2 [256, 1024]; OUT: w2 [256, K]; WIDE: w1 [N, K], N a multiple of 256);
// r: residual stream [M, 256] fp32, updated in place (FF / OUT); n_out: bf16 output with row pitch n_pitch elements
// (FF / OUT: 256 columns at n_out; WIDE: N columns).  Returns "" or an error text.
const char* make_flow_blk_launch(FlowBlkLaunch* out, int mode, const void* a, int K, const void* w1, const float* b1,
                                 const void* w2, const float* b2, float* r, const float* gamma, const float* beta,
                                 int ln, void* n_out, int n_pitch, int N, int M, int T, int max_ctas,
                                 void* vt = nullptr, int vt_col0 = 0, int vt_tp = 0, const float* r_in = nullptr);
// the two causal 3-tap convs of a ResNet block with LayerNorm + Mish (+ the block's time bias, given per launch):
// a: [B2, T, C_in] bf16, w: packed [256, 3 * C_in] bf16 (tap-major K); output bf16 [B2 * T, 256] or fp32 [B2 * T, 256]
const char* make_flow_conv_launch(FlowBlkLaunch* out, const void* a, int C_in, const void* w, const float* bias,
                                  const float* gamma, const float* beta, void* out_bf16, float* out_f32, int B2, int T,
                                  int max_ctas);
// FB_OUTFF: o [M, K] bf16 (attention output), w3 [256, K] + b3 and LayerNorm (g3, be3), then the feed-forward as in FB_FF
// w4 != NULL: the NEXT block's q/k/v projection [1536, 256] runs on the LayerNorm output (which then never leaves the SM):
// qkv_out [M, 1536] bf16 (q | k), vt [M / T * 8, 64, vt_tp] (V transposed); n_out is not written in that case.
// w1 == NULL: no feed-forward (projection + residual r_in -> r + LayerNorm(g3, be3) + the q/k/v tail)
const char* make_flow_outff_launch(FlowBlkLaunch* out, const void* o, int K, const void* w3, const float* b3, const float* g3,
                                   const float* be3, const void* w1, const float* b1, const void* w2, const float* b2, float* r,
                                   const float* gamma, const float* beta, int ln, void* n_out, int n_pitch, int M, int T,
                                   int max_ctas, const void* w4 = nullptr, void* qkv_out = nullptr, void* vt = nullptr,
                                   int vt_tp = 0, const float* r_in = nullptr);
cudaError_t launch_flow_blk(const FlowBlkLaunch& L, const int* lengths, cudaStream_t st, const float* tbias = nullptr);
cudaError_t flow_blk_init();
int flow_blk_read_trace(unsigned long long* out, int cap);   // tuning: CTA 0's timeline of the last traced launch

}  // namespace gnv
