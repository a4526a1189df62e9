// Self-attention of the CFM flow estimator (8 heads x 64, full attention over the valid keys of an utterance) on tcgen05.
//
// Work item = (utterance b, head h, 256 queries): two 128-query tiles side by side, one per softmax warpgroup, sharing
// every K / V tile.  Per 128-key tile j and warpgroup g:
//     S_g = Q_g K_j^T            tcgen05.mma, M = 128, N = 128, K = 64: S in TMEM (128 columns)
//     softmax warpgroup g        thread = query row: row max and row sum are thread-local (no shuffles); the running
//                                output O_g (TMEM, 64 columns) is rescaled in place; P = 2^(s - m) goes to shared memory as
//                                bf16 in the K-major SWIZZLE_128B operand layout
//     O_g += P_g V_j             tcgen05.mma, M = 128, N = 64, K = 128
// While warpgroup 0 runs its softmax, the tensor core works for warpgroup 1 and vice versa (the two never share a
// barrier).  K comes straight from the fused q/k/v buffer [2B, T, 1536]; V is read from Vt [2B * 8, 64, Tp] — the q/k/v
// GEMM's epilogue writes its V columns TRANSPOSED (keys contiguous), which makes V^T a plain K-major B operand.
// Bound: one MUFU.EX2 per score (16 per clock per SM).  Replaces the mma.sync kernel (flow_kernels.cu) on the bf16 path.
// Warp roles (608 threads): 0..7 softmax warpgroup 0, 8..15 warpgroup 1 (two warps per 32 query rows: one per key half),
// 16 TMEM + barriers, 17 TMA producer, 18 MMA issuer.
#pragma once
#include "conv_tc2.cuh"

namespace gnv {

constexpr int kFaStages = 2;             // K / V ring depth (32 KB per stage)
constexpr int kFaThreads = 608;

struct FlowAttnParams {
  int B2, T, nqp, items;                 // utterances, frames, 256-query groups per (b, h), work items
  const int* lengths;
  __nv_bfloat16* o;                      // [B2, T, 512] (plain stores for the rows of an empty utterance)
  float sc2;                             // softmax scale * log2(e)
  int dbg;                               // GONOVA_FB_DBG = 8 with GONOVA_FB_TRACE_MODE = 3: CTA 0's timeline
  uint32_t idesc_s, idesc_pv;
  uint32_t off_q, off_kv, off_p, off_bar;
};

struct FlowAttnMaps { CUtensorMap QK, Vt, O; };

#ifdef __CUDACC__
namespace tc2 {
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace tc2

constexpr int kFaTraceCap = 4 * 2048;
static __device__ unsigned long long g_fa_trace[kFaTraceCap];
__device__ __forceinline__ void fa_trace(int on, int role, int a, int b, int ev, unsigned int& idx) {
  if (!on) return;
  unsigned int c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  if (idx < 2048u)
    g_fa_trace[role * 2048 + idx] = ((unsigned long long)(role + 1) << 56) | ((unsigned long long)(a & 255) << 48) |
                                    ((unsigned long long)(b & 255) << 40) | ((unsigned long long)ev << 32) | c;
  ++idx;
}

template <int kVariant>
__global__ void __launch_bounds__(kFaThreads, 1)
flow_attn_tc_kernel(const FlowAttnMaps* __restrict__ maps_g, const __grid_constant__ FlowAttnParams p) {
  using namespace tc2;
  typedef __nv_bfloat16 E;
  const FlowAttnMaps& maps = *maps_g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base + p.off_q, sKV = smem_base + p.off_kv, sP = smem_base + p.off_p;
  const uint32_t bar0 = smem_base + p.off_bar;
  const uint32_t b_q_full = bar0, b_q_empty = bar0 + 16u;     // [2] each: the Q tiles of the next item load under this one
  const uint32_t b_kv_full = bar0 + 32u, b_kv_empty = b_kv_full + 8u * kFaStages;
  const uint32_t b_s_full = b_kv_empty + 8u * kFaStages;       // [2]
  const uint32_t b_p_ready = b_s_full + 16u, b_pv_done = b_p_ready + 16u;
  const uint32_t tmem_slot = b_pv_done + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tr = ((p.dbg & 8) && blockIdx.x == 0 && lane == 0) ? 1 : 0;
  unsigned int tri = 0;

  if (warp == 17 && lane == 0) {
    prefetch_tmap(&maps.QK);
    prefetch_tmap(&maps.Vt);
  }
  if (warp == 16) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) { mbar_init(b_q_full + 8u * s, 1); mbar_init(b_q_empty + 8u * s, 1); }
      for (int s = 0; s < kFaStages; ++s) { mbar_init(b_kv_full + 8u * s, 1); mbar_init(b_kv_empty + 8u * s, 1); }
      for (int g = 0; g < 2; ++g) {
        mbar_init(b_s_full + 8u * g, 1);
        mbar_init(b_p_ready + 8u * g, 8);
        mbar_init(b_pv_done + 8u * g, 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait_then_release();

  // item -> (b, h, qp); n_kv = 128-key tiles of the utterance
  auto decode = [&](int item, int& b, int& h, int& qp, int& len) {
    qp = item % p.nqp;
    const int bh = item / p.nqp;
    h = bh & 7; b = bh >> 3;
    len = p.lengths ? min(p.T, max(0, p.lengths[b])) : p.T;
  };

  if (warp == 17) {
    // ===== TMA producer =====
    Ring rkv;
    uint32_t nq = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int b, h, qp, len;
      decode(item, b, h, qp, len);
      const int n_kv = (len + 127) >> 7;
      if (n_kv == 0) continue;
      const uint32_t qb = nq & 1u;
      mbar_wait(b_q_empty + 8u * qb, ((nq >> 1) & 1u) ^ 1u, 1);
      if (elect_one()) {
        mbar_expect_tx(b_q_full + 8u * qb, 2u * 16384u);
        tma_load_3d(&maps.QK, b_q_full + 8u * qb, sQ + qb * 32768u, h * 64, qp * 256, b);
        tma_load_3d(&maps.QK, b_q_full + 8u * qb, sQ + qb * 32768u + 16384u, h * 64, qp * 256 + 128, b);
      }
      __syncwarp();
      ++nq;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(b_kv_empty + 8u * rkv.slot, rkv.phase ^ 1u, 1);
        if (elect_one()) {
          const uint32_t dst = sKV + rkv.slot * 32768u;
          mbar_expect_tx(b_kv_full + 8u * rkv.slot, 32768u);
          tma_load_3d(&maps.QK, b_kv_full + 8u * rkv.slot, dst, 512 + h * 64, j * 128, b);
          tma_load_3d(&maps.Vt, b_kv_full + 8u * rkv.slot, dst + 16384u, j * 128, 0, b * 8 + h);
          tma_load_3d(&maps.Vt, b_kv_full + 8u * rkv.slot, dst + 16384u + 8192u, j * 128 + 64, 0, b * 8 + h);
        }
        __syncwarp();
        rkv.advance(kFaStages);
      }
    }
  } else if (warp == 18) {
    // ===== MMA issuer =====
    const uint64_t q_desc0 = umma_desc_sw128(sQ), kv_desc0 = umma_desc_sw128(sKV), p_desc0 = umma_desc_sw128(sP);
    Ring rkv;
    uint32_t np = 0;                                     // p_ready completions consumed so far (same for both warpgroups)
    uint32_t nq = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int b, h, qp, len;
      decode(item, b, h, qp, len);
      const int n_kv = (len + 127) >> 7;
      if (n_kv == 0) continue;
      const uint32_t qb = nq & 1u;
      auto issue_s = [&](int g, int slot) {
        if (elect_one()) {
          const uint64_t ad = q_desc0 + (uint64_t)(qb * 2u + (uint32_t)g) * (16384u >> 4), bd = kv_desc0 + (uint64_t)slot * (32768u >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma<E>(tmem_base + (uint32_t)g * 192u, ad + 2u * kk, bd + 2u * kk, p.idesc_s, kk ? 1u : 0u);
          umma_commit(b_s_full + 8u * g);
        }
        __syncwarp();
      };
      mbar_wait(b_q_full + 8u * qb, (nq >> 1) & 1u, 2);
      ++nq;
      mbar_wait(b_kv_full + 8u * rkv.slot, rkv.phase, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_s(0, rkv.slot);
      issue_s(1, rkv.slot);
      for (int j = 0; j < n_kv; ++j) {
        const int slot = rkv.slot;
        Ring nxt = rkv;
        nxt.advance(kFaStages);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(b_p_ready + 8u * g, np & 1u, 2);     // P_g(j) is in shared memory, O_g rescaled, S_g read
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          fa_trace(tr, 2, j, g, 1, tri);
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t ad = p_desc0 + (uint64_t)(g * 2 + kb) * (16384u >> 4);
              const uint64_t bd = kv_desc0 + (uint64_t)slot * (32768u >> 4) + (uint64_t)((16384u + kb * 8192u) >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma<E>(tmem_base + (uint32_t)g * 192u + 128u, ad + 2u * kk, bd + 2u * kk, p.idesc_pv, (j | kb | kk) ? 1u : 0u);
            }
            umma_commit(b_pv_done + 8u * g);
          }
          __syncwarp();
          if (j + 1 < n_kv) {
            if (g == 0) {
              mbar_wait(b_kv_full + 8u * nxt.slot, nxt.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            issue_s(g, nxt.slot);
          }
          fa_trace(tr, 2, j, g, 3, tri);
        }
        if (elect_one()) {
          umma_commit(b_kv_empty + 8u * slot);
          if (j + 1 == n_kv) umma_commit(b_q_empty + 8u * qb);
        }
        __syncwarp();
        rkv = nxt;
        ++np;
      }
    }
  } else if (warp < 16) {
    // ===== softmax warpgroups: warpgroup g = warp / 8; inside it lane quarter q = warp % 4 and key half hh = (warp / 4) % 2 =====
    // Two warps share each 32 query rows: warp hh takes keys [64 hh, 64 hh + 64) of a tile (= one K block of P) and 32 of O's 64
    // columns.  (With ONE warp per scheduler and warpgroup the exp section ran at 1 450 cycles against its 1 024-cycle MUFU
    // floor: nothing to issue while an EX2 result is on its way.)  The row maximum is exchanged through shared memory.
    const int g = warp >> 3, q = warp & 3, hh = (warp >> 2) & 1;
    const int erow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 192u;
    const uint32_t sPg = sP + (uint32_t)g * 32768u;
    const uint32_t p_row = sPg + (uint32_t)erow * 128u;
    float* xch = reinterpret_cast<float*>(smem_gen + p.off_bar + 1024);       // [16 warps][32 lanes]
    float* mine = xch + warp * 32 + lane;
    const float* theirs = xch + (warp ^ 4) * 32 + lane;                       // the warp with the other key half of my rows
    const int pair_bar = 4 + g * 4 + q;                                       // named barrier of the two warps (64 threads)
    uint32_t ns = 0;                                     // S tiles consumed so far by this warpgroup
    if (g == 1) asm volatile("bar.arrive %0, 512;" ::"r"(2) : "memory");   // warpgroup 0 goes first
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int b, h, qp, len;
      decode(item, b, h, qp, len);
      const int n_kv = (len + 127) >> 7;
      const int t_row = qp * 256 + g * 128 + erow;       // this thread's query
      if (n_kv == 0) {
        // an empty utterance: its rows of O are zero
        if (t_row < p.T) {
          uint4* dst = reinterpret_cast<uint4*>(p.o + ((size_t)b * p.T + t_row) * 512 + h * 64 + hh * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++ns) {
        mbar_wait(b_s_full + 8u * g, ns & 1u, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        fa_trace(tr && q == 0 && hh == 0, g, j, 0, 1, tri);
        const int kmax = len - j * 128 - hh * 64;        // keys of my half that exist (may be <= 0 on the last tile)
        uint32_t sr[2][32];
        tmem_ld32_issue(lane_base + (uint32_t)(hh * 64), sr[0]);
        tmem_ld32_issue(lane_base + (uint32_t)(hh * 64 + 32), sr[1]);
        tmem_wait_ld();
        tmem_ld_pin32(sr[0]);
        tmem_ld_pin32(sr[1]);
        if (kmax < 64) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) if (c * 32 + i >= kmax) sr[c][i] = 0xff800000u;    // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
        }
        *mine = fmaxf(mx0, mx1);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float mx = fmaxf(fmaxf(mx0, mx1), *theirs);        // finite: key 0 of every tile exists
        // Lazy rescale: the reference maximum moves only when some row of the warp would otherwise see P > 2^8 (bf16 / fp32
        // have the headroom); then the O tile in TMEM is left alone.  (Both warps of a pair see the same 32 maxima: same vote.)
        const float m_cand = fmaxf(m, mx * p.sc2);
        const bool move = (j == 0) || __any_sync(0xffffffffu, m_cand - m > 8.0f);
        const float m_new = move ? m_cand : m;
        const float corr = move ? ex2_fast(m - m_new) : 1.0f;   // 0 on the first tile (m = -inf)
        // The exponentials of the two warpgroups take turns (named barriers 2 + g): side by side they just share the MUFU
        // units IN PHASE and then both wait for their MMAs together.  Alternating, one warpgroup's TMEM loads, maximum,
        // P stores and MMAs run under the other's exponentials.
        fa_trace(tr && q == 0 && hh == 0, g, j, 0, 2, tri);
        asm volatile("bar.sync %0, 512;" ::"r"(2 + g) : "memory");
        fa_trace(tr && q == 0 && hh == 0, g, j, 0, 3, tri);
        // P = 2^(s sc2 - m_new) -> bf16 -> K block hh of the warpgroup's P operand (16-byte chunks XOR row & 7); masked: 2^-inf = 0
        float l0 = 0.f, l1 = 0.f;
        const uint32_t rowa = p_row + (uint32_t)hh * 16384u;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ex2_fast(fmaf(__uint_as_float(sr[c][i]), p.sc2, -m_new));
#pragma unroll
          for (int i = 0; i < 32; i += 2) { l0 += v[i]; l1 += v[i + 1]; }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            sts128u(rowa + ((((uint32_t)(c * 4 + k4)) ^ ((uint32_t)erow & 7u)) << 4), ElemIO<E>::pack2(v[8 * k4], v[8 * k4 + 1]),
                    ElemIO<E>::pack2(v[8 * k4 + 2], v[8 * k4 + 3]), ElemIO<E>::pack2(v[8 * k4 + 4], v[8 * k4 + 5]),
                    ElemIO<E>::pack2(v[8 * k4 + 6], v[8 * k4 + 7]));
        }
        l = fmaf(l, corr, l0 + l1);                      // (this warp's half of the row sum; the halves meet at the end)
        asm volatile("bar.arrive %0, 512;" ::"r"(3 - g) : "memory");      // the other warpgroup's turn
        fa_trace(tr && q == 0 && hh == 0, g, j, 0, 4, tri);
        if (j > 0 && move) {
          // the previous tile's P V has landed in O: rescale my 32 columns of it to the new maximum
          mbar_wait(b_pv_done + 8u * g, (ns - 1u) & 1u, 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          float v[32];
          tmem_ld32(lane_base + 128u + (uint32_t)(hh * 32), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= corr;
          tmem_st32(lane_base + 128u + (uint32_t)(hh * 32), v);
          tmem_wait_st();
        }
        m = m_new;
        fence_async_smem();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (elect_one()) mbar_arrive(b_p_ready + 8u * g);
        fa_trace(tr && q == 0 && hh == 0, g, j, 0, 5, tri);
      }
      // ---- output: O / l -> bf16 -> my 32 of the 64 columns of the pair's 32 x 64 box (staged in the P buffer) -> TMA store ----
      mbar_wait(b_pv_done + 8u * g, (ns - 1u) & 1u, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      *mine = l;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float lt = l + *theirs;
      const float inv = (t_row < len && lt > 0.f) ? 1.f / lt : 0.f;
      {
        float v[32];
        tmem_ld32(lane_base + 128u + (uint32_t)(hh * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= inv;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          sts128u(p_row + ((((uint32_t)(hh * 4 + k4)) ^ ((uint32_t)erow & 7u)) << 4), ElemIO<E>::pack2(v[8 * k4], v[8 * k4 + 1]),
                  ElemIO<E>::pack2(v[8 * k4 + 2], v[8 * k4 + 3]), ElemIO<E>::pack2(v[8 * k4 + 4], v[8 * k4 + 5]),
                  ElemIO<E>::pack2(v[8 * k4 + 6], v[8 * k4 + 7]));
      }
      fence_async_smem();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // both halves of the box are staged (and both have read the row sums)
      if (hh == 0 && elect_one()) {
        tma_store_3d(&maps.O, sPg + (uint32_t)q * 4096u, h * 64, qp * 256 + g * 128 + q * 32, b);
        bulk_commit();
        bulk_wait_read<0>();                              // the box has left shared memory before the next item's P lands there
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#endif  // __CUDACC__

struct FlowAttnLaunch {
  FlowAttnMaps maps;
  const FlowAttnMaps* d_maps = nullptr;
  FlowAttnParams p;
  int grid = 0;
  size_t smem_bytes = 0;
};

// qkv: [B2, T, 1536] bf16 (q | k | unused); vt: [B2 * 8, 64, Tp] bf16 (V transposed, Tp = T rounded up to 8);
// out: [B2, T, 512] bf16
const char* make_flow_attn_launch(FlowAttnLaunch* out, const void* qkv, const void* vt, int Tp, void* o, int B2, int T,
                                  float scale, int max_ctas);
cudaError_t launch_flow_attn_tc(const FlowAttnLaunch& L, const int* lengths, cudaStream_t st);
cudaError_t flow_attn_init();
int flow_attn_read_trace(unsigned long long* out, int cap);

}  // namespace gnv
