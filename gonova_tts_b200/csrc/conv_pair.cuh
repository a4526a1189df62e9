// Fused ResBlock pair for sm_100a:   y = conv2( act_mid( conv1(xt) + b1 ) ) + b2 + res
// (one dilation step of a HiFT ResBlock: conv1 k taps / dilation d, Snake, conv2 k taps / dilation 1,
// residual add) in ONE persistent tcgen05 kernel.  The intermediate activation never leaves the SM:
//
//   TMA x slab -> conv1 MMAs -> TMEM acc1 -> epilogue 1 (bias, Snake, bf16) -> shared-memory "h slab"
//   in the K-major SWIZZLE_128B operand layout -> conv2 MMAs (taps = descriptor row offsets into the
//   h slab) -> TMEM acc2 -> epilogue 2 (the shared epi_finish_item: residual, scale, running sum,
//   fp32 + activated outputs through TMA stores).
//
// Against two conv_tc2 launches this removes one full write + read of the stage tensor (the E3
// buffer), one launch, and the whole memory-side epilogue of conv1.  A tile produces
// Mo = 128*mh - (k-1) output rows: conv1 is evaluated on 128*mh rows (the Mo rows plus conv2's halo),
// so the halo recompute is (k-1)/(128*mh) of conv1 only (<= 4 % at mh = 2).
//
// Schedule (all roles walk the same static tile list t0, t1, ...):
//   MMA issuer : M1(0) M1(1);  then per tile i:  wait E1(i) done -> M2(i) -> M1(i+2)
//   epilogue   : per tile i:  E1(i) ; E2(i-1)          (E1(i) never waits: M1(i) ran during tile i-2 / i-1)
// TMEM: acc1 and acc2, two buffers each: 4*mh*C <= 512 columns (conv1 runs two tiles ahead of conv2).
// Used for C <= 128 (stages 1 and 2, where the convs are memory / epilogue bound); C = 256 keeps the
// two-launch path, which already runs at > 1 PFLOP/s.
#pragma once
#include "conv_tc2.cuh"

namespace gnv {

struct ConvPairParams {
  int B, L;                        // utterances, rows per utterance (input length == output length)
  int C;                           // channels in = channels out
  int k, d1;                       // taps; conv1 dilation (conv2 has dilation 1)
  int p1, p2;                      // conv1 / conv2 padding: d1*(k-1)/2, (k-1)/2
  int n_chunks;                    // 128-byte K blocks per tap
  int mh, Mo, tiles_m, total_tiles;
  int a_box_rows, a_n_boxes, slab_bytes;
  int h_kb_bytes;                  // one K block of the h slab: 128*mh rows x 128 B
  int w_bytes;                     // one weight tile: C rows x 128 B
  int w_group, w_slot_bytes;       // taps per weight barrier / ring slot
  int sa, sw, n_epi_wg, out_bufs, in_ring;
  int cta2;                        // CTA pairs: tiles_m / total_tiles then count PAIRS of CTA tiles
  int slabs_first;                 // producer order: the next conv1's x slab before the weight groups of the conv2 in between
  int dbg;                         // timing experiments only (GONOVA_PAIR_DBG): 1 no epilogue-1 math/stores, 2 no epilogue-2 finish
  int mma_order;                   // 0: alternate the two accumulators per MMA, 1: four k-steps per accumulator in a row
  uint32_t idesc;
  int n_in, has_raw, n_act, act_bytes, c_tab;
  int mid_kind;                    // activation between the two convs (ACT_SNAKE_FAST | ACT_SNAKE)
  const float* bias1;
  const float* alpha_mid;
  uint32_t off_a, off_w, off_h, off_in, off_out, off_tab, off_bar;
  EpiParams ep;                    // conv2's fused epilogue (bias = conv2 bias)
};

struct ConvPairMaps {
  CUtensorMap X, W1, W2;
  CUtensorMap epi[2][6];           // stores: [0] 32-row boxes, [1] (32 - (k-1))-row boxes (last warp of a tile's last half); loads: 128 rows
};

#ifdef __CUDACC__

// CTA2 = true: CTA pairs (cluster of two, tcgen05.mma.cta_group::2): the two CTAs work on ADJACENT tiles of the same
// utterance, each with its own x slab, h slab, accumulators and epilogues; they share every weight tile (each stages
// half of its rows), and the leader's MMA thread issues both CTAs' MMAs (M = 256 across the pair).
template <typename E, bool CTA2, bool RAGGED>
__global__ void __launch_bounds__(384, 1)
conv_pair_kernel(const ConvPairMaps* __restrict__ maps_g, const __grid_constant__ ConvPairParams p) {
  using namespace tc2;
  const ConvPairMaps& maps = *maps_g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base + p.off_a, sW = smem_base + p.off_w, sH = smem_base + p.off_h;
  const uint32_t sIn = smem_base + p.off_in, sOut = smem_base + p.off_out;
  float* tab = reinterpret_cast<float*>(smem_gen + p.off_tab);   // conv2 tables (EpiCtx layout), then b1, alpha_mid, inv_mid
  const int Cp = p.c_tab;
  float* tab1 = tab + (1 + 2 * p.n_act) * Cp;
  // barriers
  const uint32_t bar0 = smem_base + p.off_bar;
  const uint32_t b_a_full = bar0, b_a_empty = b_a_full + 8u * p.sa;
  const uint32_t b_w_full = b_a_empty + 8u * p.sa, b_w_empty = b_w_full + 8u * p.sw;
  const uint32_t b_acc1_full = b_w_empty + 8u * p.sw; // two buffers
  const uint32_t b_e1_done = b_acc1_full + 16u;       // epilogue 1 finished: h slab valid AND its acc1 buffer free
  const uint32_t b_h_empty = b_e1_done + 8u;          // conv2 MMAs retired: h slab may be overwritten
  const uint32_t b_acc2_full = b_h_empty + 8u, b_acc2_empty = b_acc2_full + 16u;
  const uint32_t b_in_full = b_acc2_empty + 16u, b_in_empty = b_in_full + 8u * kMaxInSlots;
  const uint32_t tmem_slot = b_in_empty + 8u * kMaxInSlots;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KBE = KBLK_BYTES / (int)sizeof(E);
  constexpr int kSub = CTA2 ? 2 : 1;                               // CTA tiles per scheduled tile
  const int crank = CTA2 ? (int)cluster_ctarank() : 0;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;

  if (warp == kWarpProducer && lane == 0) {
    prefetch_tmap(&maps.X);
    prefetch_tmap(&maps.W1);
    prefetch_tmap(&maps.W2);
  }
  if (warp == kWarpLoader && lane == 0) {
    for (int s = 0; s < p.sa; ++s) { mbar_init(b_a_full + 8u * s, 1); mbar_init(b_a_empty + 8u * s, 1); }
    for (int s = 0; s < p.sw; ++s) { mbar_init(b_w_full + 8u * s, 1); mbar_init(b_w_empty + 8u * s, 1); }
    mbar_init(b_acc1_full, 1);
    mbar_init(b_acc1_full + 8u, 1);
    mbar_init(b_e1_done, 4 * p.n_epi_wg * kSub);                 // pair: both CTAs' epilogues arrive on the leader
    mbar_init(b_h_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_acc2_full + 8u * s, 1);
      mbar_init(b_acc2_empty + 8u * s, 4 * p.n_epi_wg * kSub);
    }
    for (int s = 0; s < kMaxInSlots; ++s) {
      mbar_init(b_in_full + 8u * s, 1);
      mbar_init(b_in_empty + 8u * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpTmem) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  {
    const int C = p.C;
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      tab[c] = (c < C && p.ep.bias) ? p.ep.bias[c] : 0.f;
      for (int a = 0; a < p.n_act; ++a) {
        const float al = (c < C && p.ep.act_alpha[a]) ? p.ep.act_alpha[a][c] : 1.f;
        tab[(1 + 2 * a) * Cp + c] = al;
        tab[(2 + 2 * a) * Cp + c] = 1.0f / (al + 1e-9f);
      }
      const float am = (c < C && p.alpha_mid) ? p.alpha_mid[c] : 1.f;
      tab1[c] = (c < C && p.bias1) ? p.bias1[c] : 0.f;
      tab1[Cp + c] = am;
      tab1[2 * Cp + c] = 1.0f / (am + 1e-9f);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait_then_release();                     // the prologue above may overlap the previous kernel's tail
  // TMEM columns: acc1 buffers at 0 and mh*C, acc2 buffers at 2*mh*C and 3*mh*C
  const uint32_t acc2_col0 = (uint32_t)(2 * p.mh * p.C);

  const int n_epi_chunks = p.C / kEpiCols;
  const int G = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;    // tile stride of this CTA (pair)
  // Ragged batches (RAGGED instantiation, launched when `lengths` is given): tiles wholly kDeadMargin rows or more past
  // their utterance's valid length are skipped by every role (see conv_tc2_kernel); `next_tile` walks this CTA's live
  // tiles.  The dense instantiation compiles to the plain strided walk.
  auto tile_live = [&](int t) -> bool {
    if constexpr (!RAGGED) return true;
    const int b = t / p.tiles_m;
    const long valid = (long)p.ep.lengths[b] * p.ep.len_mul + p.ep.len_add;
    return (long)(t % p.tiles_m) * kSub * p.Mo < valid + kDeadMargin;
  };
  auto next_tile = [&](int t) -> int {
    if constexpr (!RAGGED) {
      return t + G;
    } else {
      do { t += G; } while (t < p.total_tiles && !tile_live(t));
      return t;
    }
  };
  int tile_first = tile0, tile_second = tile0;
  if constexpr (RAGGED) {
    if (tile0 < p.total_tiles && !tile_live(tile0)) tile_first = next_tile(tile0);
    tile_second = tile_first < p.total_tiles ? next_tile(tile_first) : tile_first;
  }

  if (warp == kWarpProducer) {
    {
      // ===== TMA producer: x slabs + W1 tile groups for M1, W2 tile groups for M2, in the issuer's order =====
      // (whole warp in the loops, one elected lane issues)
      Ring ra, rw;
      auto load_w_groups = [&](const CUtensorMap* wm, int ch) {
        for (int tap = 0; tap < p.k; tap += p.w_group) {
          const int ng = min(p.w_group, p.k - tap);
          mbar_wait(b_w_empty + 8u * rw.slot, rw.phase ^ 1u, 1);
          if (elect_one()) {
            if (crank == 0) mbar_expect_tx(b_w_full + 8u * rw.slot, (uint32_t)(ng * p.w_bytes) * kSub);
            for (int g = 0; g < ng; ++g) {
              if constexpr (CTA2)   // this CTA's half of the weight rows, completing on the leader's barrier
                tma_load_2d_2sm(wm, (b_w_full + 8u * rw.slot) & kPeerBitMask, sW + rw.slot * p.w_slot_bytes + g * p.w_bytes,
                                ((tap + g) * p.n_chunks + ch) * KBE, crank * (p.C >> 1));
              else
                tma_load_2d(wm, b_w_full + 8u * rw.slot, sW + rw.slot * p.w_slot_bytes + g * p.w_bytes,
                            ((tap + g) * p.n_chunks + ch) * KBE, 0);
            }
          }
          __syncwarp();
          rw.advance(p.sw);
        }
      };
      // x slab(s) of a tile and conv1's weight groups are separate rings: `slabs_first` issues the slab loads of the NEXT
      // conv1 before the weight groups of the conv2 in between (a slab comes from HBM, ~1.5 us; issued after conv2's
      // weights it arrived late and conv1 waited for it)
      auto load_slab = [&](int t, int ch) {
        const int m_tile = (t % p.tiles_m) * kSub + crank, b = t / p.tiles_m;
        const int r0 = m_tile * p.Mo - p.p2 - p.p1;                 // first x row of the slab
        mbar_wait(b_a_empty + 8u * ra.slot, ra.phase ^ 1u, 1);
        const uint32_t dst = sA + ra.slot * p.slab_bytes;
        if (elect_one()) {
          if (crank == 0)
            mbar_expect_tx(b_a_full + 8u * ra.slot, (uint32_t)(p.a_n_boxes * p.a_box_rows) * KBLK_BYTES * kSub);
          for (int bx = 0; bx < p.a_n_boxes; ++bx) {
            if constexpr (CTA2)
              tma_load_3d_2sm(&maps.X, (b_a_full + 8u * ra.slot) & kPeerBitMask,
                              dst + (uint32_t)(bx * p.a_box_rows) * KBLK_BYTES, ch * KBE, r0 + bx * p.a_box_rows, b);
            else
              tma_load_3d(&maps.X, b_a_full + 8u * ra.slot, dst + (uint32_t)(bx * p.a_box_rows) * KBLK_BYTES, ch * KBE,
                          r0 + bx * p.a_box_rows, b);
          }
        }
        __syncwarp();
        ra.advance(p.sa);
      };
      const bool slabs_first = p.slabs_first != 0 && p.n_chunks <= p.sa;
      auto load_m1 = [&](int t, bool slabs_done) {
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          if (!slabs_done) load_slab(t, ch);
          load_w_groups(&maps.W1, ch);
        }
      };
      auto load_m2 = [&]() {
        for (int ch = 0; ch < p.n_chunks; ++ch) load_w_groups(&maps.W2, ch);
      };
      // the issuer runs conv1 two tiles ahead of conv2: M1(0) M1(1) | M2(0) M1(2) | M2(1) M1(3) | ...
      // (the dense walk is kept in its plain strided form: ptxas keeps the issuer's descriptors in uniform registers
      // for it, and loses that — R2UR before every MMA, 16 % slower k = 11 pairs — with the rotating live-tile walk)
      if constexpr (RAGGED) {
        if (tile_first < p.total_tiles) load_m1(tile_first, false);
        if (tile_second < p.total_tiles) load_m1(tile_second, false);
        for (int t = tile_first, t1 = tile_second; t < p.total_tiles;) {
          const int t2 = t1 < p.total_tiles ? next_tile(t1) : t1;
          if (slabs_first && t2 < p.total_tiles)
            for (int ch = 0; ch < p.n_chunks; ++ch) load_slab(t2, ch);
          load_m2();
          if (t2 < p.total_tiles) load_m1(t2, slabs_first);
          t = t1; t1 = t2;
        }
      } else {
        if (tile0 < p.total_tiles) load_m1(tile0, false);
        if (tile0 + G < p.total_tiles) load_m1(tile0 + G, false);
        for (int t = tile0; t < p.total_tiles; t += G) {
          if (slabs_first && t + 2 * G < p.total_tiles)
            for (int ch = 0; ch < p.n_chunks; ++ch) load_slab(t + 2 * G, ch);
          load_m2();
          if (t + 2 * G < p.total_tiles) load_m1(t + 2 * G, slabs_first);
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (crank == 0) {
      // ===== MMA issuer (pair: the leader issues both CTAs' MMAs) =====
      // whole warp in the loops, one elected lane issues (see conv_tc2_kernel): uniform registers, UTCHMMAs back to back
      const uint64_t a_desc0 = umma_desc_sw128(sA), w_desc0 = umma_desc_sw128(sW), h_desc0 = umma_desc_sw128(sH);
      Ring ra, rw;
      // one K block (channel block ch) of a conv: every tap group's weights against row-shifted views of `a_base`.
      // The issue path is the critical path (tools/mma_issue_bench.cu: an N = 64 MMA retires every 48 cycles when the
      // issuer keeps up), so descriptors advance by loop-invariant strides — no multiplies, no parameter loads per tap.
      const uint32_t k_taps = (uint32_t)p.k, w_group = (uint32_t)p.w_group;
      const uint64_t w_slot_units = (uint64_t)((uint32_t)p.w_slot_bytes >> 4), w_units = (uint64_t)((uint32_t)p.w_bytes >> 4);
      const uint32_t idesc = p.idesc, colsC = (uint32_t)p.C;
      const bool alt = p.mma_order == 0 && p.mh == 2;
      const int mh = p.mh;
      auto issue_taps = [&](uint64_t a_base, int row_step, uint32_t acc0, uint32_t& accum) {
        const uint64_t a_step = (uint64_t)((uint32_t)row_step * (KBLK_BYTES >> 4));   // descriptor units per tap
        uint64_t ad_tap = a_base;
        for (uint32_t tap = 0; tap < k_taps; tap += w_group) {
          const uint32_t ng = min(w_group, k_taps - tap);
          mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint64_t bd_g0 = w_desc0 + (uint64_t)rw.slot * w_slot_units;
            if (alt) {
              // All descriptors of the group first, then up to 32 MMAs back to back.  The tensor queue is shallow: while
              // the issuing thread computes addresses between taps the pipe drains and idles (ncu, B = 64: operand pipe
              // busy 54 % of the time, the issuer blocked behind it for the same 54 %, address math and waits the rest).
              // k-step outer, half inner: consecutive MMAs alternate between the two accumulators, so an MMA never
              // waits for the previous one's accumulate into the same TMEM tile.
              uint64_t ads[4], bds[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                ads[g] = ad_tap + (uint64_t)g * a_step;
                bds[g] = bd_g0 + (uint64_t)g * w_units;
              }
              const uint32_t acc1 = acc0 + colsC;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if ((uint32_t)g < ng) {
                  const uint64_t ad0 = ads[g], ad1 = ads[g] + (uint64_t)(BLOCK_M * (KBLK_BYTES >> 4)), bd = bds[g];
#pragma unroll
                  for (int kk = 0; kk < KBLK_BYTES / 32; ++kk) {
                    const uint32_t ac = (g == 0 && kk == 0) ? accum : 1u;
                    if constexpr (CTA2) {
                      umma_2sm<E>(acc0, ad0 + 2u * kk, bd + 2u * kk, idesc, ac);
                      umma_2sm<E>(acc1, ad1 + 2u * kk, bd + 2u * kk, idesc, ac);
                    } else {
                      umma<E>(acc0, ad0 + 2u * kk, bd + 2u * kk, idesc, ac);
                      umma<E>(acc1, ad1 + 2u * kk, bd + 2u * kk, idesc, ac);
                    }
                  }
                }
              }
            } else {
              uint64_t bd = bd_g0;
              uint64_t ad0 = ad_tap;
              uint32_t ac0 = accum;
              for (uint32_t g = 0; g < ng; ++g) {
                for (int h = 0; h < mh; ++h) {
                  const uint64_t ad = ad0 + (uint64_t)(h * BLOCK_M * (KBLK_BYTES >> 4));
                  const uint32_t acc = acc0 + (uint32_t)h * colsC;
                  if constexpr (CTA2) {
                    umma_2sm<E>(acc, ad, bd, idesc, ac0);
#pragma unroll
                    for (int kk = 1; kk < KBLK_BYTES / 32; ++kk) umma_2sm<E>(acc, ad + 2u * kk, bd + 2u * kk, idesc, 1u);
                  } else {
                    umma<E>(acc, ad, bd, idesc, ac0);
#pragma unroll
                    for (int kk = 1; kk < KBLK_BYTES / 32; ++kk) umma<E>(acc, ad + 2u * kk, bd + 2u * kk, idesc, 1u);
                  }
                }
                ac0 = 1u;
                bd += w_units;
                ad0 += a_step;
              }
            }
            if constexpr (CTA2) umma_commit_2sm(b_w_empty + 8u * rw.slot); else umma_commit(b_w_empty + 8u * rw.slot);
          }
          __syncwarp();
          accum = 1u;
          ad_tap += (uint64_t)ng * a_step;
          rw.advance(p.sw);
        }
      };
      auto issue_m1 = [&](int j) {                        // conv1 of this CTA's j-th tile into acc1 buffer j & 1
        uint32_t accum = 0u;
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          mbar_wait(b_a_full + 8u * ra.slot, ra.phase, 2);
          issue_taps(a_desc0 + (uint64_t)((uint32_t)(ra.slot * p.slab_bytes) >> 4), p.d1,
                     tmem_base + (uint32_t)((j & 1) * p.mh * p.C), accum);
          if (elect_one()) {
            if constexpr (CTA2) umma_commit_2sm(b_a_empty + 8u * ra.slot); else umma_commit(b_a_empty + 8u * ra.slot);
          }
          __syncwarp();
          ra.advance(p.sa);
        }
        if (elect_one()) {
          if constexpr (CTA2) umma_commit_2sm(b_acc1_full + 8u * (j & 1)); else umma_commit(b_acc1_full + 8u * (j & 1));
        }
        __syncwarp();
      };
      auto issue_m2 = [&](int i) {
        const int buf = i & 1;
        mbar_wait(b_acc2_empty + 8u * buf, (uint32_t)((i >> 1) & 1) ^ 1u, 2);
        uint32_t accum = 0u;
        for (int ch = 0; ch < p.n_chunks; ++ch)
          issue_taps(h_desc0 + (uint64_t)((uint32_t)(ch * p.h_kb_bytes) >> 4), 1,
                     tmem_base + acc2_col0 + (uint32_t)(buf * p.mh * p.C), accum);
        if (elect_one()) {
          if constexpr (CTA2) umma_commit_2sm(b_acc2_full + 8u * buf); else umma_commit(b_acc2_full + 8u * buf);
          if constexpr (CTA2) umma_commit_2sm(b_h_empty); else umma_commit(b_h_empty);
        }
        __syncwarp();
      };
      int i = 0;
      if constexpr (RAGGED) {
        if (tile_first < p.total_tiles) issue_m1(0);
        if (tile_second < p.total_tiles) issue_m1(1);
        for (int t = tile_first, t1 = tile_second; t < p.total_tiles; ++i) {
          mbar_wait(b_e1_done, (uint32_t)(i & 1), 2);        // h slab of tile i is valid, acc1[i & 1] is free
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_m2(i);
          const int t2 = t1 < p.total_tiles ? next_tile(t1) : t1;
          if (t2 < p.total_tiles) issue_m1(i + 2);
          t = t1; t1 = t2;
        }
      } else {
        if (tile0 < p.total_tiles) issue_m1(0);
        if (tile0 + G < p.total_tiles) issue_m1(1);
        for (int t = tile0; t < p.total_tiles; t += G, ++i) {
          mbar_wait(b_e1_done, (uint32_t)(i & 1), 2);        // h slab of tile i is valid, acc1[i & 1] is free
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_m2(i);
          if (t + 2 * G < p.total_tiles) issue_m1(i + 2);
        }
      }
    }
  } else if (warp == kWarpLoader) {
    if (p.n_in > 0) {
      // ===== epilogue-2 input loader (same ring protocol as conv_tc2; whole warp, elected lane issues) =====
      int cnt[2] = {0, 0};
      const int n_items = p.mh * n_epi_chunks;
      for (int t = RAGGED ? tile_first : tile0; t < p.total_tiles; t = RAGGED ? next_tile(t) : t + G) {
        const int m_tile = (t % p.tiles_m) * kSub + crank, b = t / p.tiles_m;
        for (int item = 0; item < n_items; ++item) {
          const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
          const int mrow = m_tile * p.Mo + h * BLOCK_M;
          const int wgi = p.n_epi_wg == 2 ? (item & 1) : 0;
          const int k = cnt[wgi]++;
          const int slot = wgi * p.in_ring + k % p.in_ring;
          mbar_wait(b_in_empty + 8u * slot, (uint32_t)((k / p.in_ring) & 1) ^ 1u, 3);
          if (elect_one()) {
            mbar_expect_tx(b_in_full + 8u * slot, (uint32_t)p.n_in * (BLOCK_M * kEpiCols * 4));
            for (int j = 0; j < p.n_in; ++j)
              tma_load_3d(&maps.epi[0][EPI_IN0 + j], b_in_full + 8u * slot,
                          sIn + (slot * p.n_in + j) * (BLOCK_M * kEpiCols * 4), cc * kEpiCols, mrow, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8 && (warp >> 2) < p.n_epi_wg) {
    // ===== epilogue warpgroups: E1(i) then E2(i-1) =====
    const int wg = warp >> 2;
    const int q = warp & 3;
    const int erow = q * 32 + lane;
    const bool elected = (threadIdx.x & 127) == 0;
    const int out_stride = (p.has_raw ? BLOCK_M * kEpiCols * 4 : 0) + p.n_act * p.act_bytes;
    Ring rin;
    int ob = 0;
    const int n_items = p.mh * n_epi_chunks;
    EpiCtx ectx;
    ectx.ep = &p.ep; ectx.tab = tab; ectx.c_tab = Cp; ectx.n_in = p.n_in; ectx.has_raw = p.has_raw;
    ectx.n_act = p.n_act; ectx.act_bytes = p.act_bytes; ectx.n_epi_wg = p.n_epi_wg; ectx.out_bufs = p.out_bufs;
    ectx.smem_in = smem_gen + p.off_in; ectx.b_in_full = b_in_full; ectx.b_in_empty = b_in_empty; ectx.in_ring = p.in_ring;
    ectx.obase_wg = sOut + wg * p.out_bufs * out_stride; ectx.out_stride = out_stride;
    ectx.wg = wg; ectx.erow = erow; ectx.lane = lane; ectx.elected = elected;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);

    auto valid_rows_of = [&](int b) {
      int vr = p.L;
      if (p.ep.lengths) {
        const int lv = p.ep.lengths[b] * p.ep.len_mul + p.ep.len_add;
        vr = lv < vr ? lv : vr;
      }
      return vr;
    };
    // E1: acc1 -> bias1 -> Snake -> E -> h slab (K-major SWIZZLE_128B operand layout)
    auto epilogue1 = [&](int t, int i) {
      const int m_tile = (t % p.tiles_m) * kSub + crank, b = t / p.tiles_m;
      const int g0 = m_tile * p.Mo - p.p2;                          // global row of h-slab row 0
      const int vr = valid_rows_of(b);
      mbar_wait(b_acc1_full + 8u * (i & 1), (uint32_t)((i >> 1) & 1), 4);
      if (i > 0) mbar_wait(b_h_empty, (uint32_t)((i - 1) & 1), 4);  // conv2 of the previous tile has read the slab
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int item = (p.n_epi_wg == 2 ? wg : 0); item < n_items; item += p.n_epi_wg) {
        const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
        const int srow = h * BLOCK_M + erow;                        // h-slab row
        const int g = g0 + srow;
        const bool live1 = g >= 0 && g < vr;                        // conv2 sees zeros outside the utterance
        float v[32];
        tmem_ld32(lane_base + (uint32_t)(((i & 1) * p.mh + h) * p.C + cc * kEpiCols), v);
        const int c0 = cc * kEpiCols;
        const float4* bt = reinterpret_cast<const float4*>(tab1 + c0);
        const float4* al = reinterpret_cast<const float4*>(tab1 + Cp + c0);
        const float4* iv = reinterpret_cast<const float4*>(tab1 + 2 * Cp + c0);
        float y[32];
        if (p.dbg & 1) {
          if (v[0] != 12345.f) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = v[j];
        } else if (p.mid_kind == ACT_SNAKE_FAST) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bt[j], a4 = al[j], i4 = iv[j];
            float x, s;
            x = v[4 * j] + b4.x;     s = __sinf(x * a4.x); y[4 * j]     = fmaf(i4.x, s * s, x);
            x = v[4 * j + 1] + b4.y; s = __sinf(x * a4.y); y[4 * j + 1] = fmaf(i4.y, s * s, x);
            x = v[4 * j + 2] + b4.z; s = __sinf(x * a4.z); y[4 * j + 2] = fmaf(i4.z, s * s, x);
            x = v[4 * j + 3] + b4.w; s = __sinf(x * a4.w); y[4 * j + 3] = fmaf(i4.w, s * s, x);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bt[j], a4 = al[j], i4 = iv[j];
            y[4 * j]     = snake_precise(v[4 * j] + b4.x, a4.x, i4.x);
            y[4 * j + 1] = snake_precise(v[4 * j + 1] + b4.y, a4.y, i4.y);
            y[4 * j + 2] = snake_precise(v[4 * j + 2] + b4.z, a4.z, i4.z);
            y[4 * j + 3] = snake_precise(v[4 * j + 3] + b4.w, a4.w, i4.w);
          }
        }
        if (!live1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = 0.f;
        }
        // channels c0 .. c0+31 live in K block kb at 16-byte chunk cb .. of a 128-byte row
        const int kb = c0 / KBE;
        const int cb = (c0 - kb * KBE) * (int)sizeof(E) / 16;
        const uint32_t rowa = sH + (uint32_t)kb * p.h_kb_bytes + (uint32_t)srow * KBLK_BYTES;
        if constexpr (sizeof(E) == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128u(rowa + ((((uint32_t)(cb + j)) ^ ((uint32_t)srow & 7u)) << 4), ElemIO<E>::pack2(y[8 * j], y[8 * j + 1]),
                    ElemIO<E>::pack2(y[8 * j + 2], y[8 * j + 3]), ElemIO<E>::pack2(y[8 * j + 4], y[8 * j + 5]),
                    ElemIO<E>::pack2(y[8 * j + 6], y[8 * j + 7]));
        } else {
          if (p.ep.round_tf32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = round_tf32(y[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(rowa + ((((uint32_t)(cb + j)) ^ ((uint32_t)srow & 7u)) << 4), y[4 * j], y[4 * j + 1], y[4 * j + 2],
                   y[4 * j + 3]);
        }
      }
      // make the slab visible to the tensor core (async proxy), release acc1 and publish the slab
      fence_async_smem();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        if constexpr (CTA2) mbar_arrive_cluster(b_e1_done & kPeerBitMask); else mbar_arrive(b_e1_done);
      }
    };
    // E2: acc2 -> bias2 -> shared epilogue (residual, running sum, outputs)
    auto epilogue2 = [&](int t, int i) {
      const int m_tile = (t % p.tiles_m) * kSub + crank, b = t / p.tiles_m;
      const int m0 = m_tile * p.Mo;
      const int vr = valid_rows_of(b);
      const int buf = i & 1;
      mbar_wait(b_acc2_full + 8u * buf, (uint32_t)((i >> 1) & 1), 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int item = (p.n_epi_wg == 2 ? wg : 0); item < n_items; item += p.n_epi_wg) {
        const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
        const int r = h * BLOCK_M + erow;
        const int m = m0 + r;
        const bool live = r < p.Mo && m < vr;
        float v[32];
        tmem_ld32(lane_base + acc2_col0 + (uint32_t)((buf * p.mh + h) * p.C + cc * kEpiCols), v);
        const int c0 = cc * kEpiCols;
        const float4* bt = reinterpret_cast<const float4*>(tab + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = bt[j];
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
        // the tile's last half stores only its first 128 - (k-1) rows (the rest belongs to the next tile): every warp
        // stores its own 32 rows, the last warp of the last half a (32 - (k-1))-row box
        const CUtensorMap* m6 = &maps.epi[(h == p.mh - 1 && q == 3) ? 1 : 0][0];
        if (p.dbg & 2) {
          if (p.n_in > 0) {
            const int in_slot = wg * p.in_ring + rin.slot;
            mbar_wait(b_in_full + 8u * in_slot, rin.phase, 5);
            __syncwarp();
            if (elect_one()) mbar_arrive(b_in_empty + 8u * in_slot);
            rin.advance(p.in_ring);
          }
          continue;
        }
        epi_finish_item<E>(ectx, v, live, c0, m6, c0, m0 + h * BLOCK_M, b, rin, ob);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        if constexpr (CTA2) mbar_arrive_cluster((b_acc2_empty + 8u * buf) & kPeerBitMask);
        else mbar_arrive(b_acc2_empty + 8u * buf);
      }
    };
    int i = 0, prev = -1;
    for (int t = RAGGED ? tile_first : tile0; t < p.total_tiles; t = RAGGED ? next_tile(t) : t + G, ++i) {
      epilogue1(t, i);
      if (prev >= 0) epilogue2(prev, i - 1);
      prev = t;
    }
    if (prev >= 0) epilogue2(prev, i - 1);
    if (elect_one()) bulk_wait_read<0>();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();
  if (warp == kWarpTmem) {
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#endif  // __CUDACC__

struct ConvPairLaunch {
  ConvPairMaps maps;
  const ConvPairMaps* d_maps = nullptr;
  ConvPairParams p;
  int grid;
  size_t smem_bytes;
  int elem_bytes;
};

// x: [B, L, C_ld] (E), w1/w2: packed [C, k*C_ld] (E).  `ep` is conv2's epilogue (bias = conv2 bias,
// res / raw / act_out tensors are [B, L, C]).  Returns "" or the reason the pair cannot be fused.
const char* make_conv_pair_launch(ConvPairLaunch* out, int elem_bytes, const void* x, const void* w1, const void* w2,
                                  int B, int L, int C, int C_ld, int k, int d1, const float* bias1,
                                  const float* alpha_mid, int mid_kind, const EpiParams& ep, int max_ctas, int mh_opt,
                                  int cta2_opt);
cudaError_t launch_conv_pair(const ConvPairLaunch& L, const int* lengths, cudaStream_t st);
cudaError_t conv_pair_init();

}  // namespace gnv
