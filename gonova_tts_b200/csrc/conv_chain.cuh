// Whole-ResBlock fusion for sm_100a: the three dilation steps of a HiFT ResBlock
//     for d in (0, 1, 2):  x = x + conv2_d( snake2_d( conv1_d( snake1_d(x) ) ) )
// (upstream hifigan.ResBlock.forward; conv1_d has dilation dil[d], conv2_d dilation 1, both k taps) in ONE persistent
// tcgen05 kernel.  The residual stream never leaves the SM between the six convs:
//
//   F2 tile (fp32, TMA) --E0--> TMEM acc2 := x            (tcgen05.st)        slab := snake1_0(x)      (bf16 / tf32 operand)
//   per d:   M1: acc1  = conv1_d(slab)                    E1: slab := snake2_d(acc1 + b1_d)
//            M2: acc2 += conv2_d(slab)   <- the residual add IS the accumulate of the MMA into the resident x
//            E2: slab := snake1_{d+1}(acc2 + sum_{e<=d} b2_e)          (d = 2: the output epilogue instead)
//   output:  y = out_scale * (acc2 + b2_0 + b2_1 + b2_2)  -> fp32, staged in the (now idle) slab, TMA store
//
// HBM traffic per element-channel: 4 B in + 4 B out, against 12 B per dilation step (36 B per block) of three fused
// pairs (conv_pair_kernel) and 16 B per step of two plain launches: the k = 3 blocks were pure HBM time.
// A tile is 128*mh = 256 rows of which Mo = 256 - 2*halo are output rows; halo = sum_d (dil[d] + 1) * (k-1)/2 rows per
// side are recomputed (k = 3: 12 rows, 10 %).  Rows whose inputs would come from outside the tile hold garbage that
// never reaches an output row (the halo is exactly the receptive field); rows outside [0, valid length) are forced
// to zero in the slab at every step, which is each conv's zero padding at the utterance's ends.
//
// TMEM: per lane acc1 + acc2 = 2*mh*C columns.  C = 64: TWO tiles ("lanes") are in flight per CTA, lane 1 running
// `lane_lag` steps behind lane 0, so that the MMAs of one lane run under the epilogue of the other (and the F2 tiles
// of the two lanes are wanted half a tile apart: the loader's ring of one tile is enough).  C = 128 (or tf32 operands,
// whose slab is twice as large): one lane.
//
// All roles walk the same static schedule (run_schedule): global step n -> for each lane (tile, phase e = 0..6);
// phase e's epilogue (E0, E1_0, E2_0, E1_1, E2_1, E1_2, output) is followed, for e < 6, by MMA phase e (conv j = e).
// Warp roles (640 threads): 0..15 FOUR epilogue warpgroups (warp w reads TMEM lane quarter w % 4; warpgroup g takes the
// (half, 32-column chunk) items g, g + 4, ...), 16 TMEM allocator, 17 F2 loader, 18 weight producer, 19 MMA issuer.
// The epilogue is what bounds this kernel (seven element-wise passes per tile, one MUFU.SIN per element in six of
// them): with two warpgroups the schedulers issued 46 % of the time (ncu: two warps per scheduler cannot cover the
// MUFU / shared-memory latencies of their own dependent chains), hence four, at <= 102 registers per thread.
#pragma once
#include "conv_tc2.cuh"

namespace gnv {

constexpr int kChainPhases = 7;

struct ConvChainParams {
  int B, L, C, k;
  int dil[3];
  int n_chunks;                    // 128-byte K blocks per tap (C / KBE)
  int lanes, lane_lag;
  int halo, margin, Mo, tiles_m, total_tiles;
  int slab_rows, slab_kb_bytes, slab_bytes;   // slab: n_chunks K blocks of slab_rows rows x 128 B; tile row r = slab row r + margin
  int w_bytes, w_group, w_slot_bytes, sw;
  int in_ring;                     // F2 chunk slots per epilogue warpgroup
  uint32_t idesc;
  int mid_kind, round_tf32;
  float out_scale;
  const float* bias1[3];
  const float* bias2[3];
  const float* alpha1[3];          // Snake before conv1_d
  const float* alpha2[3];          // Snake before conv2_d
  const int* lengths;
  int len_mul, len_add;
  float* out;                      // fp32 [B, L, C]
  int out_accum;                   // out += ... (the stage's running sum already holds the earlier ResBlocks)
  int dbg;                         // timing experiments only (GONOVA_CHAIN_DBG): 1 no Snake math, 2 no slab stores, 4 no TMEM loads
  int c_tab;
  uint32_t off_slab, off_w, off_in, off_tab, off_bar;
};

struct ConvChainMaps {
  CUtensorMap W[6];                // conv1_0, conv2_0, conv1_1, conv2_1, conv1_2, conv2_2
  CUtensorMap IN;                  // F2: fp32 [B, L, C], 128-row x 32-column boxes
};

constexpr int kChainThreads = 640;
constexpr int kChainEpiWarps = 16, kChainEpiWg = 4;
constexpr int kChainWarpTmem = 16, kChainWarpLoader = 17, kChainWarpProducer = 18, kChainWarpMma = 19;

#ifdef __CUDACC__

// Timeline of one CTA for tuning (GONOVA_CHAIN_DBG & 8): (role << 56 | lane << 48 | phase << 40 | event << 32 | clock32) words
constexpr int kChainTraceCap = 8192;                 // three roles x 2048 words + spare
static __device__ unsigned long long g_chain_trace[kChainTraceCap];
static __device__ unsigned int g_chain_trace_n;
// fire-and-forget store into the role's own region (no atomics: an atomic's round trip would stall the traced warp)
__device__ __forceinline__ void chain_trace(int on, int role, int s, int e, int ev, unsigned int& idx) {
  if (!on) return;
  unsigned int c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  if (idx < 2048u)
    g_chain_trace[role * 2048 + idx] = ((unsigned long long)(role + 1) << 56) | ((unsigned long long)s << 48) |
                                       ((unsigned long long)e << 40) | ((unsigned long long)ev << 32) | c;
  ++idx;
}

template <typename E, int C, bool RAGGED>
__global__ void __launch_bounds__(kChainThreads, 1)
conv_chain_kernel(const ConvChainMaps* __restrict__ maps_g, const __grid_constant__ ConvChainParams p) {
  using namespace tc2;
  const ConvChainMaps& maps = *maps_g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sSlab = smem_base + p.off_slab, sW = smem_base + p.off_w, sIn = smem_base + p.off_in;
  // tables (pitch Cp): b1[3], cb2[3] (cumulative conv2 biases), then (alpha, 1/(alpha + 1e-9)) for alpha1[0..2], alpha2[0..2]
  float* tab = reinterpret_cast<float*>(smem_gen + p.off_tab);
  const int Cp = p.c_tab;
  const uint32_t bar0 = smem_base + p.off_bar;
  const uint32_t b_w_full = bar0, b_w_empty = b_w_full + 8u * p.sw;
  const uint32_t b_slab_ready = b_w_empty + 8u * p.sw;          // [lane]: the slab holds the next conv's operand
  const uint32_t b_acc_full = b_slab_ready + 16u;               // [lane][conv1 / conv2]
  const uint32_t b_in_full = b_acc_full + 32u, b_in_empty = b_in_full + 8u * kMaxInSlots;
  const uint32_t tmem_slot = b_in_empty + 8u * kMaxInSlots;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KBE = KBLK_BYTES / (int)sizeof(E);
  constexpr int kMh = 2;
  const int G = (int)gridDim.x;
  const int tr = ((p.dbg & 8) && blockIdx.x == 0 && lane == 0) ? 1 : 0;
  unsigned int tri = 0;

  if (warp == kChainWarpProducer && lane == 0) {
    for (int j = 0; j < 6; ++j) prefetch_tmap(&maps.W[j]);
    prefetch_tmap(&maps.IN);
  }
  if (warp == kChainWarpLoader && lane == 0) {
    for (int s = 0; s < p.sw; ++s) { mbar_init(b_w_full + 8u * s, 1); mbar_init(b_w_empty + 8u * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_slab_ready + 8u * s, kChainEpiWarps);
      mbar_init(b_acc_full + 16u * s, 1);
      mbar_init(b_acc_full + 16u * s + 8u, 1);
    }
    for (int s = 0; s < kMaxInSlots; ++s) {
      mbar_init(b_in_full + 8u * s, 1);
      mbar_init(b_in_empty + 8u * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kChainWarpTmem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = (int)threadIdx.x; c < Cp; c += (int)blockDim.x) {
    float cum = 0.f;
    for (int d = 0; d < 3; ++d) {
      tab[d * Cp + c] = (c < C && p.bias1[d]) ? p.bias1[d][c] : 0.f;
      cum += (c < C && p.bias2[d]) ? p.bias2[d][c] : 0.f;
      tab[(3 + d) * Cp + c] = cum;
      const float a1 = (c < C && p.alpha1[d]) ? p.alpha1[d][c] : 1.f;
      const float a2 = (c < C && p.alpha2[d]) ? p.alpha2[d][c] : 1.f;
      tab[(6 + 2 * d) * Cp + c] = a1;
      tab[(7 + 2 * d) * Cp + c] = 1.0f / (a1 + 1e-9f);
      tab[(12 + 2 * d) * Cp + c] = a2;
      tab[(13 + 2 * d) * Cp + c] = 1.0f / (a2 + 1e-9f);
    }
  }
  // slab margins (rows a conv reads beyond the tile): zero once, so that the halo rows are at least deterministic
  {
    const int n16 = p.lanes * p.slab_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
      *reinterpret_cast<uint4*>(smem_gen + p.off_slab + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait_then_release();

  // ---- the tile walk and the schedule every role follows ----
  auto tile_live = [&](int t) -> bool {
    if constexpr (!RAGGED) return true;
    const int b = t / p.tiles_m;
    const long valid = (long)p.lengths[b] * p.len_mul + p.len_add;
    return (long)(t % p.tiles_m) * p.Mo < valid + kDeadMargin;
  };
  auto next_tile = [&](int t) -> int {
    if constexpr (!RAGGED) {
      return t + G;
    } else {
      do { t += G; } while (t < p.total_tiles && !tile_live(t));
      return t;
    }
  };
  auto advance_lanes = [&](int t) -> int {              // the tile `lanes` positions further in this CTA's live list
    for (int i = 0; i < p.lanes && t < p.total_tiles; ++i) t = next_tile(t);
    return t;
  };
  int first[2];
  first[0] = (int)blockIdx.x;
  if constexpr (RAGGED) {
    if (first[0] < p.total_tiles && !tile_live(first[0])) first[0] = next_tile(first[0]);
  }
  first[1] = (p.lanes > 1 && first[0] < p.total_tiles) ? next_tile(first[0]) : p.total_tiles;
  // f(lane, tile, phase) in the one order all roles share: global step n; lane 1 runs lane_lag steps behind lane 0
  auto run_schedule = [&](auto&& f) {
    int t[2] = {first[0], first[1]};
    int e[2] = {0, 0};
    int lag = p.lane_lag;
    while (t[0] < p.total_tiles || t[1] < p.total_tiles) {
      if (t[0] < p.total_tiles) {
        f(0, t[0], e[0]);
        if (++e[0] == kChainPhases) { e[0] = 0; t[0] = advance_lanes(t[0]); }
      }
      if (lag > 0 && t[0] < p.total_tiles) { --lag; continue; }   // (lane 0 exhausted: lane 1 no longer waits)
      if (t[1] < p.total_tiles) {
        f(1, t[1], e[1]);
        if (++e[1] == kChainPhases) { e[1] = 0; t[1] = advance_lanes(t[1]); }
      }
    }
  };
  constexpr int lane_cols = 2 * kMh * C;                // TMEM columns per lane: acc1 then acc2
  constexpr int n_epi_chunks = C / kEpiCols;
  constexpr int n_items = kMh * n_epi_chunks;
  const int pad2 = (p.k - 1) / 2;

  if (warp == kChainWarpProducer) {
    // ===== weight producer: the taps of conv j for every MMA phase, in the issuer's order =====
    Ring rw;
    run_schedule([&](int, int, int e) {
      if (e >= 6) return;
      for (int ch = 0; ch < p.n_chunks; ++ch)
        for (int tap = 0; tap < p.k; tap += p.w_group) {
          const int ng = min(p.w_group, p.k - tap);
          mbar_wait(b_w_empty + 8u * rw.slot, rw.phase ^ 1u, 1);
          chain_trace(tr, 2, 0, e, tap, tri);
          if (elect_one()) {
            mbar_expect_tx(b_w_full + 8u * rw.slot, (uint32_t)(ng * p.w_bytes));
            for (int g = 0; g < ng; ++g)
              tma_load_2d(&maps.W[e], b_w_full + 8u * rw.slot, sW + rw.slot * p.w_slot_bytes + g * p.w_bytes,
                          ((tap + g) * p.n_chunks + ch) * KBE, 0);
          }
          __syncwarp();
          rw.advance(p.sw);
        }
    });
  } else if (warp == kChainWarpMma) {
    // ===== MMA issuer (whole warp walks the loops, one elected lane issues) =====
    const uint64_t slab_desc0 = umma_desc_sw128(sSlab), w_desc0 = umma_desc_sw128(sW);
    const uint64_t w_slot_units = (uint64_t)((uint32_t)p.w_slot_bytes >> 4), w_units = (uint64_t)((uint32_t)p.w_bytes >> 4);
    const uint32_t idesc = p.idesc;
    Ring rw;
    uint32_t n_ready[2] = {0u, 0u};
    run_schedule([&](int s, int, int e) {
      if (e >= 6) return;
      const bool conv1 = (e & 1) == 0;
      const int dl = conv1 ? p.dil[e >> 1] : 1;
      chain_trace(tr, 1, s, e, 0, tri);
      mbar_wait(b_slab_ready + 8u * s, n_ready[s] & 1u, 2);
      ++n_ready[s];
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      chain_trace(tr, 1, s, e, 1, tri);
      const uint32_t acc0 = tmem_base + (uint32_t)(s * lane_cols + (conv1 ? 0 : kMh * C));
      const uint32_t acc1 = acc0 + (uint32_t)C;
      uint32_t first_mma = conv1 ? 0u : 1u;             // conv2 accumulates onto the resident residual from its first MMA
      const uint64_t a_step = (uint64_t)((uint32_t)dl * (KBLK_BYTES >> 4));
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        // tile row r, tap j reads slab row r + margin + (j - (k-1)/2) * dl
        uint64_t ad_tap = slab_desc0 + (uint64_t)((uint32_t)(s * p.slab_bytes + ch * p.slab_kb_bytes +
                                                             (p.margin - pad2 * dl) * KBLK_BYTES) >> 4);
        for (int tap = 0; tap < p.k; tap += p.w_group) {
          const int ng = min(p.w_group, p.k - tap);
          mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          chain_trace(tr, 1, s, e, 3, tri);
          if (elect_one()) {
            uint64_t bd = w_desc0 + (uint64_t)rw.slot * w_slot_units;
            uint64_t ad0 = ad_tap;
            for (int g = 0; g < ng; ++g) {
              const uint64_t ad1 = ad0 + (uint64_t)(BLOCK_M * (KBLK_BYTES >> 4));
#pragma unroll
              for (int kk = 0; kk < KBLK_BYTES / 32; ++kk) {   // k-step outer, half inner: alternate the two accumulators
                const uint32_t ac = kk == 0 ? first_mma : 1u;
                umma<E>(acc0, ad0 + 2u * kk, bd + 2u * kk, idesc, ac);
                umma<E>(acc1, ad1 + 2u * kk, bd + 2u * kk, idesc, ac);
              }
              first_mma = 1u;
              bd += w_units;
              ad0 += a_step;
            }
            umma_commit(b_w_empty + 8u * rw.slot);
          }
          __syncwarp();
          first_mma = 1u;
          ad_tap += (uint64_t)ng * a_step;
          rw.advance(p.sw);
        }
      }
      if (elect_one()) umma_commit(b_acc_full + 16u * s + (conv1 ? 0u : 8u));
      __syncwarp();
      chain_trace(tr, 1, s, e, 2, tri);
    });
  } else if (warp == kChainWarpLoader) {
    // ===== F2 loader: the tile's fp32 chunks for phase 0, each into the ring of the warpgroup that consumes it =====
    int cnt[kChainEpiWg] = {0, 0, 0, 0};
    run_schedule([&](int, int t, int e) {
      if (e != 0) return;
      const int m_tile = t % p.tiles_m, b = t / p.tiles_m;
      const int row0 = m_tile * p.Mo - p.halo;
      for (int item = 0; item < n_items; ++item) {
        const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
        const int wgi = item % kChainEpiWg;
        const int k = cnt[wgi]++;
        const int slot = wgi * p.in_ring + k % p.in_ring;
        mbar_wait(b_in_empty + 8u * slot, (uint32_t)((k / p.in_ring) & 1) ^ 1u, 3);
        if (elect_one()) {
          mbar_expect_tx(b_in_full + 8u * slot, (uint32_t)(BLOCK_M * kEpiCols * 4));
          tma_load_3d(&maps.IN, b_in_full + 8u * slot, sIn + slot * (BLOCK_M * kEpiCols * 4), cc * kEpiCols,
                      row0 + h * BLOCK_M, b);
        }
        __syncwarp();
      }
    });
  } else if (warp < kChainEpiWarps) {
    // ===== epilogue warpgroups =====
    const int wg = warp >> 2, q = warp & 3;
    const int erow = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    Ring rin;
    uint32_t n_acc1[2] = {0u, 0u}, n_acc2[2] = {0u, 0u};   // completions consumed per lane: conv1 / conv2 accumulators
    const bool fast = p.mid_kind == ACT_SNAKE_FAST;

    // y = snake(v + bias) for 32 channels c0.. of tile row r, written straight into the slab (K-major SWIZZLE_128B operand
    // layout): eight channels at a time, so only the accumulator row itself stays in registers
    auto snake_to_slab = [&](const float (&v)[32], const float* bias_t, const float* al_t, uint32_t slab, int r, int c0,
                             bool live) {
      const int srow = r + p.margin;
      const int kb = c0 / KBE;
      const int cb = (c0 - kb * KBE) * (int)sizeof(E) / 16;
      const uint32_t rowa = slab + (uint32_t)kb * p.slab_kb_bytes + (uint32_t)srow * KBLK_BYTES;
      const uint32_t sx = (uint32_t)srow & 7u;
      const float4* bt = reinterpret_cast<const float4*>((bias_t ? bias_t : al_t) + c0);
      const float4* al = reinterpret_cast<const float4*>(al_t + c0);
      const float4* iv = reinterpret_cast<const float4*>(al_t + Cp + c0);
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        float y[8];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * j8 + jj;
          const float4 a4 = al[j], i4 = iv[j];
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias_t) b4 = bt[j];
          const float x0 = v[4 * j] + b4.x, x1 = v[4 * j + 1] + b4.y, x2 = v[4 * j + 2] + b4.z, x3 = v[4 * j + 3] + b4.w;
          if (p.dbg & 1) {
            y[4 * jj] = x0; y[4 * jj + 1] = x1; y[4 * jj + 2] = x2; y[4 * jj + 3] = x3;
          } else if (fast) {
            float s;
            s = __sinf(x0 * a4.x); y[4 * jj]     = fmaf(i4.x, s * s, x0);
            s = __sinf(x1 * a4.y); y[4 * jj + 1] = fmaf(i4.y, s * s, x1);
            s = __sinf(x2 * a4.z); y[4 * jj + 2] = fmaf(i4.z, s * s, x2);
            s = __sinf(x3 * a4.w); y[4 * jj + 3] = fmaf(i4.w, s * s, x3);
          } else {
            y[4 * jj]     = snake_precise(x0, a4.x, i4.x);
            y[4 * jj + 1] = snake_precise(x1, a4.y, i4.y);
            y[4 * jj + 2] = snake_precise(x2, a4.z, i4.z);
            y[4 * jj + 3] = snake_precise(x3, a4.w, i4.w);
          }
        }
        if (!live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = 0.f;
        }
        if (p.dbg & 2) {
          if (y[0] == 12345.f) sts128(rowa, y[0], y[1], y[2], y[3]);
        } else if constexpr (sizeof(E) == 2) {
          sts128u(rowa + ((((uint32_t)(cb + j8)) ^ sx) << 4), ElemIO<E>::pack2(y[0], y[1]), ElemIO<E>::pack2(y[2], y[3]),
                  ElemIO<E>::pack2(y[4], y[5]), ElemIO<E>::pack2(y[6], y[7]));
        } else {
          if (p.round_tf32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = round_tf32(y[i]);
          }
          sts128(rowa + ((((uint32_t)(cb + 2 * j8)) ^ sx) << 4), y[0], y[1], y[2], y[3]);
          sts128(rowa + ((((uint32_t)(cb + 2 * j8 + 1)) ^ sx) << 4), y[4], y[5], y[6], y[7]);
        }
      }
    };
    auto publish_slab = [&](int s) {
      fence_async_smem();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) mbar_arrive(b_slab_ready + 8u * s);
    };

    run_schedule([&](int s, int t, int e) {
      const int m_tile = t % p.tiles_m, b = t / p.tiles_m;
      const int g0 = m_tile * p.Mo - p.halo;            // global row of tile row 0
      int vr = p.L;
      if (p.lengths) {
        const int lv = p.lengths[b] * p.len_mul + p.len_add;
        vr = lv < vr ? lv : vr;
      }
      const uint32_t slab = sSlab + (uint32_t)(s * p.slab_bytes);
      const uint32_t acc1_col = (uint32_t)(s * lane_cols), acc2_col = acc1_col + (uint32_t)(kMh * C);
      const int tre = tr && warp == 0;
      chain_trace(tre, 0, s, e, 0, tri);
      if (e == 0) {
        // ---- E0: x := F2 tile -> TMEM acc2; slab := snake1_0(x) ----
#pragma unroll
        for (int item = wg; item < n_items; item += kChainEpiWg) {
          const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
          const int r = h * BLOCK_M + erow;
          const int g = g0 + r;
          const bool live = g >= 0 && g < vr;
          const int slot = wg * p.in_ring + rin.slot;
          const uint8_t* in_tile = smem_gen + p.off_in + (size_t)slot * (BLOCK_M * kEpiCols * 4);
          mbar_wait(b_in_full + 8u * slot, rin.phase, 5);
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(in_tile + erow * 128 + ((j ^ (erow & 7)) << 4));
            v[4 * j] = live ? r4.x : 0.f; v[4 * j + 1] = live ? r4.y : 0.f;
            v[4 * j + 2] = live ? r4.z : 0.f; v[4 * j + 3] = live ? r4.w : 0.f;
          }
          __syncwarp();
          if (elect_one()) mbar_arrive(b_in_empty + 8u * slot);
          rin.advance(p.in_ring);
          const int c0 = cc * kEpiCols;
          tmem_st32(lane_base + acc2_col + (uint32_t)(h * C + c0), v);
          snake_to_slab(v, nullptr, tab + 6 * Cp, slab, r, c0, true);       // (rows that are not live hold zeros: snake(0) = 0)
        }
        tmem_wait_st();
        publish_slab(s);
      } else if (e < 6) {
        // ---- E1_d (odd e): slab := snake2_d(acc1 + b1_d);  E2_d (even e): slab := snake1_{d+1}(acc2 + cum. b2) ----
        const bool after_conv1 = (e & 1) == 1;
        const int d = (e - 1) >> 1;
        if (after_conv1) {
          mbar_wait(b_acc_full + 16u * s, n_acc1[s] & 1u, 4);
          ++n_acc1[s];
        } else {
          mbar_wait(b_acc_full + 16u * s + 8u, n_acc2[s] & 1u, 4);
          ++n_acc2[s];
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        chain_trace(tre, 0, s, e, 1, tri);
        const float* bias_t = tab + (after_conv1 ? d : 3 + d) * Cp;
        const float* al_t = tab + (after_conv1 ? 12 + 2 * d : 6 + 2 * (d + 1)) * Cp;
#pragma unroll
        for (int item = wg; item < n_items; item += kChainEpiWg) {
          const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
          const int r = h * BLOCK_M + erow;
          const int g = g0 + r;
          const bool live = g >= 0 && g < vr;
          const int c0 = cc * kEpiCols;
          float v[32];
          if (p.dbg & 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (float)(i + r);
          } else {
            tmem_ld32(lane_base + (after_conv1 ? acc1_col : acc2_col) + (uint32_t)(h * C + c0), v);
          }
          snake_to_slab(v, bias_t, al_t, slab, r, c0, live);
        }
        publish_slab(s);
      } else {
        // ---- output: y = out_scale * (acc2 + b2_0 + b2_1 + b2_2).  Every thread owns 32 consecutive channels of one row:
        // one whole, aligned 128-byte line of the fp32 tensor, written from registers (no staging, no store to wait for)
        mbar_wait(b_acc_full + 16u * s + 8u, n_acc2[s] & 1u, 4);
        ++n_acc2[s];
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* bias_t = tab + 5 * Cp;
#pragma unroll
        for (int item = wg; item < n_items; item += kChainEpiWg) {
          const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
          const int r = h * BLOCK_M + erow;
          const int g = g0 + r;
          const int c0 = cc * kEpiCols;
          const bool store = r >= p.halo && r < p.halo + p.Mo && g < p.L;
          const bool live = g < vr;
          float4* dst = reinterpret_cast<float4*>(p.out + ((size_t)b * p.L + (store ? g : 0)) * C + c0);
          float v[32];
          tmem_ld32(lane_base + acc2_col + (uint32_t)(h * C + c0), v);
          if (store) {
            const float4* bt = reinterpret_cast<const float4*>(bias_t + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bt[j];
              float4 o;
              o.x = live ? (v[4 * j] + b4.x) * p.out_scale : 0.f;
              o.y = live ? (v[4 * j + 1] + b4.y) * p.out_scale : 0.f;
              o.z = live ? (v[4 * j + 2] + b4.z) * p.out_scale : 0.f;
              o.w = live ? (v[4 * j + 3] + b4.w) * p.out_scale : 0.f;
              // adding into the stage's running sum: a vector reduction at L2 (one writer per element, so the result is the
              // same fp32 add a load-add-store would do) — no load to wait for, no registers held for it
              if (p.out_accum)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                             : "memory");
              else
                dst[j] = o;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      chain_trace(tre, 0, s, e, 2, tri);
    });
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kChainWarpTmem)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#endif  // __CUDACC__

struct ConvChainLaunch {
  ConvChainMaps maps;
  const ConvChainMaps* d_maps = nullptr;
  ConvChainParams p;
  int grid;
  size_t smem_bytes;
  int elem_bytes;
};

// copies the timeline words recorded by CTA 0 (GONOVA_CHAIN_DBG & 8) to host memory and resets the counter
int conv_chain_read_trace(unsigned long long* out, int cap);

struct ConvChainSpec {
  const void* w1[3];               // packed [C, k*C] (E)
  const void* w2[3];
  const float* bias1[3];
  const float* bias2[3];
  const float* alpha1[3];
  const float* alpha2[3];
  int dil[3];
};

// in / out: fp32 [B, L, C].  Returns "" or the reason the block cannot run fused.
const char* make_conv_chain_launch(ConvChainLaunch* out, int elem_bytes, const float* in, float* out_raw,
                                   const ConvChainSpec& spec, int B, int L, int C, int k, float out_scale, int out_accum,
                                   int snake_kind, int round_tf32, int len_mul, int len_add, int max_ctas);
cudaError_t launch_conv_chain(const ConvChainLaunch& L, const int* lengths, cudaStream_t st);
cudaError_t conv_chain_init();

}  // namespace gnv
