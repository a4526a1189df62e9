// Persistent tcgen05 implicit-GEMM Conv1d / polyphase ConvTranspose1d for sm_100a — version 2.
//
//   D[m, n] = sum_{tap} sum_{ci} A[b, m + off0 + tap*tap_step, ci] * W[n, tap*C_in + ci]
//
// What changed against conv_tc.cuh (kept as the simple reference kernel, GNV_FLAG_TC_V1):
//   * persistent: one CTA per SM walks a static round-robin list of (utterance, m-tile, n-tile)
//     tiles, so barrier init / TMEM alloc / descriptor prefetch are paid once per launch, and the
//     epilogue of tile i overlaps the main loop of tile i+1 (accumulators double-buffered in TMEM);
//   * A "slab": one TMA box brings 128*mh + (taps-1)*dilation consecutive time rows of one 64-channel
//     block; every tap's MMA reads the SAME slab through a shared-memory descriptor whose start
//     address is advanced by tap*dilation rows (128 B per row in the K-major SWIZZLE_128B layout).
//     L2 -> SMEM traffic for activations drops by the number of taps;
//   * mh = 2: two 128-row accumulators share every weight tile (halves the weight traffic per FLOP);
//   * the epilogue never touches global memory with per-thread strided accesses: the residual /
//     running-sum tiles are prefetched by TMA into shared memory by a dedicated warp while the main
//     loop runs, results are staged in (swizzled) shared memory and leave through TMA stores, which
//     also clip rows past the tensor end.  The polyphase scatter of the ConvTranspose is a tensor map
//     per phase (row stride = up * C), so it is a plain box store as well.
//
// Warp roles (384 threads): 0..3 and 4..7 = two epilogue warpgroups (one TMEM lane quarter per warp)
// that take alternate 32-column chunks of a tile, 8 = TMEM allocator, 9 = epilogue-input TMA loader,
// 10 = TMA producer (A slabs + W tiles), 11 = MMA issuer.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace gnv {

constexpr int kMaxPhase = 8;
constexpr int kMaxSlab = 12;
constexpr int kEpiCols = 32;    // columns per epilogue chunk (one 128-byte fp32 row)
// rows past an utterance's valid length that are still written (as zeros) in a ragged batch: covers the halo any
// consumer reads ((k - 1) * dilation / 2 <= 25 rows) with room to spare
constexpr int kDeadMargin = 64;
constexpr int kMaxInSlots = 8;  // epilogue-input slots (residual / running-sum tiles prefetched by the loader warp)
// Warp roles of the persistent kernels (384 threads).  The single-lane control warps get the HIGHEST
// warp ids: the SM's warp arbiter favours higher ids, and the MMA issuer / TMA producer must never
// queue behind the (issue-bound) epilogue warps that share their scheduler.
constexpr int kWarpTmem = 8, kWarpLoader = 9, kWarpProducer = 10, kWarpMma = 11;

enum { EPI_IN0 = 0, EPI_IN1 = 1, EPI_RAW = 2, EPI_ACT0 = 3 };   // index into ConvTc2Maps::epi[phase][.]

struct ConvTc2Params {
  int B, M_rows;
  int n_taps, n_chunks;
  int block_n, n_tiles_n, mh, tiles_m, total_tiles;
  int transposed;                  // epilogue tensor maps are per phase (= n tile), columns start at 0
  int cta2;                        // 1: CTA pairs (cluster of 2): tcgen05.mma.cta_group::2, M = 256 across the pair,
                                   // each CTA stages its own 128*mh rows of A and HALF of every weight tile
  int row_adj[kMaxPhase];          // transposed: TMA row coordinate = m - row_adj[phase]
  // A slabs: slab s of a channel block covers taps [slab_tap0[s], slab_tap0[s+1]) and holds
  // a_n_boxes * a_box_rows rows starting at tile row m0 + slab_row0[s]; tap j reads tile rows m0 + tap_row[j] ...
  int n_slabs;
  int slab_tap0[kMaxSlab + 1];
  int slab_row0[kMaxSlab];
  int tap_row[kMaxSlab];
  int a_box_rows, a_n_boxes;       // every slab is a_n_boxes TMA boxes of a_box_rows rows
  int a_base_offset_mode;          // 1: descriptor base_offset = (start address >> 7) & 7
  int sa, sw, n_epi_wg, acc_bufs;   // n_epi_wg: epilogue warpgroups in use (1 or 2)
  int out_bufs;                     // output staging buffers per warpgroup (1 or 2)
  int in_ring;                      // epilogue-input slots per warpgroup (prefetch depth of the loader warp)
  int slab_bytes, w_bytes;
  int w_group, w_slot_bytes;       // taps per weight barrier; bytes of one weight ring slot (w_group * w_bytes)
  int tmem_cols;
  uint32_t idesc;
  int n_in, has_raw, n_act;
  int act_bytes;                   // bytes of one staged activation tile (128 rows x 32 cols x sizeof(E))
  int c_tab;                       // pitch of the per-channel tables (C_out rounded up to 32)
  uint32_t off_a, off_w, off_in, off_out, off_tab, off_bar;   // shared-memory carve-up
  EpiParams ep;
};

struct ConvTc2Maps {
  CUtensorMap A, W;
  CUtensorMap epi[kMaxPhase][6];   // in0 (residual), in1 (running sum), raw, act0..2
};

#ifdef __CUDACC__
namespace tc2 {
using namespace tc;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// after tcgen05.wait::ld: names the registers as in-out operands of an (empty) volatile asm, so that nothing that reads
// them is scheduled before the wait
__device__ __forceinline__ void tmem_ld_pin32(uint32_t (&r)[32]) {
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// 16-byte shared-memory accesses into a [128 rows][32 cols] tile as TMA lays it out:
// fp32 rows are 128 B (SWIZZLE_128B: chunk ^= row & 7), bf16 rows are 64 B (SWIZZLE_64B: chunk ^= (row >> 1) & 3).
__device__ __forceinline__ uint32_t tile_addr_f32(uint32_t tile, int row, int chunk) {
  return tile + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint32_t tile_addr_b16(uint32_t tile, int row, int chunk) {
  return tile + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// ---- CTA-pair (cta_group::2) helpers -------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's leader CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on a barrier of the pair's other CTA.  Default (.release.cta) semantics: the `.release.cluster` form compiles to
// MEMBAR.ALL.CTA + ERRBAR and cost 2.6 k cycles per use on the timeline of flow_blk_kernel.  What these arrivals publish
// is either a drained TMEM buffer (ordered by tcgen05.fence::before_thread_sync) or shared memory of the arriving CTA
// itself that its own tensor core will read (ordered by fence.proxy.async) — nothing the other CTA's threads load.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
template <typename E>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  const uint32_t z = 0u;
  if constexpr (sizeof(E) == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(z)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(z)
        : "memory");
  }
}
// commit of a CTA pair's MMAs: arrives on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

// griddepcontrol.wait returns when every prerequisite grid has completed and flushed its memory (immediately when
// the kernel was not launched with the programmatic-serialization attribute); launch_dependents lets the NEXT
// kernel in the stream begin its own prologue as soon as SMs free up.
__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// elect.sync: one lane of a converged warp; ptxas knows the guarded block runs in a single thread
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

struct Ring {
  int slot = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int depth) {
    if (++slot == depth) { slot = 0; phase ^= 1u; }
  }
};

}  // namespace tc2

struct ConvTc2Params;
namespace tc2 {
// Rare, long code paths stay out of line so that the epilogue loop fits the instruction cache.
static __device__ __noinline__ float snake_precise(float x, float alpha, float inv) {
  const float s = sinf(x * alpha);
  return x + inv * (s * s);
}
static __device__ __noinline__ float elu_precise(float x) { return x > 0.f ? x : expm1f(x); }

// The part of the epilogue every conv kernel shares: given one 128-row x 32-column chunk of
// (accumulator + bias) in registers, add the residual / running-sum tiles the loader warp staged,
// scale, zero dead rows, stage the fp32 result and its activated copies in (swizzled) shared memory
// and hand them to TMA stores.  Called by all threads of one epilogue warpgroup.
struct EpiCtx {
  const EpiParams* ep;
  const float* tab;            // bias[c_tab], then (alpha[c_tab], 1/(alpha+1e-9)[c_tab]) per activation
  int c_tab, n_in, has_raw, n_act, act_bytes, n_epi_wg, out_bufs;
  int in_ring;                 // input slots per warpgroup (the loader runs that many chunks ahead of it)
  uint8_t* smem_in;            // the loader's input slots: warpgroup g owns slots [g*in_ring, (g+1)*in_ring)
  uint32_t b_in_full, b_in_empty;
  uint32_t obase_wg;           // this warpgroup's staging buffers
  int out_stride;
  int wg, erow, lane;
  bool elected;
};

template <typename E>
__device__ __forceinline__ void epi_finish_item(const EpiCtx& c, float (&v)[32], bool live, int c0,
                                                const CUtensorMap* maps6, int cs, int mrow, int b, Ring& rin, int& ob) {
    // ---- residual, scale, running sum ----
    if (c.n_in > 0) {
      const int in_slot = c.wg * c.in_ring + rin.slot;
      const uint8_t* in_tile = c.smem_in + (size_t)in_slot * c.n_in * (BLOCK_M * kEpiCols * 4);
      mbar_wait(c.b_in_full + 8u * in_slot, rin.phase, 5);
      if (c.ep->res) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = *reinterpret_cast<const float4*>(in_tile + c.erow * 128 + ((j ^ (c.erow & 7)) << 4));
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
      if (c.ep->raw_scale != 1.0f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= c.ep->raw_scale;
      }
      if (c.ep->raw_accum) {
        const uint8_t* in1 = in_tile + (c.ep->res ? BLOCK_M * kEpiCols * 4 : 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = *reinterpret_cast<const float4*>(in1 + c.erow * 128 + ((j ^ (c.erow & 7)) << 4));
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
      __syncwarp();
      if (elect_one()) mbar_arrive(c.b_in_empty + 8u * in_slot);
      rin.advance(c.in_ring);
    } else if (c.ep->raw_scale != 1.0f) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= c.ep->raw_scale;
    }
    if (!live) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    // ---- stage the outputs (the store that last used this staging buffer must have drained) ----
    const uint32_t obase = c.obase_wg + ob * c.out_stride;
    // Each WARP stages and stores its own 32 rows of the chunk (its lane 0 owns the bulk-store groups), so the four
    // warps of a warpgroup never wait for each other.
    if (elect_one()) {                                     // (always the same lane: it owns this warp's bulk groups)
      if (c.out_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
    }
    __syncwarp();
    uint32_t o = obase;
    if (c.has_raw) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(tile_addr_f32(o, c.erow, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      o += BLOCK_M * kEpiCols * 4;
    }
    for (int a = 0; a < c.n_act; ++a) {
      // The activation kind is uniform for the launch: one tight loop per kind (a per-element
      // switch makes the compiler evaluate every variant and select).  Every activation maps
      // 0 -> 0, so rows that are not live (v == 0) need no extra select.
      const int kind = c.ep->act_kind[a];
      const float4* al = reinterpret_cast<const float4*>(c.tab + (1 + 2 * a) * c.c_tab + c0);
      const float4* iv = reinterpret_cast<const float4*>(c.tab + (2 + 2 * a) * c.c_tab + c0);
      float y[32];
      if (kind == ACT_NONE) {
        // (first, and a real branch: with the plain copy as the chain's last `else` ptxas if-converted it together with
        // the GELU arm, and the q/k/v projection of the flow estimator evaluated an erff per element it then threw away)
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = v[i];
      } else if (kind == ACT_SNAKE_FAST) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 a4 = al[j], i4 = iv[j];
          float s;
          s = __sinf(v[4 * j] * a4.x);     y[4 * j]     = fmaf(i4.x, s * s, v[4 * j]);
          s = __sinf(v[4 * j + 1] * a4.y); y[4 * j + 1] = fmaf(i4.y, s * s, v[4 * j + 1]);
          s = __sinf(v[4 * j + 2] * a4.z); y[4 * j + 2] = fmaf(i4.z, s * s, v[4 * j + 2]);
          s = __sinf(v[4 * j + 3] * a4.w); y[4 * j + 3] = fmaf(i4.w, s * s, v[4 * j + 3]);
        }
      } else if (kind == ACT_LRELU) {
        const float slope = c.ep->act_slope[a];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = v[i] > 0.f ? v[i] : v[i] * slope;
      } else if (kind == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = gelu_erf_fast(v[i]);
      } else if (kind == ACT_SILU) {
        // x sigmoid(x) through MUFU.EX2 + MUFU.RCP (the Conformer feed-forward of the flow front)
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __fdividef(v[i], 1.0f + __expf(-v[i]));
      } else if (kind == ACT_ELU_FAST) {
        // (the f0 predictor's five conv layers: with the out-of-line expm1f they were EPILOGUE-bound — ncu: tensor pipe
        // active 22 %, ~50 instructions per element — at 0.11 ms per layer against 0.04 ms of MMA time)
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = v[i] > 0.f ? v[i] : __expf(v[i]) - 1.0f;
      } else if (kind == ACT_SNAKE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 a4 = al[j], i4 = iv[j];
          y[4 * j]     = snake_precise(v[4 * j], a4.x, i4.x);
          y[4 * j + 1] = snake_precise(v[4 * j + 1], a4.y, i4.y);
          y[4 * j + 2] = snake_precise(v[4 * j + 2], a4.z, i4.z);
          y[4 * j + 3] = snake_precise(v[4 * j + 3], a4.w, i4.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = elu_precise(v[i]);
      }
      if constexpr (sizeof(E) == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(tile_addr_b16(o, c.erow, j), ElemIO<E>::pack2(y[8 * j], y[8 * j + 1]),
                  ElemIO<E>::pack2(y[8 * j + 2], y[8 * j + 3]), ElemIO<E>::pack2(y[8 * j + 4], y[8 * j + 5]),
                  ElemIO<E>::pack2(y[8 * j + 6], y[8 * j + 7]));
      } else {
        if (c.ep->round_tf32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = round_tf32(y[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(tile_addr_f32(o, c.erow, j), y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      }
      o += c.act_bytes;
    }
    fence_async_smem();
    __syncwarp();
    if (elect_one()) {
      const int wq = c.erow >> 5;                          // this warp's 32-row quarter of the 128-row chunk
      uint32_t src = obase;
      if (c.has_raw) {
        tma_store_3d(maps6 + EPI_RAW, src + wq * (32 * kEpiCols * 4), cs, mrow + 32 * wq, b);
        src += BLOCK_M * kEpiCols * 4;
      }
      for (int a = 0; a < c.n_act; ++a) {
        tma_store_3d(maps6 + EPI_ACT0 + a, src + wq * (c.act_bytes >> 2), cs, mrow + 32 * wq, b);
        src += c.act_bytes;
      }
      bulk_commit();
    }
    if (++ob == c.out_bufs) ob = 0;
}

template <typename E>
__device__ __noinline__ void emit_row0(const ConvTc2Params& p, const float* tab, const float (&v)[32], int b, int c0,
                                       bool live0);
}  // namespace tc2

// CTA2 = true: CTA pairs (cluster of two).  A separate instantiation, because a kernel image that contains
// cta_group::2 instructions can only be launched with an even cluster size ("cluster misconfiguration" otherwise).
template <typename E, bool CTA2, bool RAGGED>
__global__ void __launch_bounds__(384, 1)
conv_tc2_kernel(const ConvTc2Maps* __restrict__ maps_g, const __grid_constant__ ConvTc2Params p) {
  using namespace tc2;
  const ConvTc2Maps& maps = *maps_g;          // tensor maps live in global memory (6.4 KB: too big for the
                                              // 4 KB of classic parameter space the TMA unit can always reach)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base + p.off_a, sW = smem_base + p.off_w;
  const uint32_t sIn = smem_base + p.off_in, sOut = smem_base + p.off_out;
  float* tab = reinterpret_cast<float*>(smem_gen + p.off_tab);      // bias[C_out], then per act: alpha[C_out], inv[C_out]
  // barriers (8 B each): a_full[sa] a_empty[sa] w_full[sw] w_empty[sw] acc_full[2] acc_empty[2] in_full[2] in_empty[2]
  const uint32_t bar0 = smem_base + p.off_bar;
  const uint32_t b_a_full = bar0, b_a_empty = b_a_full + 8u * p.sa;
  const uint32_t b_w_full = b_a_empty + 8u * p.sa, b_w_empty = b_w_full + 8u * p.sw;
  const uint32_t b_acc_full = b_w_empty + 8u * p.sw, b_acc_empty = b_acc_full + 16u;
  const uint32_t b_in_full = b_acc_empty + 16u, b_in_empty = b_in_full + 8u * kMaxInSlots;
  const uint32_t tmem_slot = b_in_empty + 8u * kMaxInSlots;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KBE = KBLK_BYTES / (int)sizeof(E);
  const int crank = CTA2 ? (int)cluster_ctarank() : 0;     // rank inside the CTA pair; 0 = leader (issues the MMAs)
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // first tile / tile stride of this CTA (pair)
  const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == kWarpProducer && lane == 0) {
    prefetch_tmap(&maps.A);
    prefetch_tmap(&maps.W);
  }
  if (warp == kWarpLoader && lane == 0) {
    for (int s = 0; s < p.sa; ++s) { mbar_init(b_a_full + 8u * s, 1); mbar_init(b_a_empty + 8u * s, 1); }
    for (int s = 0; s < p.sw; ++s) { mbar_init(b_w_full + 8u * s, 1); mbar_init(b_w_empty + 8u * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_acc_full + 8u * s, 1);
      mbar_init(b_acc_empty + 8u * s, 4 * p.n_epi_wg * (CTA2 ? 2 : 1));   // pair: both CTAs' epilogues arrive on the leader
    }
    for (int s = 0; s < kMaxInSlots; ++s) {
      mbar_init(b_in_full + 8u * s, 1);
      mbar_init(b_in_empty + 8u * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpTmem) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // per-channel epilogue tables: bias, then (alpha, 1/(alpha + 1e-9)) per fused activation
  {
    const int C = p.ep.C_out, Cp = p.c_tab;
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
      tab[c] = (c < C && p.ep.bias) ? p.ep.bias[c] : 0.f;
      for (int a = 0; a < p.n_act; ++a) {
        const float al = (c < C && p.ep.act_alpha[a]) ? p.ep.act_alpha[a][c] : 1.f;
        tab[(1 + 2 * a) * Cp + c] = al;
        tab[(2 + 2 * a) * Cp + c] = 1.0f / (al + 1e-9f);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();          // the peer's barriers are initialised before any remote arrive / TMA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, per-channel tables from the
  // static weights) may overlap the previous kernel's tail; from here on we read what it produced.
  pdl_wait_then_release();

  const int rows_per_cta = BLOCK_M * p.mh;
  const int rows_per_tile = rows_per_cta * (CTA2 ? 2 : 1);   // a CTA pair's tile: the leader's rows, then the peer's
  const int n_epi_chunks = p.block_n / kEpiCols;
  // Ragged batches (RAGGED instantiation, launched when `lengths` is given): a tile whose every output row lies
  // kDeadMargin or more rows past its utterance's valid length is skipped by ALL roles (same predicate, so the barrier
  // sequences stay in step).  Rows in [valid, valid + margin) are still written (as zeros): that is the zero padding the
  // next layer's last valid rows read.  The dense instantiation compiles to the plain strided walk.
  auto tile_live = [&](int t) -> bool {
    if constexpr (!RAGGED) return true;
    const int qt = t / p.n_tiles_n;
    const int m_tile = qt % p.tiles_m, b = qt / p.tiles_m;
    const long valid = (long)p.ep.lengths[b] * p.ep.len_mul + p.ep.len_add;
    const long first_row = (long)m_tile * rows_per_tile * p.ep.up - p.ep.pad_out + p.ep.shift;
    return first_row < valid + kDeadMargin;
  };
  auto next_tile = [&](int t) -> int {
    if constexpr (!RAGGED) {
      return t + tile_step;
    } else {
      do { t += tile_step; } while (t < p.total_tiles && !tile_live(t));
      return t;
    }
  };
  const int tile_first = (RAGGED && tile0 < p.total_tiles && !tile_live(tile0)) ? next_tile(tile0) : tile0;

  if (warp == kWarpProducer) {
    {
      // ===== TMA producer: A slabs and W tile groups, in the order the MMA issuer consumes them =====
      // (whole warp in the loops, one elected lane issues: see the MMA issuer)
      Ring ra, rw;
      for (int t = tile_first; t < p.total_tiles; t = next_tile(t)) {
        int q = t;
        const int n_tile = q % p.n_tiles_n; q /= p.n_tiles_n;
        const int m_tile = q % p.tiles_m;
        const int b = q / p.tiles_m;
        // polyphase tiles start at the phase's first valid GEMM row (row_adj), so that every TMA store
        // coordinate is non-negative
        const int m0 = m_tile * rows_per_tile + crank * rows_per_cta + (p.transposed ? p.row_adj[n_tile] : 0), n0 = n_tile * p.block_n;
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          for (int s = 0; s < p.n_slabs; ++s) {
            mbar_wait(b_a_empty + 8u * ra.slot, ra.phase ^ 1u, 1);
            const uint32_t dst = sA + ra.slot * p.slab_bytes;
            // a TMA box holds at most 256 rows: taller slabs arrive as two boxes of a_box_rows rows
            // CTA pair: both CTAs' loads complete on the LEADER's barrier, which expects both byte counts
            const uint32_t a_bytes = (uint32_t)(p.a_n_boxes * p.a_box_rows) * KBLK_BYTES;
            const int r0 = m0 + p.slab_row0[s];
            if (elect_one()) {
            if (crank == 0) mbar_expect_tx(b_a_full + 8u * ra.slot, CTA2 ? 2u * a_bytes : a_bytes);
            for (int bx = 0; bx < p.a_n_boxes; ++bx) {
              if constexpr (CTA2)
                tma_load_3d_2sm(&maps.A, (b_a_full + 8u * ra.slot) & kPeerBitMask,
                                dst + (uint32_t)(bx * p.a_box_rows) * KBLK_BYTES, ch * KBE, r0 + bx * p.a_box_rows, b);
              else
                tma_load_3d(&maps.A, b_a_full + 8u * ra.slot, dst + (uint32_t)(bx * p.a_box_rows) * KBLK_BYTES, ch * KBE,
                            r0 + bx * p.a_box_rows, b);
            }
            }
            __syncwarp();
            ra.advance(p.sa);
            // weight tiles travel in groups of up to w_group taps per barrier (fewer waits for small tiles)
            for (int tap = p.slab_tap0[s]; tap < p.slab_tap0[s + 1]; tap += p.w_group) {
              const int ng = min(p.w_group, p.slab_tap0[s + 1] - tap);
              mbar_wait(b_w_empty + 8u * rw.slot, rw.phase ^ 1u, 1);
              // CTA pair: w_bytes is this CTA's HALF of the weight tile (rows n0 + crank*block_n/2 ...)
              if (elect_one()) {
              if (crank == 0) mbar_expect_tx(b_w_full + 8u * rw.slot, (uint32_t)(ng * p.w_bytes) * (CTA2 ? 2u : 1u));
              for (int g = 0; g < ng; ++g) {
                if constexpr (CTA2)
                  tma_load_2d_2sm(&maps.W, (b_w_full + 8u * rw.slot) & kPeerBitMask,
                                  sW + rw.slot * p.w_slot_bytes + g * p.w_bytes, ((tap + g) * p.n_chunks + ch) * KBE,
                                  n0 + crank * (p.block_n >> 1));
                else
                  tma_load_2d(&maps.W, b_w_full + 8u * rw.slot, sW + rw.slot * p.w_slot_bytes + g * p.w_bytes,
                              ((tap + g) * p.n_chunks + ch) * KBE, n0);
              }
              }
              __syncwarp();
              rw.advance(p.sw);
            }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (crank == 0) {
      // ===== MMA issuer (CTA pair: the leader issues for both CTAs) =====
      // The WHOLE warp walks the loops (warp-uniform control flow, waits included) and one elected lane issues:
      // ptxas then keeps descriptors and loop state in uniform registers and emits the UTCHMMAs back to back.  Under
      // `if (lane == 0)` every MMA was wrapped in an ELECT / BRA.U.ANY loop of its own — and the issue path is on
      // the critical path of the N <= 128 layers (one extra R2UR per MMA measured 16 % on the k = 11 pairs).
      // Descriptors are built once; per MMA only the 14-bit start-address field moves (low word add).
      const uint64_t a_desc0 = umma_desc_sw128(sA), w_desc0 = umma_desc_sw128(sW);
      // loop-invariant strides in descriptor units (16 B): the issue path is the critical path of the N <= 128 layers
      const uint64_t w_slot_units = (uint64_t)((uint32_t)p.w_slot_bytes >> 4), w_units = (uint64_t)((uint32_t)p.w_bytes >> 4);
      const uint64_t slab_units = (uint64_t)((uint32_t)p.slab_bytes >> 4);
      const uint32_t idesc = p.idesc, cols_n = (uint32_t)p.block_n;
      const bool alt = p.mh == 2 && !p.a_base_offset_mode;
      const int mh = p.mh, w_group = p.w_group;
      Ring ra, rw, racc;
      for (int t = tile_first; t < p.total_tiles; t = next_tile(t)) {
        mbar_wait(b_acc_empty + 8u * racc.slot, racc.phase ^ 1u, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc0 = tmem_base + (uint32_t)(racc.slot * mh) * cols_n;
        uint32_t accum = 0u;                                   // 0 for the tile's first MMA of each half
        for (int ch = 0; ch < p.n_chunks; ++ch) {
          for (int s = 0; s < p.n_slabs; ++s) {
            mbar_wait(b_a_full + 8u * ra.slot, ra.phase, 2);
            const uint64_t a_slab = a_desc0 + (uint64_t)ra.slot * slab_units;
            const int row0 = p.slab_row0[s];
            for (int tap = p.slab_tap0[s]; tap < p.slab_tap0[s + 1]; tap += w_group) {
              const int ng = min(w_group, p.slab_tap0[s + 1] - tap);
              mbar_wait(b_w_full + 8u * rw.slot, rw.phase, 2);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                uint64_t bd = w_desc0 + (uint64_t)rw.slot * w_slot_units;
                uint32_t ac0 = accum;
                for (int g = 0; g < ng; ++g) {
                  const uint64_t ad0 = a_slab + (uint64_t)((uint32_t)(p.tap_row[tap + g] - row0) * (KBLK_BYTES >> 4));
                  if (alt) {
                    // k-step outer, half inner: consecutive MMAs alternate between the two accumulators, so an MMA
                    // never queues behind the previous one's accumulate into the same TMEM tile (measured +11 % on
                    // the N = 64 layers)
                    const uint64_t ad1 = ad0 + (uint64_t)(BLOCK_M * (KBLK_BYTES >> 4));
                    const uint32_t acc1 = acc0 + cols_n;
#pragma unroll
                    for (int k = 0; k < KBLK_BYTES / 32; ++k) {
                      const uint32_t ac = k == 0 ? ac0 : 1u;
                      if constexpr (CTA2) {
                        umma_2sm<E>(acc0, ad0 + 2u * k, bd + 2u * k, idesc, ac);
                        umma_2sm<E>(acc1, ad1 + 2u * k, bd + 2u * k, idesc, ac);
                      } else {
                        umma<E>(acc0, ad0 + 2u * k, bd + 2u * k, idesc, ac);
                        umma<E>(acc1, ad1 + 2u * k, bd + 2u * k, idesc, ac);
                      }
                    }
                  } else {
                    for (int h = 0; h < mh; ++h) {
                      uint64_t ad = ad0 + (uint64_t)(h * BLOCK_M * (KBLK_BYTES >> 4));
                      if (p.a_base_offset_mode) ad |= (uint64_t)((((uint32_t)ad & 0x3FFFu) >> 3) & 7u) << 49;
                      const uint32_t acc = acc0 + (uint32_t)h * cols_n;
                      if constexpr (CTA2) {
                        umma_2sm<E>(acc, ad, bd, idesc, ac0);
#pragma unroll
                        for (int k = 1; k < KBLK_BYTES / 32; ++k) umma_2sm<E>(acc, ad + 2u * k, bd + 2u * k, idesc, 1u);
                      } else {
                        umma<E>(acc, ad, bd, idesc, ac0);
#pragma unroll
                        for (int k = 1; k < KBLK_BYTES / 32; ++k) umma<E>(acc, ad + 2u * k, bd + 2u * k, idesc, 1u);
                      }
                    }
                  }
                  ac0 = 1u;
                  bd += w_units;
                }
                if constexpr (CTA2) umma_commit_2sm(b_w_empty + 8u * rw.slot); else umma_commit(b_w_empty + 8u * rw.slot);
              }
              __syncwarp();
              accum = 1u;
              rw.advance(p.sw);
            }
            if (elect_one()) {
              if constexpr (CTA2) umma_commit_2sm(b_a_empty + 8u * ra.slot); else umma_commit(b_a_empty + 8u * ra.slot);
            }
            __syncwarp();
            ra.advance(p.sa);
          }
        }
        if (elect_one()) {
          if constexpr (CTA2) umma_commit_2sm(b_acc_full + 8u * racc.slot); else umma_commit(b_acc_full + 8u * racc.slot);
        }
        __syncwarp();
        racc.advance(p.acc_bufs);
      }
    }
  } else if (warp == kWarpLoader) {
    if (p.n_in > 0) {
      // ===== epilogue-input loader: residual / running-sum tiles -> shared memory (whole warp, elected lane issues) =====
      // Each warpgroup owns a ring of in_ring slots; chunk i of a tile goes to warpgroup i % 2 (the one that will
      // consume it), so the loader runs in_ring chunks ahead of every warpgroup.
      int cnt[2] = {0, 0};                               // chunks handed to each warpgroup so far
      const int n_items = p.mh * n_epi_chunks;
      for (int t = tile_first; t < p.total_tiles; t = next_tile(t)) {
        int q = t;
        const int n_tile = q % p.n_tiles_n; q /= p.n_tiles_n;
        const int m_tile = q % p.tiles_m;
        const int b = q / p.tiles_m;
        const int n0 = n_tile * p.block_n;
        const int ph = p.transposed ? n_tile : 0;
        const int cbase = p.transposed ? 0 : n0;
        for (int item = 0; item < n_items; ++item) {
          const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
          const int mrow = m_tile * rows_per_tile + crank * rows_per_cta + h * BLOCK_M;   // TMA row coordinate (GEMM row - row_adj)
          const int wgi = p.n_epi_wg == 2 ? (item & 1) : 0;                               // the warpgroup that will consume it
          const int k = cnt[wgi]++;
          const int slot = wgi * p.in_ring + k % p.in_ring;
          mbar_wait(b_in_empty + 8u * slot, (uint32_t)((k / p.in_ring) & 1) ^ 1u, 3);
          if (elect_one()) {
            mbar_expect_tx(b_in_full + 8u * slot, (uint32_t)p.n_in * (BLOCK_M * kEpiCols * 4));
            for (int i = 0; i < p.n_in; ++i)
              tma_load_3d(&maps.epi[ph][EPI_IN0 + i], b_in_full + 8u * slot,
                          sIn + (slot * p.n_in + i) * (BLOCK_M * kEpiCols * 4), cbase + cc * kEpiCols, mrow, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 8 && (warp >> 2) < p.n_epi_wg) {
    // ===== epilogue: TMEM -> registers -> bias / residual / activations -> shared -> TMA store =====
    // Up to two warpgroups of 4 warps; the (half, 32-column chunk) items of a tile alternate between
    // them.  Each warpgroup owns one input slot (the loader's ring of two), one output staging buffer,
    // one named barrier and its own bulk-store groups, so the two never synchronise with each other.
    const int wg = warp >> 2;
    const int q = warp & 3;
    const int erow = q * 32 + lane;                     // row inside a 128-row half == TMEM lane
    const bool elected = (threadIdx.x & 127) == 0;
    const int out_stride = (p.has_raw ? BLOCK_M * kEpiCols * 4 : 0) + p.n_act * p.act_bytes;
    const uint32_t obase_wg = sOut + wg * p.out_bufs * out_stride;
    int ob = 0;
    Ring racc, rin;                                     // rin: this warpgroup's view of the input ring
    const int n_items = p.mh * n_epi_chunks;
    EpiCtx ectx;
    ectx.ep = &p.ep; ectx.tab = tab; ectx.c_tab = p.c_tab; ectx.n_in = p.n_in; ectx.has_raw = p.has_raw;
    ectx.n_act = p.n_act; ectx.act_bytes = p.act_bytes; ectx.n_epi_wg = p.n_epi_wg; ectx.out_bufs = p.out_bufs;
    ectx.smem_in = smem_gen + p.off_in; ectx.b_in_full = b_in_full; ectx.b_in_empty = b_in_empty; ectx.in_ring = p.in_ring;
    ectx.obase_wg = obase_wg; ectx.out_stride = out_stride; ectx.wg = wg; ectx.erow = erow; ectx.lane = lane;
    ectx.elected = elected;
    for (int t = tile_first; t < p.total_tiles; t = next_tile(t)) {
      int qq = t;
      const int n_tile = qq % p.n_tiles_n; qq /= p.n_tiles_n;
      const int m_tile = qq % p.tiles_m;
      const int b = qq / p.tiles_m;
      const int ph = p.transposed ? n_tile : 0;
      const int m0 = m_tile * rows_per_tile + crank * rows_per_cta + (p.transposed ? p.row_adj[ph] : 0), n0 = n_tile * p.block_n;
      const int cbase = p.transposed ? 0 : n0;           // channel of the tile's first column
      int valid_rows = p.ep.L_out;
      if (p.ep.lengths) {
        const int lv = p.ep.lengths[b] * p.ep.len_mul + p.ep.len_add;
        valid_rows = lv < valid_rows ? lv : valid_rows;
      }
      mbar_wait(b_acc_full + 8u * racc.slot, racc.phase, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int item = (p.n_epi_wg == 2 ? wg : 0); item < n_items; item += p.n_epi_wg) {
        const int h = item / n_epi_chunks, cc = item - h * n_epi_chunks;
        const int m = m0 + h * BLOCK_M + erow;
        const int p0 = m * p.ep.up + (p.transposed ? n_tile : 0) - p.ep.pad_out;
        const int row = p0 + p.ep.shift;
        const bool live = (m < p.M_rows) && (p0 >= 0) && (p0 < p.ep.L_store) && (row < valid_rows);
        const int mrow = m_tile * rows_per_tile + crank * rows_per_cta + h * BLOCK_M;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)((racc.slot * p.mh + h) * p.block_n);
        float v[32];
        tmem_ld32(trow + (uint32_t)(cc * kEpiCols), v);
        const int c0 = cbase + cc * kEpiCols;          // channel index of v[0]
        // ---- bias ----   (tables are padded to 32 columns: conv_post has C_out = 18)
        {
          const float4* bt = reinterpret_cast<const float4*>(tab + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bt[j];
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        // ReflectionPad1d((1,0)): the conv row p0 == dup_row is ALSO output row 0, where it meets the
        // residual of row 0 (x = pad(ups(x)); x = x + si).  One thread per utterance and channel chunk;
        // out of line so that its math is never speculated into the hot path.
        if (p.ep.dup_row >= 0 && p0 == p.ep.dup_row && m < p.M_rows) {
          float tmp[32];                                   // a copy: v itself must stay in registers
#pragma unroll
          for (int i = 0; i < 32; ++i) tmp[i] = v[i];
          tc2::emit_row0<E>(p, tab, tmp, b, c0, 0 < valid_rows);
        }
        tc2::epi_finish_item<E>(ectx, v, live, c0, &maps.epi[ph][0], cbase + cc * kEpiCols, mrow, b, rin, ob);
      }
      // accumulator drained: hand the TMEM buffer back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        if constexpr (CTA2) mbar_arrive_cluster((b_acc_empty + 8u * racc.slot) & kPeerBitMask);   // the leader's issuer waits for both CTAs
        else mbar_arrive(b_acc_empty + 8u * racc.slot);
      }
      racc.advance(p.acc_bufs);
    }
    if (elect_one()) bulk_wait_read<0>();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();          // the peer may still be the target of multicast commits / remote arrives
  if (warp == kWarpTmem) {
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
  }
}

namespace tc2 {
template <typename E>
__device__ __noinline__ void emit_row0(const ConvTc2Params& p, const float* tab, const float (&v)[32], int b, int c0,
                                       bool live0) {
  const int C = p.ep.C_out;
  const size_t g0 = (size_t)b * p.ep.L_out * p.ep.C_pitch + c0;
  for (int i = 0; i < 32 && c0 + i < C; ++i) {
    float u = v[i];
    if (p.ep.res) u += p.ep.res[g0 + i];
    u *= p.ep.raw_scale;
    if (p.ep.raw_accum) u += p.ep.raw[g0 + i];
    if (!live0) u = 0.f;
    if (p.has_raw) p.ep.raw[g0 + i] = u;
    for (int a = 0; a < p.n_act; ++a) {
      const float y0 = live0 ? act_apply(p.ep.act_kind[a], u, tab[(1 + 2 * a) * p.c_tab + c0 + i], p.ep.act_slope[a]) : 0.f;
      E* dst = reinterpret_cast<E*>(p.ep.act_out[a]) + g0 + i;
      if constexpr (sizeof(E) == 4) ElemIO<E>::store(dst, p.ep.round_tf32 ? round_tf32(y0) : y0);
      else ElemIO<E>::store(dst, y0);
    }
  }
}
}  // namespace tc2

#endif  // __CUDACC__

// ---- host side --------------------------------------------------------------------------------
// Launch of a persistent kernel: optional cluster of two (CTA pairs) and programmatic dependent launch.
bool pdl_enabled();      // GONOVA_PDL=0 switches programmatic dependent launch off
int in_slots_cap();      // GONOVA_IN_SLOTS caps the epilogue-input prefetch slots
template <typename Kern, typename... Args>
inline cudaError_t launch_persistent(Kern kernel, int grid, size_t smem, cudaStream_t st, bool cluster2, int threads,
                                     Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster2) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = n ? attr : nullptr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

struct ConvTc2Launch {
  ConvTc2Maps maps;                 // host copy; the caller uploads it and sets d_maps before launching
  const ConvTc2Maps* d_maps = nullptr;
  ConvTc2Params p;
  int grid;
  size_t smem_bytes;
  int elem_bytes;
};

struct ConvTc2Options {
  int slab_mode = 1;     // 0: one slab per tap; 1: one slab per channel block, taps by descriptor row offset (base_offset 0 — measured correct); 2: same with base_offset set (measured WRONG on B200, kept for the record)
  int mh = 0;            // 0 = choose
  int w_group = 0;       // 0 = choose (taps per weight barrier)
  int cta2 = -1;         // CTA pairs: -1 = choose, 0 = never, 1 = whenever the layer allows it
  int max_ctas = 148;
  int narrow_small = 1;  // halve the N tile (down to 64) while the layer has fewer tiles than half the SMs
};

// `act` = A tensor [B, L_in, C_in_ld] (E); `w` = packed weights [w_rows_alloc, n_taps*C_in_ld] (E).
// Output tensors are [B, ep.L_out, ep.C_out] with channel pitch `c_pitch_out` elements (== C_out except conv_post).
const char* make_conv_tc2_launch(ConvTc2Launch* out, int elem_bytes, const void* act, const void* w, int w_rows_alloc,
                                 const ConvGeom& g, const EpiParams& ep, int c_pitch_out, const ConvTc2Options& opt);
cudaError_t launch_conv_tc2(const ConvTc2Launch& L, const int* lengths, cudaStream_t st);
cudaError_t conv_tc2_init();
const char* conv_tc2_cluster_probe(int smem_bytes, int grid, int* max_clusters);

}  // namespace gnv
