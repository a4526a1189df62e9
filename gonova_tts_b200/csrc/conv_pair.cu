// Host side of the fused ResBlock-pair kernel (conv_pair.cuh).
#include "conv_pair.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gnv {

static constexpr size_t kMaxDynSmemPair = 227 * 1024;

cudaError_t conv_pair_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  const auto set = [](auto kernel) {
    preload_kernel((const void*)kernel);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemPair);
  };
  if ((e = set(conv_pair_kernel<__nv_bfloat16, false, false>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<__nv_bfloat16, false, true>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<float, false, false>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<float, false, true>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<__nv_bfloat16, true, false>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<__nv_bfloat16, true, true>)) != cudaSuccess) return e;
  if ((e = set(conv_pair_kernel<float, true, false>)) != cudaSuccess) return e;
  return set(conv_pair_kernel<float, true, true>);
}

namespace {
inline uint32_t up1024(uint32_t x) { return (x + 1023u) & ~1023u; }

const char* encode_rows_map(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int elem_bytes, int C, int L, int B,
                            int box_rows) {
  if (!base) return "conv_pair: epilogue tensor is NULL";
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (((uintptr_t)base & 15) || ((size_t)C * elem_bytes) % 16) return "conv_pair: epilogue tensor is not 16-byte aligned";
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * elem_bytes, (cuuint64_t)L * C * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)kEpiCols, (cuuint32_t)box_rows, 1u};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (kEpiCols * elem_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? "" : "cuTensorMapEncodeTiled failed for an epilogue tensor";
}
}  // namespace

const char* make_conv_pair_launch(ConvPairLaunch* out, int elem_bytes, const void* x, const void* w1, const void* w2,
                                  int B, int L, int C, int C_ld, int k, int d1, const float* bias1,
                                  const float* alpha_mid, int mid_kind, const EpiParams& ep, int max_ctas, int mh_opt,
                                  int cta2_opt) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  const int kbe = 128 / elem_bytes;
  if (C != C_ld || C % kbe) return "conv_pair: channels must fill whole 128-byte K blocks";
  if (C % kEpiCols || C > 128) return "conv_pair: C must be a multiple of 32 and <= 128";
  if (!(k & 1) || k < 1 || k > 11) return "conv_pair: odd kernel sizes up to 11";
  if (mid_kind != ACT_SNAKE_FAST && mid_kind != ACT_SNAKE) return "conv_pair: the middle activation must be Snake";
  if (ep.up != 1 || ep.shift != 0 || ep.dup_row >= 0) return "conv_pair: plain conv epilogue only";
  if (ep.C_pitch != C) return "conv_pair: dense channel pitch only";
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  ConvPairParams& p = out->p;
  p.ep = ep;
  p.B = B; p.L = L; p.C = C; p.k = k; p.d1 = d1;
  p.p1 = d1 * (k - 1) / 2; p.p2 = (k - 1) / 2;
  p.n_chunks = C / kbe;
  p.bias1 = bias1; p.alpha_mid = alpha_mid; p.mid_kind = mid_kind;

  int mh = mh_opt;
  if (mh != 1 && mh != 2) {
    mh = 4 * 2 * C <= 512 ? 2 : 1;
    const long tiles2 = (long)B * ((L + (256 - (k - 1)) - 1) / (256 - (k - 1)));
    if (tiles2 < 2L * max_ctas) mh = 1;
  }
  if (4 * mh * C > 512) mh = 1;
  if (4 * mh * C > 512) return "conv_pair: accumulators do not fit TMEM";
  p.mh = mh;
  p.Mo = 128 * mh - (k - 1);
  const int cta_tiles = (L + p.Mo - 1) / p.Mo;
  // CTA pairs share every weight tile and halve the B-operand reads; they need enough tiles to fill 74 pairs
  const bool cta2 = cta2_opt == 1 || (cta2_opt != 0 && (long)B * ((cta_tiles + 1) / 2) >= 2L * (max_ctas / 2));
  p.cta2 = cta2 ? 1 : 0;
  p.mma_order = getenv("GONOVA_MMA_ORDER") ? atoi(getenv("GONOVA_MMA_ORDER")) : 0;
  p.dbg = getenv("GONOVA_PAIR_DBG") ? atoi(getenv("GONOVA_PAIR_DBG")) : 0;
  p.slabs_first = getenv("GONOVA_PAIR_SLABS_FIRST") ? atoi(getenv("GONOVA_PAIR_SLABS_FIRST")) : 0;   // measured: inside run-to-run noise (+-3 %)
  p.tiles_m = cta2 ? (cta_tiles + 1) / 2 : cta_tiles;
  p.total_tiles = B * p.tiles_m;
  const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | (((cta2 ? 256u : 128u) >> 4) << 24);

  const int slab_rows = 128 * mh + (k - 1) * d1;
  if (slab_rows <= 256) {
    p.a_n_boxes = 1; p.a_box_rows = slab_rows;
  } else {
    p.a_n_boxes = 2; p.a_box_rows = (((slab_rows + 1) / 2) + 7) & ~7;
    if (p.a_box_rows > 256) return "conv_pair: x slab taller than two TMA boxes";
  }
  p.slab_bytes = (int)up1024((uint32_t)(p.a_n_boxes * p.a_box_rows) * 128u);
  p.w_bytes = (cta2 ? C / 2 : C) * 128;
  p.w_group = std::max(1, std::min(4, 32768 / p.w_bytes));
  if (const char* v = getenv("GONOVA_PAIR_WGROUP")) p.w_group = std::max(1, std::min(4, atoi(v)));
  p.w_slot_bytes = p.w_group * p.w_bytes;
  p.h_kb_bytes = 128 * mh * 128;
  // conv2's taps read up to k-1 rows past the slab's last K block: keep that inside the allocation
  const uint32_t h_bytes = up1024((uint32_t)p.n_chunks * p.h_kb_bytes + (uint32_t)(k - 1) * 128u);

  p.n_in = (ep.res ? 1 : 0) + (ep.raw_accum ? 1 : 0);
  p.has_raw = ep.raw ? 1 : 0;
  p.n_act = ep.n_act;
  if (ep.raw_accum && !ep.raw) return "conv_pair: raw_accum needs a raw output";
  if (!p.has_raw && p.n_act == 0) return "conv_pair: layer has no output";
  p.act_bytes = 128 * kEpiCols * elem_bytes;
  p.c_tab = (C + 31) & ~31;
  const uint32_t in_slot_bytes = (uint32_t)p.n_in * (128 * kEpiCols * 4);
  int in_slots = 2;                      // total epilogue-input slots (warpgroups x ring depth)
  const uint32_t out_buf = (uint32_t)(p.has_raw ? 128 * kEpiCols * 4 : 0) + (uint32_t)p.n_act * p.act_bytes;
  const uint32_t tab_bytes = up1024((uint32_t)(1 + 2 * p.n_act + 3) * p.c_tab * 4);
  const uint32_t bar_bytes = 1024;

  int sa = 1, sw = 2, nob = 1;
  auto total = [&](int sa_, int sw_, int nob_) {
    return (size_t)sa_ * p.slab_bytes + (size_t)sw_ * p.w_slot_bytes + h_bytes + (size_t)in_slots * in_slot_bytes + (size_t)nob_ * out_buf + tab_bytes +
           bar_bytes + 1024;
  };
  while (total(sa, sw, nob) > kMaxDynSmemPair && p.w_group > 1) {       // smaller weight ring slots first
    p.w_group >>= 1;
    p.w_slot_bytes = p.w_group * p.w_bytes;
  }
  if (total(sa, sw, nob) > kMaxDynSmemPair) return "conv_pair: shared memory budget exceeded";
  const int max_sa = std::min(3, p.n_chunks + 1);
  const int max_sw = 6;
  bool grew = true;
  while (grew) {
    grew = false;
    if (nob < 2 && total(sa, sw, nob + 1) <= kMaxDynSmemPair) { ++nob; grew = true; }
    if (sw < 3 && total(sa, sw + 1, nob) <= kMaxDynSmemPair) { ++sw; grew = true; }
    if (sa < 2 && sa < max_sa && total(sa + 1, sw, nob) <= kMaxDynSmemPair) { ++sa; grew = true; }
    if (nob == 2 && total(sa, sw, 4) <= kMaxDynSmemPair) { nob = 4; grew = true; }
    if (!grew && sw < max_sw && total(sa, sw + 1, nob) <= kMaxDynSmemPair) { ++sw; grew = true; }
    if (!grew && sa < max_sa && total(sa + 1, sw, nob) <= kMaxDynSmemPair) { ++sa; grew = true; }
  }
  // leftover shared memory: a second input slot per warpgroup (measured: -10 % on HBM-bound conv2 layers; it must not
  // take space from the weight ring, which costs the tensor-bound layers more)
  if (p.n_in > 0 && in_slots_cap() >= 4 && nob >= 2) {
    in_slots = 4;
    if (total(sa, sw, nob) > kMaxDynSmemPair) in_slots = 2;
  }
  p.sa = sa; p.sw = sw; p.n_epi_wg = nob >= 2 ? 2 : 1; p.out_bufs = nob == 4 ? 2 : 1;
  p.in_ring = in_slots / p.n_epi_wg;
  if (getenv("GONOVA_PAIR_DEBUG"))
    fprintf(stderr, "[gonova] pair C=%d k=%d mh=%d: x slabs %d x %d B, W ring %d x %d taps (%d B), staging bufs %d, in slots %d, total %zu B\n",
            C, k, mh, sa, p.slab_bytes, sw, p.w_group, p.w_slot_bytes, nob, in_slots, total(sa, sw, nob));
  uint32_t off = 0;
  p.off_a = off; off += (uint32_t)sa * p.slab_bytes;
  p.off_w = off; off += (uint32_t)sw * p.w_slot_bytes;
  p.off_h = off; off += h_bytes;
  p.off_in = off; off += (uint32_t)in_slots * in_slot_bytes;
  p.off_out = off; off += (uint32_t)nob * out_buf;
  off = up1024(off);
  p.off_tab = off; off += tab_bytes;
  p.off_bar = off; off += bar_bytes;
  out->smem_bytes = (size_t)off + 1024;
  if (out->smem_bytes > kMaxDynSmemPair) return "conv_pair: shared memory budget exceeded";
  if (8 * (2 * sa + 2 * sw + 13) + 16 > (int)bar_bytes) return "conv_pair: barrier area too small";

  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (((uintptr_t)x & 15) || ((uintptr_t)w1 & 15) || ((uintptr_t)w2 & 15)) return "conv_pair: operand pointers must be 16-byte aligned";
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * elem_bytes, (cuuint64_t)L * C * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)kbe, (cuuint32_t)p.a_box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&out->maps.X, dt, 3, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation tensor";
  }
  for (int i = 0; i < 2; ++i) {
    const int K = k * C;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)C};
    cuuint64_t strides[1] = {(cuuint64_t)K * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)(cta2 ? C / 2 : C)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(i == 0 ? &out->maps.W1 : &out->maps.W2, dt, 2, const_cast<void*>(i == 0 ? w1 : w2), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for a weight tensor";
  }
  for (int v = 0; v < 2; ++v) {
    const int box_rows = v == 0 ? 32 : 32 - (k - 1);      // stores are per warp (32 rows); loads stay 128 rows
    const char* e = "";
    if (ep.res) e = encode_rows_map(enc, &out->maps.epi[v][EPI_IN0], ep.res, 4, C, L, B, 128);
    if (*e) return e;
    if (ep.raw_accum) e = encode_rows_map(enc, &out->maps.epi[v][EPI_IN0 + (ep.res ? 1 : 0)], ep.raw, 4, C, L, B, 128);
    if (*e) return e;
    if (ep.raw) e = encode_rows_map(enc, &out->maps.epi[v][EPI_RAW], ep.raw, 4, C, L, B, box_rows);
    if (*e) return e;
    for (int a = 0; a < ep.n_act; ++a) {
      e = encode_rows_map(enc, &out->maps.epi[v][EPI_ACT0 + a], ep.act_out[a], elem_bytes, C, L, B, box_rows);
      if (*e) return e;
    }
  }
  out->grid = cta2 ? 2 * std::max(1, std::min(p.total_tiles, max_ctas / 2)) : std::max(1, std::min(p.total_tiles, max_ctas));
  out->elem_bytes = elem_bytes;
  return "";
}

cudaError_t launch_conv_pair(const ConvPairLaunch& L, const int* lengths, cudaStream_t st) {
  if (!L.d_maps) return cudaErrorInvalidValue;
  ConvPairParams p = L.p;
  p.ep.lengths = lengths;
  const ConvPairMaps* dm = L.d_maps;
  // the RAGGED instantiation (dead tiles skipped) only when the batch carries lengths: the dense walk stays as it was
  const auto go = [&](auto dense, auto ragged, bool pair) {
    return lengths ? launch_persistent(ragged, L.grid, L.smem_bytes, st, pair, 384, dm, p)
                   : launch_persistent(dense, L.grid, L.smem_bytes, st, pair, 384, dm, p);
  };
  if (!p.cta2) {
    if (L.elem_bytes == 2) return go(conv_pair_kernel<__nv_bfloat16, false, false>, conv_pair_kernel<__nv_bfloat16, false, true>, false);
    return go(conv_pair_kernel<float, false, false>, conv_pair_kernel<float, false, true>, false);
  }
  if (L.elem_bytes == 2) return go(conv_pair_kernel<__nv_bfloat16, true, false>, conv_pair_kernel<__nv_bfloat16, true, true>, true);
  return go(conv_pair_kernel<float, true, false>, conv_pair_kernel<float, true, true>, true);
}

}  // namespace gnv
