// Host side of the persistent tcgen05 conv kernel (conv_tc2.cuh): tile / slab / shared-memory
// planning and tensor-map encoding for one layer.
#include "conv_tc2.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace gnv {

static constexpr size_t kMaxDynSmem2 = 227 * 1024;

int in_slots_cap() {      // GONOVA_IN_SLOTS: total epilogue-input slots (2 = no prefetch ahead of a warpgroup)
  static const int cap = [] {
    const char* v = getenv("GONOVA_IN_SLOTS");
    int c = v ? atoi(v) : 6;
    return c < 2 ? 2 : (c > kMaxInSlots ? kMaxInSlots : c);
  }();
  return cap;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("GONOVA_PDL");
    return !(v && atoi(v) == 0);
  }();
  return on;
}

cudaError_t conv_tc2_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  const auto set = [](auto kernel) {
    preload_kernel((const void*)kernel);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem2);
  };
  if ((e = set(conv_tc2_kernel<__nv_bfloat16, false, false>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<__nv_bfloat16, false, true>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<float, false, false>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<float, false, true>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<__nv_bfloat16, true, false>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<__nv_bfloat16, true, true>)) != cudaSuccess) return e;
  if ((e = set(conv_tc2_kernel<float, true, false>)) != cudaSuccess) return e;
  return set(conv_tc2_kernel<float, true, true>);
}

// Diagnostics: how many 2-CTA clusters of conv_tc2_kernel<bf16> can be resident with `smem_bytes` of dynamic
// shared memory per CTA (0 or an error text means cluster launches of that configuration are refused).
const char* conv_tc2_cluster_probe(int smem_bytes, int grid, int* max_clusters) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  *max_clusters = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(max_clusters, conv_tc2_kernel<__nv_bfloat16, true, false>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
  return "";
}

namespace {

inline uint32_t up1024(uint32_t x) { return (x + 1023u) & ~1023u; }

// [B, rows, C] view of an epilogue tensor for one polyphase `phase` (or the whole tensor when up == 1):
// coordinate j of dim 1 is output row  base_row + j * up.
const char* encode_epi_map(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int elem_bytes, int C_valid,
                           int C_pitch, int L_out, int B, int up, int base_row, int n_rows, int box_rows) {
  if (!base) return "conv_tc2: epilogue tensor is NULL";
  if (n_rows <= 0) n_rows = 1;   // degenerate phase: every box is clipped (coordinates never reach it)
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const char* p = static_cast<const char*>(base) + (size_t)base_row * C_pitch * elem_bytes;
  if (((uintptr_t)p & 15) || ((size_t)C_pitch * elem_bytes) % 16) return "conv_tc2: epilogue tensor is not 16-byte aligned";
  cuuint64_t dims[3] = {(cuuint64_t)C_valid, (cuuint64_t)n_rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)up * C_pitch * elem_bytes, (cuuint64_t)L_out * C_pitch * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)kEpiCols, (cuuint32_t)box_rows, 1u};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (kEpiCols * elem_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, dt, 3, const_cast<char*>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? "" : "cuTensorMapEncodeTiled failed for an epilogue tensor";
}

}  // namespace

const char* make_conv_tc2_launch(ConvTc2Launch* out, int elem_bytes, const void* act, const void* w, int w_rows_alloc,
                                 const ConvGeom& g, const EpiParams& ep, int c_pitch_out, const ConvTc2Options& opt) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  const int kbe = 128 / elem_bytes;
  if (g.C_in_ld % kbe) return "conv_tc2: channel stride is not a multiple of the 128-byte K block";
  if (g.C_in_w != g.C_in_ld) return "conv_tc2: weight K stride must equal the activation channel stride";
  if (g.in_stride != 1) return "conv_tc2: strided input rows are not supported on the tensor-core path";
  if (g.n_taps < 1 || g.n_taps > kMaxSlab) return "conv_tc2: too many taps";
  if (((uintptr_t)act & 15) || ((uintptr_t)w & 15)) return "conv_tc2: operand pointers must be 16-byte aligned";
  const bool transposed = ep.up > 1;
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  ConvTc2Params& p = out->p;
  p.ep = ep;
  p.ep.C_pitch = c_pitch_out;
  p.B = g.B; p.M_rows = g.M_rows; p.n_taps = g.n_taps;
  p.n_chunks = g.C_in_ld / kbe;
  p.transposed = transposed ? 1 : 0;

  // ---- N tiling ----
  int block_n;
  if (transposed) {
    block_n = ep.C_out;
    if (g.N_total != ep.up * ep.C_out) return "conv_tc2: polyphase GEMM must have N == up * C_out";
    if (ep.up > kMaxPhase) return "conv_tc2: too many polyphase phases";
  } else {
    block_n = g.N_total <= 256 ? g.N_total : 256;
    // Small problems (one sentence, a first chunk): narrower tiles, so that the weight stream and the k loop of a
    // layer spread over more SMs instead of 2-8 CTAs walking all of K alone (latency, not throughput, is the metric)
    long tiles = (long)g.B * ((g.M_rows + 127) / 128) * (g.N_total / block_n);
    while (opt.narrow_small && block_n >= 128 && tiles * 2 <= opt.max_ctas && g.N_total % (block_n / 2) == 0) {
      block_n /= 2;
      tiles *= 2;
    }
  }
  if (block_n % kEpiCols || block_n < 32 || block_n > 256) return "conv_tc2: N tile must be a multiple of 32 in [32,256]";
  if (g.N_total % block_n) return "conv_tc2: N_total must be a multiple of the N tile";
  if (w_rows_alloc < g.N_total) return "conv_tc2: packed weights have fewer rows than N_total";
  p.block_n = block_n;
  p.n_tiles_n = g.N_total / block_n;

  // ---- CTA pairs (tcgen05.mma.cta_group::2): M = 256 across two SMs, each CTA stages half of every weight tile.
  // Halves the weight traffic L2 -> SMEM and the B-operand shared-memory reads per MMA; worth it where the
  // layer is tensor / operand-feed bound (N >= 128) and there are enough tiles to fill 74 pairs twice.
  bool cta2 = false;
  if (!transposed && block_n >= 128 && opt.cta2 != 0) {
    const long pair_tiles = (long)g.B * ((g.M_rows + 255) / 256) * p.n_tiles_n;
    // measured (B=64, T=500, bf16): pairs win on tensor / operand-feed bound layers (N = 256: +13 %; N = 128 with
    // K >= 896 and no residual traffic: +10 %; N = 128 with K >= 1408 even with the residual: +9 %) and lose 5-10 %
    // on the HBM-bound ones (N = 128 conv2 at k <= 7, k = 3, source_downs)
    const int K_total = g.n_taps * g.C_in;
    const bool tensor_bound = block_n >= 256 || (K_total >= 896 && !ep.res && !ep.raw_accum) || K_total >= 1408;
    cta2 = opt.cta2 == 1 || (pair_tiles >= 2L * (opt.max_ctas / 2) && tensor_bound);
  }
  p.cta2 = cta2 ? 1 : 0;

  // ---- M tiling: two 128-row accumulators per tile when they fit TMEM twice and there is enough work ----
  int mh = opt.mh;
  if (mh != 1 && mh != 2) {
    mh = block_n <= 128 ? 2 : 1;
    const long tiles2 = (long)g.B * ((g.M_rows + 255) / 256) * p.n_tiles_n;
    if (tiles2 < 2L * opt.max_ctas) mh = 1;
    if (cta2) {
      const long tiles4 = (long)g.B * ((g.M_rows + 511) / 512) * p.n_tiles_n;
      if (tiles4 < 2L * (opt.max_ctas / 2)) mh = 1;
    }
  }
  if (mh * block_n > 512) mh = 1;
  p.mh = mh;
  const int rows_tile = 128 * mh * (cta2 ? 2 : 1);
  p.tiles_m = (g.M_rows + rows_tile - 1) / rows_tile;
  p.total_tiles = g.B * p.tiles_m * p.n_tiles_n;
  p.acc_bufs = std::min(2, 512 / (mh * block_n));
  int cols = 32;
  while (cols < p.acc_bufs * mh * block_n) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;   // BF16 : TF32
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(block_n >> 3) << 17) |
            (((cta2 ? 256u : 128u) >> 4) << 24);

  // ---- A slabs ----
  int tmin = 0, tmax = 0;
  for (int j = 0; j < g.n_taps; ++j) {
    const int t = g.off0 + j * g.tap_step;
    p.tap_row[j] = t;
    if (j == 0 || t < tmin) tmin = t;
    if (j == 0 || t > tmax) tmax = t;
  }
  int slab_rows;
  if (opt.slab_mode == 0 || g.n_taps == 1) {
    p.n_slabs = g.n_taps;
    for (int j = 0; j < g.n_taps; ++j) { p.slab_tap0[j] = j; p.slab_row0[j] = p.tap_row[j]; }
    p.slab_tap0[g.n_taps] = g.n_taps;
    slab_rows = 128 * mh;
    p.a_base_offset_mode = 0;
  } else {
    p.n_slabs = 1;
    p.slab_tap0[0] = 0; p.slab_tap0[1] = g.n_taps;
    p.slab_row0[0] = tmin;
    slab_rows = 128 * mh + (tmax - tmin);
    p.a_base_offset_mode = opt.slab_mode == 2 ? 1 : 0;
  }
  if (slab_rows <= 256) {
    p.a_n_boxes = 1; p.a_box_rows = slab_rows;
  } else {
    p.a_n_boxes = 2; p.a_box_rows = (((slab_rows + 1) / 2) + 7) & ~7;
    if (p.a_box_rows > 256) return "conv_tc2: A slab taller than two TMA boxes";
  }
  p.slab_bytes = (int)up1024((uint32_t)(p.a_n_boxes * p.a_box_rows) * 128u);
  p.w_bytes = (cta2 ? block_n / 2 : block_n) * 128;      // CTA pair: this CTA's half of the weight tile
  p.w_group = std::max(1, std::min(4, 32768 / p.w_bytes));
  if (opt.w_group > 0) p.w_group = std::min(opt.w_group, 4);
  p.w_slot_bytes = p.w_group * p.w_bytes;

  // ---- epilogue staging ----
  p.n_in = (ep.res ? 1 : 0) + (ep.raw_accum ? 1 : 0);
  p.has_raw = ep.raw ? 1 : 0;
  p.n_act = ep.n_act;
  if (ep.raw_accum && !ep.raw) return "conv_tc2: raw_accum needs a raw output";
  if (!p.has_raw && p.n_act == 0) return "conv_tc2: layer has no output";
  p.act_bytes = 128 * kEpiCols * elem_bytes;
  p.c_tab = (ep.C_out + 31) & ~31;
  const uint32_t in_slot_bytes = (uint32_t)p.n_in * (128 * kEpiCols * 4);
  int in_slots = 2;                      // total epilogue-input slots (warpgroups x ring depth)
  const uint32_t out_buf = (uint32_t)(p.has_raw ? 128 * kEpiCols * 4 : 0) + (uint32_t)p.n_act * p.act_bytes;
  const uint32_t tab_bytes = up1024((uint32_t)(1 + 2 * p.n_act) * p.c_tab * 4);
  const uint32_t bar_bytes = 1024;

  // ---- shared-memory budget: grow the rings while they fit ----
  const int k_iters = p.n_chunks * g.n_taps;
  int sa = 1, sw = 2, nob = 1;          // nob = staging buffers in total = warpgroups * buffers per warpgroup
  auto total = [&](int sa_, int sw_, int nob_) {
    return (size_t)sa_ * p.slab_bytes + (size_t)sw_ * p.w_slot_bytes + (size_t)in_slots * in_slot_bytes + (size_t)nob_ * out_buf + tab_bytes +
           bar_bytes + 1024 /*alignment slack*/;
  };
  if (total(sa, sw, nob) > kMaxDynSmem2) return "conv_tc2: shared memory budget exceeded";
  const int max_sa = std::min(p.n_slabs == 1 ? 3 : 6, p.n_chunks * p.n_slabs + 1);
  const int max_sw = std::min(6, (k_iters + p.w_group - 1) / p.w_group + 1);
  bool grew = true;
  while (grew) {
    grew = false;
    if (nob < 2 && total(sa, sw, nob + 1) <= kMaxDynSmem2) { ++nob; grew = true; }
    if (sa < 2 && sa < max_sa && total(sa + 1, sw, nob) <= kMaxDynSmem2) { ++sa; grew = true; }
    if (sw < 3 && sw < max_sw && total(sa, sw + 1, nob) <= kMaxDynSmem2) { ++sw; grew = true; }
    if (nob == 2 && total(sa, sw, 4) <= kMaxDynSmem2) { nob = 4; grew = true; }
    if (!grew && sw < max_sw && total(sa, sw + 1, nob) <= kMaxDynSmem2) { ++sw; grew = true; }
    if (!grew && sa < max_sa && total(sa + 1, sw, nob) <= kMaxDynSmem2) { ++sa; grew = true; }
  }
  // leftover shared memory: a second input slot per warpgroup (measured: -10 % on HBM-bound conv2 layers; it must not
  // take space from the weight ring, which costs the tensor-bound layers more)
  if (p.n_in > 0 && in_slots_cap() >= 4 && nob >= 2) {
    in_slots = 4;
    if (total(sa, sw, nob) > kMaxDynSmem2) in_slots = 2;
  }
  p.sa = sa; p.sw = sw; p.n_epi_wg = nob >= 2 ? 2 : 1; p.out_bufs = nob == 4 ? 2 : 1;
  p.in_ring = in_slots / p.n_epi_wg;
  uint32_t off = 0;
  p.off_a = off; off += (uint32_t)sa * p.slab_bytes;
  p.off_w = off; off += (uint32_t)sw * p.w_slot_bytes;
  p.off_in = off; off += (uint32_t)in_slots * in_slot_bytes;
  p.off_out = off; off += (uint32_t)nob * out_buf;
  off = up1024(off);
  p.off_tab = off; off += tab_bytes;
  p.off_bar = off; off += bar_bytes;
  out->smem_bytes = (size_t)off + 1024;
  if (out->smem_bytes > kMaxDynSmem2) return "conv_tc2: shared memory budget exceeded";
  if (8 * (2 * sa + 2 * sw + 8) + 16 > (int)bar_bytes) return "conv_tc2: barrier area too small";

  // ---- tensor maps ----
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    const long long rs = g.a_row_stride ? g.a_row_stride : g.C_in_ld;
    const long long bs = g.a_batch_stride ? g.a_batch_stride : (long long)g.L_in * g.C_in_ld;
    if ((rs * elem_bytes) % 16 || (bs * elem_bytes) % 16) return "conv_tc2: A view strides must be multiples of 16 bytes";
    cuuint64_t dims[3] = {(cuuint64_t)g.C_in_ld, (cuuint64_t)g.L_in, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)rs * elem_bytes, (cuuint64_t)bs * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)kbe, (cuuint32_t)p.a_box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&out->maps.A, dt, 3, const_cast<void*>(act), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation tensor";
  }
  {
    const int K = g.n_taps * g.C_in_ld;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)w_rows_alloc};
    cuuint64_t strides[1] = {(cuuint64_t)K * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)(cta2 ? block_n / 2 : block_n)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&out->maps.W, dt, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the weight tensor";
  }
  const int n_phase = transposed ? ep.up : 1;
  for (int ph = 0; ph < n_phase; ++ph) {
    int base_row = 0, n_rows = ep.L_out, up = 1;
    p.row_adj[ph] = 0;
    if (transposed) {
      // phase r: p0 = m*up + r - pad; the first valid m is 0 when r >= pad, else 1
      const int adj = ph < ep.pad_out ? 1 : 0;
      const int first_p0 = adj * ep.up + ph - ep.pad_out;
      p.row_adj[ph] = adj;
      up = ep.up;
      base_row = first_p0 + ep.shift;
      n_rows = first_p0 <= ep.L_store - 1 ? (ep.L_store - 1 - first_p0) / ep.up + 1 : 0;
    }
    const char* e = "";
    if (ep.res) e = encode_epi_map(enc, &out->maps.epi[ph][EPI_IN0], ep.res, 4, ep.C_out, c_pitch_out, ep.L_out, g.B, up, base_row, n_rows, 128);
    if (*e) return e;
    if (ep.raw_accum) e = encode_epi_map(enc, &out->maps.epi[ph][EPI_IN0 + (ep.res ? 1 : 0)], ep.raw, 4, ep.C_out, c_pitch_out, ep.L_out, g.B, up, base_row, n_rows, 128);
    if (*e) return e;
    if (ep.raw) e = encode_epi_map(enc, &out->maps.epi[ph][EPI_RAW], ep.raw, 4, ep.C_out, c_pitch_out, ep.L_out, g.B, up, base_row, n_rows, 32);
    if (*e) return e;
    for (int a = 0; a < ep.n_act; ++a) {
      e = encode_epi_map(enc, &out->maps.epi[ph][EPI_ACT0 + a], ep.act_out[a], elem_bytes, ep.C_out, c_pitch_out, ep.L_out, g.B, up, base_row, n_rows, 32);
      if (*e) return e;
    }
  }
  out->grid = std::max(1, std::min(p.total_tiles, opt.max_ctas));
  if (cta2) out->grid = 2 * std::max(1, std::min(p.total_tiles, opt.max_ctas / 2));
  out->elem_bytes = elem_bytes;
  return "";
}

cudaError_t launch_conv_tc2(const ConvTc2Launch& L, const int* lengths, cudaStream_t st) {
  if (!L.d_maps) return cudaErrorInvalidValue;
  ConvTc2Params p = L.p;
  p.ep.lengths = lengths;
  const ConvTc2Maps* dm = L.d_maps;
  // the RAGGED instantiation (dead tiles skipped) only when the batch carries lengths: the dense walk stays as it was
  const auto go = [&](auto dense, auto ragged, bool pair) {
    return lengths ? launch_persistent(ragged, L.grid, L.smem_bytes, st, pair, 384, dm, p)
                   : launch_persistent(dense, L.grid, L.smem_bytes, st, pair, 384, dm, p);
  };
  if (!p.cta2) {
    if (L.elem_bytes == 2) return go(conv_tc2_kernel<__nv_bfloat16, false, false>, conv_tc2_kernel<__nv_bfloat16, false, true>, false);
    return go(conv_tc2_kernel<float, false, false>, conv_tc2_kernel<float, false, true>, false);
  }
  if (L.elem_bytes == 2) return go(conv_tc2_kernel<__nv_bfloat16, true, false>, conv_tc2_kernel<__nv_bfloat16, true, true>, true);
  return go(conv_tc2_kernel<float, true, false>, conv_tc2_kernel<float, true, true>, true);
}

}  // namespace gnv
