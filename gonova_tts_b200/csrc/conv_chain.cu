// Host side of the whole-ResBlock kernel (conv_chain.cuh): tile / shared-memory planning and tensor maps.
#include "conv_chain.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gnv {

static constexpr size_t kMaxDynSmemChain = 227 * 1024;

cudaError_t conv_chain_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  const auto set = [](auto kernel) {
    preload_kernel((const void*)kernel);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmemChain);
  };
  if ((e = set(conv_chain_kernel<__nv_bfloat16, 64, false>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<__nv_bfloat16, 64, true>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<__nv_bfloat16, 128, false>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<__nv_bfloat16, 128, true>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<float, 64, false>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<float, 64, true>)) != cudaSuccess) return e;
  if ((e = set(conv_chain_kernel<float, 128, false>)) != cudaSuccess) return e;
  return set(conv_chain_kernel<float, 128, true>);
}

int conv_chain_read_trace(unsigned long long* out, int cap) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  static unsigned long long host[kChainTraceCap];
  if (cudaMemcpyFromSymbol(host, g_chain_trace, sizeof(host)) != cudaSuccess) return -1;
  int n = 0;
  for (int i = 0; i < kChainTraceCap && n < cap; ++i)
    if (host[i]) out[n++] = host[i];
  memset(host, 0, sizeof(host));
  cudaMemcpyToSymbol(g_chain_trace, host, sizeof(host));
  return n;
}

namespace {
inline uint32_t up1024(uint32_t x) { return (x + 1023u) & ~1023u; }

const char* encode_rows(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int C, int L, int B, int box_rows) {
  if (!base) return "conv_chain: tensor is NULL";
  if (((uintptr_t)base & 15) || ((size_t)C * 4) % 16) return "conv_chain: tensor is not 16-byte aligned";
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)L * C * 4};
  cuuint32_t box[3] = {(cuuint32_t)kEpiCols, (cuuint32_t)box_rows, 1u};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? "" : "cuTensorMapEncodeTiled failed for a residual-stream tensor";
}
}  // namespace

const char* make_conv_chain_launch(ConvChainLaunch* out, int elem_bytes, const float* in, float* out_raw,
                                   const ConvChainSpec& spec, int B, int L, int C, int k, float out_scale, int out_accum,
                                   int snake_kind, int round_tf32, int len_mul, int len_add, int max_ctas) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  const int kbe = 128 / elem_bytes;
  if (C != 64 && C != 128) return "conv_chain: built for C = 64 and C = 128";
  if (!(k & 1) || k < 3 || k > 7) return "conv_chain: kernel sizes 3, 5, 7";
  if (snake_kind != ACT_SNAKE_FAST && snake_kind != ACT_SNAKE) return "conv_chain: activations must be Snake";
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  ConvChainParams& p = out->p;
  p.B = B; p.L = L; p.C = C; p.k = k;
  p.n_chunks = C / kbe;
  int halo = 0, dmax = 1;
  for (int d = 0; d < 3; ++d) {
    p.dil[d] = spec.dil[d];
    halo += (spec.dil[d] + 1) * (k - 1) / 2;
    dmax = std::max(dmax, spec.dil[d]);
    p.bias1[d] = spec.bias1[d]; p.bias2[d] = spec.bias2[d];
    p.alpha1[d] = spec.alpha1[d]; p.alpha2[d] = spec.alpha2[d];
  }
  p.halo = halo;
  p.margin = (dmax * (k - 1) / 2 + 7) & ~7;
  p.Mo = 256 - 2 * halo;
  if (p.Mo < 128) return "conv_chain: the receptive field leaves too few output rows per tile";
  p.tiles_m = (L + p.Mo - 1) / p.Mo;
  p.total_tiles = B * p.tiles_m;
  p.slab_rows = 256 + 2 * p.margin;
  p.slab_kb_bytes = p.slab_rows * 128;
  p.slab_bytes = (int)up1024((uint32_t)(p.n_chunks * p.slab_kb_bytes));
  p.mid_kind = snake_kind; p.round_tf32 = round_tf32; p.out_scale = out_scale;
  p.len_mul = len_mul; p.len_add = len_add;
  p.c_tab = (C + 31) & ~31;
  const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
  // two tiles in flight when both lanes' accumulators (2 * 2*mh*C columns) fit TMEM and both slabs fit shared memory
  p.lanes = (8 * C <= 512) ? 2 : 1;
  if (const char* v = getenv("GONOVA_CHAIN_LANES")) p.lanes = std::max(1, std::min(p.lanes, atoi(v)));
  p.w_bytes = C * 128;
  p.w_group = std::max(1, std::min(4, std::min(k, 32768 / p.w_bytes)));
  p.w_slot_bytes = p.w_group * p.w_bytes;
  p.in_ring = 1;          // F2 chunk slots per warpgroup (C = 64: exactly one tile ahead)
  p.out = out_raw;
  p.out_accum = out_accum ? 1 : 0;
  p.dbg = getenv("GONOVA_CHAIN_DBG") ? atoi(getenv("GONOVA_CHAIN_DBG")) : 0;
  const uint32_t tab_bytes = up1024((uint32_t)18 * p.c_tab * 4);
  const uint32_t bar_bytes = 1024;
  const uint32_t in_bytes = (uint32_t)kChainEpiWg * p.in_ring * (128 * kEpiCols * 4);
  auto total = [&](int lanes, int sw) {
    return (size_t)lanes * p.slab_bytes + (size_t)sw * p.w_slot_bytes + in_bytes + tab_bytes + bar_bytes + 1024;
  };
  if (total(p.lanes, 2) > kMaxDynSmemChain) p.lanes = 1;
  if (total(p.lanes, 2) > kMaxDynSmemChain) return "conv_chain: shared memory budget exceeded";
  int sw = 2;
  while (sw < 6 && total(p.lanes, sw + 1) <= kMaxDynSmemChain) ++sw;
  p.sw = sw;
  p.lane_lag = p.lanes > 1 ? 3 : 0;
  if (const char* v = getenv("GONOVA_CHAIN_LAG")) p.lane_lag = std::max(0, std::min(6, atoi(v)));
  if (p.lanes == 1) p.lane_lag = 0;
  uint32_t off = 0;
  p.off_slab = off; off += (uint32_t)p.lanes * p.slab_bytes;
  p.off_w = off; off += (uint32_t)sw * p.w_slot_bytes;
  p.off_in = off; off += in_bytes;
  off = up1024(off);
  p.off_tab = off; off += tab_bytes;
  p.off_bar = off; off += bar_bytes;
  out->smem_bytes = (size_t)off + 1024;
  if (out->smem_bytes > kMaxDynSmemChain) return "conv_chain: shared memory budget exceeded";
  if (8 * (2 * sw + 6 + 2 * kMaxInSlots) + 16 > (int)bar_bytes) return "conv_chain: barrier area too small";
  if (getenv("GONOVA_PAIR_DEBUG"))
    fprintf(stderr, "[gonova] chain C=%d k=%d: lanes %d (lag %d), halo %d, Mo %d, slab %d B, W ring %d x %d taps (%d B), total %zu B\n",
            C, k, p.lanes, p.lane_lag, p.halo, p.Mo, p.slab_bytes, sw, p.w_group, p.w_slot_bytes, out->smem_bytes);

  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  for (int j = 0; j < 6; ++j) {
    const void* w = (j & 1) ? spec.w2[j >> 1] : spec.w1[j >> 1];
    if (!w || ((uintptr_t)w & 15)) return "conv_chain: weight pointers must be 16-byte aligned";
    const int K = k * C;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)C};
    cuuint64_t strides[1] = {(cuuint64_t)K * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)C};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&out->maps.W[j], dt, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for a weight tensor";
  }
  const char* e = encode_rows(enc, &out->maps.IN, in, C, L, B, 128);
  if (*e) return e;
  if (!out_raw || ((uintptr_t)out_raw & 127)) return "conv_chain: the output tensor must be 128-byte aligned";
  // every CTA keeps `lanes` tiles in flight; with fewer tiles than that per SM one lane per CTA spreads them wider
  out->grid = std::max(1, std::min(p.total_tiles, max_ctas));
  out->elem_bytes = elem_bytes;
  return "";
}

cudaError_t launch_conv_chain(const ConvChainLaunch& L, const int* lengths, cudaStream_t st) {
  if (!L.d_maps) return cudaErrorInvalidValue;
  ConvChainParams p = L.p;
  p.lengths = lengths;
  const ConvChainMaps* dm = L.d_maps;
  const auto go = [&](auto dense, auto ragged) {
    return lengths ? launch_persistent(ragged, L.grid, L.smem_bytes, st, false, kChainThreads, dm, p)
                   : launch_persistent(dense, L.grid, L.smem_bytes, st, false, kChainThreads, dm, p);
  };
  if (L.elem_bytes == 2) {
    if (p.C == 64) return go(conv_chain_kernel<__nv_bfloat16, 64, false>, conv_chain_kernel<__nv_bfloat16, 64, true>);
    return go(conv_chain_kernel<__nv_bfloat16, 128, false>, conv_chain_kernel<__nv_bfloat16, 128, true>);
  }
  if (p.C == 64) return go(conv_chain_kernel<float, 64, false>, conv_chain_kernel<float, 64, true>);
  return go(conv_chain_kernel<float, 128, false>, conv_chain_kernel<float, 128, true>);
}

}  // namespace gnv
