// Host side of flow_blk_kernel (flow_blk.cuh): tensor maps, shared-memory carve-up, launch.
#include "flow_blk.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace gnv {

static constexpr size_t kFbMaxDynSmem = 227 * 1024;

cudaError_t flow_blk_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  const auto set = [](auto kernel) {
    preload_kernel((const void*)kernel);
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFbMaxDynSmem);
  };
  if ((e = set(flow_blk_kernel<FB_FF>)) != cudaSuccess) return e;
  if ((e = set(flow_blk_kernel<FB_OUT>)) != cudaSuccess) return e;
  if ((e = set(flow_blk_kernel<FB_CONV>)) != cudaSuccess) return e;
  if ((e = set(flow_blk_kernel<FB_OUTFF>)) != cudaSuccess) return e;
  return set(flow_blk_kernel<FB_WIDE>);
}

int flow_blk_read_trace(unsigned long long* out, int cap) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  static unsigned long long host[kFbTraceCap];
  if (cudaMemcpyFromSymbol(host, g_fb_trace, sizeof(host)) != cudaSuccess) return -1;
  int n = 0;
  for (int i = 0; i < kFbTraceCap && n < cap; ++i)
    if (host[i]) out[n++] = host[i];
  memset(host, 0, sizeof(host));
  cudaMemcpyToSymbol(g_fb_trace, host, sizeof(host));
  return n;
}

namespace {

const char* encode_2d(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int elem_bytes, bool is_float, long long cols,
                      long long rows, long long pitch_elems, int box_cols, int box_rows, CUtensorMapSwizzle sw,
                      CUtensorMapL2promotion l2) {
  if (!base) return "flow_blk: tensor is NULL";
  if (((uintptr_t)base & 15) || (pitch_elems * elem_bytes) % 16) return "flow_blk: tensor is not 16-byte aligned";
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? "" : "flow_blk: cuTensorMapEncodeTiled failed";
}

}  // namespace

const char* make_flow_blk_launch(FlowBlkLaunch* out, int mode, const void* a, int K, const void* w1, const float* b1,
                                 const void* w2, const float* b2, float* r, const float* gamma, const float* beta,
                                 int ln, void* n_out, int n_pitch, int N, int M, int T, int max_ctas, void* vt, int vt_col0,
                                 int vt_tp, const float* r_in) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  if (mode != FB_FF && mode != FB_OUT && mode != FB_WIDE) return "flow_blk: bad mode";
  if (M <= 0 || T <= 0 || K <= 0 || K % 64) return "flow_blk: bad shape";
  if (mode == FB_FF && (K != 256 || N != 256)) return "flow_blk: the feed-forward is 256 -> 1024 -> 256";
  if (mode == FB_OUT && N != 256) return "flow_blk: the output projection has 256 columns";
  if (mode == FB_WIDE && (N % 256 || N > 1536 || K != 256)) return "flow_blk: wide GEMM: K = 256, N a multiple of 256 up to 1536";
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  out->mode = mode;
  FlowBlkParams& p = out->p;
  p.M = M; p.T = T; p.tiles = (M + 255) / 256;
  p.kb_a = K / 64;
  p.n_tiles = mode == FB_WIDE ? N / 256 : 1;
  p.n_bias1 = mode == FB_FF ? 1024 : (mode == FB_WIDE && b1 ? N : 0);
  p.ln = ln;
  {
    static const int dbg = [] { const char* v = getenv("GONOVA_FB_DBG"); return v ? atoi(v) : 0; }();
    static const int dbg_mode = [] { const char* v = getenv("GONOVA_FB_TRACE_MODE"); return v ? atoi(v) : 0; }();
    p.dbg = (dbg & 8) ? ((mode == dbg_mode ? 8 : 0) | (dbg & ~8)) : dbg;   // (FB_OUTFF is built as FB_FF: trace mode 0)
  }
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta;
  p.r = r; p.r_in = r_in ? r_in : r; p.n_out = (__nv_bfloat16*)n_out; p.n_pitch = n_pitch;
  if (r_in && ((uintptr_t)r_in & 15)) return "flow_blk: residual input must be 16-byte aligned";
  p.vt = mode == FB_WIDE ? (__nv_bfloat16*)vt : nullptr; p.vt_col0 = vt_col0; p.vt_tp = vt_tp;
  if (p.vt && (vt_col0 % 256 || vt_tp < T)) return "flow_blk: bad transposed-V geometry";
  if (mode != FB_WIDE && (!r || !n_out || ((uintptr_t)r & 31) || ((uintptr_t)n_out & 31) || n_pitch % 16))
    return "flow_blk: residual / output rows must be 32-byte aligned";
  // WIDE: n tiles per work unit — the unit's A tile stays resident; fewer per unit when there are not enough m tiles to
  // give every CTA pair work
  p.npu = 1;
  if (mode == FB_WIDE) {
    const int pairs_avail = std::max(1, max_ctas / 2);
    for (int cand : {6, 3, 2, 1})
      if (p.n_tiles % cand == 0 && (long)p.tiles * (p.n_tiles / cand) >= pairs_avail) { p.npu = cand; break; }
  }
  const uint32_t fmt = 1u;   // bf16
  const auto idesc = [&](uint32_t n) { return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24); };
  p.idesc128 = idesc(128);
  p.idesc256 = idesc(256);
  p.sw = 5;
  p.off_x = 0;
  p.off_h = 4 * kFbSlot;
  p.off_w = p.off_h + 4 * kFbSlot;
  p.off_sc = p.off_w + (uint32_t)p.sw * kFbSlot;     // (no separate scratch any more: the epilogue's scratch is the H region)
  p.off_tab = p.off_sc;
  p.off_bar = p.off_tab + (((uint32_t)kFbTab * 4u + 1023u) & ~1023u);
  out->smem_bytes = (size_t)p.off_bar + 1024 + 1024;
  if (out->smem_bytes > kFbMaxDynSmem) return "flow_blk: shared memory budget exceeded";

  const CUtensorMapL2promotion l2a = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, l2w = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const char* e = encode_2d(enc, &out->maps.A, a, 2, false, K, M, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, l2a);
  if (*e) return e;
  if (mode == FB_FF) {
    e = encode_2d(enc, &out->maps.W1, w1, 2, false, 256, 1024, 256, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, l2w);
    if (*e) return e;
    e = encode_2d(enc, &out->maps.W2, w2, 2, false, 1024, 256, 1024, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, l2w);
    if (*e) return e;
  } else if (mode == FB_OUT) {
    e = encode_2d(enc, &out->maps.W2, w2, 2, false, K, 256, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, l2w);
    if (*e) return e;
    out->maps.W1 = out->maps.W2;
  } else {
    e = encode_2d(enc, &out->maps.W1, w1, 2, false, K, N, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, l2w);
    if (*e) return e;
    out->maps.W2 = out->maps.W1;
  }
  if (mode == FB_WIDE) {
    e = encode_2d(enc, &out->maps.Nout, n_out, 2, false, N, M, n_pitch, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, l2a);
    if (*e) return e;
  } else {
    out->maps.Nout = out->maps.A;
  }
  const int items = p.tiles * (p.n_tiles / p.npu);
  const int pairs = std::max(1, std::min(items, std::max(1, max_ctas / 2)));
  out->grid = 2 * pairs;
  return "";
}

const char* make_flow_conv_launch(FlowBlkLaunch* out, const void* a, int C_in, const void* w, const float* bias,
                                  const float* gamma, const float* beta, void* out_bf16, float* out_f32, int B2, int T,
                                  int max_ctas) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  if (B2 <= 0 || T <= 0 || C_in <= 0 || C_in % 64) return "flow_conv: bad shape";
  if (!a || !w || (!out_bf16 && !out_f32)) return "flow_conv: NULL tensor";
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  out->mode = FB_CONV;
  FlowBlkParams& p = out->p;
  p.M = B2 * T; p.T = T;
  p.conv_tpb = (T + 255) / 256;
  p.tiles = B2 * p.conv_tpb;
  p.conv_nch = C_in / 64;
  p.kb_a = 3 * p.conv_nch;
  p.n_tiles = 1; p.npu = 1; p.n_bias1 = 0; p.ln = 1;
  p.b1 = nullptr; p.b2 = bias; p.gamma = gamma; p.beta = beta;
  p.r = out_f32; p.r_in = out_f32; p.n_out = (__nv_bfloat16*)out_bf16; p.n_pitch = 256;
  p.out_f32 = out_bf16 ? 0 : 1;
  if (((uintptr_t)a & 15) || (out_bf16 && ((uintptr_t)out_bf16 & 15)) || (out_f32 && ((uintptr_t)out_f32 & 15)))
    return "flow_conv: tensors must be 16-byte aligned";
  const auto idesc = [&](uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24); };
  p.idesc128 = idesc(128);
  p.idesc256 = idesc(256);
  p.sw = 5;
  p.off_x = 0;
  p.off_h = 4 * kFbSlot;
  p.off_w = p.off_h + 4 * kFbSlot;
  p.off_sc = p.off_w + (uint32_t)p.sw * kFbSlot;
  p.off_tab = p.off_sc;
  p.off_bar = p.off_tab + (((uint32_t)kFbTab * 4u + 1023u) & ~1023u);
  out->smem_bytes = (size_t)p.off_bar + 1024 + 1024;
  if (out->smem_bytes > kFbMaxDynSmem) return "flow_conv: shared memory budget exceeded";
  {
    cuuint64_t dims[3] = {(cuuint64_t)C_in, (cuuint64_t)T, (cuuint64_t)B2};
    cuuint64_t strides[2] = {(cuuint64_t)C_in * 2, (cuuint64_t)T * C_in * 2};
    cuuint32_t box[3] = {64u, 128u, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&out->maps.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "flow_conv: cuTensorMapEncodeTiled failed";
  }
  const char* e = encode_2d(enc, &out->maps.W2, w, 2, false, 3LL * C_in, 256, 3LL * C_in, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (*e) return e;
  out->maps.W1 = out->maps.W2;
  out->maps.Nout = out->maps.W2;
  const int pairs = std::max(1, std::min(p.tiles, std::max(1, max_ctas / 2)));
  out->grid = 2 * pairs;
  return "";
}

const char* make_flow_outff_launch(FlowBlkLaunch* out, const void* o, int K, const void* w3, const float* b3, const float* g3,
                                   const float* be3, const void* w1, const float* b1, const void* w2, const float* b2, float* r,
                                   const float* gamma, const float* beta, int ln, void* n_out, int n_pitch, int M, int T,
                                   int max_ctas, const void* w4, void* qkv_out, void* vt, int vt_tp, const float* r_in) {
  // the feed-forward's launch, then the out-projection's operands on top of it.  w1 == NULL: no feed-forward (projection +
  // residual + LayerNorm + the q/k/v tail): the maps of a dummy feed-forward over the projection's own weights are never used
  const bool noff = w1 == nullptr;
  if (noff && !w4) return "flow_outff: without a feed-forward the q/k/v tail is the launch's only output";
  const char* e = make_flow_blk_launch(out, FB_FF, o, 256, noff ? w3 : w1, b1, noff ? w3 : w2, b2, r, gamma, beta, ln,
                                       noff ? (void*)qkv_out : n_out, noff ? 1536 : n_pitch, 256, M, T, max_ctas, nullptr, 0, 0, nullptr);
  if (*e) return e;
  out->p.noff = noff ? 1 : 0;
  if (r_in) {
    if ((uintptr_t)r_in & 15) return "flow_outff: residual input must be 16-byte aligned";
    out->p.r_in = r_in;
  }
  if (K <= 0 || K % 64 || !w3) return "flow_outff: bad out-projection";
  PFN_encodeTiled enc = get_encode_tiled();
  out->mode = FB_OUTFF;
  FlowBlkParams& p = out->p;
  p.kb_a = K / 64;
  p.b3 = b3; p.g3 = g3; p.be3 = be3;
  e = encode_2d(enc, &out->maps.A, o, 2, false, K, M, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (*e) return e;
  e = encode_2d(enc, &out->maps.W3, w3, 2, false, K, 256, K, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (*e) return e;
  if (w4) {
    if (!ln || !qkv_out || !vt || vt_tp < T) return "flow_outff: the q/k/v tail needs the LayerNorm and its outputs";
    p.qkv = 1;
    p.vt = (__nv_bfloat16*)vt; p.vt_col0 = 1024; p.vt_tp = vt_tp;
    e = encode_2d(enc, &out->maps.W4, w4, 2, false, 256, 1536, 256, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (*e) return e;
    e = encode_2d(enc, &out->maps.Nout, qkv_out, 2, false, 1536, M, 1536, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (*e) return e;
  } else {
    out->maps.W4 = out->maps.W3;
  }
  return "";
}

cudaError_t launch_flow_blk(const FlowBlkLaunch& L, const int* lengths, cudaStream_t st, const float* tbias) {
  if (!L.d_maps) return cudaErrorInvalidValue;
  FlowBlkParams p = L.p;
  p.lengths = lengths;
  if (L.mode == FB_CONV) {
    p.b1 = tbias;                      // the kernel's b1 table holds the ResNet block's time bias (zeros when NULL)
    p.n_bias1 = tbias ? 256 : 0;
  }
  const FlowBlkMaps* dm = L.d_maps;
  switch (L.mode) {
    case FB_FF:  return launch_persistent(flow_blk_kernel<FB_FF>, L.grid, L.smem_bytes, st, true, kFbThreads, dm, p);
    case FB_OUT: return launch_persistent(flow_blk_kernel<FB_OUT>, L.grid, L.smem_bytes, st, true, kFbThreads, dm, p);
    case FB_CONV: return launch_persistent(flow_blk_kernel<FB_CONV>, L.grid, L.smem_bytes, st, true, kFbThreads, dm, p);
    case FB_OUTFF: return launch_persistent(flow_blk_kernel<FB_OUTFF>, L.grid, L.smem_bytes, st, true, kFbThreads, dm, p);
    default:     return launch_persistent(flow_blk_kernel<FB_WIDE>, L.grid, L.smem_bytes, st, true, kFbThreads, dm, p);
  }
}

}  // namespace gnv
