// Host side of flow_attn_tc_kernel (flow_attn.cuh): tensor maps, shared-memory carve-up, launch.
#include "flow_attn.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace gnv {

static constexpr size_t kFaMaxDynSmem = 227 * 1024;

cudaError_t flow_attn_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  preload_kernel((const void*)flow_attn_tc_kernel<0>);
  return cudaFuncSetAttribute(flow_attn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFaMaxDynSmem);
}

int flow_attn_read_trace(unsigned long long* out, int cap) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  static unsigned long long host[kFaTraceCap];
  if (cudaMemcpyFromSymbol(host, g_fa_trace, sizeof(host)) != cudaSuccess) return -1;
  int n = 0;
  for (int i = 0; i < kFaTraceCap && n < cap; ++i)
    if (host[i]) out[n++] = host[i];
  memset(host, 0, sizeof(host));
  cudaMemcpyToSymbol(g_fa_trace, host, sizeof(host));
  return n;
}

namespace {
const char* encode_3d(PFN_encodeTiled enc, CUtensorMap* m, const void* base, long long d0, long long d1, long long d2,
                      long long s1_elems, long long s2_elems, int b0, int b1) {
  if (!base) return "flow_attn: tensor is NULL";
  if (((uintptr_t)base & 15) || (s1_elems * 2) % 16 || (s2_elems * 2) % 16) return "flow_attn: tensor is not 16-byte aligned";
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1_elems * 2, (cuuint64_t)s2_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1u};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? "" : "flow_attn: cuTensorMapEncodeTiled failed";
}
}  // namespace

const char* make_flow_attn_launch(FlowAttnLaunch* out, const void* qkv, const void* vt, int Tp, void* o, int B2, int T,
                                  float scale, int max_ctas) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  if (B2 <= 0 || T <= 0 || Tp < T || Tp % 8) return "flow_attn: bad shape";
  memset(&out->maps, 0, sizeof(out->maps));
  memset(&out->p, 0, sizeof(out->p));
  out->d_maps = nullptr;
  FlowAttnParams& p = out->p;
  p.B2 = B2; p.T = T;
  p.nqp = (T + 255) / 256;
  p.items = B2 * 8 * p.nqp;
  p.o = (__nv_bfloat16*)o;
  p.sc2 = scale * 1.4426950408889634f;
  {
    const char* v = getenv("GONOVA_FB_DBG");
    const char* m = getenv("GONOVA_FB_TRACE_MODE");
    p.dbg = (v && (atoi(v) & 8) && m && atoi(m) == 3) ? 8 : 0;
  }
  const auto idesc = [&](uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); };
  p.idesc_s = idesc(128);
  p.idesc_pv = idesc(64);
  p.off_q = 0;
  p.off_kv = 2 * 32768;
  p.off_p = p.off_kv + kFaStages * 32768;
  p.off_bar = p.off_p + 2 * 32768;
  out->smem_bytes = (size_t)p.off_bar + 1024 + 2048 + 1024;      // barriers | row-maximum exchange | alignment slack
  if (out->smem_bytes > kFaMaxDynSmem) return "flow_attn: shared memory budget exceeded";
  const char* e = encode_3d(enc, &out->maps.QK, qkv, 1536, T, B2, 1536, (long long)T * 1536, 64, 128);
  if (*e) return e;
  e = encode_3d(enc, &out->maps.Vt, vt, T, 64, (long long)B2 * 8, Tp, 64LL * Tp, 64, 64);
  if (*e) return e;
  e = encode_3d(enc, &out->maps.O, o, 512, T, B2, 512, (long long)T * 512, 64, 32);
  if (*e) return e;
  out->grid = std::max(1, std::min(p.items, max_ctas));
  return "";
}

cudaError_t launch_flow_attn_tc(const FlowAttnLaunch& L, const int* lengths, cudaStream_t st) {
  if (!L.d_maps) return cudaErrorInvalidValue;
  FlowAttnParams p = L.p;
  p.lengths = lengths;
  return launch_persistent(flow_attn_tc_kernel<0>, L.grid, L.smem_bytes, st, false, kFaThreads, L.d_maps, p);
}

}  // namespace gnv
