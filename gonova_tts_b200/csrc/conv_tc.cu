// Host side of the tcgen05 conv kernel: tensor-map encoding and launch.
#include "conv_tc.cuh"

#include <cudaTypedefs.h>

#include <cstdio>
#include <cstring>
#include <mutex>

namespace gnv {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

void preload_kernel(const void* kernel) {
  typedef CUresult (*PFN_funcLoad)(CUfunction);
  static PFN_funcLoad fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuFuncLoad", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_funcLoad>(p);
    return static_cast<PFN_funcLoad>(nullptr);
  }();
  if (!fn) return;
  cudaFunction_t f = nullptr;
  if (cudaGetFuncBySymbol(&f, kernel) == cudaSuccess && f) fn(reinterpret_cast<CUfunction>(f));
  cudaGetLastError();                                        // (best effort: never leaves an error behind)
}

static constexpr size_t kMaxDynSmem = 227 * 1024;

static uint32_t* g_dbg_host = nullptr;

cudaError_t tc_debug_device_ptr(uint32_t** out) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaHostAlloc((void**)&g_dbg_host, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err != cudaSuccess) { g_dbg_host = nullptr; return; }
    memset(g_dbg_host, 0, 64);
  });
  if (err != cudaSuccess || !g_dbg_host) return err != cudaSuccess ? err : cudaErrorMemoryAllocation;
  return cudaHostGetDevicePointer((void**)out, g_dbg_host, 0);
}

const char* tc_debug_string() {
  static thread_local char buf[160];
  if (!g_dbg_host || (g_dbg_host[0] >> 16) != 0xDEAD) return "";
  snprintf(buf, sizeof(buf), " [barrier wait timed out: role tag %u, block %u, barrier smem 0x%x, parity %u]",
           g_dbg_host[0] & 0xFFFFu, g_dbg_host[1], g_dbg_host[2], g_dbg_host[3]);
  return buf;
}

cudaError_t conv_tc_init() {
  uint32_t* dptr = nullptr;
  cudaError_t e = tc_debug_device_ptr(&dptr);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(tc::g_tc_debug, &dptr, sizeof(dptr));
  if (e != cudaSuccess) return e;
  preload_kernel((const void*)conv_tc_kernel<__nv_bfloat16>);
  preload_kernel((const void*)conv_tc_kernel<float>);
  e = cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kMaxDynSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
}

const char* make_conv_tc_launch(ConvTcLaunch* out, int elem_bytes, const void* act, const void* w, int w_rows_alloc,
                                const ConvGeom& g, const EpiParams& ep, int block_n) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  const int kbe = 128 / elem_bytes;
  if (g.C_in_ld % kbe) return "conv_tc: channel stride is not a multiple of the 128-byte K block";
  if (g.C_in_w != g.C_in_ld) return "conv_tc: weight K stride must equal the activation channel stride";
  if (g.in_stride != 1) return "conv_tc: strided input rows are not supported on the tensor-core path";
  if (g.a_row_stride || g.a_batch_stride) return "conv_tc: strided A views need the persistent kernel";
  if (block_n % 16 || block_n < 16 || block_n > 256) return "conv_tc: block_n must be a multiple of 16 in [16,256]";
  if (g.N_total % block_n) return "conv_tc: N_total must be a multiple of block_n";
  if (w_rows_alloc < g.N_total) return "conv_tc: packed weights have fewer rows than N_total";
  if (((uintptr_t)act & 15) || ((uintptr_t)w & 15)) return "conv_tc: operand pointers must be 16-byte aligned";

  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.C_in_ld, (cuuint64_t)g.L_in, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)g.C_in_ld * elem_bytes, (cuuint64_t)g.L_in * g.C_in_ld * elem_bytes};
    cuuint32_t box[3] = {(cuuint32_t)kbe, 128u, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&out->tmA, dt, 3, const_cast<void*>(act), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation tensor";
  }
  {
    const int K = g.n_taps * g.C_in_ld;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)w_rows_alloc};
    cuuint64_t strides[1] = {(cuuint64_t)K * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&out->tmW, dt, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the weight tensor";
  }
  ConvTcParams& p = out->p;
  p.g = g;
  p.ep = ep;
  p.block_n = block_n;
  p.n_chunks = g.C_in_ld / kbe;
  int cols = 32;
  while (cols < block_n) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;   // BF16 : TF32
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(block_n >> 3) << 17) | ((128u >> 4) << 24);
  const int stage_bytes = 128 * 128 + block_n * 128;
  // <= ~100 KiB keeps two CTAs resident per SM; wide tiles take the whole SM.
  const int budget = block_n <= 128 ? 100 * 1024 : 200 * 1024;
  int stages = budget / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  const int total_k = g.n_taps * p.n_chunks;
  if (stages > total_k) stages = total_k < 1 ? 1 : total_k;
  p.stages = stages;
  out->smem_bytes = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 8 * (2 * stages + 1) + 16;
  if (out->smem_bytes > kMaxDynSmem) return "conv_tc: shared memory budget exceeded";
  out->grid = dim3((g.M_rows + 127) / 128, g.N_total / block_n, g.B);
  out->elem_bytes = elem_bytes;
  return "";
}

cudaError_t launch_conv_tc(const ConvTcLaunch& L, cudaStream_t st) {
  if (L.elem_bytes == 2)
    conv_tc_kernel<__nv_bfloat16><<<L.grid, 256, L.smem_bytes, st>>>(L.tmA, L.tmW, L.p);
  else
    conv_tc_kernel<float><<<L.grid, 256, L.smem_bytes, st>>>(L.tmA, L.tmW, L.p);
  return cudaGetLastError();
}

}  // namespace gnv
