// C ABI of libgonova_hift.so (include/gonova_hift.h): weight re-packing, the per-(B,T) launch plan
// and the decode schedule.  Nothing here allocates, synchronises or aborts inside the decode calls;
// every failure becomes a non-zero status plus gnv_last_error() text.
//
// Decode schedule for one [B,80,T] batch (HiFTGenerator.decode), E = bf16 | tf32 | fp32:
//   mel NCT -> melE [B,T,Cld]                               pack kernel
//   s -> spec [B,F,18] fp32                                 STFT kernel
//   conv_pre            melE -> X0 = lrelu(.)               conv GEMM, activation fused
//   per stage i (C = 256/128/64, L = 8T/40T/120T+1):
//     source_downs[i]   spec -> F1 (raw), E0 = snake        GEMM over k consecutive STFT frames (K = k*Cs)
//     source_resblock   6 conv GEMMs, residual stream F1 updated in place
//     ups[i]            X_i -> F2 = convT + bias + F1,  E0/E1/E2 = snake_j(F2)   (polyphase GEMM)
//     resblocks 3i+j    6 conv GEMMs each on (E_j, E3, F1); the last one accumulates F3 += x/3 and,
//                       for j = 2, emits X_{i+1} = lrelu(F3) straight from the epilogue
//   conv_post           X_3 -> P [B,F,18] fp32
//   iSTFT head          P -> wav                            exp/sin/iDFT/overlap-add/clamp kernel
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/gonova_hift.h"
#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_tc2.cuh"
#include "conv_pair.cuh"
#include "conv_chain.cuh"
#include "flow_blk.cuh"
#include "flow_attn.cuh"
#include "kernels.h"
#include "flow_kernels.h"
#include "flow_enc_kernels.h"

using namespace gnv;

namespace {

thread_local std::string tl_error;
thread_local std::string tl_error_ret;   // what gnv_last_error(h) hands out: a private copy of the handle's message

constexpr int kSPF = 480;
constexpr int kMelC = 80;
constexpr int kSpecFront = 8;    // zero rows before / after the STFT frames of each utterance (conv + GEMM-K padding)
constexpr int kSpecBack = 40;
constexpr int kPostPitch = 20;   // conv_post output [B, F, 18] is stored with a 16-byte-multiple channel pitch (TMA)

struct HostT {
  const float* data = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};
using WeightMap = std::map<std::string, HostT>;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline float host_round_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u = (u + 0x1000u) & ~0x1FFFu;   // cvt.rna.tf32.f32: round to nearest, ties away from zero
  float y;
  memcpy(&y, &u, 4);
  return y;
}

struct ConvLayer {
  int C_in = 0, C_out = 0, k = 1, stride = 1, pad = 0, dil = 1;
  bool transposed = false;
  bool strided_simt = false;   // source_downs on CUDA cores: strided rows, K = 18*k
  bool flat = false;           // source_downs as a GEMM over k consecutive STFT frames (tensor cores)
  bool causal = false;         // all of the conv's padding on the left (the flow decoder's CausalConv1d): L_out = L_in
  bool ahead = false;          // all of the padding on the right (the flow encoder's pre-lookahead conv): L_out = L_in
  int conv_k = 0, conv_stride = 0, conv_pad = 0;   // the real conv geometry behind a flat layer
  double flops_per_row = 0.0;  // algorithmic flops per output row (0 = 2*C_out*C_in*k)
  int C_in_ld = 0;             // channel stride of the A operand (and of each tap in the packed W)
  int n_taps = 0;
  int N_valid = 0, N_total = 0, block_n = 0;
  int w_elem = 4;              // bytes per packed weight element
  void* w = nullptr;           // device [N_total, n_taps * C_in_ld]
  float* bias = nullptr;       // device [C_out]
};

struct ResBlockW {
  ConvLayer c1[3], c2[3];
  float* a1[3] = {nullptr, nullptr, nullptr};
  float* a2[3] = {nullptr, nullptr, nullptr};
};

enum SimtVariant { SV_FFF = 0, SV_BBB = 1, SV_FFB = 2 };

struct ConvOp {
  bool tc = false;
  int tcv = 1;           // 1: conv_tc_kernel (one tile per CTA), 2: conv_tc2_kernel (persistent), 3: conv_pair_kernel,
                         // 4: conv_chain_kernel (a whole ResBlock)
  ConvTcLaunch tcl;
  ConvTc2Launch tc2l;
  ConvPairLaunch pairl;
  ConvChainLaunch chainl;
  FlowBlkLaunch blkl;    // tcv 5: flow_blk_kernel (a fused transformer-block step of the flow estimator)
  FlowAttnLaunch attnl;  // tcv 6: flow_attn_tc_kernel (self-attention of the flow estimator)
  ConvGeom g;
  EpiParams ep;
  const void* A = nullptr;
  const void* W = nullptr;
  int variant = SV_FFF;
  std::string name;      // upstream module path of the layer ("resblocks.4.convs1.2")
  int branch = 0;        // 1: source branch (source_downs.i, source_resblocks.i.*): depends on the STFT only
  int stage = -1;        // upsampling stage of a source-branch op / of ups.i / of a ResBlock op
  int rb = -1;           // ResBlock j = 0..2 of its stage (the three run in parallel in fork mode)
  bool rb_last = false;  // the launch that adds the ResBlock into the stage's running sum: these stay ordered 0, 1, 2
  double flops = 0.0;    // algorithmic: 2 * B * L_out * C_out * C_in * k (convT: 2 * B * L_in * ...)
};

// Optional per-launch timing (gnv_inference_profile): one CUDA event after every launch on the
// launching stream; durations are differences of consecutive events.
struct Profiler {
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  std::vector<int> kinds;       // GNV_LAUNCH_*
  std::vector<double> flops;
  cudaError_t err = cudaSuccess;
  void begin(cudaStream_t s) { st = s; push_event(); }
  void push_event() {
    cudaEvent_t e = nullptr;
    cudaError_t r = cudaEventCreate(&e);
    if (r == cudaSuccess) r = cudaEventRecord(e, st);
    if (r != cudaSuccess && err == cudaSuccess) err = r;
    ev.push_back(e);
  }
  void mark(const std::string& name, int kind, double fl) {
    names.push_back(name); kinds.push_back(kind); flops.push_back(fl);
    push_event();
  }
  ~Profiler() { for (auto e : ev) if (e) cudaEventDestroy(e); }
};
inline void prof_mark(Profiler* p, const char* name, int kind, double fl = 0.0) { if (p) p->mark(name, kind, fl); }

// Offsets (bytes) into the caller's workspace for one (B, T).
struct WsLayout {
  size_t melE = 0, spec = 0, X0 = 0, P = 0, H0 = 0, H1 = 0, f0 = 0;
  size_t F1[3], F2[3], F3[3], E[3][4];
  size_t F1x[3][2], E3x[3][2];    // fork mode: residual stream / ping-pong buffer of ResBlocks 1 and 2 of a stage (else = F1 / E[3])
  size_t total = 0;
};

// Device home of one plan's tensor maps.  Slots belong to the decoder handle and are recycled when the plan cache
// evicts a plan (a service sees a new sentence length on almost every call: no cudaMalloc / cudaFree per call and
// no growth).  `last_use` is recorded after every launch sequence that reads the slot.
struct MapsSlot {
  void* d = nullptr;           // device buffer
  char* staging = nullptr;     // pinned host mirror the upload is issued from
  size_t bytes = 0;
  cudaEvent_t last_use = nullptr;
  bool used = false;
  bool owns = true;            // false: `d` / `staging` are slices of a chunk the decoder handle owns
};

struct Plan {
  std::vector<ConvOp> decode_ops;
  std::vector<ConvOp> f0_ops;
  WsLayout lay;
  MapsSlot slot;                   // tensor maps of every persistent-kernel op
  unsigned long long last_tick = 0;
  bool pinned = false;             // used inside a stream capture: a CUDA graph points at `slot` for good, never evicted
  Plan() = default;
  Plan(const Plan&) = delete;
  Plan& operator=(const Plan&) = delete;
};

void free_slot(MapsSlot& sl) {
  if (sl.last_use) cudaEventDestroy(sl.last_use);
  if (sl.owns && sl.staging) cudaFreeHost(sl.staging);
  if (sl.owns && sl.d) cudaFree(sl.d);
  sl = MapsSlot();
}

// Where a decoder handle's slots come from: chunks of kSlotChunk slots (one cudaMalloc + one cudaHostAlloc per chunk).
// Allocating under traffic is what costs — measured with 16 connections: a plan build takes 0.5 ms, but 13 ms (p90)
// to 127 ms (max) when it also has to allocate its slot while the GPU is busy.
struct SlotChunks {
  std::vector<void*> dev, host;
};
constexpr int kSlotChunk = 32;

// Uploads the tensor maps of every persistent-kernel op to the plan's slot (taken from `free_slots` when one is large
// enough) and points the ops at it.  The copy is issued on `st` from pinned memory and waited for.
std::string upload_maps(std::vector<ConvOp*>& ops, std::vector<MapsSlot>& free_slots, MapsSlot* slot, cudaStream_t st,
                        SlotChunks* chunks = nullptr) {
  size_t bytes = 0;
  for (ConvOp* op : ops) {
    if (op->tc && op->tcv == 2) bytes += sizeof(ConvTc2Maps);
    if (op->tc && op->tcv == 3) bytes += sizeof(ConvPairMaps);
    if (op->tc && op->tcv == 4) bytes += sizeof(ConvChainMaps);
    if (op->tc && op->tcv == 5) bytes += sizeof(FlowBlkMaps);
    if (op->tc && op->tcv == 6) bytes += sizeof(FlowAttnMaps);
  }
  if (!bytes) return "";
  MapsSlot sl;
  for (size_t i = 0; i < free_slots.size(); ++i)
    if (free_slots[i].bytes >= bytes) { sl = free_slots[i]; free_slots.erase(free_slots.begin() + i); break; }
  if (!sl.d && chunks) {
    const size_t pitch = align_up(bytes, 1024);
    void *dbase = nullptr, *hbase = nullptr;
    cudaError_t e = cudaMalloc(&dbase, pitch * kSlotChunk);
    if (e == cudaSuccess) e = cudaHostAlloc(&hbase, pitch * kSlotChunk, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      if (dbase) cudaFree(dbase);
      return std::string("tensor-map slot chunk: ") + cudaGetErrorString(e);
    }
    chunks->dev.push_back(dbase);
    chunks->host.push_back(hbase);
    for (int i = 0; i < kSlotChunk; ++i) {
      MapsSlot c;
      c.d = (char*)dbase + (size_t)i * pitch;
      c.staging = (char*)hbase + (size_t)i * pitch;
      c.bytes = pitch;
      c.owns = false;
      if (cudaEventCreateWithFlags(&c.last_use, cudaEventDisableTiming) != cudaSuccess) return "tensor-map slot event";
      if (i == 0) sl = c; else free_slots.push_back(c);
    }
  }
  if (!sl.d) {
    cudaError_t e = cudaMalloc(&sl.d, bytes);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&sl.staging, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.last_use, cudaEventDisableTiming);
    if (e != cudaSuccess) { free_slot(sl); return std::string("tensor-map slot: ") + cudaGetErrorString(e); }
    sl.bytes = bytes;
  } else if (sl.used) {
    // the evicted plan's launches (and its own upload) must be done before the mirror is rewritten; it was the least
    // recently used plan, so this returns at once in practice
    cudaError_t e = cudaEventSynchronize(sl.last_use);
    if (e != cudaSuccess) { free_slots.push_back(sl); return std::string("tensor-map slot wait: ") + cudaGetErrorString(e); }
  }
  size_t off = 0;
  for (ConvOp* op : ops) {
    if (op->tc && op->tcv == 2) { memcpy(sl.staging + off, &op->tc2l.maps, sizeof(ConvTc2Maps)); off += sizeof(ConvTc2Maps); }
    if (op->tc && op->tcv == 3) { memcpy(sl.staging + off, &op->pairl.maps, sizeof(ConvPairMaps)); off += sizeof(ConvPairMaps); }
    if (op->tc && op->tcv == 4) { memcpy(sl.staging + off, &op->chainl.maps, sizeof(ConvChainMaps)); off += sizeof(ConvChainMaps); }
    if (op->tc && op->tcv == 5) { memcpy(sl.staging + off, &op->blkl.maps, sizeof(FlowBlkMaps)); off += sizeof(FlowBlkMaps); }
    if (op->tc && op->tcv == 6) { memcpy(sl.staging + off, &op->attnl.maps, sizeof(FlowAttnMaps)); off += sizeof(FlowAttnMaps); }
  }
  cudaError_t e = cudaMemcpyAsync(sl.d, sl.staging, bytes, cudaMemcpyHostToDevice, st);
  // `st` is the handle's own upload stream (or the test hook's stream): the copy is complete before the plan is
  // published, so a later cache hit on ANY stream — or inside a stream capture — finds the maps in place
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaEventRecord(sl.last_use, st);
  if (e != cudaSuccess) { free_slots.push_back(sl); return std::string("cudaMemcpyAsync(tensor maps): ") + cudaGetErrorString(e); }
  sl.used = true;
  off = 0;
  for (ConvOp* op : ops) {
    if (op->tc && op->tcv == 2) { op->tc2l.d_maps = reinterpret_cast<const ConvTc2Maps*>((char*)sl.d + off); off += sizeof(ConvTc2Maps); }
    if (op->tc && op->tcv == 3) { op->pairl.d_maps = reinterpret_cast<const ConvPairMaps*>((char*)sl.d + off); off += sizeof(ConvPairMaps); }
    if (op->tc && op->tcv == 4) { op->chainl.d_maps = reinterpret_cast<const ConvChainMaps*>((char*)sl.d + off); off += sizeof(ConvChainMaps); }
    if (op->tc && op->tcv == 5) { op->blkl.d_maps = reinterpret_cast<const FlowBlkMaps*>((char*)sl.d + off); off += sizeof(FlowBlkMaps); }
    if (op->tc && op->tcv == 6) { op->attnl.d_maps = reinterpret_cast<const FlowAttnMaps*>((char*)sl.d + off); off += sizeof(FlowAttnMaps); }
  }
  *slot = sl;
  return "";
}

using PlanKey = std::tuple<int, int, const void*>;

}  // namespace

struct gnv_decoder {
  int device = 0, dtype = GNV_DTYPE_TF32;
  unsigned flags = 0;
  int eb = 4;            // bytes per activation element E
  bool use_tc = true;
  int tc_version = 2;
  bool fuse_pairs = true;   // conv1 + Snake + conv2 + residual of a ResBlock step in one kernel
  int pair_cta2 = -1;       // CTA pairs in the fused kernel: -1 = choose (enough tiles for 74 pairs twice), 0 = never, 1 = always.
                            // Round 1 measured pairs 0.15-0.2 ms SLOWER per C=64 pair and kept them off; the cost was the
                            // `mbarrier.arrive.release.cluster` of every epilogue (MEMBAR.ALL.CTA + ERRBAR, 2.6 k cycles, found on
                            // flow_blk_kernel's timeline).  With default-semantics arrives the step is 2.3 % faster in pairs
                            // (three interleaved runs: 26.83 -> 26.20 ms), mostly as clock: half the weight traffic into shared
                            // memory, and the step runs into the board's power cap.
  int fuse_max_c = 64;      // ... for stages with at most this many channels.  Measured (B=64, T=500, bf16): C=64
                            // pairs are 5-28 % faster fused; C=128 pairs must drop to 128-row tiles to fit TMEM
                            // (3*mh*C <= 512), which doubles the weight traffic and makes k=7/11 pairs 15-30 % slower.
  int fuse_k3_max_c = 128;  // ... and for k = 3 ResBlocks up to this many channels (GONOVA_FUSE_K3_MAX_C).  Measured at C = 128:
                            // the three fused k = 3 pairs take 1.51 ms against 1.59 ms in six launches; with CTA pairs in the
                            // fused kernel the step is 0.9 % shorter (three interleaved runs: 25.95 -> 25.72 ms)
  int chain_max_k = 3;      // whole-ResBlock kernel (conv_chain_kernel) for the blocks that open a stage's running sum (j = 0)
  int chain_max_c = 64;     // ... with at most this many taps / channels (GONOVA_CHAIN_MAX_K / _MAX_C; 0 = never)
  std::map<std::pair<int, int>, int> launch_counts;   // (B, T) -> conv launches of the last plan built
  ConvTc2Options tc2opt;
  int snake_kind = ACT_SNAKE;
  std::string err;
  std::vector<void*> allocs;
  ConvLayer conv_pre, ups[3], sdown[3], sdown_flat[3], conv_post, f0c[5];
  int spec_cs = 20;      // channel pitch of the STFT buffer (18 -> 24 bf16 / 20 fp32: 16-byte multiples)
  ResBlockW rb[9], srb[3];
  float *f0_w = nullptr, *f0_b = nullptr, *lin_w = nullptr, *lin_b = nullptr;
  std::map<PlanKey, std::shared_ptr<Plan>> plans;   // shared: a call keeps its plan alive if another thread evicts the cache
  std::vector<MapsSlot> free_slots;                 // tensor-map slots of evicted plans, reused by the next plan built
  SlotChunks slot_chunks;                           // the memory behind all of them
  size_t max_plans = 128;                           // LRU bound of `plans` (GONOVA_MAX_PLANS); pinned plans do not count out
  unsigned long long tick = 0;
  unsigned long long plans_built = 0, slots_allocated = 0;
  // small problems: the source branch of every stage runs on a side stream beside conv_pre / the previous stages
  cudaStream_t upload_st = nullptr;                 // tensor-map uploads of new plans (synchronised at build time)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  cudaStream_t rbs[2] = {nullptr, nullptr};         // ResBlocks 1 and 2 of a stage (ResBlock 0 stays on the caller's stream)
  cudaEvent_t ev_ups = nullptr, ev_pre[2] = {nullptr, nullptr};
  int fork_max_frames = 4096;                       // B * T up to which the fork is used (GONOVA_FORK_MAX_FRAMES; 0 = never): measured
                                                    // at T = 500: B = 1 -9 %, B = 4 -6 %, B = 6 -8 %, B = 8 -3 %, B >= 12 nothing
  std::mutex mu;
  std::mutex err_mu;                                // guards `err`
};

namespace {

int fail(gnv_handle h, const std::string& msg) {
  tl_error = msg;
  if (h) {
    std::lock_guard<std::mutex> lk(h->err_mu);
    h->err = msg;
  }
  return 1;
}
int fail_cuda(gnv_handle h, const char* what, cudaError_t e) {
  return fail(h, std::string(what) + ": " + cudaGetErrorString(e) + tc_debug_string());
}

// Tuning knobs for experiments on the GPU box (defaults are the production configuration).
ConvTc2Options tc2_options_from_env() {
  ConvTc2Options o;
  if (const char* v = getenv("GONOVA_TC2_SLAB")) o.slab_mode = atoi(v);
  if (const char* v = getenv("GONOVA_TC2_MH")) o.mh = atoi(v);
  if (const char* v = getenv("GONOVA_TC2_CTAS")) o.max_ctas = atoi(v);
  if (const char* v = getenv("GONOVA_TC2_WGROUP")) o.w_group = atoi(v);
  if (const char* v = getenv("GONOVA_TC2_CTA2")) o.cta2 = atoi(v);
  if (const char* v = getenv("GONOVA_TC2_NARROW")) o.narrow_small = atoi(v);
  if (o.slab_mode < 0 || o.slab_mode > 2) o.slab_mode = 1;
  if (o.max_ctas < 1) o.max_ctas = 148;
  return o;
}

const int kStageC[3] = {256, 128, 64};
const int kUpRate[3] = {8, 5, 3};
const int kUpK[3] = {16, 11, 7};
const int kRbK[3] = {3, 7, 11};
const int kSrbK[3] = {7, 7, 11};
const int kDil[3] = {1, 3, 5};
const int kSdK[3] = {30, 6, 1};
const int kSdS[3] = {15, 3, 1};
const int kSdP[3] = {7, 1, 0};

inline int stage_len(int i, int T) { return i == 0 ? 8 * T : (i == 1 ? 40 * T : 120 * T + 1); }

int choose_block_n(int N) {
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 32;
}

// ---- weights ------------------------------------------------------------------------------------
struct Uploader {
  gnv_handle h;
  bool ok = true;
  std::string msg;
  void* put(const void* host, size_t bytes) {
    if (!ok) return nullptr;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, align_up(bytes, 256));
    if (e != cudaSuccess) { ok = false; msg = std::string("cudaMalloc: ") + cudaGetErrorString(e); return nullptr; }
    h->allocs.push_back(d);
    e = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { ok = false; msg = std::string("cudaMemcpy: ") + cudaGetErrorString(e); return nullptr; }
    return d;
  }
};

const HostT* find(const WeightMap& wm, const std::string& name, std::string* err) {
  auto it = wm.find(name);
  if (it == wm.end()) { *err = "missing weight '" + name + "'"; return nullptr; }
  return &it->second;
}

// Packs one Conv1d [Cout,Cin,k] / ConvTranspose1d [Cin,Cout,k] into the GEMM B operand
// W[n, tap*C_in_ld + ci] (K-major rows) in the element type of the chosen arithmetic.
bool pack_layer(gnv_handle h, Uploader& up, const WeightMap& wm, const std::string& prefix, ConvLayer& L,
                int C_in, int C_out, int k, int stride, int pad, int dil, bool transposed, bool strided_simt,
                std::string* err) {
  const HostT* w = find(wm, prefix + ".weight", err);
  const HostT* b = find(wm, prefix + ".bias", err);
  if (!w || !b) return false;
  const int64_t want = (int64_t)C_in * C_out * k;
  const bool shape_ok = w->shape.size() == 3 && w->numel() == want &&
                        w->shape[0] == (transposed ? C_in : C_out) && w->shape[1] == (transposed ? C_out : C_in) &&
                        w->shape[2] == k && b->numel() == C_out;
  if (!shape_ok) { *err = "bad shape for '" + prefix + "'"; return false; }
  L.C_in = C_in; L.C_out = C_out; L.k = k; L.stride = stride; L.pad = pad; L.dil = dil;
  L.transposed = transposed; L.strided_simt = strided_simt;
  const int eb = h->eb;
  const int kbe = 128 / eb;
  L.C_in_ld = strided_simt ? C_in : (int)align_up(C_in, kbe);
  L.w_elem = eb;
  if (transposed) {
    L.n_taps = (k + stride - 1) / stride;
    L.N_valid = stride * C_out;
  } else {
    L.n_taps = k;
    L.N_valid = C_out;
  }
  L.block_n = choose_block_n(L.N_valid);
  L.N_total = (int)align_up(L.N_valid, L.block_n);
  const size_t K = (size_t)L.n_taps * L.C_in_ld;
  std::vector<float> P((size_t)L.N_total * K, 0.f);
  if (transposed) {
    for (int r = 0; r < stride; ++r)
      for (int j = 0; j < L.n_taps; ++j) {
        const int kk = r + stride * j;
        if (kk >= k) continue;
        for (int co = 0; co < C_out; ++co)
          for (int ci = 0; ci < C_in; ++ci)
            P[((size_t)r * C_out + co) * K + (size_t)j * L.C_in_ld + ci] = w->data[((size_t)ci * C_out + co) * k + kk];
      }
  } else {
    for (int co = 0; co < C_out; ++co)
      for (int ci = 0; ci < C_in; ++ci)
        for (int kk = 0; kk < k; ++kk)
          P[(size_t)co * K + (size_t)kk * L.C_in_ld + ci] = w->data[((size_t)co * C_in + ci) * k + kk];
  }
  if (eb == 2) {
    std::vector<__nv_bfloat16> Q(P.size());
    for (size_t i = 0; i < P.size(); ++i) Q[i] = __float2bfloat16_rn(P[i]);
    L.w = up.put(Q.data(), Q.size() * 2);
  } else {
    if (h->dtype == GNV_DTYPE_TF32)
      for (auto& v : P) v = host_round_tf32(v);
    L.w = up.put(P.data(), P.size() * 4);
  }
  L.bias = (float*)up.put(b->data, (size_t)C_out * 4);
  return up.ok;
}

// source_downs[i] as a plain GEMM: the STFT buffer is [rows, Cs] with Cs*elem a 16-byte multiple, so
// the k frames a strided conv reads for output row m are ONE contiguous run of k*Cs elements starting
// at frame (m*stride - pad).  The A operand is the overlapping-row view {row stride = stride*Cs};
// W[n, j*Cs + c] = w[n, c, j], zero for the padding channels and up to the 128-byte K block.
bool pack_flat(gnv_handle h, Uploader& up, const WeightMap& wm, const std::string& prefix, ConvLayer& L, int C_in,
               int C_out, int k, int stride, int pad, std::string* err) {
  const HostT* w = find(wm, prefix + ".weight", err);
  const HostT* b = find(wm, prefix + ".bias", err);
  if (!w || !b) return false;
  if (w->numel() != (int64_t)C_in * C_out * k || b->numel() != C_out) { *err = "bad shape for '" + prefix + "'"; return false; }
  const int eb = h->eb, kbe = 128 / eb, Cs = h->spec_cs;
  const int K = (int)align_up((size_t)k * Cs, kbe);
  L = ConvLayer();
  L.C_in = K; L.C_out = C_out; L.k = 1; L.stride = 1; L.pad = 0; L.dil = 1;
  L.flat = true; L.conv_k = k; L.conv_stride = stride; L.conv_pad = pad;
  L.flops_per_row = 2.0 * C_out * C_in * k;
  L.C_in_ld = K; L.w_elem = eb; L.n_taps = 1;
  L.N_valid = C_out; L.block_n = choose_block_n(C_out); L.N_total = (int)align_up(C_out, L.block_n);
  std::vector<float> P((size_t)L.N_total * K, 0.f);
  for (int co = 0; co < C_out; ++co)
    for (int ci = 0; ci < C_in; ++ci)
      for (int j = 0; j < k; ++j) P[(size_t)co * K + (size_t)j * Cs + ci] = w->data[((size_t)co * C_in + ci) * k + j];
  if (eb == 2) {
    std::vector<__nv_bfloat16> Q(P.size());
    for (size_t i = 0; i < P.size(); ++i) Q[i] = __float2bfloat16_rn(P[i]);
    L.w = up.put(Q.data(), Q.size() * 2);
  } else {
    if (h->dtype == GNV_DTYPE_TF32)
      for (auto& v : P) v = host_round_tf32(v);
    L.w = up.put(P.data(), P.size() * 4);
  }
  L.bias = (float*)up.put(b->data, (size_t)C_out * 4);
  return up.ok;
}

bool put_vec(Uploader& up, const WeightMap& wm, const std::string& name, int64_t n, float** out, std::string* err) {
  const HostT* t = find(wm, name, err);
  if (!t) return false;
  if (t->numel() != n) { *err = "bad shape for '" + name + "'"; return false; }
  *out = (float*)up.put(t->data, (size_t)n * 4);
  return up.ok;
}

bool pack_resblock(gnv_handle h, Uploader& up, const WeightMap& wm, const std::string& prefix, ResBlockW& R, int C,
                   int k, std::string* err) {
  for (int d = 0; d < 3; ++d) {
    const std::string sd = std::to_string(d);
    if (!pack_layer(h, up, wm, prefix + ".convs1." + sd, R.c1[d], C, C, k, 1, (k * kDil[d] - kDil[d]) / 2, kDil[d],
                    false, false, err))
      return false;
    if (!pack_layer(h, up, wm, prefix + ".convs2." + sd, R.c2[d], C, C, k, 1, (k - 1) / 2, 1, false, false, err))
      return false;
    if (!put_vec(up, wm, prefix + ".activations1." + sd + ".alpha", C, &R.a1[d], err)) return false;
    if (!put_vec(up, wm, prefix + ".activations2." + sd + ".alpha", C, &R.a2[d], err)) return false;
  }
  return true;
}

// ---- workspace ----------------------------------------------------------------------------------
// Small problems run the independent branches of the decoder on forked streams (run_decode).
inline bool fork_mode(const gnv_decoder* h, int B, int T) {
  return h->fork_max_frames > 0 && (long)B * T <= h->fork_max_frames;
}

WsLayout make_layout(const gnv_decoder* h, int B, int T) {
  WsLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t eb = h->eb;
  const int F = 120 * T + 1;
  const bool fork = fork_mode(h, B, T);
  w.melE = take((size_t)B * T * h->conv_pre.C_in_ld * eb);
  w.spec = take((size_t)B * (F + kSpecFront + kSpecBack) * h->spec_cs * eb);
  w.X0 = take((size_t)B * T * 512 * eb);
  w.P = take((size_t)B * F * kPostPitch * 4);
  w.H0 = take((size_t)B * T * 512 * eb);
  w.H1 = take((size_t)B * T * 512 * eb);
  w.f0 = take((size_t)B * T * 4);
  for (int i = 0; i < 3; ++i) {
    const size_t n = (size_t)B * stage_len(i, T) * kStageC[i];
    w.F1[i] = take(n * 4);
    w.F2[i] = take(n * 4);
    w.F3[i] = take(n * 4);
    for (int j = 0; j < 4; ++j) w.E[i][j] = take(n * eb);
    for (int j = 0; j < 2; ++j) {
      w.F1x[i][j] = fork ? take(n * 4) : w.F1[i];
      w.E3x[i][j] = fork ? take(n * eb) : w.E[i][3];
    }
  }
  w.total = off;
  return w;
}

// ---- op construction ----------------------------------------------------------------------------
struct ActSpec {
  int kind = ACT_NONE;
  const float* alpha = nullptr;
  float slope = 0.f;
  void* out = nullptr;
};

struct EpiSpec {
  const float* res = nullptr;
  float* raw = nullptr;
  float raw_scale = 1.f;
  int raw_accum = 0;
  std::vector<ActSpec> acts;
  int len_mul = 1, len_add = 0;
  bool reflect_front = false;   // ReflectionPad1d((1,0)) after the last upsampling
  int c_pitch = 0;              // output channel pitch; 0 = C_out
};

// Builds the GEMM geometry + fused epilogue of one layer on input [B, L_in, C_in_ld].
std::string make_op(const gnv_decoder* h, const ConvLayer& L, const void* A, int B, int L_in, const EpiSpec& es,
                    ConvOp* op, long long a_row_stride = 0, long long a_batch_stride = 0) {
  ConvGeom& g = op->g;
  EpiParams& ep = op->ep;
  memset(&g, 0, sizeof(g));
  memset(&ep, 0, sizeof(ep));
  g.B = B; g.L_in = L_in; g.C_in = L.C_in; g.C_in_ld = L.C_in_ld; g.C_in_w = L.C_in_ld;
  g.N_total = L.N_total; g.n_taps = L.n_taps;
  g.a_row_stride = a_row_stride; g.a_batch_stride = a_batch_stride;
  ep.C_out = L.C_out; ep.N_valid = L.N_valid;
  ep.C_pitch = es.c_pitch > 0 ? es.c_pitch : L.C_out;
  if (L.transposed) {
    g.off0 = 0; g.tap_step = -1; g.in_stride = 1;
    ep.up = L.stride; ep.pad_out = L.pad;
    ep.L_store = L_in * L.stride;
    g.M_rows = (ep.L_store - 1 + L.pad) / L.stride + 1;
    ep.shift = es.reflect_front ? 1 : 0;
    ep.dup_row = es.reflect_front ? 1 : -1;
    ep.L_out = ep.L_store + ep.shift;
  } else {
    const int L_out = (L.causal || L.ahead) ? L_in : (L_in + 2 * L.pad - L.dil * (L.k - 1) - 1) / L.stride + 1;
    g.off0 = L.causal ? -L.dil * (L.k - 1) : (L.ahead ? 0 : -L.pad); g.tap_step = L.dil; g.in_stride = L.stride;
    g.M_rows = L_out;
    ep.up = 1; ep.pad_out = 0; ep.shift = 0; ep.dup_row = -1;
    ep.L_store = L_out; ep.L_out = L_out;
  }
  ep.lengths = nullptr; ep.len_mul = es.len_mul; ep.len_add = es.len_add;
  ep.bias = L.bias; ep.res = es.res; ep.raw = es.raw; ep.raw_scale = es.raw_scale; ep.raw_accum = es.raw_accum;
  if ((int)es.acts.size() > kMaxAct) return "too many fused activations";
  ep.n_act = (int)es.acts.size();
  for (int i = 0; i < ep.n_act; ++i) {
    ep.act_out[i] = es.acts[i].out;
    ep.act_alpha[i] = es.acts[i].alpha;
    ep.act_kind[i] = es.acts[i].kind;
    ep.act_slope[i] = es.acts[i].slope;
  }
  ep.round_tf32 = h->dtype == GNV_DTYPE_TF32 ? 1 : 0;
  op->A = A; op->W = L.w;
  if (L.strided_simt) {
    op->tc = false;
    op->variant = h->eb == 2 ? SV_BBB : SV_FFF;
    return "";
  }
  if (h->use_tc) {
    op->tc = true;
    op->tcv = h->tc_version;
    if (h->tc_version == 2)
      return make_conv_tc2_launch(&op->tc2l, h->eb, A, L.w, L.N_total, g, ep, ep.C_pitch, h->tc2opt);
    return make_conv_tc_launch(&op->tcl, h->eb, A, L.w, L.N_total, g, ep, L.block_n);
  }
  op->tc = false;
  op->variant = h->eb == 2 ? SV_BBB : SV_FFF;
  return "";
}

// Geometry + epilogue parameters of a layer without building a launch (the fused pair builds its own).
std::string make_epi_only(const gnv_decoder* h, const ConvLayer& L, int B, int L_in, const EpiSpec& es, ConvOp* op) {
  gnv_decoder tmp_h;
  tmp_h.dtype = h->dtype; tmp_h.eb = h->eb; tmp_h.use_tc = false;   // no launch object
  ConvLayer Lc = L;
  Lc.strided_simt = false;
  return make_op(&tmp_h, Lc, nullptr, B, L_in, es, op);
}

cudaError_t run_op(const ConvOp& op, const int* lengths, cudaStream_t st) {
  if (op.tc) {
    if (op.tcv == 6) return launch_flow_attn_tc(op.attnl, lengths, st);
    if (op.tcv == 5) return launch_flow_blk(op.blkl, lengths, st);
    if (op.tcv == 4) return launch_conv_chain(op.chainl, lengths, st);
    if (op.tcv == 3) return launch_conv_pair(op.pairl, lengths, st);
    if (op.tcv == 2) return launch_conv_tc2(op.tc2l, lengths, st);
    if (!lengths) return launch_conv_tc(op.tcl, st);
    ConvTcLaunch L = op.tcl;
    L.p.ep.lengths = lengths;
    return launch_conv_tc(L, st);
  }
  EpiParams ep = op.ep;
  ep.lengths = lengths;
  switch (op.variant) {
    case SV_BBB: return launch_conv_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(op.A, op.W, op.g, ep, st);
    case SV_FFB: return launch_conv_simt<float, float, __nv_bfloat16>(op.A, op.W, op.g, ep, st);
    default:     return launch_conv_simt<float, float, float>(op.A, op.W, op.g, ep, st);
  }
}

std::string build_plan(gnv_decoder* h, int B, int T, void* ws, Plan* plan, cudaStream_t st) {
  plan->lay = make_layout(h, B, T);
  const WsLayout& w = plan->lay;
  char* base = (char*)ws;
  auto P = [&](size_t off) { return (void*)(base + off); };
  auto Fp = [&](size_t off) { return (float*)(base + off); };
  std::string e;
  auto add = [&](std::vector<ConvOp>& ops, const ConvLayer& L, const void* A, int L_in, const EpiSpec& es,
                 const std::string& name, long long a_row_stride = 0, long long a_batch_stride = 0) -> bool {
    if (!e.empty()) return false;
    ConvOp op;
    std::string me = make_op(h, L, A, B, L_in, es, &op, a_row_stride, a_batch_stride);
    if (!me.empty()) { e = name + ": " + me; return false; }
    op.name = name;
    const double rows = L.transposed ? (double)L_in : (double)op.ep.L_store;
    op.flops = L.flops_per_row > 0 ? B * rows * L.flops_per_row : 2.0 * B * rows * L.C_out * L.C_in * L.k;
    ops.push_back(op);
    return true;
  };
  const int snake = h->snake_kind;
  const int F = 120 * T + 1;

  // ---- f0 predictor: 5 x (conv k3 + ELU) ----
  {
    const void* in = P(w.melE);
    size_t outs[2] = {w.H0, w.H1};
    for (int i = 0; i < 5; ++i) {
      EpiSpec es;
      es.acts.push_back({h->snake_kind == ACT_SNAKE_FAST ? ACT_ELU_FAST : ACT_ELU, nullptr, 0.f, P(outs[i & 1])});
      add(plan->f0_ops, h->f0c[i], in, T, es, "f0_predictor.condnet." + std::to_string(2 * i));
      in = P(outs[i & 1]);
    }
  }
  // ---- decode ----
  auto& ops = plan->decode_ops;
  {
    EpiSpec es;
    es.acts.push_back({ACT_LRELU, nullptr, 0.1f, P(w.X0)});
    add(ops, h->conv_pre, P(w.melE), T, es, "conv_pre");
  }
  const void* X = P(w.X0);
  int Lx = T;
  for (int i = 0; i < 3; ++i) {
    const int Ls = stage_len(i, T);
    const int lm = i == 0 ? 8 : (i == 1 ? 40 : 120), la = i == 2 ? 1 : 0;
    float *F1 = Fp(w.F1[i]), *F2 = Fp(w.F2[i]), *F3 = Fp(w.F3[i]);
    void* E[4] = {P(w.E[i][0]), P(w.E[i][1]), P(w.E[i][2]), P(w.E[i][3])};
    // j = 0 opens the stage's running sum, j = 1 adds into it; j = 2 also emits the next stage's input (an activated
    // copy the chain kernel does not write) and stays on the pair path
    auto chain_block = [&](const ResBlockW& R, int stage, int j, bool final_to_sum) {
      return h->use_tc && h->tc_version == 2 && final_to_sum && j <= 1 && R.c1[0].k <= h->chain_max_k &&
             kStageC[stage] <= h->chain_max_c && R.c1[0].k >= 3;
    };
    auto resblock = [&](const ResBlockW& R, void* EA, const float* res_first, float* raw_stream, bool final_to_sum,
                        int j, const std::string& rbname, void* E3) {
      // ---- the whole block in one kernel (conv_chain_kernel): reads the fp32 stage input, keeps the residual stream in
      // TMEM across the six convs, writes out_scale * x.  For the block that OPENS the stage's running sum (j = 0: no
      // accumulate, no successor activation).
      if (chain_block(R, i, j, final_to_sum) && e.empty()) {
        ConvChainSpec cs;
        for (int d = 0; d < 3; ++d) {
          cs.w1[d] = R.c1[d].w; cs.w2[d] = R.c2[d].w;
          cs.bias1[d] = R.c1[d].bias; cs.bias2[d] = R.c2[d].bias;
          cs.alpha1[d] = R.a1[d]; cs.alpha2[d] = R.a2[d];
          cs.dil[d] = R.c1[d].dil;
        }
        ConvOp op;
        memset(&op.g, 0, sizeof(op.g));
        memset(&op.ep, 0, sizeof(op.ep));
        const char* ce = make_conv_chain_launch(&op.chainl, h->eb, res_first, F3, cs, B, Ls, kStageC[i], R.c1[0].k,
                                                1.f / 3.f, j > 0 ? 1 : 0, snake, h->dtype == GNV_DTYPE_TF32 ? 1 : 0, lm, la,
                                                h->tc2opt.max_ctas);
        if (!*ce) {
          op.tc = true; op.tcv = 4;
          op.name = rbname + ".chain";
          op.flops = 6.0 * (2.0 * B * (double)Ls * kStageC[i] * kStageC[i] * R.c1[0].k);
          ops.push_back(op);
          return;
        }
      }
      void* cur = EA;                 // the (activated) input of the next dilation step
      for (int d = 0; d < 3; ++d) {
        EpiSpec es; es.len_mul = lm; es.len_add = la;
        es.res = d == 0 ? res_first : raw_stream;
        ActSpec next_act;             // what the step emits for its successor (none for the last step of j < 2)
        bool has_next = false;
        if (d < 2) {
          es.raw = raw_stream;
          next_act = {snake, R.a1[d + 1], 0.f, nullptr};
          has_next = true;
        } else if (!final_to_sum) {
          es.raw = raw_stream;
        } else {
          es.raw = F3; es.raw_scale = 1.f / 3.f; es.raw_accum = j > 0 ? 1 : 0;
          if (j == 2) { next_act = {ACT_LRELU, nullptr, i == 2 ? 0.01f : 0.1f, E[0]}; has_next = true; }
        }
        // ---- fused: one kernel, the intermediate stays in shared memory.  The step reads `cur` with a
        // halo while other tiles write the successor's input, so input and output ping-pong (cur <-> E3).
        if (h->use_tc && h->tc_version == 2 && h->fuse_pairs && e.empty() &&
            (kStageC[i] <= h->fuse_max_c || (kStageC[i] <= h->fuse_k3_max_c && R.c1[d].k == 3))) {
          EpiSpec ef = es;
          void* out_buf = (cur == E3) ? EA : E3;
          if (has_next) {
            ActSpec a = next_act;
            if (!a.out) a.out = out_buf;
            ef.acts.push_back(a);
          }
          ConvOp op;
          // build conv2's epilogue parameters through make_op's common path (geometry of a plain k-tap conv)
          ConvOp tmp_op;
          std::string me = make_epi_only(h, R.c2[d], B, Ls, ef, &tmp_op);
          if (me.empty()) {
            const char* pe = make_conv_pair_launch(&op.pairl, h->eb, cur, R.c1[d].w, R.c2[d].w, B, Ls, R.c1[d].C_in,
                                                   R.c1[d].C_in_ld, R.c1[d].k, R.c1[d].dil, R.c1[d].bias, R.a2[d], snake,
                                                   tmp_op.ep, h->tc2opt.max_ctas, h->tc2opt.mh, h->pair_cta2);
            if (!*pe) {
              op.tc = true; op.tcv = 3;
              op.g = tmp_op.g; op.ep = tmp_op.ep;
              op.name = rbname + ".pair" + std::to_string(d);
              op.flops = 2.0 * (2.0 * B * (double)Ls * R.c1[d].C_out * R.c1[d].C_in * R.c1[d].k);
              ops.push_back(op);
              if (has_next && next_act.out == nullptr) cur = out_buf;
              continue;
            }
          }
        }
        // ---- two launches: conv1 -> E3, conv2 -> (raw, next act written over the step's own input)
        {
          EpiSpec e1; e1.len_mul = lm; e1.len_add = la;
          void* mid = (cur == E3) ? EA : E3;
          e1.acts.push_back({snake, R.a2[d], 0.f, mid});
          add(ops, R.c1[d], cur, Ls, e1, rbname + ".convs1." + std::to_string(d));
          if (has_next) {
            ActSpec a = next_act;
            if (!a.out) a.out = cur;
            es.acts.push_back(a);
          }
          add(ops, R.c2[d], mid, Ls, es, rbname + ".convs2." + std::to_string(d));
        }
      }
    };
    // source branch
    {
      EpiSpec es; es.len_mul = lm; es.len_add = la;
      es.raw = F1;
      es.acts.push_back({snake, h->srb[i].a1[0], 0.f, E[0]});
      // STFT buffer: [B, front + F + back, Cs] of E; frame f of utterance b is row front + f
      const int Cs = h->spec_cs, Fp = F + kSpecFront + kSpecBack;
      const char* spec0 = (const char*)P(w.spec);
      const std::string nm = "source_downs." + std::to_string(i);
      bool done = false;
      if (h->use_tc && h->tc_version == 2 && e.empty()) {       // (an EARLIER layer's error must survive: only this layer's own is cleared)
        const ConvLayer& Lf = h->sdown_flat[i];
        const void* Av = spec0 + (size_t)(kSpecFront - Lf.conv_pad) * Cs * h->eb;
        done = add(ops, Lf, Av, Ls, es, nm, (long long)Lf.conv_stride * Cs, (long long)Fp * Cs);
        if (!done) e.clear();          // tensor map refused (overlapping rows): fall back to the CUDA-core kernel
      }
      if (!done) add(ops, h->sdown[i], spec0 + (size_t)kSpecFront * Cs * h->eb, F, es, nm, Cs, (long long)Fp * Cs);
    }
    resblock(h->srb[i], E[0], F1, F1, false, 0, "source_resblocks." + std::to_string(i), E[3]);
    // upsampling + fuse
    {
      EpiSpec es; es.len_mul = lm; es.len_add = la;
      es.res = F1; es.raw = F2; es.reflect_front = (i == 2);
      for (int j = 0; j < 3; ++j)     // (a chained block applies its own first Snake to the F2 tile it loads anyway)
        if (!chain_block(h->rb[3 * i + j], i, j, true)) es.acts.push_back({snake, h->rb[3 * i + j].a1[0], 0.f, E[j]});
      add(ops, h->ups[i], X, Lx, es, "ups." + std::to_string(i));
    }
    for (int j = 0; j < 3; ++j)       // fork mode: every ResBlock of the stage has its own residual stream and ping-pong buffer
      resblock(h->rb[3 * i + j], E[j], F2, j == 0 ? F1 : Fp(w.F1x[i][j - 1]), true, j, "resblocks." + std::to_string(3 * i + j),
               j == 0 ? E[3] : P(w.E3x[i][j - 1]));
    X = E[0];
    Lx = Ls;
  }
  {
    EpiSpec es; es.len_mul = 120; es.len_add = 1;
    es.raw = Fp(w.P);
    es.c_pitch = kPostPitch;
    add(ops, h->conv_post, X, Lx, es, "conv_post");
  }
  if (!e.empty()) return e;
  for (ConvOp& op : plan->decode_ops) {
    for (int i = 0; i < 3; ++i) {
      const std::string si = std::to_string(i);
      if (op.name == "source_downs." + si || op.name.rfind("source_resblocks." + si + ".", 0) == 0) { op.branch = 1; op.stage = i; }
      if (op.name == "ups." + si) op.stage = i;
    }
    if (op.name.rfind("resblocks.", 0) == 0) {
      const int N = atoi(op.name.c_str() + 10);
      op.stage = N / 3; op.rb = N % 3;
      const std::string tail = op.name.substr(op.name.rfind('.') + 1);
      op.rb_last = tail == "pair2" || tail == "chain" || (tail == "2" && op.name.find(".convs2.") != std::string::npos);
    }
  }
  std::vector<ConvOp*> all;
  for (ConvOp& op : plan->f0_ops) all.push_back(&op);
  for (ConvOp& op : plan->decode_ops) all.push_back(&op);
  if (h->launch_counts.size() > 4096) h->launch_counts.clear();
  h->launch_counts[{B, T}] = (int)plan->decode_ops.size();
  const size_t free_before = h->free_slots.size();
  const size_t chunks_before = h->slot_chunks.dev.size();
  std::string ue = upload_maps(all, h->free_slots, &plan->slot, st, &h->slot_chunks);
  h->slots_allocated += (h->slot_chunks.dev.size() - chunks_before) * kSlotChunk;
  (void)free_before;
  return ue;
}

// Plan cache: LRU over (B, T, workspace), bounded by h->max_plans.  The reference service decodes one sentence at a
// time (services/tts/server.py:118-182), so nearly every call brings a new T: a miss must be cheap and must not grow
// the process.  An evicted plan hands its tensor-map slot to the next plan built.
int get_plan(gnv_handle h, int B, int T, void* ws, size_t ws_bytes, cudaStream_t st, std::shared_ptr<Plan>* out) {
  if (B <= 0 || T <= 0) return fail(h, "B and T must be positive");
  if (!ws) return fail(h, "workspace is NULL");
  if (((uintptr_t)ws & 1023) != 0) return fail(h, "workspace must be 1024-byte aligned");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  std::lock_guard<std::mutex> lk(h->mu);
  PlanKey key(B, T, ws);
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    if (cap != cudaStreamCaptureStatusNone)
      return fail(h, "the first call for a new (B, T, workspace) builds its launch plan and cannot run inside a stream "
                     "capture: call once outside the capture first");
    while (h->plans.size() >= h->max_plans) {
      auto victim = h->plans.end();
      for (auto j = h->plans.begin(); j != h->plans.end(); ++j)
        if (!j->second->pinned && j->second.use_count() == 1 &&
            (victim == h->plans.end() || j->second->last_tick < victim->second->last_tick))
          victim = j;
      if (victim == h->plans.end()) break;             // everything is pinned or in use: let the cache grow
      if (victim->second->slot.d) h->free_slots.push_back(victim->second->slot);
      h->plans.erase(victim);
    }
    auto p = std::make_shared<Plan>();
    const auto t0 = std::chrono::steady_clock::now();
    if (!h->upload_st) {
      cudaError_t ce = cudaStreamCreateWithFlags(&h->upload_st, cudaStreamNonBlocking);
      if (ce != cudaSuccess) return fail_cuda(h, "upload stream", ce);
    }
    // the tensor maps go up on the handle's own stream and are waited for (about 20 us of a 0.4 ms build): not behind
    // whatever the caller's stream still has queued, and visible to every stream that uses the plan afterwards
    std::string e = build_plan(h, B, T, ws, p.get(), h->upload_st);
    if (!e.empty()) {
      if (p->slot.d) h->free_slots.push_back(p->slot);
      return fail(h, e);
    }
    ++h->plans_built;
    if (getenv("GONOVA_PLAN_TIMING"))
      fprintf(stderr, "[gonova] plan B=%d T=%d built in %.0f us (%zu cached, %llu slots)\n", B, T,
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(),
              h->plans.size() + 1, h->slots_allocated);
    it = h->plans.emplace(key, std::move(p)).first;
  }
  if (ws_bytes < it->second->lay.total) return fail(h, "workspace too small for (B, T)");
  it->second->last_tick = ++h->tick;
  if (cap != cudaStreamCaptureStatusNone) it->second->pinned = true;
  *out = it->second;
  return 0;
}

// After the launches of one call: marks the plan's tensor-map slot as in use up to this point of the stream.
void plan_used(Plan* plan, cudaStream_t st) {
  if (!plan->slot.last_use || plan->pinned) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
  cudaEventRecord(plan->slot.last_use, st);
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

#define GNV_CK(h, what, expr)                                  \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return fail_cuda(h, what, _e);      \
  } while (0)

int run_f0(gnv_handle h, Plan* plan, const float* mel, const int* lengths, int B, int T, float* f0, char* ws,
           bool pack_mel, cudaStream_t st, Profiler* prof = nullptr) {
  const WsLayout& w = plan->lay;
  if (pack_mel) {
    GNV_CK(h, "pack mel", launch_nct_to_nlc(mel, B, kMelC, T, lengths, ws + w.melE, h->conv_pre.C_in_ld, h->eb,
                                            h->dtype == GNV_DTYPE_TF32, st));
    prof_mark(prof, "pack_mel", GNV_LAUNCH_AUX);
  }
  for (const ConvOp& op : plan->f0_ops) {
    GNV_CK(h, "f0 conv", run_op(op, lengths, st));
    prof_mark(prof, op.name.c_str(), op.tc ? GNV_LAUNCH_CONV_TC : GNV_LAUNCH_CONV_SIMT, op.flops);
  }
  const void* hl = ws + (plan->f0_ops.size() % 2 ? w.H0 : w.H1);
  GNV_CK(h, "f0 head", launch_f0_head(hl, h->eb, B * T, T, lengths, 512, h->f0_w, h->f0_b, f0, st));
  prof_mark(prof, "f0_predictor.classifier", GNV_LAUNCH_AUX, 2.0 * B * T * 512);
  return 0;
}

int run_decode(gnv_handle h, Plan* plan, const float* mel, const float* s, const int* lengths, int B, int T,
               float* wav, char* ws, bool pack_mel, cudaStream_t st, Profiler* prof = nullptr) {
  const WsLayout& w = plan->lay;
  if (pack_mel) {
    GNV_CK(h, "pack mel", launch_nct_to_nlc(mel, B, kMelC, T, lengths, ws + w.melE, h->conv_pre.C_in_ld, h->eb,
                                            h->dtype == GNV_DTYPE_TF32, st));
    prof_mark(prof, "pack_mel", GNV_LAUNCH_AUX);
  }
  GNV_CK(h, "stft", launch_stft(s, B, T * kSPF, lengths, ws + w.spec, h->eb, h->dtype == GNV_DTYPE_TF32, h->spec_cs,
                                kSpecFront, 120 * T + 1 + kSpecFront + kSpecBack, st));
  prof_mark(prof, "stft", GNV_LAUNCH_AUX);
  // Small problems are bound by the chain of ~75 dependent launches, not by the SMs.  The source branch of a stage
  // (source_downs.i + source_resblocks.i: 18 launches in all) depends on the STFT only and meets the main path at
  // ups.i, so it runs on a side stream beside conv_pre and the previous stages' ResBlocks.  Stage buffers are distinct,
  // and the fork / join events are ordinary stream dependencies, also under stream capture.
  const bool fork = fork_mode(h, B, T) && !prof;          // (a profiled run times launches on one stream)
  if (fork) {
    if (!h->side) {
      std::lock_guard<std::mutex> lk(h->mu);
      if (!h->side) {
        cudaStream_t sd = nullptr;
        for (int i = 0; i < 2; ++i) GNV_CK(h, "side stream", cudaStreamCreateWithFlags(&h->rbs[i], cudaStreamNonBlocking));
        GNV_CK(h, "fork event", cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        GNV_CK(h, "fork event", cudaEventCreateWithFlags(&h->ev_ups, cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) GNV_CK(h, "join event", cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) GNV_CK(h, "join event", cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming));
        GNV_CK(h, "side stream", cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
        h->side = sd;
      }
    }
    const std::vector<ConvOp>& ops = plan->decode_ops;
    const size_t n_ops = ops.size();
    auto launch = [&](const ConvOp& op, cudaStream_t s2) -> int {
      GNV_CK(h, "conv", run_op(op, lengths, s2));
      prof_mark(prof, op.name.c_str(), op.tc ? GNV_LAUNCH_CONV_TC : GNV_LAUNCH_CONV_SIMT, op.flops);
      return 0;
    };
    // (a) the source branch of all three stages, on its own stream, right behind the STFT
    GNV_CK(h, "fork", cudaEventRecord(h->ev_fork, st));
    GNV_CK(h, "fork", cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    for (size_t k = 0; k < n_ops; ++k) {
      const ConvOp& op = ops[k];
      if (op.branch != 1) continue;
      if (int rc = launch(op, h->side)) return rc;
      bool last_of_stage = true;
      for (size_t m = k + 1; m < n_ops; ++m)
        if (ops[m].branch == 1 && ops[m].stage == op.stage) { last_of_stage = false; break; }
      if (last_of_stage) GNV_CK(h, "join", cudaEventRecord(h->ev_join[op.stage], h->side));
    }
    // (b) the main path; the three ResBlocks of a stage side by side up to the launch that adds each into the running
    // sum F3 — those three stay on the caller's stream in the order 0, 1, 2 (same arithmetic as the serial schedule)
    for (size_t k = 0; k < n_ops;) {
      const ConvOp& op = ops[k];
      if (op.branch == 1) { ++k; continue; }
      if (op.rb < 0) {
        if (op.stage >= 0) GNV_CK(h, "join", cudaStreamWaitEvent(st, h->ev_join[op.stage], 0));   // ups.i adds the source branch
        if (int rc = launch(op, st)) return rc;
        ++k;
        continue;
      }
      // ops[k ...) up to the end of this stage's ResBlocks
      std::vector<const ConvOp*> pre[3];
      const ConvOp* last[3] = {nullptr, nullptr, nullptr};
      const int stage = op.stage;
      size_t m = k;
      for (; m < n_ops && ops[m].branch == 0 && ops[m].rb >= 0 && ops[m].stage == stage; ++m) {
        if (ops[m].rb_last) last[ops[m].rb] = &ops[m]; else pre[ops[m].rb].push_back(&ops[m]);
      }
      GNV_CK(h, "fork", cudaEventRecord(h->ev_ups, st));
      for (int j = 1; j < 3; ++j) GNV_CK(h, "fork", cudaStreamWaitEvent(h->rbs[j - 1], h->ev_ups, 0));
      const size_t depth = std::max(pre[0].size(), std::max(pre[1].size(), pre[2].size()));
      for (size_t d = 0; d < depth; ++d)                    // round robin, so that the host feeds all three streams early
        for (int j = 0; j < 3; ++j)
          if (d < pre[j].size())
            if (int rc = launch(*pre[j][d], j == 0 ? st : h->rbs[j - 1])) return rc;
      for (int j = 1; j < 3; ++j) GNV_CK(h, "join", cudaEventRecord(h->ev_pre[j - 1], h->rbs[j - 1]));
      for (int j = 0; j < 3; ++j) {
        if (j > 0) GNV_CK(h, "join", cudaStreamWaitEvent(st, h->ev_pre[j - 1], 0));
        if (last[j]) if (int rc = launch(*last[j], st)) return rc;
      }
      k = m;
    }
  } else {
    for (const ConvOp& op : plan->decode_ops) {
      GNV_CK(h, "conv", run_op(op, lengths, st));
      prof_mark(prof, op.name.c_str(), op.tc ? GNV_LAUNCH_CONV_TC : GNV_LAUNCH_CONV_SIMT, op.flops);
    }
  }
  GNV_CK(h, "istft", launch_istft((const float*)(ws + w.P), B, 120 * T + 1, kPostPitch, lengths, 0.99f, wav, st));
  prof_mark(prof, "istft_head", GNV_LAUNCH_AUX);
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int gnv_abi_version(void) { return GNV_ABI_VERSION; }

const char* gnv_last_error(gnv_handle h) {
  // always a thread-local string: another thread may be rewriting h->err
  if (h) {
    std::lock_guard<std::mutex> lk(h->err_mu);
    tl_error_ret = h->err;
    return tl_error_ret.c_str();
  }
  return tl_error.c_str();
}

void gnv_destroy(gnv_handle h) {
  if (!h) return;
  {
    DeviceGuard dg(h->device);
    for (auto& kv : h->plans) free_slot(kv.second->slot);
    for (MapsSlot& sl : h->free_slots) free_slot(sl);
    for (void* q : h->slot_chunks.dev) cudaFree(q);
    for (void* q : h->slot_chunks.host) cudaFreeHost(q);
    if (h->upload_st) cudaStreamDestroy(h->upload_st);
    if (h->side) cudaStreamDestroy(h->side);
    for (int i = 0; i < 2; ++i) if (h->rbs[i]) cudaStreamDestroy(h->rbs[i]);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_ups) cudaEventDestroy(h->ev_ups);
    for (int i = 0; i < 3; ++i) if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    for (int i = 0; i < 2; ++i) if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
    for (void* p : h->allocs) cudaFree(p);
  }
  delete h;
}

int gnv_create(const GnvWeight* weights, int n_weights, int device, int dtype, unsigned flags, gnv_handle* out) {
  if (!out) return fail(nullptr, "out handle is NULL");
  *out = nullptr;
  if (!weights || n_weights <= 0) return fail(nullptr, "no weights given");
  if (dtype != GNV_DTYPE_TF32 && dtype != GNV_DTYPE_BF16 && dtype != GNV_DTYPE_FP32)
    return fail(nullptr, "unknown dtype");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess) return fail_cuda(nullptr, "cudaGetDeviceCount", ce);
  if (device < 0 || device >= ndev) return fail(nullptr, "no such CUDA device");
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, device);
  if (ce != cudaSuccess) return fail_cuda(nullptr, "cudaGetDeviceProperties", ce);
  // the library holds sm_100a cubins only (arch-specific code is not forward compatible, not even to sm_103)
  if (prop.major != 10 || prop.minor != 0) return fail(nullptr, "libgonova_hift is built for sm_100a (B200) only; found sm_" +
                                                 std::to_string(prop.major) + std::to_string(prop.minor));
  DeviceGuard dg(device);
  if (!dg.ok) return fail(nullptr, "cudaSetDevice failed");

  WeightMap wm;
  for (int i = 0; i < n_weights; ++i) {
    const GnvWeight& g = weights[i];
    if (!g.name || !g.data || g.ndim < 1 || g.ndim > 4) return fail(nullptr, "malformed GnvWeight entry");
    HostT t;
    t.data = g.data;
    for (int d = 0; d < g.ndim; ++d) t.shape.push_back(g.shape[d]);
    wm[g.name] = t;
  }

  gnv_decoder* h = new gnv_decoder();
  h->device = device; h->dtype = dtype; h->flags = flags;
  h->eb = dtype == GNV_DTYPE_BF16 ? 2 : 4;
  h->spec_cs = dtype == GNV_DTYPE_BF16 ? 24 : 20;
  h->use_tc = dtype != GNV_DTYPE_FP32 && !(flags & GNV_FLAG_SIMT_CONV);
  h->tc_version = (flags & GNV_FLAG_TC_V1) ? 1 : 2;
  h->tc2opt = tc2_options_from_env();
  if (const char* v = getenv("GONOVA_FUSE_PAIRS")) h->fuse_pairs = atoi(v) != 0;
  if (const char* v = getenv("GONOVA_FUSE_MAX_C")) h->fuse_max_c = atoi(v);
  if (const char* v = getenv("GONOVA_FUSE_K3_MAX_C")) h->fuse_k3_max_c = atoi(v);
  if (const char* v = getenv("GONOVA_PAIR_CTA2")) h->pair_cta2 = atoi(v);
  if (const char* v = getenv("GONOVA_CHAIN_MAX_K")) h->chain_max_k = atoi(v);
  if (const char* v = getenv("GONOVA_CHAIN_MAX_C")) h->chain_max_c = atoi(v);
  if (const char* v = getenv("GONOVA_FORK_MAX_FRAMES")) h->fork_max_frames = atoi(v);
  if (const char* v = getenv("GONOVA_MAX_PLANS")) h->max_plans = (size_t)(atoi(v) > 0 ? atoi(v) : 1);
  h->snake_kind = (dtype == GNV_DTYPE_FP32 || (flags & GNV_FLAG_PRECISE_ACT)) ? ACT_SNAKE : ACT_SNAKE_FAST;
  Uploader up{h};
  std::string err;
  bool ok = true;
  ok = ok && pack_layer(h, up, wm, "conv_pre", h->conv_pre, 80, 512, 7, 1, 3, 1, false, false, &err);
  for (int i = 0; ok && i < 3; ++i) {
    const std::string si = std::to_string(i);
    const int cin = 512 >> i, cout = 256 >> i;
    ok = ok && pack_layer(h, up, wm, "ups." + si, h->ups[i], cin, cout, kUpK[i], kUpRate[i],
                          (kUpK[i] - kUpRate[i]) / 2, 1, true, false, &err);
    ok = ok && pack_layer(h, up, wm, "source_downs." + si, h->sdown[i], 18, cout, kSdK[i], kSdS[i], kSdP[i], 1, false,
                          true, &err);
    ok = ok && pack_flat(h, up, wm, "source_downs." + si, h->sdown_flat[i], 18, cout, kSdK[i], kSdS[i], kSdP[i], &err);
    ok = ok && pack_resblock(h, up, wm, "source_resblocks." + si, h->srb[i], cout, kSrbK[i], &err);
    for (int j = 0; ok && j < 3; ++j)
      ok = ok && pack_resblock(h, up, wm, "resblocks." + std::to_string(3 * i + j), h->rb[3 * i + j], cout, kRbK[j],
                               &err);
  }
  ok = ok && pack_layer(h, up, wm, "conv_post", h->conv_post, 64, 18, 7, 1, 3, 1, false, false, &err);
  for (int i = 0; ok && i < 5; ++i)
    ok = ok && pack_layer(h, up, wm, "f0_predictor.condnet." + std::to_string(2 * i), h->f0c[i], i == 0 ? 80 : 512, 512,
                          3, 1, 1, 1, false, false, &err);
  ok = ok && put_vec(up, wm, "f0_predictor.classifier.weight", 512, &h->f0_w, &err);
  ok = ok && put_vec(up, wm, "f0_predictor.classifier.bias", 1, &h->f0_b, &err);
  ok = ok && put_vec(up, wm, "m_source.l_linear.weight", 9, &h->lin_w, &err);
  ok = ok && put_vec(up, wm, "m_source.l_linear.bias", 1, &h->lin_b, &err);
  if (!ok || !up.ok) {
    std::string m = !err.empty() ? err : up.msg;
    gnv_destroy(h);
    return fail(nullptr, "gnv_create: " + m);
  }
  if (h->use_tc) {
    ce = conv_tc_init();
    if (ce == cudaSuccess) ce = conv_tc2_init();
    if (ce == cudaSuccess) ce = conv_pair_init();
    if (ce == cudaSuccess) ce = conv_chain_init();
    if (ce != cudaSuccess) {
      gnv_destroy(h);
      return fail_cuda(nullptr, "conv_tc_init", ce);
    }
    if (!get_encode_tiled()) {
      gnv_destroy(h);
      return fail(nullptr, "driver does not export cuTensorMapEncodeTiled");
    }
  }
  *out = h;
  return 0;
}

int gnv_workspace_bytes(gnv_handle h, int B, int T, size_t* out_bytes) {
  if (!h || !out_bytes) return fail(h, "NULL argument");
  if (B <= 0 || T <= 0) return fail(h, "B and T must be positive");
  *out_bytes = make_layout(h, B, T).total;
  return 0;
}

int gnv_f0(gnv_handle h, const float* mel, const int32_t* lengths, int B, int T, float* f0, void* workspace,
           size_t workspace_bytes, void* stream) {
  if (!h || !mel || !f0) return fail(h, "NULL argument");
  DeviceGuard dg(h->device);
  std::shared_ptr<Plan> plan_ref;
  if (int rc = get_plan(h, B, T, workspace, workspace_bytes, (cudaStream_t)stream, &plan_ref)) return rc;
  Plan* plan = plan_ref.get();
  const int rc = run_f0(h, plan, mel, lengths, B, T, f0, (char*)workspace, true, (cudaStream_t)stream);
  plan_used(plan, (cudaStream_t)stream);
  return rc;
}

int gnv_source(gnv_handle h, const float* f0, int B, int T, uint64_t seed, const float* phase_vec, const float* noise,
               float* s, void* stream) {
  if (!h || !f0 || !s) return fail(h, "NULL argument");
  if (B <= 0 || T <= 0) return fail(h, "B and T must be positive");
  DeviceGuard dg(h->device);
  GNV_CK(h, "source", launch_source(f0, B, T, seed, phase_vec, noise, h->lin_w, h->lin_b, s, (cudaStream_t)stream));
  return 0;
}

int gnv_source_stream(gnv_handle h, const float* f0, int B, int T, uint64_t seed, int64_t sample0, const double* f0_sum0,
                      float* s, double* f0_sum_out, void* stream) {
  if (!h || !f0 || !s) return fail(h, "NULL argument");
  if (B <= 0 || T <= 0 || sample0 < 0) return fail(h, "B and T must be positive, sample0 non-negative");
  DeviceGuard dg(h->device);
  GNV_CK(h, "source", launch_source(f0, B, T, seed, nullptr, nullptr, h->lin_w, h->lin_b, s, (cudaStream_t)stream,
                                    f0_sum0, (long long)sample0, f0_sum_out));
  return 0;
}

int gnv_decode(gnv_handle h, const float* mel, const float* s, const int32_t* lengths, int B, int T, float* wav,
               void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !mel || !s || !wav) return fail(h, "NULL argument");
  DeviceGuard dg(h->device);
  std::shared_ptr<Plan> plan_ref;
  if (int rc = get_plan(h, B, T, workspace, workspace_bytes, (cudaStream_t)stream, &plan_ref)) return rc;
  Plan* plan = plan_ref.get();
  const int rc = run_decode(h, plan, mel, s, lengths, B, T, wav, (char*)workspace, true, (cudaStream_t)stream);
  plan_used(plan, (cudaStream_t)stream);
  return rc;
}

static int inference_impl(gnv_handle h, const float* mel, const float* cache_source, int cache_len,
                          const int32_t* lengths, int B, int T, uint64_t seed, float* wav, float* s_out,
                          void* workspace, size_t workspace_bytes, cudaStream_t st, Profiler* prof,
                          uint64_t* seed_dev = nullptr, int seed_per_row = 0) {
  if (!h || !mel || !wav || !s_out) return fail(h, "NULL argument");
  if (cache_len < 0 || (cache_len > 0 && !cache_source)) return fail(h, "bad cache_source");
  DeviceGuard dg(h->device);
  std::shared_ptr<Plan> plan_ref;
  if (int rc = get_plan(h, B, T, workspace, workspace_bytes, st, &plan_ref)) return rc;
  Plan* plan = plan_ref.get();
  struct Used { Plan* p; cudaStream_t s; ~Used() { plan_used(p, s); } } used{plan, st};
  char* ws = (char*)workspace;
  float* f0 = (float*)(ws + plan->lay.f0);
  if (prof) prof->begin(st);
  if (int rc = run_f0(h, plan, mel, lengths, B, T, f0, ws, true, st, prof)) return rc;
  GNV_CK(h, "source", launch_source(f0, B, T, seed, nullptr, nullptr, h->lin_w, h->lin_b, s_out, st, nullptr, 0, nullptr,
                                    seed_dev, seed_per_row));
  prof_mark(prof, "m_source", GNV_LAUNCH_AUX);
  const size_t L = (size_t)T * kSPF;
  if (cache_len > 0) {
    const size_t n = (size_t)cache_len < L ? (size_t)cache_len : L;
    GNV_CK(h, "cache_source copy", cudaMemcpy2DAsync(s_out, L * 4, cache_source, (size_t)cache_len * 4, n * 4, B,
                                                     cudaMemcpyDeviceToDevice, st));
    prof_mark(prof, "cache_source_copy", GNV_LAUNCH_AUX);
  }
  return run_decode(h, plan, mel, s_out, lengths, B, T, wav, ws, false, st, prof);
}

int gnv_inference(gnv_handle h, const float* mel, const float* cache_source, int cache_len, const int32_t* lengths,
                  int B, int T, uint64_t seed, float* wav, float* s_out, void* workspace, size_t workspace_bytes,
                  void* stream) {
  return inference_impl(h, mel, cache_source, cache_len, lengths, B, T, seed, wav, s_out, workspace, workspace_bytes,
                        (cudaStream_t)stream, nullptr);
}

int gnv_inference_dseed(gnv_handle h, const float* mel, const float* cache_source, int cache_len, const int32_t* lengths,
                        int B, int T, uint64_t* seed_dev, int per_row, float* wav, float* s_out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!seed_dev) return fail(h, "seed_dev is NULL");
  return inference_impl(h, mel, cache_source, cache_len, lengths, B, T, 0, wav, s_out, workspace, workspace_bytes,
                        (cudaStream_t)stream, nullptr, seed_dev, per_row ? 1 : 0);
}

int gnv_inference_profile(gnv_handle h, const float* mel, const int32_t* lengths, int B, int T, uint64_t seed,
                          float* wav, float* s_out, void* workspace, size_t workspace_bytes, void* stream,
                          int capacity, float* ms_out, int32_t* kind_out, double* flops_out, char* names_out,
                          int* n_out) {
  if (!ms_out || !kind_out || !flops_out || !names_out || !n_out) return fail(h, "NULL argument");
  Profiler prof;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = inference_impl(h, mel, nullptr, 0, lengths, B, T, seed, wav, s_out, workspace, workspace_bytes, st,
                              &prof))
    return rc;
  if (prof.err != cudaSuccess) return fail_cuda(h, "profile events", prof.err);
  GNV_CK(h, "profile sync", cudaStreamSynchronize(st));
  const int n = (int)prof.names.size();
  if (n > capacity) return fail(h, "profile arrays too small");
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    GNV_CK(h, "cudaEventElapsedTime", cudaEventElapsedTime(&ms, prof.ev[i], prof.ev[i + 1]));
    ms_out[i] = ms;
    kind_out[i] = prof.kinds[i];
    flops_out[i] = prof.flops[i];
    snprintf(names_out + (size_t)i * GNV_LAUNCH_NAME_LEN, GNV_LAUNCH_NAME_LEN, "%s", prof.names[i].c_str());
  }
  *n_out = n;
  return 0;
}

int gnv_pcm_tail(const float* cur, int64_t cur_stride, const float* prev_tail, const float* fade_w, int rows, int n,
                 int fade, float limit, int16_t* out_i16, float* out_f32, int64_t out_stride, void* stream) {
  if (!cur || (!out_i16 && !out_f32)) return fail(nullptr, "NULL argument");
  if (rows < 0 || n < 0 || fade < 0) return fail(nullptr, "negative size");
  if (fade > n) fade = n;
  cudaError_t e = launch_pcm_tail(cur, cur_stride, prev_tail, fade_w, rows, n, fade, limit, out_i16, out_f32,
                                  out_stride, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(nullptr, "pcm_tail", e);
  return 0;
}

int gnv_pcm_mulaw(const int16_t* pcm, int64_t n, uint8_t* out, void* stream) {
  if (!pcm || !out) return fail(nullptr, "NULL argument");
  if (n < 0) return fail(nullptr, "negative size");
  cudaError_t e = launch_mulaw(pcm, (long long)n, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(nullptr, "pcm_mulaw", e);
  return 0;
}

// ---- unit-test hooks ----------------------------------------------------------------------------
namespace {
struct Scratch {
  std::vector<void*> p;
  ~Scratch() { for (void* q : p) cudaFree(q); }
  void* get(size_t bytes, cudaError_t* e) {
    void* d = nullptr;
    *e = cudaMalloc(&d, align_up(bytes ? bytes : 1, 1024));
    if (*e == cudaSuccess) p.push_back(d);
    return d;
  }
};
}  // namespace

int gnv_stft(const float* s, int B, int L, float* spec_nct, void* stream) {
  if (!s || !spec_nct || B <= 0 || L < 16 || L % 4) return fail(nullptr, "gnv_stft: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int F = L / 4 + 1;
  Scratch sc;
  cudaError_t e;
  float* nlc = (float*)sc.get((size_t)B * F * 18 * 4, &e);
  if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
  GNV_CK(nullptr, "stft", launch_stft(s, B, L, nullptr, nlc, 4, 0, 18, 0, F, st));
  GNV_CK(nullptr, "unpack", launch_nlc_to_nct(nlc, B, F, 18, 18, 4, spec_nct, st));
  GNV_CK(nullptr, "sync", cudaStreamSynchronize(st));
  return 0;
}

int gnv_istft(const float* x_nct, int B, int F, float limit, float* wav, void* stream) {
  if (!x_nct || !wav || B <= 0 || F < 2) return fail(nullptr, "gnv_istft: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  cudaError_t e;
  float* nlc = (float*)sc.get((size_t)B * F * 18 * 4, &e);
  if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
  GNV_CK(nullptr, "pack", launch_nct_to_nlc(x_nct, B, 18, F, nullptr, nlc, 18, 4, 0, st));
  GNV_CK(nullptr, "istft", launch_istft(nlc, B, F, 18, nullptr, limit, wav, st));
  GNV_CK(nullptr, "sync", cudaStreamSynchronize(st));
  return 0;
}

int gnv_conv1d(int device, int dtype, unsigned flags, int transposed, const float* x, int B, int Cin, int Lin,
               const float* w_host, const float* bias_host, int Cout, int k, int stride, int pad, int dil, int act,
               const float* alpha_host, float slope, const float* res_nct, float* y_nct, int Lout, void* stream) {
  if (!x || !w_host || !y_nct) return fail(nullptr, "gnv_conv1d: NULL argument");
  if (dtype != GNV_DTYPE_TF32 && dtype != GNV_DTYPE_BF16 && dtype != GNV_DTYPE_FP32)
    return fail(nullptr, "unknown dtype");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(nullptr, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  gnv_decoder tmp;
  tmp.device = device; tmp.dtype = dtype; tmp.flags = flags;
  tmp.eb = dtype == GNV_DTYPE_BF16 ? 2 : 4;
  const bool strided = !transposed && stride != 1;
  tmp.use_tc = dtype != GNV_DTYPE_FP32 && !(flags & GNV_FLAG_SIMT_CONV) && !strided;
  tmp.tc_version = (flags & GNV_FLAG_TC_V1) ? 1 : 2;
  tmp.tc2opt = tc2_options_from_env();
  const int Cp = (Cout + 3) & ~3;             // channel pitch of the outputs: rows are 16-byte multiples
  struct Cleanup {
    gnv_decoder* d;
    ~Cleanup() { for (void* p : d->allocs) cudaFree(p); }
  } cleanup{&tmp};
  if (tmp.use_tc) {
    cudaError_t ce = conv_tc_init();
    if (ce == cudaSuccess) ce = conv_tc2_init();
    if (ce != cudaSuccess) return fail_cuda(nullptr, "conv_tc_init", ce);
  }
  std::vector<float> zero_bias((size_t)Cout, 0.f);
  WeightMap wm;
  HostT tw, tb;
  tw.data = w_host;
  tw.shape = transposed ? std::vector<int64_t>{Cin, Cout, k} : std::vector<int64_t>{Cout, Cin, k};
  tb.data = bias_host ? bias_host : zero_bias.data();
  tb.shape = {Cout};
  wm["l.weight"] = tw;
  wm["l.bias"] = tb;
  Uploader up{&tmp};
  ConvLayer L;
  std::string err;
  if (!pack_layer(&tmp, up, wm, "l", L, Cin, Cout, k, stride, pad, dil, transposed != 0, strided, &err))
    return fail(nullptr, "gnv_conv1d: " + (err.empty() ? up.msg : err));
  const int expect = transposed ? Lin * stride : (Lin + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  if (transposed && (k - stride) != 2 * pad) return fail(nullptr, "gnv_conv1d: convT needs pad == (k - stride) / 2");
  if (expect != Lout) return fail(nullptr, "gnv_conv1d: Lout does not match the layer geometry");
  float* alpha_d = nullptr;
  if (alpha_host) alpha_d = (float*)up.put(alpha_host, (size_t)Cout * 4);
  if (!up.ok) return fail(nullptr, up.msg);
  Scratch sc;
  cudaError_t e;
  const int a_eb = tmp.eb;
  void* A = sc.get((size_t)B * Lin * L.C_in_ld * a_eb, &e);
  if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
  float* raw = (float*)sc.get((size_t)B * Lout * Cp * 4, &e);
  if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
  void* actb = sc.get((size_t)B * Lout * Cp * tmp.eb, &e);
  if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
  float* res = nullptr;
  if (res_nct) {
    res = (float*)sc.get((size_t)B * Lout * Cp * 4, &e);
    if (e != cudaSuccess) return fail_cuda(nullptr, "cudaMalloc", e);
    GNV_CK(nullptr, "pack res", launch_nct_to_nlc(res_nct, B, Cout, Lout, nullptr, res, Cp, 4, 0, st));
  }
  GNV_CK(nullptr, "pack x", launch_nct_to_nlc(x, B, Cin, Lin, nullptr, A, L.C_in_ld, a_eb,
                                              dtype == GNV_DTYPE_TF32 ? 1 : 0, st));
  EpiSpec es;
  es.res = res;
  es.raw = (act != GNV_ACT_NONE && (flags & GNV_FLAG_HOOK_NO_RAW)) ? nullptr : raw;
  es.c_pitch = Cp;
  if (act != GNV_ACT_NONE) es.acts.push_back({act, alpha_d, slope, actb});
  ConvOp op;
  std::string me = make_op(&tmp, L, A, B, Lin, es, &op);
  if (!me.empty()) return fail(nullptr, "gnv_conv1d: " + me);
  std::vector<ConvOp*> one{&op};
  std::vector<MapsSlot> no_free;
  MapsSlot slot;
  struct SlotGuard { MapsSlot* s; ~SlotGuard() { free_slot(*s); } } slot_guard{&slot};   // the hook syncs before returning
  me = upload_maps(one, no_free, &slot, st);
  if (!me.empty()) return fail(nullptr, "gnv_conv1d: " + me);
  GNV_CK(nullptr, "conv", run_op(op, nullptr, st));
  if (act != GNV_ACT_NONE)
    GNV_CK(nullptr, "unpack", launch_nlc_to_nct(actb, B, Lout, Cout, Cp, tmp.eb, y_nct, st));
  else
    GNV_CK(nullptr, "unpack", launch_nlc_to_nct(raw, B, Lout, Cout, Cp, 4, y_nct, st));
  GNV_CK(nullptr, "sync", cudaStreamSynchronize(st));
  return 0;
}

int gnv_debug_tap(gnv_handle h, const char* name, int B, int T, void* workspace, float* out_nct,
                  size_t out_capacity_elems, int64_t* out_shape3, void* stream) {
  if (!h || !name || !workspace || !out_nct || !out_shape3) return fail(h, "NULL argument");
  DeviceGuard dg(h->device);
  const WsLayout w = make_layout(h, B, T);
  char* ws = (char*)workspace;
  const std::string n(name);
  const float* src = nullptr;
  int L = 0, C = 0, Cld = 0;
  const int F = 120 * T + 1;
  if (n == "s_stft") {
    if ((size_t)B * F * 18 > out_capacity_elems) return fail(h, "tap output buffer too small");
    out_shape3[0] = B; out_shape3[1] = 18; out_shape3[2] = F;
    const int Cs = h->spec_cs;
    GNV_CK(h, "tap", launch_nlc_to_nct(ws + w.spec + (size_t)kSpecFront * Cs * h->eb, B, F, 18, Cs, h->eb, out_nct,
                                       (cudaStream_t)stream, (long long)(F + kSpecFront + kSpecBack) * Cs));
    return 0;
  }
  else if (n == "conv_post") { src = (const float*)(ws + w.P); L = F; C = 18; Cld = kPostPitch; }
  else if (n.size() == 5 && n.compare(0, 4, "fuse") == 0 && n[4] >= '0' && n[4] <= '2') {
    const int i = n[4] - '0';
    src = (const float*)(ws + w.F2[i]); L = stage_len(i, T); C = kStageC[i];
  } else if (n.size() == 6 && n.compare(0, 5, "stage") == 0 && n[5] >= '0' && n[5] <= '2') {
    const int i = n[5] - '0';
    src = (const float*)(ws + w.F3[i]); L = stage_len(i, T); C = kStageC[i];
  } else {
    return fail(h, "unknown tap '" + n + "'");
  }
  if ((size_t)B * L * C > out_capacity_elems) return fail(h, "tap output buffer too small");
  out_shape3[0] = B; out_shape3[1] = C; out_shape3[2] = L;
  GNV_CK(h, "tap", launch_nlc_to_nct(src, B, L, C, Cld ? Cld : C, 4, out_nct, (cudaStream_t)stream));
  return 0;
}

int gnv_debug_chain_trace(unsigned long long* out, int cap, int* n_out) {
  if (!out || !n_out) return fail(nullptr, "NULL argument");
  const int n = conv_chain_read_trace(out, cap);
  if (n < 0) return fail(nullptr, "chain trace: CUDA error");
  *n_out = n;
  return 0;
}

int gnv_debug_cluster_probe(int smem_bytes, int grid, int* max_clusters) {
  if (!max_clusters) return fail(nullptr, "NULL argument");
  cudaError_t ce = conv_tc2_init();
  if (ce != cudaSuccess) return fail_cuda(nullptr, "conv_tc2_init", ce);
  const char* e = conv_tc2_cluster_probe(smem_bytes, grid, max_clusters);
  if (*e) return fail(nullptr, std::string("cluster probe: ") + e);
  return 0;
}

int gnv_decode_launches(gnv_handle h, int B, int T, int* out) {
  if (!h || !out) return fail(h, "NULL argument");
  // pack mel + STFT + iSTFT head + the conv launches of the plan (unfused: conv_pre + 3 x (source_down + 6 +
  // ups + 18) + conv_post = 80; fused ResBlock pairs in stages 1 and 2 make it 56)
  int convs = 1 + 3 * (1 + 6 + 1 + 18) + 1;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->launch_counts.find({B, T});
    if (it != h->launch_counts.end()) convs = it->second;
    else if (h->use_tc && h->tc_version == 2 && h->fuse_pairs) convs -= 4 * 3 * (h->fuse_max_c >= 128 ? 2 : 1);
  }
  *out = 3 + convs;
  return 0;
}

int gnv_plan_stats(gnv_handle h, uint64_t out[4]) {
  if (!h || !out) return fail(h, "NULL argument");
  std::lock_guard<std::mutex> lk(h->mu);
  out[0] = h->plans.size();
  out[1] = h->plans_built;
  out[2] = h->slots_allocated;
  out[3] = 0;
  for (auto& kv : h->plans) out[3] += kv.second->pinned ? 1 : 0;
  return 0;
}

int gnv_inference_launches(gnv_handle h, int B, int T, int* out) {
  if (!h || !out) return fail(h, "NULL argument");
  int d = 0;
  gnv_decode_launches(h, B, T, &d);
  *out = d + 5 /*f0 convs*/ + 1 /*f0 head*/ + 1 /*source*/;   // the mel pack is shared
  return 0;
}

}  // extern "C"

#include "flow_api.inc"
#include "flow_enc_api.inc"
