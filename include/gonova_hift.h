/*
 * gonova_hift.h — C ABI of libgonova_hift.so, the B200 (sm_100a) waveform decoder.
 *
 * This is the drop-in boundary for ONE path of websines/gonova-tts: the neural vocoder the TTS
 * service reaches through
 *     self.model.generate(...)                     services/tts/core/synthesizer.py:344-350
 *       -> ChatterboxTTS.generate -> S3Gen.inference -> self.mel2wav.inference(speech_feat, cache_source)
 * (the engine is imported at synthesizer.py:167 and constructed at synthesizer.py:185; it is a
 * third-party package that is not vendored, so the callee side is cited by its upstream name).
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; gnv_last_error() gives the text.
 *     Nothing aborts; nothing calls cudaDeviceSynchronize(); once a shape's launch plan exists (first call,
 *     see gnv_plan_stats) nothing allocates inside the decode calls (safe to capture in a CUDA graph).
 *   - the caller owns all input, output and workspace buffers and passes raw DEVICE pointers
 *     (torch: tensor.data_ptr()).  The handle owns only re-packed weights.
 *   - `stream` is a cudaStream_t passed as void* (torch: torch.cuda.current_stream().cuda_stream).  All work is
 *     ordered on it: small batches (B*T <= 4096 frames) also use streams the handle owns for independent branches,
 *     forked from and joined back into `stream` by events, so to the caller — and to a stream capture — a call
 *     still is one piece of work on `stream`.  ONE CALL AT A TIME PER HANDLE: the fork streams and events belong to the
 *     handle, so two threads (or two streams) driving one handle concurrently must serialise their calls themselves
 *     (the Python shim holds a lock); calls on different streams one after the other are fine — a launch plan built
 *     on one stream is waited for by the first call that uses it from another.
 *   - tensors are fp32, contiguous, in the layouts of the upstream PyTorch module:
 *     mel [B,80,T]  source s [B,1,480*T]  wav [B,480*T]  f0 [B,T].
 */
#ifndef GONOVA_HIFT_H_
#define GONOVA_HIFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNV_ABI_VERSION 2

/* arithmetic of the conv GEMMs (activations/weights as stored in HBM; accumulation is fp32) */
#define GNV_DTYPE_TF32 0      /* tcgen05 kind::tf32 — the "fp32 parity" path                       */
#define GNV_DTYPE_BF16 1      /* tcgen05 kind::f16 (bf16) — the throughput path                    */
#define GNV_DTYPE_FP32 2      /* exact fp32 FMA on CUDA cores (validation path, no tensor cores)   */

/* gnv_create flags */
#define GNV_FLAG_SIMT_CONV   1u /* run the conv layers on the CUDA-core kernel in the chosen dtype */
#define GNV_FLAG_PRECISE_ACT 2u /* Snake with libdevice sinf instead of MUFU.SIN (always on for FP32) */
#define GNV_FLAG_HOOK_NO_RAW 8u /* gnv_conv1d only: do not write the fp32 (pre-activation) output tensor */
#define GNV_FLAG_TC_V1       4u /* use the simple one-tile-per-CTA tcgen05 kernel instead of the persistent one */

/* activation kinds accepted by gnv_conv1d (unit-test hook) */
#define GNV_ACT_NONE  0
#define GNV_ACT_SNAKE 1
#define GNV_ACT_LRELU 2
#define GNV_ACT_ELU   3
#define GNV_ACT_SNAKE_FAST 4  /* Snake through MUFU.SIN, what the bf16 / tf32 decode uses */

typedef struct gnv_decoder* gnv_handle;

/* One named host tensor of the decoder's effective (weight-norm already folded) parameters.
 * `name` is the upstream module path below `mel2wav.` — e.g. "conv_pre.weight", "ups.0.bias",
 * "resblocks.4.convs1.2.weight", "resblocks.4.activations2.0.alpha",
 * "source_downs.1.weight", "f0_predictor.condnet.0.weight", "f0_predictor.classifier.weight",
 * "m_source.l_linear.weight".  `data` is HOST fp32, PyTorch layout
 * (Conv1d [Cout,Cin,k]; ConvTranspose1d [Cin,Cout,k]; Linear [out,in]). */
typedef struct {
  const char*  name;
  const float* data;
  int32_t      ndim;
  int64_t      shape[4];
} GnvWeight;

/* replaces: HiFTGenerator.__init__ + load_state_dict of `mel2wav.*` (reached from
 * ChatterboxTTS.from_pretrained, synthesizer.py:185). */
int gnv_create(const GnvWeight* weights, int n_weights, int device, int dtype, unsigned flags,
               gnv_handle* out);
void gnv_destroy(gnv_handle h);
/* h may be NULL: returns the calling thread's last error (e.g. from a failed gnv_create). */
const char* gnv_last_error(gnv_handle h);
int gnv_abi_version(void);

/* bytes of scratch the decode calls need for a [B, 80, T] batch */
int gnv_workspace_bytes(gnv_handle h, int B, int T, size_t* out_bytes);

/* replaces: ConvRNNF0Predictor.forward (first statement of HiFTGenerator.inference). */
int gnv_f0(gnv_handle h, const float* mel, const int32_t* lengths, int B, int T, float* f0,
           void* workspace, size_t workspace_bytes, void* stream);

/* replaces: f0_upsamp + SourceModuleHnNSF/SineGen (second statement of HiFTGenerator.inference).
 * phase_vec [B,9] and noise [B,9,480*T] are optional (NULL -> counter-based Philox from `seed`);
 * they exist so a test can drive the kernel and the oracle with the same random numbers. */
int gnv_source(gnv_handle h, const float* f0, int B, int T, uint64_t seed,
               const float* phase_vec, const float* noise, float* s, void* stream);

/* No counterpart in the reference, which ships one chunk per sentence (services/tts/core/synthesizer.py:320-321,
 * :352-357); upstream's token streaming carries `cache_source` across chunks for the same purpose.
 * The same source for a piece of a longer utterance (intra-sentence streaming): frames
 * [t0, t0 + T) of an utterance whose earlier frames were generated by earlier calls.  sample0 = 480 * t0 keys the
 * noise by absolute sample index; f0_sum0 [B] (fp64, device, or NULL = 0) is the sum of f0 over frames [0, t0)
 * — the running phase — and f0_sum_out [B] (may alias f0_sum0, may be NULL) receives it for the next call.  A
 * stream cut into pieces of any size gives the samples gnv_source gives for the whole utterance. */
int gnv_source_stream(gnv_handle h, const float* f0, int B, int T, uint64_t seed, int64_t sample0,
                      const double* f0_sum0, float* s, double* f0_sum_out, void* stream);

/* replaces: HiFTGenerator.decode(x=mel, s=s).  lengths [B] (int32, mel frames, device) or NULL
 * for a rectangular batch; rows past an utterance's length decode as if the utterance ended
 * there and come back as zeros: every row of a ragged batch is bit-identical to that utterance
 * decoded alone, and the work past its end is skipped, not computed and discarded.  Entries are
 * read as clamp(lengths[b], 0, T) (0 = a silent filler row).  The workspace is pure scratch: its
 * contents before the call do not matter and nothing in it survives the call. */
int gnv_decode(gnv_handle h, const float* mel, const float* s, const int32_t* lengths,
               int B, int T, float* wav, void* workspace, size_t workspace_bytes, void* stream);

/* replaces: HiFTGenerator.inference(speech_feat, cache_source): f0 -> source -> overwrite the
 * first cache_len samples of s with cache_source [B,1,cache_len] -> decode.  Writes s_out. */
int gnv_inference(gnv_handle h, const float* mel, const float* cache_source, int cache_len,
                  const int32_t* lengths, int B, int T, uint64_t seed,
                  float* wav, float* s_out, void* workspace, size_t workspace_bytes, void* stream);

/* gnv_inference with the NSF seed(s) in DEVICE memory.  Same replaced call as gnv_inference.
 *   per_row == 0: one word; the source kernel reads *seed_dev when it RUNS and a one-thread kernel adds 1 to it behind
 *     it.  For CUDA-graph capture of a fixed chunk shape (first-chunk latency, BASELINE configs[1]): a by-value seed
 *     would be baked into the graph and every replay would repeat the same noise, which upstream's
 *     `torch.randn_like` (SineGen) never does.  Replay i after *seed_dev = s equals gnv_inference(seed = s + i).
 *   per_row != 0: seed_dev[B], one seed per utterance of the batch (not modified).  Row b gets exactly the source an
 *     utterance decoded alone — gnv_inference(B = 1, seed = seed_dev[b]) — gets: a request's audio does not depend on
 *     which micro-batch (services/tts/server.py:110-186) it happened to share. */
int gnv_inference_dseed(gnv_handle h, const float* mel, const float* cache_source, int cache_len,
                        const int32_t* lengths, int B, int T, uint64_t* seed_dev, int per_row,
                        float* wav, float* s_out, void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hook (bench.py's roofline): one gnv_inference with a CUDA event recorded after every
 * launch on `stream`; synchronises the stream before returning.  For launch i < *n_out:
 * ms_out[i] device time, kind_out[i] one of GNV_LAUNCH_*, flops_out[i] the layer's algorithmic flops
 * (2 * B * L_out * C_out * C_in * k; 0 for the byte-moving kernels), names_out + i*GNV_LAUNCH_NAME_LEN
 * the upstream module path of the layer.  All four arrays are HOST memory with `capacity` entries. */
#define GNV_LAUNCH_AUX       0   /* bandwidth-bound kernels: packers, STFT, source, f0 head, iSTFT head */
#define GNV_LAUNCH_CONV_TC   1   /* tcgen05 implicit-GEMM conv kernel */
#define GNV_LAUNCH_CONV_SIMT 2   /* CUDA-core conv kernel (source_downs; validation path) */
#define GNV_LAUNCH_NAME_LEN  48
int gnv_inference_profile(gnv_handle h, const float* mel, const int32_t* lengths, int B, int T, uint64_t seed,
                          float* wav, float* s_out, void* workspace, size_t workspace_bytes, void* stream,
                          int capacity, float* ms_out, int32_t* kind_out, double* flops_out, char* names_out,
                          int* n_out);

/* The streaming tail: optional fade/crossfade of the head, clamp(+-limit), optional int16 pack.
 *   head i < fade:  v = prev_tail ? prev_tail[i]*(1-fade_w[i]) + cur[i]*fade_w[i] : cur[i]*fade_w[i]
 *   (fade_w NULL or fade 0: no fade);  v = clamp(v, -limit, limit);
 *   out_f32[i] = v;  out_i16[i] = sat_i16(rint(v * 32767.f))   (either output may be NULL).
 * cur/out rows are `n` samples, row strides in elements.  prev_tail rows are `fade` samples.
 * replaces: S3Token2Wav `wav[:, :960] *= trim_fade`, and the service's
 * `.cpu().numpy().astype(float32)` / `.tobytes()` tail (synthesizer.py:352-357, server.py:150-155). */
int gnv_pcm_tail(const float* cur, int64_t cur_stride, const float* prev_tail, const float* fade_w,
                 int rows, int n, int fade, float limit,
                 int16_t* out_i16, float* out_f32, int64_t out_stride, void* stream);

/* G.711 mu-law companding of int16 PCM (telephony wire format; the reference's `phone` extra, pyproject.toml:55-57):
 * out[i] = audioop.lin2ulaw(pcm[i]) bit for bit.  pcm 16-byte aligned, out 8-byte aligned, device pointers.
 * replaces nothing in the reference (it ships float32 only, server.py:152); SURVEY 8f-3. */
int gnv_pcm_mulaw(const int16_t* pcm, int64_t n, uint8_t* out, void* stream);

/* ---- unit-test hooks (one kernel each; used by tests/, not by the service) --------------------
 * TEST ONLY: unlike every entry point above, gnv_stft / gnv_istft / gnv_conv1d allocate scratch with cudaMalloc and
 * synchronise the stream before returning.  Never call them from a serving path or inside a stream capture. */

/* HiFTGenerator._stft + cat(real, imag): s [B, L] -> spec [B, 18, L/4+1] (fp32, NCT). */
int gnv_stft(const float* s, int B, int L, float* spec_nct, void* stream);
/* exp / min(.,100) / sin / HiFTGenerator._istft / clamp: x [B, 18, F] (NCT) -> wav [B, 4*(F-1)]. */
int gnv_istft(const float* x_nct, int B, int F, float limit, float* wav, void* stream);

/* One Conv1d / ConvTranspose1d layer through the same kernels gnv_decode uses.
 *   x [B,Cin,Lin] fp32 NCT, w PyTorch layout (HOST), bias (HOST or NULL), alpha (HOST, Snake) ->
 *   y [B,Cout,Lout] fp32 NCT = act( conv(x) + bias + (res ? res : 0) ).  dtype/flags as gnv_create. */
int gnv_conv1d(int device, int dtype, unsigned flags, int transposed,
               const float* x, int B, int Cin, int Lin,
               const float* w_host, const float* bias_host, int Cout, int k, int stride, int pad, int dil,
               int act, const float* alpha_host, float slope, const float* res_nct,
               float* y_nct, int Lout, void* stream);

/* Copy a named intermediate of the LAST gnv_decode on (B, T, workspace) into out as fp32 NCT.
 * names: "s_stft", "fuse{0,1,2}" (x + source branch, the input of stage i's ResBlocks),
 * "stage{0,1,2}" (mean of the three ResBlocks), "conv_post".  The other intermediates are
 * updated in place by later layers and no longer exist when the decode returns. */
int gnv_debug_tap(gnv_handle h, const char* name, int B, int T, void* workspace, float* out_nct,
                  size_t out_capacity_elems, int64_t* out_shape3, void* stream);

/* diagnostics (tuning only, env GONOVA_CHAIN_DBG=8): per-phase clock stamps CTA 0 of the whole-ResBlock kernel recorded since
 * the last call; synchronises the device.  Word = role<<56 | lane<<48 | phase<<40 | event<<32 | clock32. */
int gnv_debug_chain_trace(unsigned long long* out, int cap, int* n_out);

/* diagnostics: resident 2-CTA clusters of the persistent conv kernel for a dynamic shared-memory size */
int gnv_debug_cluster_probe(int smem_bytes, int grid, int* max_clusters);

/* number of kernel launches one gnv_decode / gnv_inference (B, T) issues (bench.py's gpu_launches) */
int gnv_decode_launches(gnv_handle h, int B, int T, int* out);
int gnv_inference_launches(gnv_handle h, int B, int T, int* out);

/* Diagnostics; no counterpart in the reference (its one-request-at-a-time worker, services/tts/server.py:118-182, is
 * why the cache exists).  Launch-plan cache of a handle.  A plan (tensor maps, tile lists) is built on the first call for a new
 * (B, T, workspace) and kept in an LRU of 128 (env GONOVA_MAX_PLANS); an evicted plan's device slot is reused by the
 * next plan, so a service that sees a new sentence length on every call (services/tts/server.py:118-182) neither
 * allocates nor grows.  A plan that was used inside a stream capture is pinned (the CUDA graph points at its slot);
 * the FIRST call for a shape must be made outside a capture.
 * out[0] = plans cached, out[1] = plans built so far, out[2] = device slots ever allocated, out[3] = pinned plans. */
int gnv_plan_stats(gnv_handle h, uint64_t out[4]);

/* ---- the step before the vocoder: the CFM flow decoder ---------------------------------------------
 * replaces: CausalConditionalCFM.forward -> solve_euler -> ConditionalDecoder.forward of the engine's S3Gen flow
 * (`flow.decoder(mu, mask, spks, cond, n_timesteps)` inside S3Gen.flow_inference), reached from
 * services/tts/core/synthesizer.py:344-350 (model.generate -> S3Gen.inference).  Weights: the estimator's state dict under
 * upstream's names ("time_mlp.linear_1.weight", "down_blocks.0.0.block1.block.0.weight", "mid_blocks.3.1.2.attn1.to_q.weight",
 * "up_blocks.0.2.weight", "final_proj.bias", ...), HOST fp32 in PyTorch layout.  dtype GNV_DTYPE_BF16 or GNV_DTYPE_TF32.
 * Same conventions as above: caller-owned device buffers, caller's stream, status + gnv_last_error(NULL). */
typedef struct gnv_flow* gnv_flow_handle;
int gnv_flow_create(const GnvWeight* weights, int n_weights, int device, int dtype, unsigned flags, gnv_flow_handle* out);
void gnv_flow_destroy(gnv_flow_handle f);
int gnv_flow_workspace_bytes(gnv_flow_handle f, int B, int T, size_t* out_bytes);
/* z (the initial noise), mu, cond [B,80,T], spks [B,80] fp32 device tensors; lengths [B] int32 device or NULL (frames past
 * an utterance's length are masked like upstream's `mask`); n_timesteps Euler steps on the cosine grid with classifier-free
 * guidance `cfg_rate` (upstream: 10 and 0.7) -> mel [B,80,T] fp32. */
int gnv_flow_decode(gnv_flow_handle f, const float* z, const float* mu, const float* spks, const float* cond,
                    const int32_t* lengths, int B, int T, int n_timesteps, float cfg_rate, float* mel,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Measurement hook (bench.py's flow roofline), like gnv_inference_profile: one gnv_flow_decode with a CUDA event after every
 * launch of the Euler loop; synchronises the stream before returning.  Arrays are HOST memory with `capacity` entries. */
int gnv_flow_profile(gnv_flow_handle f, const float* z, const float* mu, const float* spks, const float* cond,
                     const int32_t* lengths, int B, int T, int n_timesteps, float cfg_rate, float* mel, void* workspace,
                     size_t workspace_bytes, void* stream, int capacity, float* ms_out, int32_t* kind_out, double* flops_out,
                     char* names_out, int* n_out);
/* Tuning hook: CTA 0's event timeline of the last flow_blk_kernel launch traced (GONOVA_FB_DBG=8, GONOVA_FB_TRACE_MODE =
 * 0 feed-forward / 1 output projection / 2 q,k,v); words (role + 1) << 56 | a << 48 | b << 40 | event << 32 | clock32. */
int gnv_debug_flow_trace(unsigned long long* out, int cap, int* n_out);
/* kernel launches of one gnv_flow_decode with the plan last built (bench bookkeeping) */
int gnv_flow_launches(gnv_flow_handle f, int n_timesteps, int* out);

/* ---- the front of the flow step: speech tokens -> mu / spks ("tokens -> mel") -----------------------------
 * replaces: the part of CausalMaskedDiffWithXvec.inference in front of `self.decoder(...)` in the engine's S3Gen flow
 * (`flow.inference(token, token_len, prompt_token, ..., embedding, finalize)`): F.normalize + spk_embed_affine_layer,
 * input_embedding, UpsampleConformerEncoder (pre-lookahead convs, 6 + 4 relative-position Conformer layers around a x2
 * upsampling conv), encoder_proj; reached from services/tts/core/synthesizer.py:344-350 through S3Gen.inference.
 * Weights: that module's state dict under upstream's names ("input_embedding.weight", "spk_embed_affine_layer.weight",
 * "encoder.encoders.0.self_attn.linear_pos.weight", "encoder.up_layer.conv.weight", "encoder_proj.bias", ...), HOST fp32. */
typedef struct gnv_flow_enc* gnv_flow_enc_handle;
int gnv_flow_enc_create(const GnvWeight* weights, int n_weights, int device, int dtype, unsigned flags, gnv_flow_enc_handle* out);
void gnv_flow_enc_destroy(gnv_flow_enc_handle f);
int gnv_flow_enc_workspace_bytes(gnv_flow_enc_handle f, int B, int L, size_t* out_bytes);
/* tokens [B, L] int32 device (the prompt's tokens followed by the utterance's; negative ids count as 0 like upstream's
 * clamp), token_len [B] int32 device or NULL, embedding [B, 192] fp32 x-vectors (NULL with spks NULL: skip the speaker
 * projection) -> mu [B, 80, 2L] fp32 (gnv_flow_decode's `mu`; frames at or past 2 * token_len[b] are zero) and
 * spks [B, 80].  An utterance of a ragged batch is encoded exactly as if it were alone (upstream runs one at a time). */
int gnv_flow_encode(gnv_flow_enc_handle f, const int32_t* tokens, const int32_t* token_len, const float* embedding, int B, int L,
                    float* mu, float* spks, void* workspace, size_t workspace_bytes, void* stream);
/* Measurement hook like gnv_flow_profile: one gnv_flow_encode with a CUDA event after every launch. */
int gnv_flow_encode_profile(gnv_flow_enc_handle f, const int32_t* tokens, const int32_t* token_len, const float* embedding, int B,
                            int L, float* mu, float* spks, void* workspace, size_t workspace_bytes, void* stream, int capacity,
                            float* ms_out, int32_t* kind_out, double* flops_out, char* names_out, int* n_out);
/* kernel launches of one gnv_flow_encode with the plan last built (bench bookkeeping) */
int gnv_flow_enc_launches(gnv_flow_enc_handle f, int* out);

#ifdef __cplusplus
}
#endif
#endif /* GONOVA_HIFT_H_ */
