/* A host in plain C: the decoder behind the C ABI with no Python and no torch anywhere.
 *
 *   c_host <weights.bin> <mel.bin> <wav_out.bin> [dtype: 0 = tf32, 1 = bf16] [seed]
 *
 * weights.bin (written by gonova_tts_b200.weights.write_flat): int32 n; per tensor: int32 name_len, name bytes,
 *   int32 ndim, int64 shape[ndim], float32 data — the folded (weight-norm applied) HiFT parameters.
 * mel.bin: int32 B, int32 T, float32 mel[B][80][T].   wav_out.bin: float32 wav[B][480*T].
 *
 * This is what the reference's service would bind if it were not Python: create once, one
 * workspace from the caller's allocator, gnv_inference on the caller's stream, caller-owned buffers throughout. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gonova_hift.h"

#define CK(expr)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(e_));                       \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

static int rd(void* dst, size_t n, FILE* f) { return fread(dst, 1, n, f) == n ? 0 : -1; }

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s weights.bin mel.bin wav_out.bin [dtype] [seed]\n", argv[0]);
    return 1;
  }
  const int dtype = argc > 4 ? atoi(argv[4]) : GNV_DTYPE_BF16;
  const uint64_t seed = argc > 5 ? strtoull(argv[5], NULL, 10) : 1;

  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int32_t n = 0;
  if (rd(&n, 4, f) || n <= 0 || n > 4096) { fprintf(stderr, "bad weight file\n"); return 1; }
  GnvWeight* w = (GnvWeight*)calloc((size_t)n, sizeof(GnvWeight));
  for (int i = 0; i < n; ++i) {
    int32_t len = 0, nd = 0;
    if (rd(&len, 4, f) || len <= 0 || len > 255) { fprintf(stderr, "bad weight name\n"); return 1; }
    char* name = (char*)calloc((size_t)len + 1, 1);
    if (rd(name, (size_t)len, f) || rd(&nd, 4, f) || nd < 1 || nd > 4) { fprintf(stderr, "bad weight entry\n"); return 1; }
    size_t count = 1;
    for (int d = 0; d < nd; ++d) {
      if (rd(&w[i].shape[d], 8, f)) { fprintf(stderr, "bad weight shape\n"); return 1; }
      count *= (size_t)w[i].shape[d];
    }
    float* data = (float*)malloc(count * sizeof(float));
    if (!data || rd(data, count * sizeof(float), f)) { fprintf(stderr, "bad weight data\n"); return 1; }
    w[i].name = name; w[i].data = data; w[i].ndim = nd;
  }
  fclose(f);

  f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 1; }
  int32_t B = 0, T = 0;
  if (rd(&B, 4, f) || rd(&T, 4, f) || B <= 0 || T <= 0) { fprintf(stderr, "bad mel file\n"); return 1; }
  const size_t n_mel = (size_t)B * 80 * (size_t)T, n_wav = (size_t)B * 480 * (size_t)T;
  float* mel_h = (float*)malloc(n_mel * sizeof(float));
  if (!mel_h || rd(mel_h, n_mel * sizeof(float), f)) { fprintf(stderr, "bad mel data\n"); return 1; }
  fclose(f);

  gnv_handle h = NULL;
  if (gnv_create(w, n, 0, dtype, 0, &h)) { fprintf(stderr, "gnv_create: %s\n", gnv_last_error(NULL)); return 3; }
  size_t ws_bytes = 0;
  if (gnv_workspace_bytes(h, B, T, &ws_bytes)) { fprintf(stderr, "%s\n", gnv_last_error(h)); return 3; }

  cudaStream_t st;
  float *mel_d = NULL, *wav_d = NULL, *src_d = NULL;
  void* ws = NULL;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaMalloc((void**)&mel_d, n_mel * sizeof(float)));
  CK(cudaMalloc((void**)&wav_d, n_wav * sizeof(float)));
  CK(cudaMalloc((void**)&src_d, n_wav * sizeof(float)));
  CK(cudaMalloc(&ws, ws_bytes));                       /* cudaMalloc returns 256-byte aligned, 1024 in practice */
  if (((uintptr_t)ws & 1023) != 0) { fprintf(stderr, "workspace is not 1024-byte aligned\n"); return 2; }
  CK(cudaMemcpyAsync(mel_d, mel_h, n_mel * sizeof(float), cudaMemcpyHostToDevice, st));
  for (int rep = 0; rep < 2; ++rep) {                  /* the second call reuses the launch plan */
    if (gnv_inference(h, mel_d, NULL, 0, NULL, B, T, seed, wav_d, src_d, ws, ws_bytes, (void*)st)) {
      fprintf(stderr, "gnv_inference: %s\n", gnv_last_error(h));
      return 3;
    }
  }
  float* wav_h = (float*)malloc(n_wav * sizeof(float));
  CK(cudaMemcpyAsync(wav_h, wav_d, n_wav * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  uint64_t stats[4] = {0, 0, 0, 0};
  gnv_plan_stats(h, stats);

  f = fopen(argv[3], "wb");
  if (!f || fwrite(wav_h, sizeof(float), n_wav, f) != n_wav) { perror(argv[3]); return 1; }
  fclose(f);
  printf("decoded %d x %d frames (%.2f s of audio each), ABI %d, plans built %llu\n", B, T, T / 50.0, gnv_abi_version(),
         (unsigned long long)stats[1]);
  gnv_destroy(h);
  cudaFree(mel_d); cudaFree(wav_d); cudaFree(src_d); cudaFree(ws);
  cudaStreamDestroy(st);
  return 0;
}
