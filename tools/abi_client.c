/* Plain-C client of libegn.so (include/egn.h): what a non-Python host of the reference path would do.
 * No torch anywhere - device memory comes from the CUDA runtime, weights from blob files written by
 * egn_b200.pack.pack_state_dict (the state_dicts of gen_00000016.pt['a'] and baseline_edge_16.pkl).
 *
 *   abi_client <bdcn.blob> <esf.blob> <frames.f32 [B,1,240,320]> <B> <out_prefix>
 *
 * Runs calc_edge -> DenseNet2D forward (baseline_edge.yaml) -> argmax / centres and writes
 * <out_prefix>.argmax.u8 [B,240,320], <out_prefix>.elpred.f32 [B,10], <out_prefix>.edge.f32 [B,240,320].
 * Built by __graft_entry__.build() with gcc; tests/test_gpu_parity.py compares its files bit for bit with
 * the Python mirrors' outputs. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "egn.h"

#define HW (240 * 320)

static void* read_file(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc((size_t)sz);
  if (fread(p, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "short read %s\n", path); exit(2); }
  fclose(f);
  *n = (size_t)sz;
  return p;
}

static void write_file(const char* prefix, const char* suffix, const void* p, size_t n) {
  char path[1024];
  snprintf(path, sizeof(path), "%s%s", prefix, suffix);
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(p, 1, n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
  fclose(f);
}

#define EGN(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, egn_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc != 6) { fprintf(stderr, "usage: %s bdcn.blob esf.blob frames.f32 B out_prefix\n", argv[0]); return 2; }
  const int B = atoi(argv[4]);
  size_t nb = 0, ne = 0, nx = 0;
  void* bdcn_blob = read_file(argv[1], &nb);
  void* esf_blob = read_file(argv[2], &ne);
  float* x_host = (float*)read_file(argv[3], &nx);
  if (nx != (size_t)B * HW * sizeof(float)) { fprintf(stderr, "frames file does not hold %d frames\n", B); return 2; }

  egn_config cfg_edge = {0, 0, 0, 0, 0, 8};
  egn_config cfg_esf = {1, 0, 0, 0, 0, 8};          /* configs/baseline_edge.yaml: add_edge = 1 */
  egn_ctx *edge = NULL, *esf = NULL;
  EGN(egn_create(0, &cfg_edge, &edge));
  EGN(egn_create(0, &cfg_esf, &esf));
  EGN(egn_plan(edge, B));
  EGN(egn_plan(esf, B));
  EGN(egn_set_weights(edge, EGN_NET_BDCN, bdcn_blob, nb));
  EGN(egn_set_weights(esf, EGN_NET_ESF, esf_blob, ne));

  float *x, *e, *logits, *el_out, *latent, *el_pred;
  uint8_t* argmax;
  CU(cudaMalloc((void**)&x, (size_t)B * HW * 4));
  CU(cudaMalloc((void**)&e, (size_t)B * HW * 4));
  CU(cudaMalloc((void**)&logits, (size_t)B * 3 * HW * 4));
  CU(cudaMalloc((void**)&el_out, (size_t)B * 10 * 4));
  CU(cudaMalloc((void**)&latent, (size_t)B * 153 * 4));
  CU(cudaMalloc((void**)&el_pred, (size_t)B * 10 * 4));
  CU(cudaMalloc((void**)&argmax, (size_t)B * HW));
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  CU(cudaMemcpyAsync(x, x_host, (size_t)B * HW * 4, cudaMemcpyHostToDevice, st));
  EGN(egn_bdcn_forward(edge, x, 1, e, B, st));                       /* utils.calc_edge */
  EGN(egn_esf_forward(esf, x, e, logits, el_out, latent, B, st));    /* DenseNet2D.forward */
  EGN(egn_seg_post(esf, logits, el_out, NULL, argmax, el_pred, B, st)); /* get_predictions + seg centres */
  CU(cudaStreamSynchronize(st));

  uint8_t* am_h = (uint8_t*)malloc((size_t)B * HW);
  float* ep_h = (float*)malloc((size_t)B * 10 * 4);
  float* e_h = (float*)malloc((size_t)B * HW * 4);
  CU(cudaMemcpy(am_h, argmax, (size_t)B * HW, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ep_h, el_pred, (size_t)B * 10 * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(e_h, e, (size_t)B * HW * 4, cudaMemcpyDeviceToHost));
  write_file(argv[5], ".argmax.u8", am_h, (size_t)B * HW);
  write_file(argv[5], ".elpred.f32", ep_h, (size_t)B * 10 * 4);
  write_file(argv[5], ".edge.f32", e_h, (size_t)B * HW * 4);
  printf("abi_client: %d frames, %lld kernel launches\n", B, egn_launch_count(edge) + egn_launch_count(esf));
  EGN(egn_destroy(edge));
  EGN(egn_destroy(esf));
  return 0;
}
