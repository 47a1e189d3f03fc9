/* Minimal C host of libcrossclr_b200 (no Python, no torch): one forward + backward of the CrossCLR criterion on
 * device buffers the caller owns.  What a non-Python user of the reference's loss would link against.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_host.c \
 *       -L crossmodal_contrastive_learning_b200 -lcrossclr_b200 -L /usr/local/cuda/lib64 -lcudart -o c_host
 *
 * Replaces, for such a host: trainer/loss.py:76-114 (forward) and its autograd backward.
 */
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "crossclr_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CC(x) do { int rc_ = (x); if (rc_ != CROSSCLR_OK) { fprintf(stderr, "%s: %s\n", #x, crossclr_last_error()); return 1; } } while (0)

int main(void) {
  const int B = 1024, D = 512;
  crossclr_problem_t p = {2, B, D, 0, 2 * B, 0.03f, 0.8f};          /* one rank: [video; text] segments, all rows owned */
  const int path = crossclr_choose_path(&p, CROSSCLR_F32, 0);
  if (path < 0) { fprintf(stderr, "%s\n", crossclr_last_error()); return 1; }
  const size_t fsz = crossclr_feature_dtype(path) == CROSSCLR_F32 ? 4 : 2;
  const size_t pitch = (size_t)crossclr_feature_pitch(path, D);      /* stacked rows carry a tail on the tensor-core paths */
  const size_t S = (size_t)crossclr_segment_rows(path, B);           /* and segments are padded to a multiple of 128 rows */

  float* h = (float*)malloc((size_t)2 * B * D * sizeof(float));
  unsigned s = 12345u;
  for (size_t i = 0; i < (size_t)2 * B * D; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }

  float *video, *text, *rnorm, *stats, *coef, *scal, *dv, *dt;
  void *feat, *ws;
  double* loss;
  const size_t ws_bytes = crossclr_workspace_bytes(&p, path);
  CK(cudaMalloc((void**)&video, (size_t)B * D * 4)); CK(cudaMalloc((void**)&text, (size_t)B * D * 4));
  CK(cudaMalloc(&feat, 2 * S * pitch * fsz));   CK(cudaMalloc((void**)&rnorm, (size_t)2 * B * 4));
  CK(cudaMalloc((void**)&stats, 2 * S * 8));  CK(cudaMalloc((void**)&coef, 2 * S * 8));
  CK(cudaMalloc((void**)&scal, 16));                  CK(cudaMalloc((void**)&loss, 8));
  CK(cudaMalloc((void**)&dv, (size_t)B * D * 4));     CK(cudaMalloc((void**)&dt, (size_t)B * D * 4));
  CK(cudaMalloc(&ws, ws_bytes));
  CK(cudaMemcpy(video, h, (size_t)B * D * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(text, h + (size_t)B * D, (size_t)B * D * 4, cudaMemcpyHostToDevice));

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CC(crossclr_forward(&p, path, video, text, CROSSCLR_F32, D, D, feat, rnorm, stats, coef, scal, loss, st));
  CC(crossclr_bwd(&p, path, feat, rnorm, coef, scal, NULL, 1.0f, dv, D, dt, D, CROSSCLR_F32, ws, ws_bytes, st));
  CK(cudaStreamSynchronize(st));

  double hl;
  float g0;
  CK(cudaMemcpy(&hl, loss, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&g0, dv, 4, cudaMemcpyDeviceToHost));
  printf("loss %.9f  dv[0][0] %.6e  (%lld kernel launches)\n", hl, (double)g0, (long long)crossclr_launch_count());
  return 0;
}
