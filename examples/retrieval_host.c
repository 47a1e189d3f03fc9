/* C host of the score-tile entry points of libcrossclr_b200 (no Python, no torch): the MaxMargin_coot loss with its gradients
 * and the retrieval ranks of the same pairs, on device buffers the caller owns.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/retrieval_host.c \
 *       -L crossmodal_contrastive_learning_b200 -lcrossclr_b200 -L /usr/local/cuda/lib64 -lcudart -o retrieval_host
 *
 * Replaces, for such a host: trainer/loss.py:29-41 (MaxMargin_coot.forward over `cosine_sim`, :7-15) and its autograd
 * backward; the ranks are the hinge indicators of :34-35 at margin 0 (the reference only pictures retrieval).
 */
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "crossclr_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CC(x) do { int rc_ = (x); if (rc_ != CROSSCLR_OK) { fprintf(stderr, "%s: %s\n", #x, crossclr_last_error()); return 1; } } while (0)

int main(void) {
  const int B = 1024, D = 256;
  /* paired embeddings: s_i = 0.2 im_i + noise, rows of roughly unit length (fp32: staged as fp16 hi + lo pairs on the device) */
  float* h = (float*)malloc((size_t)2 * B * D * sizeof(float));
  unsigned s = 2024u;
  for (size_t i = 0; i < (size_t)B * D; ++i) {
    s = s * 1664525u + 1013904223u; h[i] = ((float)(s >> 8) / 8388608.0f - 1.0f) * 0.108f;
    s = s * 1664525u + 1013904223u; h[(size_t)B * D + i] = 0.2f * h[i] + ((float)(s >> 8) / 8388608.0f - 1.0f) * 0.108f;
  }
  float *im, *sd, *d_im, *d_s;
  void* ws;
  double* loss;
  int32_t *r_im2s, *r_s2im;
  const size_t ws_bytes = crossclr_maxmargin_workspace_bytes(B, D, CROSSCLR_F32);
  CK(cudaMalloc((void**)&im, (size_t)B * D * 4));   CK(cudaMalloc((void**)&sd, (size_t)B * D * 4));
  CK(cudaMalloc((void**)&d_im, (size_t)B * D * 4)); CK(cudaMalloc((void**)&d_s, (size_t)B * D * 4));
  CK(cudaMalloc(&ws, ws_bytes));                    CK(cudaMalloc((void**)&loss, 8));
  CK(cudaMalloc((void**)&r_im2s, (size_t)B * 4));   CK(cudaMalloc((void**)&r_s2im, (size_t)B * 4));
  CK(cudaMemcpy(im, h, (size_t)B * D * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(sd, h + (size_t)B * D, (size_t)B * D * 4, cudaMemcpyHostToDevice));
  printf("kernel: %s\n", crossclr_maxmargin_kernel_name(im, sd, CROSSCLR_F32, D, D, B, D));

  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  /* ranks first (they share the workspace with the loss; each forward rewrites it), then loss + gradients */
  CC(crossclr_retrieval_ranks(im, sd, CROSSCLR_F32, D, D, B, D, ws, ws_bytes, r_im2s, r_s2im, st));
  CC(crossclr_maxmargin_fwd(im, sd, CROSSCLR_F32, D, D, B, D, 0.1f, ws, ws_bytes, loss, st));
  CC(crossclr_maxmargin_bwd(im, sd, CROSSCLR_F32, D, D, B, D, 0.1f, ws, ws_bytes, NULL, d_im, D, d_s, D, CROSSCLR_F32, st));
  CK(cudaStreamSynchronize(st));

  double hl;
  float g0;
  int32_t* ranks = (int32_t*)malloc((size_t)2 * B * sizeof(int32_t));
  CK(cudaMemcpy(&hl, loss, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&g0, d_im, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ranks, r_im2s, (size_t)B * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ranks + B, r_s2im, (size_t)B * 4, cudaMemcpyDeviceToHost));
  int r1 = 0, r10 = 0;
  for (int i = 0; i < 2 * B; ++i) { r1 += ranks[i] < 1; r10 += ranks[i] < 10; }
  printf("loss %.9f  d_im[0][0] %.6e  R@1 %.3f  R@10 %.3f (both directions)\n", hl, (double)g0, r1 / (2.0 * B), r10 / (2.0 * B));
  return 0;
}
