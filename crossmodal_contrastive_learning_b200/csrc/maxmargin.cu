// MaxMargin_coot (trainer/loss.py:17-41 of the reference): the exact fp32 CUDA-core kernels for any B, D (fp32 inputs, small
// or unaligned problems) and the dispatch to the tensor-core kernels of maxmargin_tc.cu (16-bit inputs).  SURVEY.md section 8
// row f1 -- next to the CrossCLR hot path, not part of it.  The same forward with margin 0 yields the retrieval ranks (f4).
//
//   scores = im s^T (plain dot products: `cosine_sim`, :7-15, does not normalise);  d_i = scores_ii
//   loss   = (1/B^2) sum_{i != j} [ max(0, m + scores_ij - d_i) + max(0, m + scores_ij - d_j) ]          (:34-41)
//   dL/dscores_ij = (1[m + s_ij - d_i > 0] + 1[m + s_ij - d_j > 0]) / B^2   (i != j)
//   dL/dscores_ii = -(#active row hinges of i + #active column hinges of i) / B^2
//   dL/dim = G s,  dL/ds = G^T im;  G is symmetric under swapping the roles of im and s, so ONE kernel serves both.
//
// No B x B array is stored: the forward keeps d[B] and the hinge counts cnt[B]; the backward recomputes score tiles.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace crossclr {

namespace {

constexpr int MT = 32;           // score tile edge; 256 threads, each 4 scores (rows ty + 8k, column tx)

template <typename T>
__global__ void __launch_bounds__(256) mm_diag_kernel(const T* __restrict__ im, int64_t im_stride, const T* __restrict__ s,
                                                     int64_t s_stride, int B, int D, float* __restrict__ diag,
                                                     float* __restrict__ cnt, double* __restrict__ acc,
                                                     int* __restrict__ rank_row, int* __restrict__ rank_col) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) acc[0] = 0.0;
  if (row >= B) return;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32)
    dot = fmaf(to_float<T>(im[(int64_t)row * im_stride + d]), to_float<T>(s[(int64_t)row * s_stride + d]), dot);
  dot = warp_sum(dot);
  if (lane == 0) {
    diag[row] = dot; cnt[row] = 0.f;
    if (rank_row != nullptr) rank_row[row] = 0;
    if (rank_col != nullptr) rank_col[row] = 0;
  }
}

// 32 x 32 tile of A B^T (rows a0.., b0..) into sc[4] of each thread; the same summation order (d ascending, one fma per
// element) in the forward and in the backward, so the hinge indicators of the two passes agree exactly.  The diagonal
// kernel above uses a different order; only differences of scores against d enter the hinges, and an element within
// rounding of the hinge contributes ~0 to the loss and 1/B^2 to one gradient entry either way.
template <typename T>
__device__ __forceinline__ void score_tile(const T* __restrict__ A, int64_t a_stride, const T* __restrict__ Bm,
                                           int64_t b_stride, int a0, int b0, int n, int D, float (*As)[MT + 1],
                                           float (*Bs)[MT + 1], float (&sc)[4]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  sc[0] = sc[1] = sc[2] = sc[3] = 0.f;
  for (int d0 = 0; d0 < D; d0 += MT) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 8 * k, d = d0 + tx;
      As[r][tx] = (a0 + r < n && d < D) ? to_float<T>(A[(int64_t)(a0 + r) * a_stride + d]) : 0.f;
      Bs[r][tx] = (b0 + r < n && d < D) ? to_float<T>(Bm[(int64_t)(b0 + r) * b_stride + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < MT; ++d) {
      const float bv = Bs[tx][d];
#pragma unroll
      for (int k = 0; k < 4; ++k) sc[k] = fmaf(As[ty + 8 * k][d], bv, sc[k]);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mm_fwd_kernel(const T* __restrict__ im, int64_t im_stride, const T* __restrict__ s,
                                                    int64_t s_stride, int B, int D, float margin,
                                                    const float* __restrict__ diag, float* __restrict__ cnt,
                                                    double* __restrict__ acc, int* __restrict__ rank_row,
                                                    int* __restrict__ rank_col) {
  __shared__ float As[MT][MT + 1], Bs[MT][MT + 1];
  __shared__ float colcnt[MT];
  __shared__ float wsum[8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
  if (threadIdx.x < MT) colcnt[threadIdx.x] = 0.f;
  float sc[4];
  score_tile<T>(im, im_stride, s, s_stride, i0, j0, B, D, As, Bs, sc);     // ends with __syncthreads()
  const int j = j0 + tx;
  const float dj = j < B ? diag[j] : 0.f;
  float local = 0.f, ccol = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k;
    const bool valid = (i < B) && (j < B) && (i != j);
    const float di = i < B ? diag[i] : 0.f;                                // warp-uniform row
    const float cs = valid ? fmaxf(0.f, margin + sc[k] - di) : 0.f;        // trainer/loss.py:34
    const float ci = valid ? fmaxf(0.f, margin + sc[k] - dj) : 0.f;        // :35
    local += cs + ci;
    const unsigned rowmask = __ballot_sync(0xffffffffu, cs > 0.f);
    if (tx == 0 && rowmask) {                                              // active row hinges of i in this tile
      atomicAdd(&cnt[i], (float)__popc(rowmask));
      if (rank_row != nullptr) atomicAdd(&rank_row[i], __popc(rowmask));
    }
    if (ci > 0.f) ccol += 1.f;
  }
  if (ccol > 0.f) atomicAdd(&colcnt[tx], ccol);
  local = warp_sum(local);
  if (tx == 0) wsum[ty] = local;
  __syncthreads();
  if (threadIdx.x < MT && colcnt[threadIdx.x] > 0.f && j0 + (int)threadIdx.x < B) {
    atomicAdd(&cnt[j0 + threadIdx.x], colcnt[threadIdx.x]);                // active column hinges of j in this tile
    if (rank_col != nullptr) atomicAdd(&rank_col[j0 + threadIdx.x], (int)colcnt[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    if (t != 0.f) atomicAdd(acc, (double)t);
  }
}

__global__ void mm_finish_kernel(const double* __restrict__ acc, int B, double* __restrict__ loss) {
  loss[0] = acc[0] / ((double)B * (double)B);                              // trainer/loss.py:41
}

// out[a, :] = c * sum_b G[a][b] Bm[b, :] for the 32 rows a of this block; G recomputed tile by tile.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) mm_grad_kernel(const T* __restrict__ A, int64_t a_stride, const T* __restrict__ Bm,
                                                     int64_t b_stride, int B, int D, float margin,
                                                     const float* __restrict__ diag, const float* __restrict__ cnt,
                                                     const double* __restrict__ grad_out, TO* __restrict__ out,
                                                     int64_t out_stride) {
  extern __shared__ float out_s[];                                         // [MT][D] fp32 accumulators
  __shared__ float As[MT][MT + 1], Bs[MT][MT + 1];
  __shared__ float Gs[MT][MT + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int a0 = blockIdx.x * MT;
  for (int e = threadIdx.x; e < MT * D; e += blockDim.x) out_s[e] = 0.f;
  for (int b0 = 0; b0 < B; b0 += MT) {
    float sc[4];
    score_tile<T>(A, a_stride, Bm, b_stride, a0, b0, B, D, As, Bs, sc);
    const int b = b0 + tx;
    const float db = b < B ? diag[b] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int a = a0 + ty + 8 * k;
      float gval = 0.f;
      if (a < B && b < B) {
        if (a != b) gval = ((margin + sc[k] - diag[a] > 0.f) ? 1.f : 0.f) + ((margin + sc[k] - db > 0.f) ? 1.f : 0.f);
        else gval = -cnt[a];
      }
      Gs[ty + 8 * k][tx] = gval;
    }
    __syncthreads();
    // out_s[r][d] += sum_j Gs[r][j] * Bm[b0 + j][d]; thread (tx, ty) owns rows ty + 8k and columns d = tx (mod 32)
    for (int d = tx; d < D; d += 32) {
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int jj = 0; jj < MT && b0 + jj < B; ++jj) {
        const float bv = to_float<T>(Bm[(int64_t)(b0 + jj) * b_stride + d]);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc4[k] = fmaf(Gs[ty + 8 * k][jj], bv, acc4[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) out_s[(ty + 8 * k) * D + d] += acc4[k];
    }
    __syncthreads();
  }
  double c = 1.0 / ((double)B * (double)B);
  if (grad_out != nullptr) c *= grad_out[0];
  const float cf = (float)c;
  for (int e = threadIdx.x; e < MT * D; e += blockDim.x) {
    const int r = e / D, d = e - r * D;
    if (a0 + r < B) out[(int64_t)(a0 + r) * out_stride + d] = from_float<TO>(cf * out_s[e]);
  }
}

template <typename T>
int fwd_t(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int B, int D, float margin, float* diag,
          float* cnt, double* acc, int* rank_row, int* rank_col, double* loss, cudaStream_t st) {
  mm_diag_kernel<T><<<(B + 7) / 8, 256, 0, st>>>((const T*)im, im_stride, (const T*)s, s_stride, B, D, diag, cnt, acc,
                                                 rank_row, rank_col);
  int rc = check_launch("mm_diag_kernel");
  if (rc) return rc;
  dim3 grid((B + MT - 1) / MT, (B + MT - 1) / MT);
  mm_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)im, im_stride, (const T*)s, s_stride, B, D, margin, diag, cnt, acc,
                                         rank_row, rank_col);
  rc = check_launch("mm_fwd_kernel");
  if (rc || loss == nullptr) return rc;
  mm_finish_kernel<<<1, 1, 0, st>>>(acc, B, loss);
  return check_launch("mm_finish_kernel");
}

template <typename T, typename TO>
int grad_t(const void* A, int64_t a_stride, const void* Bm, int64_t b_stride, int B, int D, float margin,
           const float* diag, const float* cnt, const double* grad_out, void* out, int64_t out_stride, cudaStream_t st) {
  const size_t smem = (size_t)MT * D * sizeof(float);
  CC_CHECK_CUDA(cudaFuncSetAttribute(mm_grad_kernel<T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mm_grad_kernel<T, TO><<<(B + MT - 1) / MT, 256, smem, st>>>((const T*)A, a_stride, (const T*)Bm, b_stride, B, D, margin,
                                                             diag, cnt, grad_out, (TO*)out, out_stride);
  return check_launch("mm_grad_kernel");
}

template <typename T>
int grad_out_t(const void* A, int64_t as, const void* Bm, int64_t bs, int B, int D, float margin, const float* diag,
               const float* cnt, const double* go, void* out, int64_t os, int out_dtype, cudaStream_t st) {
  switch (out_dtype) {
    case CROSSCLR_F32: return grad_t<T, float>(A, as, Bm, bs, B, D, margin, diag, cnt, go, out, os, st);
    case CROSSCLR_F16: return grad_t<T, __half>(A, as, Bm, bs, B, D, margin, diag, cnt, go, out, os, st);
    case CROSSCLR_BF16: return grad_t<T, __nv_bfloat16>(A, as, Bm, bs, B, D, margin, diag, cnt, go, out, os, st);
    default: set_error("crossclr_maxmargin_bwd: unsupported output dtype %d", out_dtype); return CROSSCLR_EINVAL;
  }
}

// 0 = by the rule of maxmargin_tc_applies, 1 = CUDA cores, 2 = tensor cores (an error where they do not apply)
int maxmargin_path_override() {
  const char* e = getenv("CROSSCLR_MAXMARGIN_PATH");
  if (e == nullptr || !*e) return 0;
  if (!strcmp(e, "simt")) return 1;
  if (!strcmp(e, "tc")) return 2;
  return 0;
}

// workspace layout: double acc | pad | float diag[B] | float cnt[B] | (256-byte aligned) float dacc[ceil128(B)][ceil64(D)] |
// (fp32 inputs) the staged fp16 [hi | lo] rows of both tensors
size_t mm_dacc_offset(int B) { return (16 + 2 * (size_t)B * sizeof(float) + 255) / 256 * 256; }
size_t mm_stage_offset(int B, int D) { return mm_dacc_offset(B) + (maxmargin_tc_dacc_bytes(B, D) + 255) / 256 * 256; }

// *use_tc: which kernels serve this problem; an explicit "tc" request for a problem they do not take is an error
int mm_choose(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D, bool* use_tc) {
  const int ov = maxmargin_path_override();
  const bool can = maxmargin_tc_applies(im, im_stride, s, s_stride, dtype, B, D);
  if (ov == 2 && !can) {
    set_error("CROSSCLR_MAXMARGIN_PATH=tc: the tensor-core kernels need batch >= 256, dim >= 64 and, for fp16 / bf16 inputs, "
              "16-byte aligned rows (got dtype %d, batch %d, dim %d)", dtype, B, D);
    return CROSSCLR_EINVAL;
  }
  *use_tc = can && ov != 1;
  return CROSSCLR_OK;
}

}  // namespace

size_t maxmargin_workspace_bytes(int B, int D, int dtype) {
  return mm_stage_offset(B, D) + maxmargin_tc_stage_bytes(B, D, dtype);
}

const char* maxmargin_kernel_name(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D) {
  bool tc = false;
  if (mm_choose(im, im_stride, s, s_stride, dtype, B, D, &tc)) return "invalid";
  return tc ? "mm_tc_kernel" : "mm_fwd_kernel";
}

int launch_maxmargin_fwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                         float margin, void* workspace, int* rank_row, int* rank_col, double* loss, cudaStream_t st) {
  double* acc = (double*)workspace;
  float* diag = (float*)((char*)workspace + 16);
  float* cnt = diag + B;
  bool tc = false;
  int rc = mm_choose(im, im_stride, s, s_stride, dtype, B, D, &tc);
  if (rc) return rc;
  if (tc)
    return launch_maxmargin_tc_fwd(im, im_stride, s, s_stride, dtype, B, D, margin, diag, cnt, acc,
                                   (char*)workspace + mm_stage_offset(B, D), rank_row, rank_col, loss, st);
  switch (dtype) {
    case CROSSCLR_F32: return fwd_t<float>(im, im_stride, s, s_stride, B, D, margin, diag, cnt, acc, rank_row, rank_col, loss, st);
    case CROSSCLR_F16: return fwd_t<__half>(im, im_stride, s, s_stride, B, D, margin, diag, cnt, acc, rank_row, rank_col, loss, st);
    case CROSSCLR_BF16:
      return fwd_t<__nv_bfloat16>(im, im_stride, s, s_stride, B, D, margin, diag, cnt, acc, rank_row, rank_col, loss, st);
    default: set_error("crossclr_maxmargin_fwd: unsupported dtype %d", dtype); return CROSSCLR_EINVAL;
  }
}

int launch_maxmargin_bwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                         float margin, void* workspace, const double* grad_out, void* d_im, int64_t d_im_stride,
                         void* d_s, int64_t d_s_stride, int out_dtype, cudaStream_t st) {
  const float* diag = (const float*)((const char*)workspace + 16);
  const float* cnt = diag + B;
  bool tc = false;
  int rc = mm_choose(im, im_stride, s, s_stride, dtype, B, D, &tc);
  if (rc) return rc;
  if (tc)
    return launch_maxmargin_tc_bwd(im, im_stride, s, s_stride, dtype, B, D, margin, diag, cnt,
                                   (float*)((char*)workspace + mm_dacc_offset(B)), (char*)workspace + mm_stage_offset(B, D),
                                   grad_out, d_im, d_im_stride, d_s, d_s_stride, out_dtype, st);
  if ((size_t)D * MT * sizeof(float) > 200 * 1024) {
    set_error("crossclr_maxmargin_bwd: dim %d too large for the CUDA-core path", D);
    return CROSSCLR_EINVAL;
  }
#define CC_MM(T)                                                                                                           \
  rc = grad_out_t<T>(im, im_stride, s, s_stride, B, D, margin, diag, cnt, grad_out, d_im, d_im_stride, out_dtype, st);     \
  if (!rc) rc = grad_out_t<T>(s, s_stride, im, im_stride, B, D, margin, diag, cnt, grad_out, d_s, d_s_stride, out_dtype, st)
  switch (dtype) {
    case CROSSCLR_F32: CC_MM(float); break;
    case CROSSCLR_F16: CC_MM(__half); break;
    case CROSSCLR_BF16: CC_MM(__nv_bfloat16); break;
    default: set_error("crossclr_maxmargin_bwd: unsupported dtype %d", dtype); return CROSSCLR_EINVAL;
  }
#undef CC_MM
  return rc;
}

}  // namespace crossclr
