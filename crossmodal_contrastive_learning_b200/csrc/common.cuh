// Shared helpers for libcrossclr_b200 (host + device).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "crossclr_b200.h"

namespace crossclr {

constexpr float kEps = 1e-12f;              // F.normalize default eps (trainer/loss.py:79-80)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kShiftHeadroom = 96.0f;     // largest shifted log2-logit we allow: 2^96 * 2^19 rows < 2^127
// The constant shift keeps every term of a row representable only while the largest possible log2-logit, log2e max(1,|w|)/tau,
// stays below this (tau >= ~0.0073 at |w| <= 1): beyond it a weakly aligned row flushes to zero as a whole.  Such problems run
// on the exact path with per-row online maxima in the log2 domain, see Geometry::row_shift.
constexpr float kConstShiftMaxLogit = 200.0f;

// ---- error reporting -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define CC_CHECK_CUDA(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      (void)cudaGetLastError(); /* clear the non-sticky error so the next launch check is not blamed */ \
      ::crossclr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CROSSCLR_ECUDA;                                                                       \
    }                                                                                              \
  } while (0)

#define CC_REQUIRE(cond, ...)                                                                      \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::crossclr::set_error(__VA_ARGS__);                                                          \
      return CROSSCLR_EINVAL;                                                                      \
    }                                                                                              \
  } while (0)

// per-kernel event timing (api.cu); a no-op unless crossclr_timing_enable(1) was called
void timing_begin(int kernel, cudaStream_t st);
void timing_end(int kernel, cudaStream_t st);
struct TimedLaunch {
  int k; cudaStream_t st;
  TimedLaunch(int k_, cudaStream_t st_) : k(k_), st(st_) { timing_begin(k, st); }
  ~TimedLaunch() { timing_end(k, st); }
};

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return CROSSCLR_ECUDA;
  }
  return CROSSCLR_OK;
}

// ---- problem geometry (shared by every kernel) -------------------------------------------------
// Tensor-core path: every segment is padded to a multiple of 128 rows and every row to a multiple of 64 columns with zeros
// (crossclr_segment_rows / crossclr_feature_pitch), and nseg, bseg, dim, rows, row_begin, row_count below describe that PADDED
// layout -- the tile kernels never see a ragged edge.  Zero rows drop out of every product on their own (sum_j P_gj f_j with
// f_j = 0); what has to know the real extent is the forward's row sums (padded columns are masked in the segment's last
// 128-column block), finalize (padded rows carry no loss and get coef = 0), pack and grad_finish.
struct Geometry {
  int nseg, bseg, dim, rows;       // rows = nseg * bseg
  int row_begin, row_count;
  int bvalid, dvalid, rows_valid;  // the caller's rows per segment, columns, and nseg * bvalid
  int split;                       // CROSSCLR_PATH_TC_SPLIT: rows are [hi | lo | tail], the S product runs over 3 dim / 64 K chunks
  int pitch;                       // row pitch of the stacked matrix in elements (dim, or dim + CROSSCLR_ROW_TAIL on the TC path)
  int row_shift;                   // SIMT path, small temperatures: no common shift; the forward keeps an online (max, sum) per row
                                   // and stats = (log2 X_g, xpos_g), coef = (log2 Z_g, rho_g); the backward forms
                                   // 2^(x - log2 Z_g) + 2^(x - log2 Z_j).  Like the reference's max-subtracted float64 softmax
                                   // (trainer/loss.py:59-60) nothing over- or underflows, whatever the temperature.
  float k_inter;                   // log2e / tau
  float k_intra;                   // w * log2e / tau
  float shift;                     // log2-domain shift
  float inv_tau;
  float w;
};

inline float problem_max_logit(const crossclr_problem_t* p) {
  return kLog2e * fmaxf(1.0f, fabsf(p->negative_weight)) / p->temperature;
}
inline float problem_shift(const crossclr_problem_t* p) { return fmaxf(0.0f, problem_max_logit(p) - kShiftHeadroom); }
inline bool problem_needs_row_shift(const crossclr_problem_t* p) { return problem_max_logit(p) > kConstShiftMaxLogit; }

inline bool path_is_tc(int path) { return path == CROSSCLR_PATH_TC || path == CROSSCLR_PATH_TC_SPLIT; }

inline int tc_pad_rows(int bseg) { return (bseg + 127) / 128 * 128; }
inline int tc_pad_dim(int dim, bool split = false) {
  if (!split) return (dim + 63) / 64 * 64;
  return dim <= 512 ? (dim + 127) / 128 * 128 : (dim + 255) / 256 * 256;     // what the dataflow backward tiles
}

inline Geometry make_geometry(const crossclr_problem_t* p, int path = CROSSCLR_PATH_SIMT) {
  Geometry g;
  g.nseg = p->nseg; g.bseg = p->bseg; g.dim = p->dim; g.rows = p->nseg * p->bseg;
  g.row_begin = p->row_begin; g.row_count = p->row_count;
  g.bvalid = p->bseg; g.dvalid = p->dim; g.rows_valid = g.rows;
  g.split = path == CROSSCLR_PATH_TC_SPLIT ? 1 : 0;
  if (path_is_tc(path)) {
    g.bseg = tc_pad_rows(p->bseg);
    g.dim = tc_pad_dim(p->dim, g.split);
    g.rows = p->nseg * g.bseg;
    g.row_begin = p->row_begin / p->bseg * g.bseg;       // owned rows are whole segments (validate_problem)
    g.row_count = p->row_count / p->bseg * g.bseg;
  }
  g.pitch = path_is_tc(path) ? g.dim * (1 + g.split) + CROSSCLR_ROW_TAIL : g.dim;
  g.row_shift = (path == CROSSCLR_PATH_SIMT && problem_needs_row_shift(p)) ? 1 : 0;
  g.inv_tau = 1.0f / p->temperature;
  g.k_inter = kLog2e / p->temperature;
  g.k_intra = p->negative_weight * kLog2e / p->temperature;
  g.shift = problem_shift(p);
  g.w = p->negative_weight;
  return g;
}

int validate_problem(const crossclr_problem_t* p);

// modality (0 video / 1 text) and sample index of stacked row g
__host__ __device__ __forceinline__ int row_modality(int g, int bseg) { return (g / bseg) & 1; }
__host__ __device__ __forceinline__ int row_sample(int g, int bseg) {
  int s = g / bseg;
  return (s >> 1) * bseg + (g - s * bseg);
}
__host__ __device__ __forceinline__ int row_partner(int g, int bseg) {
  return ((g / bseg) & 1) ? g - bseg : g + bseg;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_float(T x);
template <> __device__ __forceinline__ float to_float<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_float<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T> __device__ __forceinline__ T from_float(float x);
template <> __device__ __forceinline__ float from_float<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_float<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// ---- entry points implemented per translation unit --------------------------------------------
// simt_kernels.cu
// out pitch: dim (fp32 rows) or dim + CROSSCLR_ROW_TAIL (16-bit rows with the residual scale in the tail)
int launch_pack(const void* x, int in_dtype, int64_t stride, int rows, int dim, void* out, int out_dtype,
                float* rnorm, cudaStream_t st);
// stats_zero / ticket_zero (optional): the [2 rows][2] row statistics and the forward's finalize ticket, zeroed in the same
// launch (or by memsets when the generic kernels run)
int launch_pack2(const void* xv, const void* xt, int in_dtype, int64_t sv, int64_t st_, int rows, int dim, void* out,
                 int out_dtype, float* rnorm, cudaStream_t st, float* stats_zero = nullptr,
                 unsigned int* ticket_zero = nullptr);
int launch_fwd_simt(const Geometry& g, const float* feat, float* stats, cudaStream_t st);
int launch_bwd_simt(const Geometry& g, const float* feat, const float* coef, float* dfhat, cudaStream_t st);
int launch_finalize(const Geometry& g, const float* stats, float* coef, double* loss, float* scal, cudaStream_t st);
// feat_dtype CROSSCLR_F32: plain normalised rows (pitch dim); CROSSCLR_F16 / BF16: (f, q) rows of the TC paths
int launch_grad_finish(const Geometry& g, const void* feat, int feat_dtype, const float* rnorm_owned,
                       const float* coef, const float* scal, bool use_sigma, const double* grad_out,
                       float grad_scale, const float* dfhat, void* dv, int64_t dv_stride, void* dt,
                       int64_t dt_stride, int out_dtype, cudaStream_t st, const float* dfhat2 = nullptr);
// tc_kernels.cu
// fused finalize (single rank, all rows owned): the last CTA of the forward writes coef / loss / scal; `ticket` must be 0
struct FwdFinalize { float* coef; double* loss; float* scal; unsigned int* ticket; };
bool fwd_tc_can_finalize(const Geometry& g);
int launch_fwd_tc(const Geometry& g, const void* feat_f16, float* stats, cudaStream_t st, const FwdFinalize* fin = nullptr);
// *two_partials: the result is dfhat + dfhat_late (grad_finish adds them)
int launch_bwd_tc(const Geometry& g, const void* feat_f16, const float* coef, const float* scal, float* dfhat,
                  float* dfhat_late, bool* two_partials, void* scratch, cudaStream_t st);
size_t bwd_pair_scratch_bytes();   // global-memory P-tile rings of the paired backward (D <= 512)
const char* bwd_tc_kernel_name(const Geometry& g);   // the backward kernel launch_bwd_tc picks for this problem
// flow_kernels.cu: dataflow backward (producer pairs -> P-tile pool -> consumer pairs), D <= 512
size_t bwd_flow_scratch_bytes(int rows, int row_count);
bool bwd_flow_applies(const Geometry& g);
int launch_scale_grad(const Geometry& g, const void* feat, int feat_dtype, const float* coef, const float* scal, bool use_sigma,
                      const double* grad_out, double mult, const float* dfhat, const float* dfhat2, double* out, cudaStream_t st);
// dfhat_late: a second [row_count][dim] partial that the kernel fills when its plan uses late consumers (*used_late)
int launch_bwd_flow(const Geometry& g, const void* feat_f16, const float* coef, const float* scal, float* dfhat,
                    float* dfhat_late, bool* used_late, void* scratch, cudaStream_t st);
int run_selftest(int variant, const uint16_t* a, const uint16_t* b, float* out, int n, int k);

// maxmargin.cu (MaxMargin_coot, trainer/loss.py:17-41; retrieval ranks = the same forward at margin 0)
size_t maxmargin_workspace_bytes(int B, int D, int dtype);
const char* maxmargin_kernel_name(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D);
// rank_row / rank_col (optional, int[B]): #{j != i : m + s_ij - s_ii > 0} and #{i != j : m + s_ij - s_jj > 0}; loss optional
int launch_maxmargin_fwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                         float margin, void* workspace, int* rank_row, int* rank_col, double* loss, cudaStream_t st);
int launch_maxmargin_bwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                         float margin, void* workspace, const double* grad_out, void* d_im, int64_t d_im_stride,
                         void* d_s, int64_t d_s_stride, int out_dtype, cudaStream_t st);
// maxmargin_tc.cu: tcgen05 score tiles with a hinge epilogue (16-bit inputs, TMA straight from the caller's tensors)
bool maxmargin_tc_applies(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D);
size_t maxmargin_tc_dacc_bytes(int B, int D);
size_t maxmargin_tc_stage_bytes(int B, int D, int dtype);   // fp32 inputs: staged fp16 [hi | lo] rows + scales, else 0
int launch_maxmargin_tc_fwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                            float margin, float* diag, float* cnt, double* acc, void* stage, int* rank_row, int* rank_col,
                            double* loss, cudaStream_t st);
int launch_maxmargin_tc_bwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                            float margin, const float* diag, const float* cnt, float* dacc, void* stage,
                            const double* grad_out, void* d_im, int64_t d_im_stride, void* d_s, int64_t d_s_stride,
                            int out_dtype, cudaStream_t st);

// peer.cu: exchange of row shards through NVLink peer memory (one kernel: peer stores + cross-rank barrier)
int peer_max_ranks();
int launch_peer_exchange(void* const* bases, uint32_t* const* flags, int n, int rank, size_t offset, size_t bytes,
                         int entry_barrier, uint32_t* state, cudaStream_t st);

}  // namespace crossclr
