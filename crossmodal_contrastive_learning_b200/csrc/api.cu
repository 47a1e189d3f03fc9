// extern "C" entry points of libcrossclr_b200 (see include/crossclr_b200.h for the contract).
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

namespace crossclr {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- per-kernel event timing ------------------------------------------------------------------------
namespace {
struct TimingRec { int kernel; cudaEvent_t a, b; };
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mu;
std::vector<TimingRec> g_timing;
thread_local cudaEvent_t tl_start = nullptr;
}  // namespace

// One thread spinning on %globaltimer for ~30 us.  Queued in front of the start event of a timed launch, it keeps the
// stream busy while the host submits the launches inside the bracket, so that the events bracket back-to-back device work
// even when the stream is otherwise starved (eager, host-bound steps); one idle thread does not move clocks or power.
__global__ void timing_spin_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

void timing_begin(int kernel, cudaStream_t st) {
  (void)kernel;
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  cudaEventCreate(&tl_start);
  timing_spin_kernel<<<1, 1, 0, st>>>(30000ull);
  cudaEventRecord(tl_start, st);
}
void timing_end(int kernel, cudaStream_t st) {
  if (tl_start == nullptr) return;
  TimingRec r{kernel, tl_start, nullptr};
  tl_start = nullptr;
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, st);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing.push_back(r);
}

int validate_problem(const crossclr_problem_t* p) {
  CC_REQUIRE(p != nullptr, "problem is NULL");
  CC_REQUIRE(p->nseg >= 2 && (p->nseg % 2) == 0, "nseg must be a positive even number (got %d)", p->nseg);
  CC_REQUIRE(p->bseg >= 1 && p->dim >= 1, "bseg and dim must be >= 1 (got %d, %d)", p->bseg, p->dim);
  CC_REQUIRE((int64_t)p->nseg * p->bseg < (int64_t)1 << 30, "too many stacked rows");
  CC_REQUIRE(p->row_count == 2 * p->bseg && p->row_begin >= 0 && p->row_begin % (2 * p->bseg) == 0 &&
                 p->row_begin + p->row_count <= p->nseg * p->bseg,
             "owned rows must be one rank's [video; text] pair of segments (row_begin %d row_count %d)",
             p->row_begin, p->row_count);
  CC_REQUIRE(p->temperature > 0.f && std::isfinite(p->temperature), "temperature must be positive and finite");
  CC_REQUIRE(std::isfinite(p->negative_weight), "negative_weight must be finite");
  return CROSSCLR_OK;
}

// `auto` takes the tensor-core path unless the zero padding (segments to 128 rows, rows to 64 columns) would dominate the work
static bool tc_worthwhile(const crossclr_problem_t* p) { return p->bseg >= 96 && p->dim >= 48; }

// CROSSCLR_PATH_TC_SPLIT runs its backward on the dataflow kernel only (flow_kernels.cu: flow_plan's shape rule, restated
// here as a pure function of the problem so that planning needs no device)
static bool split_shape_ok(const crossclr_problem_t* p) {
  const Geometry g = make_geometry(p, CROSSCLR_PATH_TC_SPLIT);
  return g.dim <= 1024 && g.rows >= 2048 && g.rows / 256 < (1 << 20);
}

// gradient accumulator of the owned rows, in the path's own (for TC: padded) geometry
static size_t dfhat_bytes(const Geometry& g) {
  return ((size_t)g.row_count * (size_t)g.dim * sizeof(float) + 255) & ~(size_t)255;
}

}  // namespace crossclr

using namespace crossclr;

extern "C" {

int crossclr_version(void) { return CROSSCLR_VERSION; }

const char* crossclr_last_error(void) { return g_err; }

int crossclr_device_supported(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    return CROSSCLR_ECUDA;
  }
  return major == 10 ? 1 : 0;
}

int crossclr_choose_path(const crossclr_problem_t* p, int in_dtype, int exact) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(in_dtype == CROSSCLR_F32 || in_dtype == CROSSCLR_F16 || in_dtype == CROSSCLR_BF16,
             "crossclr_choose_path: unsupported input dtype %d", in_dtype);
  if (!tc_worthwhile(p) || exact) return CROSSCLR_PATH_SIMT;
  if (problem_needs_row_shift(p)) return CROSSCLR_PATH_SIMT;      // temperature below the constant shift's range (~0.0073)
  if (in_dtype == CROSSCLR_F32)     // fp32 values are not rounded to fp16 unasked: hi + lo operands where they apply, else exact
    return split_shape_ok(p) ? CROSSCLR_PATH_TC_SPLIT : CROSSCLR_PATH_SIMT;
  return CROSSCLR_PATH_TC;
}

int crossclr_feature_dtype(int path) {
  if (path == CROSSCLR_PATH_SIMT) return CROSSCLR_F32;
  if (path == CROSSCLR_PATH_TC) return CROSSCLR_F16;
  if (path == CROSSCLR_PATH_TC_SPLIT) return CROSSCLR_F16X2;
  set_error("crossclr_feature_dtype: path must be SIMT, TC or TC_SPLIT");
  return CROSSCLR_EINVAL;
}

int64_t crossclr_feature_pitch(int path, int32_t dim) {
  if (path == CROSSCLR_PATH_SIMT) return dim;
  if (path == CROSSCLR_PATH_TC) return (int64_t)tc_pad_dim(dim) + CROSSCLR_ROW_TAIL;
  if (path == CROSSCLR_PATH_TC_SPLIT) return 2 * (int64_t)tc_pad_dim(dim, true) + CROSSCLR_ROW_TAIL;
  set_error("crossclr_feature_pitch: path must be SIMT, TC or TC_SPLIT");
  return CROSSCLR_EINVAL;
}

int64_t crossclr_segment_rows(int path, int32_t bseg) {
  if (path == CROSSCLR_PATH_SIMT) return bseg;
  if (path_is_tc(path)) return tc_pad_rows(bseg);
  set_error("crossclr_segment_rows: path must be SIMT, TC or TC_SPLIT");
  return CROSSCLR_EINVAL;
}

// workspace of the backward: [dfhat | dfhat of the late consumers | flag word block | kernel scratch]
static size_t ws_flag_offset(const Geometry& g) { return 2 * dfhat_bytes(g); }
static size_t ws_scratch_offset(const Geometry& g) { return 2 * dfhat_bytes(g) + 256; }
static size_t workspace_bytes(const crossclr_problem_t* p, int path) {
  const Geometry g = make_geometry(p, path);
  if (!path_is_tc(path)) return dfhat_bytes(g);
  // + P-tile scratch of the role-specialised backward kernels (cluster rings, or the dataflow kernel's pool + control words)
  return ws_scratch_offset(g) + std::max(bwd_pair_scratch_bytes(), bwd_flow_scratch_bytes(g.rows, g.row_count));
}

size_t crossclr_workspace_bytes(const crossclr_problem_t* p, int path) {
  if (validate_problem(p)) return 0;
  return workspace_bytes(p, path);
}

float crossclr_shift(const crossclr_problem_t* p) { return problem_shift(p); }

const char* crossclr_bwd_kernel_name(const crossclr_problem_t* p, int path) {
  if (validate_problem(p)) return "";
  if (path == CROSSCLR_PATH_SIMT) return "bwd_simt_kernel";
  if (!path_is_tc(path)) return "";
  return bwd_tc_kernel_name(make_geometry(p, path));
}

int64_t crossclr_launch_count(void) { return g_launches.load(); }

int crossclr_pack(const void* x, int in_dtype, int64_t x_row_stride, int32_t rows, int32_t dim, void* feat_out,
                  int feat_dtype, float* rnorm_out, void* stream) {
  CC_REQUIRE(x && feat_out && rnorm_out, "crossclr_pack: NULL pointer");
  CC_REQUIRE(rows >= 0 && dim >= 1 && x_row_stride >= dim, "crossclr_pack: bad shape/stride");
  return launch_pack(x, in_dtype, x_row_stride, rows, dim, feat_out, feat_dtype, rnorm_out, (cudaStream_t)stream);
}

int crossclr_pack2(const void* video, const void* text, int in_dtype, int64_t video_row_stride, int64_t text_row_stride,
                   int32_t rows, int32_t dim, void* feat_out, int feat_dtype, float* rnorm_out, void* stream) {
  CC_REQUIRE(video && text && feat_out && rnorm_out, "crossclr_pack2: NULL pointer");
  CC_REQUIRE(rows >= 0 && dim >= 1 && video_row_stride >= dim && text_row_stride >= dim,
             "crossclr_pack2: bad shape/stride");
  return launch_pack2(video, text, in_dtype, video_row_stride, text_row_stride, rows, dim, feat_out, feat_dtype,
                      rnorm_out, (cudaStream_t)stream);
}

int crossclr_forward(const crossclr_problem_t* p, int path, const void* video, const void* text, int in_dtype,
                     int64_t video_row_stride, int64_t text_row_stride, void* feat, float* rnorm, float* stats,
                     float* coef, float* scal, double* loss_out, void* stream) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(p->nseg == 2, "crossclr_forward is the single-rank entry point (nseg must be 2, got %d)", p->nseg);
  const int fdt = crossclr_feature_dtype(path);
  if (fdt < 0) return fdt;
  CC_REQUIRE(video && text && feat && rnorm && stats && coef && scal && loss_out, "crossclr_forward: NULL pointer");
  const Geometry g = make_geometry(p, path);
  CC_REQUIRE(!(path_is_tc(path) && problem_needs_row_shift(p)), "temperature %g is below the tensor-core path's range: use "
             "CROSSCLR_PATH_SIMT (crossclr_choose_path does)", (double)p->temperature);
  if (path_is_tc(path) && fwd_tc_can_finalize(g)) {
    // three launches become two: pack also zeroes the statistics, the forward's last CTA finalizes
    CC_REQUIRE(p->bseg >= 0 && video_row_stride >= p->dim && text_row_stride >= p->dim, "crossclr_forward: bad shape/stride");
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scal + 3);
    rc = launch_pack2(video, text, in_dtype, video_row_stride, text_row_stride, p->bseg, p->dim, feat, fdt, rnorm,
                      (cudaStream_t)stream, stats, ticket);
    if (rc) return rc;
    const FwdFinalize fin{coef, loss_out, scal, ticket};
    return launch_fwd_tc(g, feat, stats, (cudaStream_t)stream, &fin);
  }
  rc = crossclr_pack2(video, text, in_dtype, video_row_stride, text_row_stride, p->bseg, p->dim, feat, fdt, rnorm,
                      stream);
  if (rc) return rc;
  rc = crossclr_fwd(p, path, feat, stats, nullptr, 0, stream);
  if (rc) return rc;
  return crossclr_finalize(p, path, stats, coef, loss_out, scal, stream);
}

int crossclr_fwd(const crossclr_problem_t* p, int path, const void* feat, float* stats, void* workspace,
                 size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(feat && stats, "crossclr_fwd: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const Geometry g = make_geometry(p, path);
  CC_REQUIRE(path == CROSSCLR_PATH_SIMT || path_is_tc(path), "crossclr_fwd: path must be SIMT or TC (got %d)", path);
  CC_REQUIRE(!(path_is_tc(path) && problem_needs_row_shift(p)), "temperature %g is below the tensor-core path's range "
             "(log2e max(1,|w|)/tau <= %g): use CROSSCLR_PATH_SIMT (crossclr_choose_path does)", (double)p->temperature,
             (double)kConstShiftMaxLogit);
  // X accumulates with atomics across column ranges: zero the owned rows first
  CC_CHECK_CUDA(cudaMemsetAsync(stats + 2 * (size_t)g.row_begin, 0, 2 * (size_t)g.row_count * sizeof(float), st));
  if (path_is_tc(path)) return launch_fwd_tc(g, feat, stats, st);
  return launch_fwd_simt(g, (const float*)feat, stats, st);
}

int crossclr_finalize(const crossclr_problem_t* p, int path, const float* stats, float* coef, double* loss_out, float* scal,
                      void* stream) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(stats && coef && loss_out && scal, "crossclr_finalize: NULL pointer");
  CC_REQUIRE(path == CROSSCLR_PATH_SIMT || path_is_tc(path), "crossclr_finalize: path must be SIMT or TC (got %d)", path);
  return launch_finalize(make_geometry(p, path), stats, coef, loss_out, scal, (cudaStream_t)stream);
}

// stage 1 of the backward: dfhat (head of the workspace) = sum_j P~ f_j over the owned rows
static int bwd_accumulate(const crossclr_problem_t* p, int path, const Geometry& g, const void* feat, const float* coef,
                          const float* scal, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace == nullptr || workspace_bytes < ::workspace_bytes(p, path)) {
    set_error("crossclr_bwd: workspace too small (%zu < %zu)", workspace_bytes, ::workspace_bytes(p, path));
    return CROSSCLR_EWORKSPACE;
  }
  float* dfhat = (float*)workspace;
  if (path_is_tc(path)) {
    bool two = false;
    int rc = launch_bwd_tc(g, feat, coef, scal, dfhat, (float*)((char*)workspace + dfhat_bytes(g)), &two,
                           (char*)workspace + ws_scratch_offset(g), st);
    if (rc) return rc;
    // the finish stage may run from another call (crossclr_bwd_finish): leave word 0 of the flag block = partial count - 1
    CC_CHECK_CUDA(cudaMemsetAsync((char*)workspace + ws_flag_offset(g), two ? 1 : 0, 4, st));
    return CROSSCLR_OK;
  }
  CC_REQUIRE(path == CROSSCLR_PATH_SIMT, "crossclr_bwd: path must be SIMT or TC (got %d)", path);
  return launch_bwd_simt(g, (const float*)feat, coef, dfhat, st);
}

int crossclr_bwd_accumulate(const crossclr_problem_t* p, int path, const void* feat, const float* coef, const float* scal,
                            void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(feat && coef && scal, "crossclr_bwd_accumulate: NULL pointer");
  return bwd_accumulate(p, path, make_geometry(p, path), feat, coef, scal, workspace, workspace_bytes, (cudaStream_t)stream);
}

int crossclr_bwd_finish(const crossclr_problem_t* p, int path, const void* feat, const float* rnorm_owned, const float* coef,
                        const float* scal, const double* grad_out, float grad_scale, void* dv, int64_t dv_row_stride,
                        void* dt, int64_t dt_row_stride, int out_dtype, const void* workspace, void* stream) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(feat && rnorm_owned && coef && scal && dv && dt && workspace, "crossclr_bwd_finish: NULL pointer");
  CC_REQUIRE(dv_row_stride >= p->dim && dt_row_stride >= p->dim, "crossclr_bwd_finish: output row stride < dim");
  CC_REQUIRE(path == CROSSCLR_PATH_SIMT || path_is_tc(path), "crossclr_bwd_finish: path must be SIMT or TC (got %d)", path);
  const bool tc = path_is_tc(path);
  const Geometry g = make_geometry(p, path);
  // TC paths: the second partial (late consumers of the dataflow kernel) counts iff the flag word says so (read on device)
  const float* dfhat2 = tc ? (const float*)((const char*)workspace + dfhat_bytes(g)) : nullptr;
  return launch_grad_finish(g, feat, tc ? crossclr_feature_dtype(path) : CROSSCLR_F32, rnorm_owned, coef, scal, tc, grad_out,
                            grad_scale, (const float*)workspace, dv, dv_row_stride, dt, dt_row_stride, out_dtype,
                            (cudaStream_t)stream, dfhat2);
}

int crossclr_bwd_scale_grad(const crossclr_problem_t* p, int path, const void* feat, const float* coef, const float* scal,
                            const double* grad_out, float grad_scale, float logit_scale, const void* workspace,
                            double* dscale_out, void* stream) {
  int rc = validate_problem(p);
  if (rc) return rc;
  CC_REQUIRE(feat && coef && scal && workspace && dscale_out, "crossclr_bwd_scale_grad: NULL pointer");
  CC_REQUIRE(path == CROSSCLR_PATH_SIMT || path_is_tc(path), "crossclr_bwd_scale_grad: path must be SIMT or TC (got %d)", path);
  CC_REQUIRE(std::isfinite(logit_scale) && logit_scale != 0.f, "crossclr_bwd_scale_grad: logit_scale must be finite and non-zero");
  const bool tc = path_is_tc(path);
  const Geometry g = make_geometry(p, path);
  const float* dfhat2 = tc ? (const float*)((const char*)workspace + dfhat_bytes(g)) : nullptr;
  // p->temperature is the EFFECTIVE temperature tau / s the kernels ran at
  const double mult = (double)grad_scale / ((double)logit_scale * (double)g.rows_valid * (double)p->temperature);
  return launch_scale_grad(g, feat, tc ? crossclr_feature_dtype(path) : CROSSCLR_F32, coef, scal, tc, grad_out, mult,
                           (const float*)workspace, dfhat2, dscale_out, (cudaStream_t)stream);
}

int crossclr_bwd(const crossclr_problem_t* p, int path, const void* feat, const float* rnorm_owned, const float* coef,
                 const float* scal, const double* grad_out, float grad_scale, void* dv, int64_t dv_row_stride,
                 void* dt, int64_t dt_row_stride, int out_dtype, void* workspace, size_t workspace_bytes,
                 void* stream) {
  int rc = crossclr_bwd_accumulate(p, path, feat, coef, scal, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return crossclr_bwd_finish(p, path, feat, rnorm_owned, coef, scal, grad_out, grad_scale, dv, dv_row_stride, dt, dt_row_stride,
                             out_dtype, workspace, stream);
}

size_t crossclr_maxmargin_workspace_bytes(int32_t batch, int32_t dim, int dtype) {
  return (batch >= 1 && dim >= 1) ? maxmargin_workspace_bytes(batch, dim, dtype) : 0;
}

const char* crossclr_maxmargin_kernel_name(const void* im, const void* s, int dtype, int64_t im_row_stride,
                                           int64_t s_row_stride, int32_t batch, int32_t dim) {
  return maxmargin_kernel_name(im, im_row_stride, s, s_row_stride, dtype, batch, dim);
}

static int maxmargin_check(const char* who, const void* im, const void* s, int dtype, int64_t im_row_stride,
                           int64_t s_row_stride, int32_t batch, int32_t dim, const void* workspace, size_t workspace_bytes) {
  CC_REQUIRE(im && s && workspace, "%s: NULL pointer", who);
  CC_REQUIRE(batch >= 1 && dim >= 1 && im_row_stride >= dim && s_row_stride >= dim, "%s: bad shape/stride", who);
  CC_REQUIRE((int64_t)batch * batch < ((int64_t)1 << 40), "%s: batch too large", who);
  if (workspace_bytes < maxmargin_workspace_bytes(batch, dim, dtype)) {
    set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, maxmargin_workspace_bytes(batch, dim, dtype));
    return CROSSCLR_EWORKSPACE;
  }
  return CROSSCLR_OK;
}

int crossclr_maxmargin_fwd(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                           int32_t batch, int32_t dim, float margin, void* workspace, size_t workspace_bytes,
                           double* loss_out, void* stream) {
  CC_REQUIRE(loss_out, "crossclr_maxmargin_fwd: NULL pointer");
  int rc = maxmargin_check("crossclr_maxmargin_fwd", im, s, dtype, im_row_stride, s_row_stride, batch, dim, workspace, workspace_bytes);
  if (rc) return rc;
  return launch_maxmargin_fwd(im, im_row_stride, s, s_row_stride, dtype, batch, dim, margin, workspace, nullptr, nullptr,
                              loss_out, (cudaStream_t)stream);
}

int crossclr_maxmargin_bwd(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                           int32_t batch, int32_t dim, float margin, void* workspace, size_t workspace_bytes,
                           const double* grad_out, void* d_im, int64_t d_im_row_stride, void* d_s, int64_t d_s_row_stride,
                           int out_dtype, void* stream) {
  CC_REQUIRE(d_im && d_s, "crossclr_maxmargin_bwd: NULL pointer");
  CC_REQUIRE(d_im_row_stride >= dim && d_s_row_stride >= dim, "crossclr_maxmargin_bwd: bad shape/stride");
  int rc = maxmargin_check("crossclr_maxmargin_bwd", im, s, dtype, im_row_stride, s_row_stride, batch, dim, workspace, workspace_bytes);
  if (rc) return rc;
  return launch_maxmargin_bwd(im, im_row_stride, s, s_row_stride, dtype, batch, dim, margin, workspace, grad_out, d_im,
                              d_im_row_stride, d_s, d_s_row_stride, out_dtype, (cudaStream_t)stream);
}

int crossclr_retrieval_ranks(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                             int32_t batch, int32_t dim, void* workspace, size_t workspace_bytes, int32_t* rank_im2s,
                             int32_t* rank_s2im, void* stream) {
  CC_REQUIRE(rank_im2s && rank_s2im, "crossclr_retrieval_ranks: NULL pointer");
  int rc = maxmargin_check("crossclr_retrieval_ranks", im, s, dtype, im_row_stride, s_row_stride, batch, dim, workspace, workspace_bytes);
  if (rc) return rc;
  return launch_maxmargin_fwd(im, im_row_stride, s, s_row_stride, dtype, batch, dim, 0.0f, workspace, rank_im2s, rank_s2im,
                              nullptr, (cudaStream_t)stream);
}

int crossclr_peer_alloc(size_t bytes, void** ptr_out) {
  CC_REQUIRE(ptr_out && bytes > 0, "crossclr_peer_alloc: bad argument");
  void* p = nullptr;
  CC_CHECK_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) { cudaFree(p); (void)cudaGetLastError(); set_error("crossclr_peer_alloc: memset failed: %s", cudaGetErrorString(e)); return CROSSCLR_ECUDA; }
  *ptr_out = p;
  return CROSSCLR_OK;
}

int crossclr_peer_free(void* ptr) {
  if (ptr != nullptr) CC_CHECK_CUDA(cudaFree(ptr));
  return CROSSCLR_OK;
}

int crossclr_peer_export(const void* ptr, void* handle_out) {
  CC_REQUIRE(ptr && handle_out, "crossclr_peer_export: NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == CROSSCLR_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CC_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle_out, &h, sizeof(h));
  return CROSSCLR_OK;
}

int crossclr_peer_import(const void* handle, void** ptr_out) {
  CC_REQUIRE(handle && ptr_out, "crossclr_peer_import: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CC_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return CROSSCLR_OK;
}

int crossclr_peer_release(void* ptr) {
  if (ptr != nullptr) CC_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return CROSSCLR_OK;
}

int crossclr_peer_exchange(void* const* bases, uint32_t* const* flags, int32_t n_ranks, int32_t rank, size_t offset,
                           size_t bytes, int entry_barrier, uint32_t* state, void* stream) {
  CC_REQUIRE(bases && flags && state, "crossclr_peer_exchange: NULL pointer");
  CC_REQUIRE(n_ranks >= 2 && n_ranks <= peer_max_ranks() && rank >= 0 && rank < n_ranks,
             "crossclr_peer_exchange: %d ranks (rank %d) outside [2, %d]", n_ranks, rank, peer_max_ranks());
  CC_REQUIRE(offset % 16 == 0 && bytes % 16 == 0, "crossclr_peer_exchange: offset and size must be multiples of 16 bytes");
  for (int i = 0; i < n_ranks; ++i) CC_REQUIRE(bases[i] && flags[i], "crossclr_peer_exchange: NULL buffer of rank %d", i);
  return launch_peer_exchange(bases, flags, n_ranks, rank, offset, bytes, entry_barrier, state, (cudaStream_t)stream);
}

int crossclr_timing_enable(int on) {
  g_timing_on.store(on ? 1 : 0);
  return CROSSCLR_OK;
}

int crossclr_timing_read(int kernel, double* total_ms, int64_t* launches) {
  CC_REQUIRE(kernel >= 0 && kernel < CROSSCLR_K_COUNT && total_ms && launches, "crossclr_timing_read: bad argument");
  double ms = 0.0;
  int64_t n = 0;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  std::vector<TimingRec> keep;
  for (const TimingRec& r : g_timing) {
    if (r.kernel != kernel) { keep.push_back(r); continue; }
    float t = 0.f;
    CC_CHECK_CUDA(cudaEventSynchronize(r.b));
    CC_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    ms += t;
    ++n;
  }
  g_timing.swap(keep);
  *total_ms = ms;
  *launches = n;
  return CROSSCLR_OK;
}

int crossclr_selftest(int variant, const uint16_t* a_host, const uint16_t* b_host, float* out_host, int32_t n,
                      int32_t k) {
  CC_REQUIRE(a_host && b_host && out_host, "crossclr_selftest: NULL pointer");
  return run_selftest(variant, a_host, b_host, out_host, n, k);
}

}  // extern "C"
