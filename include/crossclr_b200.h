/*
 * crossclr_b200.h -- C ABI of the B200-native CrossCLR criterion (libcrossclr_b200.so).
 *
 * The reference (amazon-science/crossmodal-contrastive-learning) has no FFI layer: its boundary is the
 * Python nn.Module `trainer/loss.py:44-114  CrossCLR_onlyIntraModality`.  The Python mirror of that
 * module (crossmodal_contrastive_learning_b200/loss.py) keeps the module surface and calls the entry
 * points below through ctypes.  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CROSSCLR_E* code on failure; the message is
 *     available from crossclr_last_error() (thread-local).
 *   - all pointers are BORROWED DEVICE pointers (owned by the caller, e.g. torch's allocator) unless
 *     a parameter says "host".  No allocation, no stream/device synchronisation and no host<->device
 *     copies happen inside; `stream` is a cudaStream_t passed as void*.
 *   - "stacked matrix": the R = nseg*bseg L2-normalised feature rows laid out as nseg segments of bseg rows,
 *     segment s holding modality (s & 1) (0 = video, 1 = text) of rank (s >> 1); i.e. the layout an
 *     all-gather of per-rank [2][bseg][dim] blocks produces.  One GPU: nseg = 2 (V rows then T rows).
 *     Stacked row g belongs to sample (g / bseg >> 1) * bseg + g % bseg.
 *   - the calling rank owns stacked rows [row_begin, row_begin + row_count).
 */
#ifndef CROSSCLR_B200_H_
#define CROSSCLR_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CROSSCLR_API __attribute__((visibility("default")))
#else
#define CROSSCLR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CROSSCLR_VERSION 140          /* 0.1.4 */

/* error codes */
#define CROSSCLR_OK            0
#define CROSSCLR_EINVAL       -1      /* bad argument / unsupported shape for the requested path   */
#define CROSSCLR_ECUDA        -2      /* a CUDA runtime/driver call failed (see crossclr_last_error) */
#define CROSSCLR_EWORKSPACE   -3      /* workspace too small                                          */
#define CROSSCLR_EUNSUPPORTED -4      /* device is not sm_100                                         */

/* element types */
#define CROSSCLR_F32  0
#define CROSSCLR_F16  1
#define CROSSCLR_BF16 2
#define CROSSCLR_F16X2 3              /* stacked rows only: fp16 hi + lo pairs of CROSSCLR_PATH_TC_SPLIT                */

/* kernel families ("path") */
#define CROSSCLR_PATH_AUTO 0          /* crossclr_choose_path decides                                 */
#define CROSSCLR_PATH_SIMT 1          /* fp32 CUDA-core kernels, any B / D, fp32 stacked features     */
#define CROSSCLR_PATH_TC   2          /* tcgen05 + TMA + TMEM kernels, fp16 stacked rows, any B / D:
                                         segments are zero-padded to 128 rows, rows to 64 columns     */
#define CROSSCLR_PATH_TC_SPLIT 3      /* the same kernels on fp32 inputs kept as fp16 hi + lo pairs: the similarity
                                         product runs over K = 3 D as [hi | lo | hi] x [hi | hi | lo], i.e.
                                         hi.hi + lo.hi + hi.lo -- fp32-grade logits (the dropped lo.lo term is 2^-24).
                                         Needs >= 2048 stacked rows and dim <= 1024 (dataflow backward only)        */

/* Stacked rows of the tensor-core path.  The layout is PADDED: every segment holds crossclr_segment_rows(path, bseg) =
 * roundup(bseg, 128) rows, of which the first bseg are the caller's and the rest are zero rows (q = 1), and every row holds
 * roundup(dim, 64) feature columns (zeros beyond dim) followed by the tail: the row pitch is
 * crossclr_feature_pitch(path, dim) = roundup(dim, 64) + CROSSCLR_ROW_TAIL elements.  stats and coef are indexed by PADDED
 * stacked row (nseg * crossclr_segment_rows rows; padded rows hold no statistics and coef = 0), rnorm and the gradients by
 * the caller's rows.  A shape with bseg % 128 == 0 and dim % 64 == 0 has no padding at all.  Zero rows drop out of every
 * product; the forward masks them out of the row sums.  (Reference: trainer/loss.py:83-88 takes any [B, D].)
 * Row g is stored as f_g; the first 4 bytes of the tail hold the fp32 residual scale q_g, and the L2-normalised row is
 * q_g * f_g.
 *   16-bit inputs: f_g = x_g * 2^-e (EXACT: a power-of-two rescale into ||f|| in [1, 2), stored as fp16 -- a bf16 value
 *                  below 2 is exactly representable in fp16 down to 2^-17), q_g = 2^e / ||x_g|| in (1/2, 1]; rows inside the
 *                  eps clamp (||x|| < 1e-12) are stored as in the fp32 case.
 *                  The tensor cores then multiply the caller's own values and the cosine is formed in the fp32
 *                  epilogue as q_i q_j (f_i . f_j): no operand rounding at all (gradients ~1e-5 of the reference).
 *                  (kind::f16 MMAs need A and B of one format and the probability operand must be fp16 -- bf16 would
 *                  cost 1.6e-3 on the gradients -- hence fp16 rows for bf16 inputs too.)
 *   fp32 inputs:   the same rescale, then rounded to fp16 (operand rounding 2^-12: gradients within ~2e-5 / tau of the
 *                  reference; exact again if the values happen to be 16-bit representable).
 * CROSSCLR_PATH_TC_SPLIT rows are [hi (Dp) | lo (Dp) | tail]: hi = fp16(f), lo = fp16(f - hi) of the rescaled fp32 row f; Dp here
 * is roundup(dim, 128) (roundup(dim, 256) above 512) and the pitch 2 Dp + CROSSCLR_ROW_TAIL.
 * The SIMT path keeps plain fp32 normalised rows, pitch dim, bseg rows per segment, no tail. */
#define CROSSCLR_ROW_TAIL 64

typedef struct crossclr_problem {
  int32_t nseg;             /* segments in the stacked matrix = 2 * world_size                        */
  int32_t bseg;             /* rows per segment = local batch                                         */
  int32_t dim;              /* embedding dim D                                                        */
  int32_t row_begin;        /* first stacked row owned by the caller = 2 * rank * bseg                */
  int32_t row_count;        /* owned rows = 2 * bseg                                                  */
  float   temperature;      /* tau      (trainer/loss.py:54, read per call at :90-93)                 */
  float   negative_weight;  /* w        (trainer/loss.py:56, read per call at :99-100)                */
} crossclr_problem_t;

CROSSCLR_API int         crossclr_version(void);
CROSSCLR_API const char* crossclr_last_error(void);

/* 1 if `device` can run this library (compute capability 10.x), 0 otherwise, <0 on error. */
CROSSCLR_API int crossclr_device_supported(int device);

/* Resolve CROSSCLR_PATH_AUTO for a problem and input dtype.  16-bit inputs: CROSSCLR_PATH_TC when the temperature allows it
 * and the shape is not so small that the zero padding would dominate (bseg >= 96 and dim >= 48), else CROSSCLR_PATH_SIMT.
 * fp32 inputs are never rounded to fp16 behind the caller's back (the reference multiplies them in fp32,
 * trainer/loss.py:83-88): CROSSCLR_PATH_TC_SPLIT where it applies, else CROSSCLR_PATH_SIMT; CROSSCLR_PATH_TC on fp32 inputs
 * (operands rounded to fp16, gradients within ~2e-5 / tau) is an explicit opt-in.  exact != 0 forces the fp32 SIMT path.
 * Temperature: the TC kernels use ONE log2-domain shift for all rows, which keeps every row representable while
 * log2e * max(1,|w|) / tau <= 200 (tau >= ~0.0073 at |w| <= 1); they return CROSSCLR_EINVAL beyond.  The SIMT path switches to
 * per-row online maxima in the log2 domain there and follows the reference's max-subtracted float64 softmax
 * (trainer/loss.py:59-60) at any temperature; in that mode stats[2g+0] = log2 X_g (not X_g / 2^shift) and
 * coef[2g+0] = log2 Z_g (not 1 / Z_g). */
CROSSCLR_API int crossclr_choose_path(const crossclr_problem_t* p, int in_dtype, int exact);

/* Stacked-row element type a path consumes: SIMT -> CROSSCLR_F32, TC -> CROSSCLR_F16, TC_SPLIT -> CROSSCLR_F16X2 (fp16 storage). */
CROSSCLR_API int crossclr_feature_dtype(int path);

/* Row pitch (in elements) of the stacked matrix of a path: dim for SIMT, roundup(dim, 64) + CROSSCLR_ROW_TAIL for TC,
 * 2 Dp + CROSSCLR_ROW_TAIL for TC_SPLIT. */
CROSSCLR_API int64_t crossclr_feature_pitch(int path, int32_t dim);

/* Rows per segment of the stacked matrix (and of stats / coef) of a path: bseg for SIMT, roundup(bseg, 128) for TC / TC_SPLIT. */
CROSSCLR_API int64_t crossclr_segment_rows(int path, int32_t bseg);

/* Bytes of scratch crossclr_bwd needs for this problem and path (crossclr_fwd needs none). */
CROSSCLR_API size_t crossclr_workspace_bytes(const crossclr_problem_t* p, int path);

/*
 * L2-normalise one modality block `x` ([rows][dim], element stride 1, row stride `x_row_stride`
 * elements, dtype `in_dtype`) into its segment of the stacked matrix: `feat_out` points at the first
 * row of that segment (dtype `feat_dtype`; CROSSCLR_F32: row stride dim, receives x / max(||x||_2, 1e-12);
 * CROSSCLR_F16: the padded TC layout -- roundup(rows, 128) rows of pitch roundup(dim, 64) + CROSSCLR_ROW_TAIL are written,
 * (f, q) as described above and zeros for the padding);
 * `rnorm_out[rows]` receives 1 / max(||x||_2, 1e-12) (kept by the caller for the backward).
 * Replaces: trainer/loss.py:79-80 (F.normalize x2).
 */
CROSSCLR_API int crossclr_pack(const void* x, int in_dtype, int64_t x_row_stride, int32_t rows, int32_t dim,
                  void* feat_out, int feat_dtype, float* rnorm_out, void* stream);

/*
 * Both modality blocks of one rank in one launch: rows of `video` then rows of `text` are normalised into the
 * rank's two consecutive segments starting at `feat_out` ([2*crossclr_segment_rows][pitch], padding rows and columns
 * written as zeros when feat_dtype is the TC path's); `rnorm_out[2*rows]`.
 * Replaces: trainer/loss.py:79-80.
 */
CROSSCLR_API int crossclr_pack2(const void* video, const void* text, int in_dtype, int64_t video_row_stride,
                   int64_t text_row_stride, int32_t rows, int32_t dim, void* feat_out, int feat_dtype,
                   float* rnorm_out, void* stream);

/*
 * Single-rank forward in one call (nseg == 2): crossclr_pack2 -> crossclr_fwd -> crossclr_finalize on `stream`.
 * Buffers as in the individual calls, S = crossclr_segment_rows(path, bseg): feat [2*S][crossclr_feature_pitch(path, dim)]
 * (dtype crossclr_feature_dtype(path)), rnorm [2*bseg], stats / coef [2*S][2], scal [4], loss_out double[1].
 * Replaces: trainer/loss.py:79-114 (forward).
 */
CROSSCLR_API int crossclr_forward(const crossclr_problem_t* p, int path, const void* video, const void* text, int in_dtype,
                     int64_t video_row_stride, int64_t text_row_stride, void* feat, float* rnorm, float* stats,
                     float* coef, float* scal, double* loss_out, void* stream);

/*
 * Forward statistics of the owned rows.  For every owned stacked row g writes
 *   stats[2g+0] = X_g    = sum over the 2B-1 non-positive logits of 2^(logit*log2e - shift)
 *                          (includes the intra-modal diagonal, which is logit 0: loss.py:65,96-97)
 *   stats[2g+1] = xpos_g = positive logit * log2e - shift
 * with shift = crossclr_shift(p).  `feat` is the full stacked matrix [nseg*bseg][pitch] in the layout of
 * `path`; `stats` has room for [nseg*crossclr_segment_rows(path, bseg)][2] floats (only the owned segments are written).
 * Replaces: trainer/loss.py:83-100 (4 GEMMs, /tau, mask, weight, concat) and the row reductions of
 * :59-60; no B x B intermediate is written to memory.
 */
CROSSCLR_API int crossclr_fwd(const crossclr_problem_t* p, int path, const void* feat, float* stats, void* workspace,
                 size_t workspace_bytes, void* stream);

/*
 * Loss and backward coefficients from the (all-gathered) statistics of ALL nseg*crossclr_segment_rows(path, bseg) rows
 * (`path`: the one crossclr_fwd ran on -- it fixes the row layout and the meaning of stats[.,0]):
 *   loss_out[0] (double) = (1/2B) sum_g log1p(X_g * 2^-xpos_g)          trainer/loss.py:60,:111-114
 *   coef[2g+0] = 1/Z_g, coef[2g+1] = X_g/Z_g   with Z_g = X_g + 2^xpos_g (shifted units)
 *   scal[0] = sigma (power-of-two scale applied to the fp16 probability tiles), scal[1] = 1/sigma,
 *   scal[2] = max_g X_g/Z_g, scal[3] scratch (the single-rank forward's finalize ticket)
 */
CROSSCLR_API int crossclr_finalize(const crossclr_problem_t* p, int path, const float* stats, float* coef,
                      double* loss_out, float* scal, void* stream);

/*
 * Gradients of loss * (*grad_out) * grad_scale w.r.t. the caller's own video/text rows.
 * `rnorm_owned[row_count]`: reciprocal norms of the owned rows (video rows then text rows) from
 * crossclr_pack.  `grad_out` is a DEVICE pointer to the upstream scalar gradient (double) or NULL for
 * 1.0.  dv/dt: [bseg][dim] outputs, dtype `out_dtype`, row strides in elements.
 * Replaces: the autograd backward of trainer/loss.py:79-114 (98 aten ops, 8 mm).
 */
CROSSCLR_API int crossclr_bwd(const crossclr_problem_t* p, int path, const void* feat, const float* rnorm_owned,
                 const float* coef, const float* scal, const double* grad_out, float grad_scale, void* dv,
                 int64_t dv_row_stride, void* dt, int64_t dt_row_stride, int out_dtype, void* workspace,
                 size_t workspace_bytes, void* stream);

/*
 * Opt-in learnable temperature (the reference registers `logit_scale`, trainer/loss.py:52, and never uses it; SURVEY.md
 * section 8 f3).  If the caller ran the problem at the effective temperature p->temperature = tau / logit_scale -- i.e. with every
 * logit multiplied by the scalar s = logit_scale -- this writes d loss / d s of the OWNED rows' share of the global loss to
 * dscale_out[0] (device double; summing the ranks' values gives the gradient), times grad_out[0] (NULL = 1) and grad_scale.
 * Call it between crossclr_bwd_accumulate and the next use of `workspace` (it reads the accumulated rows): O(B D).
 */
CROSSCLR_API int crossclr_bwd_scale_grad(const crossclr_problem_t* p, int path, const void* feat, const float* coef,
                            const float* scal, const double* grad_out, float grad_scale, float logit_scale,
                            const void* workspace, double* dscale_out, void* stream);

/*
 * The two stages of crossclr_bwd, callable on their own (crossclr_bwd == accumulate, then finish, on one stream):
 *   crossclr_bwd_accumulate  the similarity / gradient kernel: workspace[0 .. row_count*dim) (fp32) = for every owned row g
 *                            q_g sigma sum_j P_gj Fhat_j without the positive-pair term -- the O(B^2 D) part, the kernel
 *                            the roofline is quoted on (bench.py times this call alone);
 *   crossclr_bwd_finish      positive-pair term, F.normalize backward, upstream gradient, cast -- O(B D).
 * Replaces: the autograd backward of trainer/loss.py:83-112 (accumulate) and of :79-80, :114 (finish).
 */
CROSSCLR_API int crossclr_bwd_accumulate(const crossclr_problem_t* p, int path, const void* feat, const float* coef,
                            const float* scal, void* workspace, size_t workspace_bytes, void* stream);
CROSSCLR_API int crossclr_bwd_finish(const crossclr_problem_t* p, int path, const void* feat, const float* rnorm_owned,
                        const float* coef, const float* scal, const double* grad_out, float grad_scale, void* dv,
                        int64_t dv_row_stride, void* dt, int64_t dt_row_stride, int out_dtype, const void* workspace,
                        void* stream);

/* The constant log2-domain shift used by fwd/bwd for this problem: max(0, log2e*max(1,|w|)/tau - 96). */
CROSSCLR_API float crossclr_shift(const crossclr_problem_t* p);

/* Name of the backward similarity/gradient kernel crossclr_bwd launches for this problem and path on the CURRENT device
 * (bench.py's roofline.kernel); "" if the path does not apply.  Static storage. */
CROSSCLR_API const char* crossclr_bwd_kernel_name(const crossclr_problem_t* p, int path);

/* Number of kernel launches issued so far by this library in this process (bench.py's gpu_launches). */
CROSSCLR_API int64_t crossclr_launch_count(void);

/*
 * Per-kernel device timing (bench.py's roofline leg).  While enabled, every kernel launch of this library is
 * bracketed by cudaEventRecord on its own stream.  crossclr_timing_read synchronises the recorded events,
 * returns the summed duration / launch count of kernel family `kernel` (CROSSCLR_K_*) since the last read of
 * that family and forgets them.  Disabled by default (no events, no overhead).
 */
#define CROSSCLR_K_PACK     0
#define CROSSCLR_K_FWD      1
#define CROSSCLR_K_FINALIZE 2
#define CROSSCLR_K_BWD      3
#define CROSSCLR_K_GRADFIN  4
#define CROSSCLR_K_COUNT    5
CROSSCLR_API int crossclr_timing_enable(int on);
CROSSCLR_API int crossclr_timing_read(int kernel, double* total_ms, int64_t* launches);

/*
 * MaxMargin_coot (trainer/loss.py:17-41; SURVEY.md section 8 row f1 -- beside the CrossCLR hot path).  No batch x batch
 * array is stored.
 *   loss_out[0] (double) = (1/B^2) sum_{i != j} [max(0, m + s_ij - s_ii) + max(0, m + s_ij - s_jj)],  s = im s^T
 * `im`, `s`: [batch][dim] device arrays of `dtype`, row strides in elements.  batch >= 256 and dim >= 64 run on the tensor
 * cores -- tcgen05 score tiles, a hinge epilogue, and for the backward the 0/1/2 indicator tile as the A operand of the
 * gradient product.  fp16 / bf16 inputs (16-byte aligned base pointers and row strides) are read by TMA straight out of the
 * caller's tensors; fp32 inputs (the reference multiplies them in fp32) are staged in the workspace as fp16 hi + lo pairs after a
 * power-of-two scale per tensor and the score product runs over K = 3 dim (hi.hi + lo.hi + hi.lo): fp32-grade scores.
 * Everything else runs on exact fp32 CUDA-core kernels (dim <= 1600 there).  crossclr_maxmargin_kernel_name reports which
 * (CROSSCLR_MAXMARGIN_PATH=simt|tc overrides).  The forward leaves the diagonal, the hinge counts and (fp32) the staged
 * rows in `workspace` (crossclr_maxmargin_workspace_bytes(batch, dim, dtype) bytes), which the backward reads and
 * extends; `grad_out` is a DEVICE pointer to the upstream scalar gradient (double) or NULL for 1.0.
 * Replaces: trainer/loss.py:29-41 (forward) and its autograd backward.
 */
CROSSCLR_API size_t crossclr_maxmargin_workspace_bytes(int32_t batch, int32_t dim, int dtype);
CROSSCLR_API const char* crossclr_maxmargin_kernel_name(const void* im, const void* s, int dtype, int64_t im_row_stride,
                                           int64_t s_row_stride, int32_t batch, int32_t dim);
CROSSCLR_API int crossclr_maxmargin_fwd(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                           int32_t batch, int32_t dim, float margin, void* workspace, size_t workspace_bytes,
                           double* loss_out, void* stream);
CROSSCLR_API int crossclr_maxmargin_bwd(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                           int32_t batch, int32_t dim, float margin, void* workspace, size_t workspace_bytes,
                           const double* grad_out, void* d_im, int64_t d_im_row_stride, void* d_s, int64_t d_s_row_stride,
                           int out_dtype, void* stream);

/*
 * Retrieval ranks over the score matrix s = im s^T of two [batch][dim] embedding blocks (the
 * reference only pictures retrieval, figures/qual_retriv.png, and scores with the plain mm of `cosine_sim`,
 * trainer/loss.py:7-15 -- pass normalised rows for cosine ranking).  The MaxMargin forward at margin 0 counts exactly this:
 *   rank_im2s[i] = #{j != i : s_ij > s_ii}   (0-based rank of item i's own partner among all s, query im_i)
 *   rank_s2im[j] = #{i != j : s_ij > s_jj}   (query s_j against all im)
 * recall@K = mean(rank < K).  Same kernels, workspace and path rule as crossclr_maxmargin_fwd; nothing batch x batch stored.
 */
CROSSCLR_API int crossclr_retrieval_ranks(const void* im, const void* s, int dtype, int64_t im_row_stride, int64_t s_row_stride,
                             int32_t batch, int32_t dim, void* workspace, size_t workspace_bytes, int32_t* rank_im2s,
                             int32_t* rank_s2im, void* stream);

/*
 * Exchange of row shards through NVLink peer memory on one node (SURVEY.md section 8 rows e / f2): what the row-sharded step
 * uses instead of the two NCCL all-gathers of trainer-side code (the reference itself is single-device; its DDP caller would
 * all-gather features before trainer/loss.py:79).  Every rank allocates buffers of one size with crossclr_peer_alloc
 * (cudaMalloc, zeroed), exports a 64-byte handle, and imports the handles of its peers (CUDA IPC, one process per GPU).
 * crossclr_peer_exchange is ONE kernel: it stores bytes [offset, offset + bytes) of the local buffer bases[rank] to the same
 * offset of every other bases[p], and its last block runs a cross-rank barrier over `flags` (flags[p] = rank p's array of
 * 2 n_ranks 32-bit words, peer-mapped like the buffers; zero before the first exchange).  When it completes in stream order,
 * every rank's slice is in this rank's buffer.  `entry_barrier` != 0 adds a barrier BEFORE the stores: no rank overwrites a
 * peer's buffer until that peer has reached the same exchange in its own stream (i.e. is done reading the previous contents).
 * `bases` / `flags` are HOST arrays of n_ranks DEVICE pointers; `state`: four zero-initialised 32-bit words of local device
 * memory (epochs, ticket).  Collective semantics: every rank issues the same sequence of exchanges.  Capturable in a CUDA
 * graph.
 */
#define CROSSCLR_PEER_HANDLE_BYTES 64
CROSSCLR_API int crossclr_peer_alloc(size_t bytes, void** ptr_out);
CROSSCLR_API int crossclr_peer_free(void* ptr);
CROSSCLR_API int crossclr_peer_export(const void* ptr, void* handle_out);
CROSSCLR_API int crossclr_peer_import(const void* handle, void** ptr_out);
CROSSCLR_API int crossclr_peer_release(void* ptr);
CROSSCLR_API int crossclr_peer_exchange(void* const* bases, uint32_t* const* flags, int32_t n_ranks, int32_t rank, size_t offset,
                           size_t bytes, int entry_barrier, uint32_t* state, void* stream);

/*
 * Hardware self-test of the tcgen05/TMA building blocks (descriptor encodings, TMEM layouts).
 * `variant` selects the block under test (0: K-major x K-major, 1: swizzled thread-written A x MN-major
 * B with b given as [k][n], 2: A from TMEM, 3: un-swizzled thread-written A, 4: one cta_group::2 MMA stream over
 * a cluster of two CTAs, M = 256, 5: un-swizzled MN-major A = the transposed read of a probability tile); host buffers hold fp16 bit patterns: a [128][k] (variant 4: [256][k]),
 * b [n][k] (variant 1: [k][n]), out [128][n] float (variant 4: [256][n]).  Synchronous (allocates, copies,
 * syncs); tests only.
 */
CROSSCLR_API int crossclr_selftest(int variant, const uint16_t* a_host, const uint16_t* b_host, float* out_host,
                      int32_t n, int32_t k);

#ifdef __cplusplus
}
#endif
#endif /* CROSSCLR_B200_H_ */
