// CUDA-core (SIMT) kernels of libcrossclr_b200:
//   * pack / rownorm / finalize / grad_finish -- the O(B*D) and O(B) stages shared by both paths
//   * fwd_simt / bwd_simt -- exact-fp32 similarity kernels for any B, D (the generic path; also the
//     on-device cross-check for the tcgen05 path).
// Math follows trainer/loss.py:79-114 of the reference as restated in SURVEY.md App. A.
#include "common.cuh"
#include "finalize.cuh"

namespace crossclr {

// ================================================================================================
// pack: one modality block -> its segment of the stacked matrix (trainer/loss.py:79-80, F.normalize with eps 1e-12) and
// the reciprocal norms for the backward.  One warp per row.  Two row formats (include/crossclr_b200.h):
//   fp32 rows (SIMT path), pitch dim: x / max(||x||, eps)
//   16-bit rows (TC paths), pitch dim + CROSSCLR_ROW_TAIL, the fp32 residual scale q in the first tail word:
//       16-bit inputs: f = fp16(x * 2^-e) (exact, also for bf16: 8 significant bits below 1), q = 2^e / max(||x||, eps)
//       fp32 inputs:   f = fp16(x / max(||x||, eps)), q = 1                                (the normalised row is q * f)
// Power-of-two split of the reciprocal norm: rn = q * pow2 with ||x * pow2|| in [1, 2) and q in (1/2, 1], so that the
// symmetric probability tile P q_g q_j of the backward never exceeds P.  Rows inside the eps clamp (||x|| < 1e-12, among
// them all-zero rows) are stored normalised instead (x * 1e12, rounded to fp16, q = 1): their logits are ~0 either way.
__device__ __forceinline__ void split_rnorm(float norm, float rn, float& pow2, float& q) {
  if (!(norm >= kEps)) { pow2 = rn; q = 1.0f; return; }
  int e = 0;
  (void)frexpf(norm, &e);                                  // norm = m * 2^e, m in [1/2, 1)
  e = min(127, e);
  pow2 = exp2f((float)(1 - e));
  q = rn * exp2f((float)(e - 1));
}

// Output rows: modality m, row r -> m * rows_pad + r, dimp columns (+ the tail when pitch > dimp).  Rows r >= rows and columns
// d >= dim are zero padding (tensor-core layout); rnorm is compact: m * rows + r.
template <typename Tin, typename Tout, bool kRaw>
__global__ void __launch_bounds__(256) pack_kernel(const Tin* __restrict__ xv, const Tin* __restrict__ xt, int64_t sv,
                                                  int64_t st_, int rows, int rows_pad, int nmod, int dim, int dimp, int split,
                                                  Tout* __restrict__ out, int64_t pitch, float* __restrict__ rnorm,
                                                  float2* __restrict__ stats_zero, unsigned int* __restrict__ ticket_zero) {
  const int fw = dimp * (1 + split);                            // feature elements per row: [hi] or [hi | lo]
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nmod * rows_pad) return;
  if (stats_zero != nullptr && lane == 0) stats_zero[row] = make_float2(0.f, 0.f);   // the forward accumulates into it
  if (ticket_zero != nullptr && row == 0 && lane == 0) *ticket_zero = 0u;
  const int mod = row >= rows_pad ? 1 : 0, r = row - mod * rows_pad;
  Tout* dst = out + (int64_t)row * pitch;
  if (r >= rows) {                                             // padding row: zero features, q = 1
    for (int d = lane; d < fw; d += 32) dst[d] = from_float<Tout>(0.f);
    if (lane == 0 && pitch > fw) *reinterpret_cast<float*>(dst + fw) = 1.0f;
    return;
  }
  const Tin* src = mod == 0 ? xv + (int64_t)r * sv : xt + (int64_t)r * st_;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const float f = to_float<Tin>(src[d]);
    ss = fmaf(f, f, ss);
  }
  ss = warp_sum(ss);
  const float norm = sqrtf(ss);
  const float rn = 1.0f / fmaxf(norm, kEps);
  float mul = rn, q = 1.0f;
  if (kRaw) split_rnorm(norm, rn, mul, q);
  for (int d = lane; d < dimp; d += 32) {
    const float f = d < dim ? to_float<Tin>(src[d]) * mul : 0.f;
    const Tout hi = from_float<Tout>(f);
    dst[d] = hi;
    if (split) dst[dimp + d] = from_float<Tout>(f - to_float<Tout>(hi));    // lo: what the fp16 rounding of f dropped
  }
  if (lane == 0) {
    rnorm[mod * rows + r] = rn;
    if (pitch > fw) *reinterpret_cast<float*>(dst + fw) = q;
  }
}

// bf16 -> fp16 rows, both modalities of one rank in one launch, 16-byte vectors held in registers between the norm and the
// scale pass (dim <= 1024, dim % 8 == 0).  The rescale is exact: a power of two.
__global__ void __launch_bounds__(256) pack2_bf16_raw_kernel(const uint4* __restrict__ xv, const uint4* __restrict__ xt,
                                                            int64_t sv_vec, int64_t st_vec, int rows, int rows_pad, int dim_vec,
                                                            int dimp_vec, uint4* __restrict__ out, int64_t pitch_vec,
                                                            float* __restrict__ rnorm, float2* __restrict__ stats_zero,
                                                            unsigned int* __restrict__ ticket_zero) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= 2 * rows_pad) return;
  if (stats_zero != nullptr && lane == 0) stats_zero[row] = make_float2(0.f, 0.f);
  if (ticket_zero != nullptr && row == 0 && lane == 0) *ticket_zero = 0u;
  const int mod = row >= rows_pad ? 1 : 0, r = row - mod * rows_pad;
  uint4* dst = out + (int64_t)row * pitch_vec;
  if (r >= rows) {                                             // padding row: zero features, q = 1
    for (int d = lane; d < dimp_vec; d += 32) dst[d] = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0) *reinterpret_cast<float*>(dst + dimp_vec) = 1.0f;
    return;
  }
  const uint4* src = mod == 0 ? xv + (int64_t)r * sv_vec : xt + (int64_t)r * st_vec;
  uint4 u[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < dim_vec) {
      u[i] = __ldg(src + d);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[i]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
  }
  ss = warp_sum(ss);
  const float norm = sqrtf(ss);
  const float rn = 1.0f / fmaxf(norm, kEps);
  float mul, q;
  split_rnorm(norm, rn, mul, q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < dim_vec) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[i]);
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        oh[k] = __floats2half2_rn(f.x * mul, f.y * mul);
      }
      dst[d] = o;
    } else if (d < dimp_vec) {
      dst[d] = make_uint4(0u, 0u, 0u, 0u);                   // padding columns
    }
  }
  if (lane == 0) {
    rnorm[mod * rows + r] = rn;
    *reinterpret_cast<float*>(dst + dimp_vec) = q;
  }
}

template <typename Tin>
static int pack_dispatch(const void* xv, const void* xt, int64_t sv, int64_t st_, int rows, int nmod, int dim,
                         void* out, int out_dtype, float* rnorm, cudaStream_t st, float* stats_zero, unsigned int* ticket_zero) {
  const bool tc = out_dtype != CROSSCLR_F32;                       // the tensor-core layout is padded (Geometry)
  const int split = out_dtype == CROSSCLR_F16X2 ? 1 : 0;           // [hi | lo] rows of CROSSCLR_PATH_TC_SPLIT
  const int rows_pad = tc ? tc_pad_rows(rows) : rows, dimp = tc ? tc_pad_dim(dim, split) : dim;
  dim3 block(256), grid((nmod * rows_pad + 7) / 8);
  const int64_t pitch = tc ? dimp * (1 + split) + CROSSCLR_ROW_TAIL : dimp;
#define CC_PACK(Tout, kRaw)                                                                                              \
  pack_kernel<Tin, Tout, kRaw><<<grid, block, 0, st>>>((const Tin*)xv, (const Tin*)xt, sv, st_, rows, rows_pad, nmod, dim, \
                                                       dimp, split, (Tout*)out, pitch, rnorm, (float2*)stats_zero, ticket_zero)
  if (out_dtype == CROSSCLR_F32) CC_PACK(float, false);
  else if (out_dtype == CROSSCLR_F16 || out_dtype == CROSSCLR_F16X2) CC_PACK(__half, true);   // power-of-two rescale: exact for 16-bit values
  else { set_error("crossclr_pack: unsupported stacked dtype %d", out_dtype); return CROSSCLR_EINVAL; }
#undef CC_PACK
  return check_launch("pack_kernel");
}

static int pack_any(const void* xv, const void* xt, int in_dtype, int64_t sv, int64_t st_, int rows, int nmod, int dim,
                    void* out, int out_dtype, float* rnorm, cudaStream_t st, float* stats_zero, unsigned int* ticket_zero) {
  if (rows == 0) return CROSSCLR_OK;
  TimedLaunch timed(CROSSCLR_K_PACK, st);
  if (nmod == 2 && in_dtype == CROSSCLR_BF16 && out_dtype == CROSSCLR_F16 && dim % 8 == 0 && dim <= 1024 && sv % 8 == 0 &&
      st_ % 8 == 0 && ((uintptr_t)xv % 16 == 0) && ((uintptr_t)xt % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
    const int rows_pad = tc_pad_rows(rows), dimp = tc_pad_dim(dim);
    dim3 block(256), grid((2 * rows_pad + 7) / 8);
    pack2_bf16_raw_kernel<<<grid, block, 0, st>>>((const uint4*)xv, (const uint4*)xt, sv / 8, st_ / 8, rows, rows_pad, dim / 8,
                                                  dimp / 8, (uint4*)out, (dimp + CROSSCLR_ROW_TAIL) / 8, rnorm,
                                                  (float2*)stats_zero, ticket_zero);
    return check_launch("pack2_bf16_raw_kernel");
  }
  switch (in_dtype) {
    case CROSSCLR_F32: return pack_dispatch<float>(xv, xt, sv, st_, rows, nmod, dim, out, out_dtype, rnorm, st, stats_zero, ticket_zero);
    case CROSSCLR_F16: return pack_dispatch<__half>(xv, xt, sv, st_, rows, nmod, dim, out, out_dtype, rnorm, st, stats_zero, ticket_zero);
    case CROSSCLR_BF16: return pack_dispatch<__nv_bfloat16>(xv, xt, sv, st_, rows, nmod, dim, out, out_dtype, rnorm, st, stats_zero, ticket_zero);
    default: set_error("crossclr_pack: unsupported input dtype %d", in_dtype); return CROSSCLR_EINVAL;
  }
}

int launch_pack2(const void* xv, const void* xt, int in_dtype, int64_t sv, int64_t st_, int rows, int dim, void* out,
                 int out_dtype, float* rnorm, cudaStream_t st, float* stats_zero, unsigned int* ticket_zero) {
  return pack_any(xv, xt, in_dtype, sv, st_, rows, 2, dim, out, out_dtype, rnorm, st, stats_zero, ticket_zero);
}

int launch_pack(const void* x, int in_dtype, int64_t stride, int rows, int dim, void* out, int out_dtype,
                float* rnorm, cudaStream_t st) {
  return pack_any(x, x, in_dtype, stride, stride, rows, 1, dim, out, out_dtype, rnorm, st, nullptr, nullptr);
}

// ================================================================================================
// fwd_simt: 64x64 similarity tiles on CUDA cores, fused scale / mask / exp2 / row-sum epilogue.
// Replaces trainer/loss.py:83-100 + the row sums of :59-60 for the owned rows.
constexpr int FT = 64;   // tile rows == tile cols
constexpr int FK = 16;   // k chunk

// log2(2^a + 2^b); -inf is the empty sum
__device__ __forceinline__ float logaddexp2(float a, float b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  return lo == -INFINITY ? hi : hi + log2f(1.0f + exp2f(lo - hi));
}
// merge a partial log2-sum into *addr (CAS loop: a handful of partials per row)
__device__ __forceinline__ void atomic_logaddexp2(float* addr, float l) {
  int* ai = reinterpret_cast<int*>(addr);
  int old = *ai, assumed;
  do {
    assumed = old;
    old = atomicCAS(ai, assumed, __float_as_int(logaddexp2(__int_as_float(assumed), l)));
  } while (old != assumed);
}

// kRowShift (Geometry::row_shift): every thread keeps a running (max, sum 2^(x - max)) per row -- an online softmax -- and the
// row's partials merge in the log2 domain, so nothing over- or underflows whatever the temperature: stats[g] = (log2 X_g, xpos_g).
template <bool kRowShift>
__global__ void __launch_bounds__(256) fwd_simt_kernel(Geometry g, const float* __restrict__ F,
                                                      float* __restrict__ stats, int col_tiles_per_block) {
  __shared__ float As[FK][FT + 4];
  __shared__ float Bs[FK][FT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = g.row_begin + blockIdx.x * FT;
  const int row_end = g.row_begin + g.row_count;
  const int n_col_tiles = (g.rows + FT - 1) / FT;
  const int ct0 = blockIdx.y * col_tiles_per_block;
  const int ct1 = min(n_col_tiles, ct0 + col_tiles_per_block);

  int gi[4], mod_i[4], samp_i[4];
  float rsum[4], rmax[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    gi[r] = row0 + ty * 4 + r;
    bool ok = gi[r] < row_end;
    int gg = ok ? gi[r] : g.row_begin;
    mod_i[r] = row_modality(gg, g.bseg);
    samp_i[r] = row_sample(gg, g.bseg);
    rsum[r] = 0.f;
    rmax[r] = -INFINITY;
    if (!ok) gi[r] = -1;
  }
  auto online_add = [&](int r, float x) {        // rsum[r] 2^rmax[r] += 2^x
    if (x > rmax[r]) { rsum[r] = fmaf(rsum[r], exp2f(rmax[r] - x), 1.0f); rmax[r] = x; }
    else rsum[r] += exp2f(x - rmax[r]);
  };

  for (int ct = ct0; ct < ct1; ++ct) {
    const int col0 = ct * FT;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int k0 = 0; k0 < g.dim; k0 += FK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int e = threadIdx.x + 256 * i;
        int rr = e >> 4, kk = e & 15;
        int ga = row0 + rr, gb = col0 + rr, kd = k0 + kk;
        As[kk][rr] = (ga < row_end && kd < g.dim) ? F[(int64_t)ga * g.dim + kd] : 0.f;
        Bs[kk][rr] = (gb < g.rows && kd < g.dim) ? F[(int64_t)gb * g.dim + kd] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < FK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx * 4 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
      }
      __syncthreads();
    }

#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = col0 + tx * 4 + c;
      if (j >= g.rows) continue;
      const int mod_j = row_modality(j, g.bseg);
      const int samp_j = row_sample(j, g.bseg);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (gi[r] < 0) continue;
        const bool same_mod = (mod_i[r] == mod_j);
        const float k = same_mod ? g.k_intra : g.k_inter;
        const float x = kRowShift ? acc[r][c] * k : fmaf(acc[r][c], k, -g.shift);
        if (samp_i[r] == samp_j) {
          if (!same_mod) stats[2 * (int64_t)gi[r] + 1] = x;   // the positive logit a_ii (loss.py:102-109)
          else if (kRowShift) online_add(r, 0.f);          // masked intra-modal diagonal: logit 0 (loss.py:65,96-97)
          else rsum[r] += exp2f(-g.shift);
        } else if (kRowShift) {
          online_add(r, x);
        } else {
          rsum[r] += exp2f(x);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (kRowShift) {
      float l = rsum[r] > 0.f ? rmax[r] + log2f(rsum[r]) : -INFINITY;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) l = logaddexp2(l, __shfl_xor_sync(0xffffffffu, l, o));
      if (tx == 0 && gi[r] >= 0 && l != -INFINITY) atomic_logaddexp2(&stats[2 * (int64_t)gi[r]], l);
      continue;
    }
    float v = rsum[r];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (tx == 0 && gi[r] >= 0) atomicAdd(&stats[2 * (int64_t)gi[r]], v);
  }
}

// Row-shift mode: the log2-domain accumulators start at the empty sum.
__global__ void row_shift_init_kernel(Geometry g, float2* __restrict__ stats) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < g.row_count) stats[g.row_begin + l] = make_float2(-INFINITY, 0.f);
}

int launch_fwd_simt(const Geometry& g, const float* feat, float* stats, cudaStream_t st) {
  TimedLaunch timed(CROSSCLR_K_FWD, st);
  if (g.row_shift) {
    row_shift_init_kernel<<<(g.row_count + 255) / 256, 256, 0, st>>>(g, reinterpret_cast<float2*>(stats));
    int rc = check_launch("row_shift_init_kernel");
    if (rc) return rc;
  }
  const int row_tiles = (g.row_count + FT - 1) / FT;
  const int n_col_tiles = (g.rows + FT - 1) / FT;
  int splits = max(1, min(n_col_tiles, (2 * 148 + row_tiles - 1) / row_tiles));
  int per = (n_col_tiles + splits - 1) / splits;
  splits = (n_col_tiles + per - 1) / per;
  dim3 grid(row_tiles, splits), block(256);
  if (g.row_shift) fwd_simt_kernel<true><<<grid, block, 0, st>>>(g, feat, stats, per);
  else fwd_simt_kernel<false><<<grid, block, 0, st>>>(g, feat, stats, per);
  return check_launch("fwd_simt_kernel");
}

// ================================================================================================
// bwd_simt: dFhat[g][d] = sum_j P_gj * F_j[d]  with
//   P_gj = 2^(x_gj) * (1/Z_g + 1/Z_j) * kappa_gj   (0 for the same-sample pair), F = normalised rows
// i.e. the un-normalised gradient w.r.t. the normalised row, without the positive-pair term (added in
// grad_finish).  Block = 32 rows x one 256-wide slab of D; similarity tiles are recomputed per slab.
constexpr int BR = 32, BJ = 32, BK = 32, BSLAB = 256;

__global__ void __launch_bounds__(256) bwd_simt_kernel(Geometry g, const float* __restrict__ F,
                                                      const float* __restrict__ coef, float* __restrict__ dfhat) {
  __shared__ float As[BK][BR + 1];
  __shared__ float Bs[BK][BJ + 1];
  __shared__ float Ps[BR][BJ + 1];
  const int t = threadIdx.x;
  const int row0 = g.row_begin + blockIdx.x * BR;
  const int row_end = g.row_begin + g.row_count;
  const int d = blockIdx.y * BSLAB + t;
  const int tys = t >> 5, txs = t & 31;

  float acc[BR];
#pragma unroll
  for (int r = 0; r < BR; ++r) acc[r] = 0.f;

  int gi[4], mod_i[4], samp_i[4];
  float iz_i[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    gi[r] = row0 + tys * 4 + r;
    bool ok = gi[r] < row_end;
    int gg = ok ? gi[r] : g.row_begin;
    mod_i[r] = row_modality(gg, g.bseg);
    samp_i[r] = row_sample(gg, g.bseg);
    iz_i[r] = coef[2 * (int64_t)gg];
    if (!ok) gi[r] = -1;
  }

  for (int j0 = 0; j0 < g.rows; j0 += BJ) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < g.dim; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int e = t + 256 * i;
        int rr = e >> 5, kk = e & 31;
        int ga = row0 + rr, gb = j0 + rr, kd = k0 + kk;
        As[kk][rr] = (ga < row_end && kd < g.dim) ? F[(int64_t)ga * g.dim + kd] : 0.f;
        Bs[kk][rr] = (gb < g.rows && kd < g.dim) ? F[(int64_t)gb * g.dim + kd] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float b = Bs[kk][txs];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r] = fmaf(As[kk][tys * 4 + r], b, s[r]);
      }
      __syncthreads();
    }
    {
      const int j = j0 + txs;
      const bool jok = j < g.rows;
      const int jj = jok ? j : 0;
      const int mod_j = row_modality(jj, g.bseg);
      const int samp_j = row_sample(jj, g.bseg);
      const float iz_j = coef[2 * (int64_t)jj];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float p = 0.f;
        if (jok && gi[r] >= 0 && samp_i[r] != samp_j) {
          const bool same_mod = (mod_i[r] == mod_j);
          const float k = same_mod ? g.k_intra : g.k_inter;
          if (g.row_shift) {                          // coef[.,0] = log2 Z: two exponentials, each against its own row's Z
            const float x = s[r] * k;
            p = (exp2f(x - iz_i[r]) + exp2f(x - iz_j)) * (same_mod ? g.w : 1.0f);
          } else {
            const float x = fmaf(s[r], k, -g.shift);
            p = exp2f(x) * (iz_i[r] + iz_j) * (same_mod ? g.w : 1.0f);
          }
        }
        Ps[tys * 4 + r][txs] = p;
      }
    }
    __syncthreads();
    if (d < g.dim) {
      const int jn = min(BJ, g.rows - j0);
      for (int jj = 0; jj < jn; ++jj) {
        const float fj = F[(int64_t)(j0 + jj) * g.dim + d];
#pragma unroll
        for (int r = 0; r < BR; ++r) acc[r] = fmaf(Ps[r][jj], fj, acc[r]);
      }
    }
    __syncthreads();
  }
  if (d < g.dim) {
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      int gr = row0 + r;
      if (gr < row_end) dfhat[(int64_t)(gr - g.row_begin) * g.dim + d] = acc[r];
    }
  }
}

int launch_bwd_simt(const Geometry& g, const float* feat, const float* coef, float* dfhat, cudaStream_t st) {
  TimedLaunch timed(CROSSCLR_K_BWD, st);
  dim3 grid((g.row_count + BR - 1) / BR, (g.dim + BSLAB - 1) / BSLAB), block(256);
  bwd_simt_kernel<<<grid, block, 0, st>>>(g, feat, coef, dfhat);
  return check_launch("bwd_simt_kernel");
}

// ================================================================================================
// finalize: loss (trainer/loss.py:60,:111-114) + backward coefficients from per-row statistics -- finalize.cuh, one block.
__global__ void __launch_bounds__(1024) finalize_kernel(Geometry g, const float* __restrict__ stats,
                                                       float* __restrict__ coef, double* __restrict__ loss,
                                                       float* __restrict__ scal) {
  __shared__ double s_sum[32];
  __shared__ float s_rho[32];
  finalize_block<false>(g, stats, coef, loss, scal, s_sum, s_rho);
  if (threadIdx.x == 0) scal[3] = 0.f;
}

int launch_finalize(const Geometry& g, const float* stats, float* coef, double* loss, float* scal,
                    cudaStream_t st) {
  TimedLaunch timed(CROSSCLR_K_FINALIZE, st);
  const int threads = std::max(32, std::min(1024, ((g.rows + 31) / 32) * 32));
  finalize_kernel<<<1, threads, 0, st>>>(g, stats, coef, loss, scal);
  return check_launch("finalize_kernel");
}

// ================================================================================================
// grad_finish: positive-pair term, F.normalize backward (projection + 1/norm), upstream scale, cast.
//   dFhat_g  = (c/tau) [ acc_g / sigma - (rho_g + rho_partner) * Fhat_partner ]      c = 1/(2B)
//   dF_g     = rnorm_g (dFhat_g - (dFhat_g . Fhat_g) Fhat_g)      (no projection if ||F_g|| < eps)
// `F` holds the normalised rows (fp32 or fp16); `rn` the reciprocal norms of the owned rows.
// One warp per owned row.
// kTail: the rows are (f, q) rows of the TC paths (pitch dim + tail, normalised row = q * f); else plain normalised rows.
template <typename TF, bool kTail>
__device__ __forceinline__ float row_scale(const TF* row, const Geometry& g) {
  return kTail ? *reinterpret_cast<const float*>(row + (g.pitch - CROSSCLR_ROW_TAIL)) : 1.0f;
}
// element d of a stored row: hi (+ lo for the split rows of CROSSCLR_PATH_TC_SPLIT)
template <typename TF>
__device__ __forceinline__ float row_elem(const TF* row, int d, const Geometry& g) {
  return g.split ? to_float<TF>(row[d]) + to_float<TF>(row[g.dim + d]) : to_float<TF>(row[d]);
}

template <typename TF, typename TO, bool kTail>
__global__ void __launch_bounds__(256) grad_finish_kernel(Geometry g, const TF* __restrict__ F,
                                                         const float* __restrict__ rn, const float* __restrict__ coef,
                                                         const float* __restrict__ scal, bool use_sigma,
                                                         const double* __restrict__ grad_out, float grad_scale,
                                                         const float* __restrict__ dfhat, TO* __restrict__ dv,
                                                         int64_t dv_stride, TO* __restrict__ dt, int64_t dt_stride,
                                                         const float* __restrict__ dfhat2, const unsigned int* __restrict__ two) {
  const int l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= g.row_count) return;
  const int gr = g.row_begin + l;
  const int r = gr % g.bseg;
  if (r >= g.bvalid) return;                   // zero-padding row of the tensor-core layout
  const int pg = row_partner(gr, g.bseg);
  const float rn_g = rn[(l / g.bseg) * g.bvalid + r];   // reciprocal norms of the OWNED rows only, compact
  const float* dh2 = (dfhat2 != nullptr && two != nullptr && *two != 0u) ? dfhat2 + (int64_t)l * g.dim : nullptr;
  const TF* fg = F + (int64_t)gr * g.pitch;
  const TF* fp = F + (int64_t)pg * g.pitch;
  const float qg = row_scale<TF, kTail>(fg, g), qp = row_scale<TF, kTail>(fp, g);
  const float acc_scale = (use_sigma ? scal[1] : 1.0f) / qg;       // TC paths accumulate sum_j (P q_g q_j) f_j = q_g dFhat_g
  const float pos_coef = -(coef[2 * (int64_t)gr + 1] + coef[2 * (int64_t)pg + 1]) * qp;
  const float* dh = dfhat + (int64_t)l * g.dim;
  float dot = 0.f;
  for (int d = lane; d < g.dim; d += 32) {
    const float acc = dh2 ? dh[d] + dh2[d] : dh[d];
    const float h = acc * acc_scale + pos_coef * row_elem<TF>(fp, d, g);
    dot = fmaf(h, row_elem<TF>(fg, d, g), dot);
  }
  dot = warp_sum(dot) * qg * qg;        // (h . Fhat_g) Fhat_g with Fhat_g = q_g f_g
  if (rn_g >= 1.0f / kEps) dot = 0.f;   // ||x|| < eps: the clamp in F.normalize is active, no norm gradient
  double m = (double)g.inv_tau * (double)grad_scale / (double)g.rows_valid;
  if (grad_out != nullptr) m *= grad_out[0];
  const float mult = (float)m * rn_g;
  const int mod = row_modality(gr, g.bseg);
  TO* out = mod == 0 ? dv + (int64_t)r * dv_stride : dt + (int64_t)r * dt_stride;
  for (int d = lane; d < g.dvalid; d += 32) {
    const float acc = dh2 ? dh[d] + dh2[d] : dh[d];
    const float h = acc * acc_scale + pos_coef * row_elem<TF>(fp, d, g);
    out[d] = from_float<TO>(mult * (h - dot * row_elem<TF>(fg, d, g)));
  }
}

// Vectorised form for the tensor-core paths (16-bit (f, q) rows, dim = 256 NV <= 1024, 16-byte aligned rows): each lane
// owns NV groups of 8 consecutive columns, read once (16-byte loads) and kept in registers between the dot-product
// pass and the output pass.  HBM-bound: reads dfhat (4 B), the row and its partner (2 x 2 B), writes the gradient.
template <typename TF> struct Half2Of;
template <> struct Half2Of<__half> { using type = __half2; static __device__ __forceinline__ float2 cvt(__half2 h) { return __half22float2(h); } };
template <> struct Half2Of<__nv_bfloat16> { using type = __nv_bfloat162; static __device__ __forceinline__ float2 cvt(__nv_bfloat162 h) { return __bfloat1622float2(h); } };

template <typename TF, typename TO, int NV>
__global__ void __launch_bounds__(256) grad_finish_vec_kernel(Geometry g, const TF* __restrict__ F,
                                                             const float* __restrict__ rn, const float* __restrict__ coef,
                                                             const float* __restrict__ scal, bool use_sigma,
                                                             const double* __restrict__ grad_out, float grad_scale,
                                                             const float* __restrict__ dfhat, TO* __restrict__ dv,
                                                             int64_t dv_stride, TO* __restrict__ dt, int64_t dt_stride,
                                                             const float* __restrict__ dfhat2, const unsigned int* __restrict__ two) {
  using H2 = typename Half2Of<TF>::type;
  const int l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= g.row_count) return;
  const int gr = g.row_begin + l;
  const int r = gr % g.bseg;
  if (r >= g.bvalid) return;                   // zero-padding row
  const int pg = row_partner(gr, g.bseg);
  const float rn_g = rn[(l / g.bseg) * g.bvalid + r];
  const TF* rowg = F + (int64_t)gr * g.pitch;
  const TF* rowp = F + (int64_t)pg * g.pitch;
  const float qg = *reinterpret_cast<const float*>(rowg + g.dim), qp = *reinterpret_cast<const float*>(rowp + g.dim);
  const float acc_scale = (use_sigma ? scal[1] : 1.0f) / qg;       // the kernels accumulate sum_j (P q_g q_j) f_j = q_g dFhat_g
  const float pos_coef = -(coef[2 * (int64_t)gr + 1] + coef[2 * (int64_t)pg + 1]) * qp;
  const uint4* fg = reinterpret_cast<const uint4*>(rowg);
  const uint4* fp = reinterpret_cast<const uint4*>(rowp);
  const float4* dh = reinterpret_cast<const float4*>(dfhat + (int64_t)l * g.dim);
  const float4* dh2 = (dfhat2 != nullptr && two != nullptr && *two != 0u) ? reinterpret_cast<const float4*>(dfhat2 + (int64_t)l * g.dim)
                                                                       : nullptr;   // partial of the late consumers
  float h[NV][8], f[NV][8];
  float dot = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = lane + 32 * v;
    const uint4 ug = __ldg(fg + idx), up = __ldg(fp + idx);
    float4 d0 = __ldg(dh + 2 * idx), d1 = __ldg(dh + 2 * idx + 1);
    if (dh2 != nullptr) {
      const float4 e0 = __ldg(dh2 + 2 * idx), e1 = __ldg(dh2 + 2 * idx + 1);
      d0.x += e0.x; d0.y += e0.y; d0.z += e0.z; d0.w += e0.w;
      d1.x += e1.x; d1.y += e1.y; d1.z += e1.z; d1.w += e1.w;
    }
    const H2* hg = reinterpret_cast<const H2*>(&ug);
    const H2* hp = reinterpret_cast<const H2*>(&up);
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 a = Half2Of<TF>::cvt(hg[q]), b = Half2Of<TF>::cvt(hp[q]);
      f[v][2 * q] = a.x; f[v][2 * q + 1] = a.y;
      h[v][2 * q] = fmaf(pos_coef, b.x, dd[2 * q] * acc_scale);
      h[v][2 * q + 1] = fmaf(pos_coef, b.y, dd[2 * q + 1] * acc_scale);
      dot = fmaf(h[v][2 * q], a.x, dot);
      dot = fmaf(h[v][2 * q + 1], a.y, dot);
    }
  }
  dot = warp_sum(dot) * qg * qg;        // (h . Fhat_g) Fhat_g with Fhat_g = q_g f_g
  if (rn_g >= 1.0f / kEps) dot = 0.f;   // ||x|| < eps: the clamp in F.normalize is active, no norm gradient
  double m = (double)g.inv_tau * (double)grad_scale / (double)g.rows_valid;
  if (grad_out != nullptr) m *= grad_out[0];
  const float mult = (float)m * rn_g;
  const int mod = row_modality(gr, g.bseg);
  TO* out = mod == 0 ? dv + (int64_t)r * dv_stride : dt + (int64_t)r * dt_stride;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = lane + 32 * v;
    alignas(16) TO o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = from_float<TO>(mult * (h[v][q] - dot * f[v][q]));
    if (sizeof(TO) == 2) {
      reinterpret_cast<uint4*>(out)[idx] = *reinterpret_cast<const uint4*>(o);
    } else {
      reinterpret_cast<uint4*>(out)[2 * idx] = reinterpret_cast<const uint4*>(o)[0];
      reinterpret_cast<uint4*>(out)[2 * idx + 1] = reinterpret_cast<const uint4*>(o)[1];
    }
  }
}

template <typename TF, typename TO>
static bool grad_finish_vec(const Geometry& g, const void* feat, const float* rnorm, const float* coef,
                            const float* scal, bool use_sigma, const double* grad_out, float grad_scale,
                            const float* dfhat, void* dv, int64_t dvs, void* dt, int64_t dts, cudaStream_t st,
                            const float* dfhat2, const unsigned int* two) {
  const size_t osz = sizeof(TO);
  if (g.dim % 256 != 0 || g.dim > 1024 || g.dvalid != g.dim || g.split || ((uintptr_t)feat | (uintptr_t)dfhat | (uintptr_t)dv | (uintptr_t)dt) % 16 != 0 ||
      (dvs * osz) % 16 != 0 || (dts * osz) % 16 != 0 || (g.pitch * sizeof(TF)) % 16 != 0)
    return false;
  dim3 block(256), grid((g.row_count + 7) / 8);
#define CC_GFV(NV)                                                                                                      \
  grad_finish_vec_kernel<TF, TO, NV><<<grid, block, 0, st>>>(g, (const TF*)feat, rnorm, coef, scal, use_sigma, grad_out, \
                                                            grad_scale, dfhat, (TO*)dv, dvs, (TO*)dt, dts, dfhat2, two)
  switch (g.dim / 256) {
    case 1: CC_GFV(1); break;
    case 2: CC_GFV(2); break;
    case 3: CC_GFV(3); break;
    default: CC_GFV(4); break;
  }
#undef CC_GFV
  return true;
}

template <typename TF, bool kTail>
static int grad_finish_out(const Geometry& g, const void* feat, const float* rnorm, const float* coef,
                           const float* scal, bool use_sigma, const double* grad_out, float grad_scale,
                           const float* dfhat, void* dv, int64_t dvs, void* dt, int64_t dts, int out_dtype,
                           cudaStream_t st, const float* dfhat2 = nullptr, const unsigned int* two = nullptr) {
  dim3 block(256), grid((g.row_count + 7) / 8);
#define CC_GF(TO)                                                                                                    \
  grad_finish_kernel<TF, TO, kTail><<<grid, block, 0, st>>>(g, (const TF*)feat, rnorm, coef, scal, use_sigma, grad_out, \
                                                           grad_scale, dfhat, (TO*)dv, dvs, (TO*)dt, dts, dfhat2, two)
  switch (out_dtype) {
    case CROSSCLR_F32: CC_GF(float); break;
    case CROSSCLR_F16: CC_GF(__half); break;
    case CROSSCLR_BF16: CC_GF(__nv_bfloat16); break;
    default: set_error("crossclr_bwd: unsupported output dtype %d", out_dtype); return CROSSCLR_EINVAL;
  }
#undef CC_GF
  return check_launch("grad_finish_kernel");
}

template <typename TF>
static int grad_finish_16(const Geometry& g, const void* feat, const float* rnorm, const float* coef, const float* scal,
                          bool use_sigma, const double* grad_out, float grad_scale, const float* dfhat, void* dv,
                          int64_t dv_stride, void* dt, int64_t dt_stride, int out_dtype, cudaStream_t st, const float* dfhat2,
                          const unsigned int* two) {
  bool done = false;
  switch (out_dtype) {
    case CROSSCLR_F32: done = grad_finish_vec<TF, float>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv, dv_stride, dt, dt_stride, st, dfhat2, two); break;
    case CROSSCLR_F16: done = grad_finish_vec<TF, __half>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv, dv_stride, dt, dt_stride, st, dfhat2, two); break;
    case CROSSCLR_BF16: done = grad_finish_vec<TF, __nv_bfloat16>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv, dv_stride, dt, dt_stride, st, dfhat2, two); break;
    default: break;
  }
  if (done) return check_launch("grad_finish_vec_kernel");
  return grad_finish_out<TF, true>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv, dv_stride, dt,
                                   dt_stride, out_dtype, st, dfhat2, two);
}

// ================================================================================================
// scale_grad: d loss / d s for logits multiplied by a scalar s (an opt-in learnable temperature: the kernels ran at the effective
// temperature tau / s).  With A_g = sum_j P_gj Fhat_j (the accumulated row, before the positive-pair term):
//   sum_{g,j} P_gj cos_gj = 2 tau sum_g sum_{j in negatives} p_gj a_gj      (P, cos, kappa symmetric; a = kappa cos / tau)
//   dL/ds = (1 / (s R)) sum_g [ sum_j p_gj a_gj - a_g,pos ] = (1 / (s R tau)) sum_g [ (A_g . Fhat_g) / 2 - rho_g (Fhat_g . Fhat_partner) ]
// (the positive enters with p_pos - 1 = -rho_g; the masked diagonal has logit 0).  One warp per owned row, one double atomic per
// warp; summing the ranks' results gives the gradient of the global loss.
template <typename TF, bool kTail>
__global__ void __launch_bounds__(256) scale_grad_kernel(Geometry g, const TF* __restrict__ F, const float* __restrict__ coef,
                                                        const float* __restrict__ scal, bool use_sigma,
                                                        const double* __restrict__ grad_out, double mult,
                                                        const float* __restrict__ dfhat, const float* __restrict__ dfhat2,
                                                        const unsigned int* __restrict__ two, double* __restrict__ out) {
  const int l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= g.row_count) return;
  const int gr = g.row_begin + l;
  if (gr % g.bseg >= g.bvalid) return;         // zero-padding row
  const int pg = row_partner(gr, g.bseg);
  const TF* fg = F + (int64_t)gr * g.pitch;
  const TF* fp = F + (int64_t)pg * g.pitch;
  const float qg = row_scale<TF, kTail>(fg, g), qp = row_scale<TF, kTail>(fp, g);
  const float* dh = dfhat + (int64_t)l * g.dim;
  const float* dh2 = (dfhat2 != nullptr && two != nullptr && *two != 0u) ? dfhat2 + (int64_t)l * g.dim : nullptr;
  float d1 = 0.f, d2 = 0.f;
  for (int d = lane; d < g.dim; d += 32) {
    const float a = dh2 ? dh[d] + dh2[d] : dh[d];
    const float f = row_elem<TF>(fg, d, g);
    d1 = fmaf(a, f, d1);
    d2 = fmaf(f, row_elem<TF>(fp, d, g), d2);
  }
  d1 = warp_sum(d1);
  d2 = warp_sum(d2);
  if (lane == 0) {
    // the accumulated row is q_g sigma A_g on the tensor-core paths (A_g . Fhat_g = acc . f_g / sigma), A_g on the exact path
    const double aa = (double)d1 * (use_sigma ? (double)scal[1] : 1.0);
    const double pp = (double)d2 * (double)qg * (double)qp;
    double c = (0.5 * aa - (double)coef[2 * (int64_t)gr + 1] * pp) * mult;
    if (grad_out != nullptr) c *= grad_out[0];
    atomicAdd(out, c);
  }
}

int launch_scale_grad(const Geometry& g, const void* feat, int feat_dtype, const float* coef, const float* scal, bool use_sigma,
                      const double* grad_out, double mult, const float* dfhat, const float* dfhat2, double* out, cudaStream_t st) {
  const size_t part_bytes = ((size_t)g.row_count * (size_t)g.dim * sizeof(float) + 255) & ~(size_t)255;
  const unsigned int* two = dfhat2 ? reinterpret_cast<const unsigned int*>(reinterpret_cast<const char*>(dfhat) + 2 * part_bytes) : nullptr;
  cudaMemsetAsync(out, 0, sizeof(double), st);
  dim3 block(256), grid((g.row_count + 7) / 8);
  if (feat_dtype == CROSSCLR_F32)
    scale_grad_kernel<float, false><<<grid, block, 0, st>>>(g, (const float*)feat, coef, scal, use_sigma, grad_out, mult, dfhat, nullptr, nullptr, out);
  else
    scale_grad_kernel<__half, true><<<grid, block, 0, st>>>(g, (const __half*)feat, coef, scal, use_sigma, grad_out, mult, dfhat, dfhat2, two, out);
  return check_launch("scale_grad_kernel");
}

int launch_grad_finish(const Geometry& g, const void* feat, int feat_dtype, const float* rnorm,
                       const float* coef, const float* scal, bool use_sigma, const double* grad_out,
                       float grad_scale, const float* dfhat, void* dv, int64_t dv_stride, void* dt,
                       int64_t dt_stride, int out_dtype, cudaStream_t st, const float* dfhat2) {
  TimedLaunch timed(CROSSCLR_K_GRADFIN, st);
  // the flag word (second partial in use?) sits right behind the two partials, 256-byte aligned as in api.cu
  const size_t part_bytes = ((size_t)g.row_count * (size_t)g.dim * sizeof(float) + 255) & ~(size_t)255;
  const unsigned int* two = dfhat2 ? reinterpret_cast<const unsigned int*>(reinterpret_cast<const char*>(dfhat) + 2 * part_bytes) : nullptr;
  if (feat_dtype == CROSSCLR_F32)
    return grad_finish_out<float, false>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv,
                                         dv_stride, dt, dt_stride, out_dtype, st);
  if (feat_dtype == CROSSCLR_F16 || feat_dtype == CROSSCLR_F16X2)
    return grad_finish_16<__half>(g, feat, rnorm, coef, scal, use_sigma, grad_out, grad_scale, dfhat, dv, dv_stride, dt,
                                  dt_stride, out_dtype, st, dfhat2, two);
  set_error("crossclr_bwd: unsupported stacked dtype %d", feat_dtype);
  return CROSSCLR_EINVAL;
}

}  // namespace crossclr
