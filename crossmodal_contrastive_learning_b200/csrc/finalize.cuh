// Loss and backward coefficients from the per-row statistics (trainer/loss.py:60,:111-114), as a block-level device
// function: run by finalize_kernel (one block; multi-rank, after the statistics all-gather) and by the last CTA of the
// single-rank tensor-core forward.  One block, fixed thread -> row mapping, fixed reduction order: the result does not
// depend on any schedule, and there is no cross-launch state.
#pragma once

#include "common.cuh"

namespace crossclr {

// loss_g = log1p(X 2^-xp) = ln2 * log2(1 + 2^t), t = log2 X - xp; fp32 pieces (1e-7 relative), double sum.  The direct
// form keeps converged rows exact (log1p(y) ~ y); the log-domain form covers y beyond fp32 range.  NaN / Inf in the
// statistics propagate into the loss (the reference returns a NaN loss for NaN features).
__device__ __forceinline__ float row_loss_piece(float X, float xp) {
  const float t = log2f(X) - xp;
  if (t < 100.f) return log1pf(xp > -120.f ? X * exp2f(-xp) : exp2f(t));
  return 0.6931471805599453f * t;                    // log2(1 + 2^t) - t < 2^-100
}

// `stats` must be read through L2 when other CTAs of the same grid wrote it (kThroughL2).  s_sum / s_rho: >= 32 entries.
template <bool kThroughL2>
__device__ __forceinline__ void finalize_block(const Geometry& g, const float* __restrict__ stats, float* __restrict__ coef,
                                               double* __restrict__ loss, float* __restrict__ scal, double* s_sum,
                                               float* s_rho) {
  double lsum = 0.0;
  float rho_max = 0.f;
  for (int i = threadIdx.x; i < g.rows; i += blockDim.x) {
    const float2 sx = kThroughL2 ? __ldcg(reinterpret_cast<const float2*>(stats) + i) : reinterpret_cast<const float2*>(stats)[i];
    if (g.bvalid != g.bseg && (i % g.bseg) >= g.bvalid) {       // zero-padding row: no loss, and a FINITE (zero) coefficient --
      reinterpret_cast<float2*>(coef)[i] = make_float2(0.f, 0.f);   // the backward multiplies it into zero features
      continue;
    }
    const float X = sx.x, xp = sx.y;
    float rho;
    if (g.row_shift) {                                 // stats = (log2 X_g, xpos_g), both absolute
      const float d = X - xp;                          // log2(X_g / e_pos)
      const float e = exp2f(-fabsf(d));
      rho = d > 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);       // X / Z
      if (d != d) rho = d;
      reinterpret_cast<float2*>(coef)[i] = make_float2(fmaxf(X, xp) + log2f(1.0f + e), rho);     // (log2 Z_g, rho_g)
      lsum += (double)(log1pf(e) + (d > 0.f ? 0.6931471805599453f * d : 0.f));                    // log(1 + X / e_pos)
    } else {
      const float Z = X + exp2f(xp);
      rho = X / Z;
      reinterpret_cast<float2*>(coef)[i] = make_float2(1.0f / Z, rho);
      lsum += (double)row_loss_piece(X, xp);
    }
    rho_max = (rho != rho || rho_max != rho_max) ? __int_as_float(0x7fc00000) : fmaxf(rho_max, rho);   // NaN is sticky
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    const float other = __shfl_xor_sync(0xffffffffu, rho_max, o);
    rho_max = (other != other || rho_max != rho_max) ? __int_as_float(0x7fc00000) : fmaxf(rho_max, other);
  }
  if (lane == 0) { s_sum[wid] = lsum; s_rho[wid] = rho_max; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    float rmax = 0.f;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) {
      tot += s_sum[w];
      rmax = (s_rho[w] != s_rho[w] || rmax != rmax) ? __int_as_float(0x7fc00000) : fmaxf(rmax, s_rho[w]);
    }
    loss[0] = tot / (double)g.rows_valid;
    // fp16 probability tiles hold sigma * 2^x (1/Z_g + 1/Z_j) kappa <= sigma * 2 rho_max max(1,|w|)
    const float bound = 2.0f * rmax * fmaxf(1.0f, fabsf(g.w));
    int ex = 0;
    if (bound > 0.f && isfinite(bound)) ex = 14 - (int)ceilf(log2f(bound));
    ex = max(-100, min(100, ex));
    scal[0] = exp2f((float)ex);
    scal[1] = exp2f((float)-ex);
    scal[2] = rmax;
  }
}

}  // namespace crossclr
