// Tensor-core path of libcrossclr_b200 (sm_100a only): TMA -> shared memory -> tcgen05.mma -> TMEM,
// with the CrossCLR epilogues fused so that no B x B intermediate ever leaves the SM.
//
// Operand: the stacked matrix Fhat [R = nseg*bseg][D] of L2-normalised rows in fp16 (written by
// crossclr_pack).  Because the rows are unit vectors the accumulators are cosines and
//     logit*log2e - shift = acc * k - shift,   k = log2e/tau (inter-modal) or w*log2e/tau (intra-modal).
// The full symmetric 2B x 2B logit matrix L = Fhat Fhat^T (blocks [[w cv, a],[a^t, w ct]]) drives both
// kernels:
//   forward : X_g = sum_j 2^(L_gj) over non-positive j           (trainer/loss.py:83-100 + :59-60)
//   backward: dFhat_g = sum_j P_gj Fhat_j, P_gj = 2^(L_gj) (1/Z_g + 1/Z_j) kappa_gj   (autograd of the above)
//
// Kernel anatomy (both kernels, one CTA per SM, 256 threads):
//   warp 0      TMA producer   : row block of A resident in smem (D <= 512) or streamed; B tiles through a ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma (128 x 128 x 16, fp16 in, fp32 accumulate in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       : tcgen05.ld 32x32b (thread = row), exp2 / row sums (fwd) or fp16 P tile (bwd)
// Synchronisation is mbarrier-only: full/empty per ring stage, full/empty per TMEM accumulator buffer.
#include "tc_common.cuh"
#include "finalize.cuh"

namespace crossclr {
using namespace ptx;

namespace {


// ================================================================================================
// Forward.  Tile = 128 rows x 256 columns of the stacked Gram matrix; persistent, one CTA per SM.
//   warp 0  TMA producer (all lanes run the loop, one elected lane issues)
//   warp 1  MMA issuer   (same)
//   warp 2  TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld (software pipelined) -> x = acc*k - shift -> ex2 -> thread-local row sums
// ================================================================================================
// 32 accumulator columns of one row -> exponentials summed into rs.  x = (f_g . f_j) * (k q_g) * q_j - shift with the column
// scales q_j broadcast from the warp's shared-memory copy (16-byte loads, four columns each); the arithmetic runs on packed
// fp32 pairs (FMUL2 / FFMA2 / FADD2).
__device__ __forceinline__ float ex2_of(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fwd_sum_chunk(const uint32_t (&v)[32], float kq, float nshift, const float4* __restrict__ qc,
                                              float (&rs)[4]) {
  const uint64_t kq2 = pack2(kq, kq), ns2 = pack2(nshift, nshift);
  uint64_t rs01 = pack2(rs[0], rs[1]), rs23 = pack2(rs[2], rs[3]);
#pragma unroll
  for (int q = 0; q < 32; q += 4) {
    const float4 c = qc[q >> 2];
    const uint64_t x01 = ffma2(fmul2(pack2(__uint_as_float(v[q + 0]), __uint_as_float(v[q + 1])), kq2), pack2(c.x, c.y), ns2);
    const uint64_t x23 = ffma2(fmul2(pack2(__uint_as_float(v[q + 2]), __uint_as_float(v[q + 3])), kq2), pack2(c.z, c.w), ns2);
    float x0, x1, x2, x3;
    unpack2(x01, x0, x1);
    unpack2(x23, x2, x3);
    rs01 = fadd2(rs01, pack2(ex2_of(x0), ex2_of(x1)));
    rs23 = fadd2(rs23, pack2(ex2_of(x2), ex2_of(x3)));
  }
  unpack2(rs01, rs[0], rs[1]);
  unpack2(rs23, rs[2], rs[3]);
}

// A 32-column chunk that needs per-column care (scalar code; a few chunks per row block):
//  * it holds the same-sample column dq (>= 0): the masked intra-modal diagonal counts as logit 0
//    (trainer/loss.py:65,96-97); the positive is kept out of X and written to stats[.,1];
//  * only its first nv columns are real rows -- the rest is the zero padding at the end of a segment (Geometry) and adds nothing.
__device__ __forceinline__ void fwd_special_chunk(const uint32_t (&v)[32], float kq, float nshift, const float* __restrict__ qc,
                                                  bool same_mod, float diag_term, int dq, int nv, int gi,
                                                  float* __restrict__ stats, float (&rs)[4]) {
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const float x = fmaf(__uint_as_float(v[q]) * kq, qc[q], nshift);
    float e = fast_exp2(x);
    if (q == dq) {
      if (same_mod) e = diag_term;
      else { e = 0.f; stats[2 * (int64_t)gi + 1] = x; }
    }
    if (q >= nv) e = 0.f;
    rs[q & 3] += e;
  }
}

// ================================================================================================
// Forward, cta_group::2.  A cluster of two CTAs (one TPC) owns a 256-row x 256-column tile of the stacked Gram matrix:
// CTA r holds rows [128 r, +128) of the row-block pair (resident for D <= 512) and stages rows [128 r, +128) of the
// 256-row column block, so every B chunk is written to shared memory once per PAIR and each SM's tensor core reads
// 4 KiB (A) + 4 KiB (its half of B) per 128-cycle MMA instead of 4 + 8: the single-CTA kernel above is bound by
// shared-memory bandwidth (TMA fill + operand reads ~160 B/clk against 128), this one is not.
// The leader (rank 0) issues every MMA (M = 256, N = 256) and multicasts the commits; all TMA loads count their
// bytes on the leader's barriers; each CTA's epilogue drains its own 128 accumulator rows from its own TMEM and
// its warps arrive (one lane each) on the leader's tempty barrier.
// ================================================================================================

constexpr int FWD_QSTAGES = 4;        // tiles of column scales in flight (shared-memory ring of the producer warp)
constexpr int FWD2_THREADS = 640;       // warps 0-3: producer / MMA / TMEM alloc / idle; warps 4-19: epilogue

// Tiles (row-block pair ib, column block jb) of one CTA pair.  Symmetric mode (single rank: the owned rows are all rows, so the
// Gram matrix is square and symmetric): only jb >= ib, i.e. the upper triangle of 256 x 256 pair tiles including the
// diagonal -- an off-diagonal tile then feeds the row sums of its rows AND, through its column sums, the row sums of its
// columns' rows.
//  * linear order (matrix fits the L2): pair p owns a contiguous range of the row-major tile list, so a row block stays
//    resident over a sweep of column blocks;
//  * blocked order (stacked matrix larger than the L2: c4 / c5): the tile list runs super-tile by super-tile (8 x 8 tiles)
//    and pair p takes every npairs-th entry, so that at any moment all pairs work inside the same one or two super-tiles and
//    their 16 row / column blocks are read from HBM once instead of once per tile.  Entries outside the matrix (edges, the
//    lower half of diagonal super-tiles) are skipped.
//  * row-interleaved order (matrix larger than the L2, row block resident: c4): pair p sweeps whole rows p, p + npairs, ...
//    (every other round in reverse pair order, which evens out the triangle's row lengths), so that the row block stays
//    resident AND all pairs move through the column blocks roughly in step.
enum { FWD_ORDER_LINEAR = 0, FWD_ORDER_BLOCKED = 1, FWD_ORDER_ROWS = 2 };

template <bool kSymW, int kOrderW>
struct FwdTileWalk {
  int ib, jb;
  bool valid;
  int t, t_end;                 // linear: index in the tile list;  blocked: index in the padded super-tile list;  rows: round
  int ncb, nrbp, npairs, nsc;   // (nrbp, npairs: blocked / rows orders; nsc: blocked; rows: nsc holds the pair index)
  static constexpr int SB = 8;
  __device__ FwdTileWalk(int pair, int npairs_, int tiles_total, int ncb_, int nrbp_) : ncb(ncb_) {
    if (kOrderW == FWD_ORDER_LINEAR) {
      t = (int)((long long)pair * tiles_total / npairs_);
      t_end = (int)((long long)(pair + 1) * tiles_total / npairs_);
      valid = t < t_end;
      if (!kSymW) { ib = t / ncb; jb = t - ib * ncb; }
      else {
        ib = 0;
        int rem = t;
        while (rem >= ncb - ib) { rem -= ncb - ib; ++ib; }
        jb = ib + rem;
      }
    } else if (kOrderW == FWD_ORDER_ROWS) {
      nrbp = nrbp_; npairs = npairs_; nsc = pair;
      t = 0;
      ib = pair;
      jb = kSymW ? ib : 0;
      valid = ib < nrbp;
    } else {
      nrbp = nrbp_; npairs = npairs_;
      nsc = (ncb + SB - 1) / SB;
      const int nsr = (nrbp + SB - 1) / SB;
      t_end = (kSymW ? nsc * (nsc + 1) / 2 : nsr * nsc) * (SB * SB);
      t = pair - npairs;
      ib = jb = 0;
      valid = true;
      advance();
    }
  }
  __device__ __forceinline__ bool advance() {
    if (kOrderW == FWD_ORDER_LINEAR) {
      if (++t >= t_end) return valid = false;
      if (++jb == ncb) { ++ib; jb = kSymW ? ib : 0; }
      return true;
    }
    if (kOrderW == FWD_ORDER_ROWS) {
      if (++jb < ncb) return true;
      ++t;                                           // next round; odd rounds run the pairs in reverse
      ib = t * npairs + ((t & 1) ? npairs - 1 - nsc : nsc);
      jb = kSymW ? ib : 0;
      return valid = ib < nrbp;
    }
    for (;;) {
      t += npairs;
      if (t >= t_end) return valid = false;
      const int st = t >> 6, w = t & 63;
      int sr, sc;
      if (!kSymW) { sr = st / nsc; sc = st - sr * nsc; }
      else {                                         // st = sr nsc - sr (sr - 1) / 2 + (sc - sr), sc >= sr
        const float bq = 2.0f * nsc + 1.0f;
        sr = (int)((bq - sqrtf(bq * bq - 8.0f * st)) * 0.5f);
        while (sr > 0 && sr * nsc - sr * (sr - 1) / 2 > st) --sr;
        while ((sr + 1) * nsc - (sr + 1) * sr / 2 <= st) ++sr;
        sc = sr + (st - (sr * nsc - sr * (sr - 1) / 2));
      }
      ib = sr * SB + (w >> 3);
      jb = sc * SB + (w & 7);
      if (ib < nrbp && jb < ncb && (!kSymW || jb >= ib)) return true;
    }
  }
};

// fire-and-forget fp32 add (REDG; atomicAdd with an unused result still compiles to ATOMG for .f32)
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// as fwd_sum_chunk, but leaves the exponentials in v (bit patterns) for the column sums
__device__ __forceinline__ void fwd_sum_chunk_keep(uint32_t (&v)[32], float kq, float nshift, const float4* __restrict__ qc,
                                                   float (&rs)[4]) {
  const uint64_t kq2 = pack2(kq, kq), ns2 = pack2(nshift, nshift);
  uint64_t rs01 = pack2(rs[0], rs[1]), rs23 = pack2(rs[2], rs[3]);
#pragma unroll
  for (int q = 0; q < 32; q += 4) {
    const float4 c = qc[q >> 2];
    const uint64_t x01 = ffma2(fmul2(pack2(__uint_as_float(v[q + 0]), __uint_as_float(v[q + 1])), kq2), pack2(c.x, c.y), ns2);
    const uint64_t x23 = ffma2(fmul2(pack2(__uint_as_float(v[q + 2]), __uint_as_float(v[q + 3])), kq2), pack2(c.z, c.w), ns2);
    float x0, x1, x2, x3;
    unpack2(x01, x0, x1);
    unpack2(x23, x2, x3);
    const float e0 = ex2_of(x0), e1 = ex2_of(x1), e2 = ex2_of(x2), e3 = ex2_of(x3);
    rs01 = fadd2(rs01, pack2(e0, e1));
    rs23 = fadd2(rs23, pack2(e2, e3));
    v[q + 0] = __float_as_uint(e0); v[q + 1] = __float_as_uint(e1);
    v[q + 2] = __float_as_uint(e2); v[q + 3] = __float_as_uint(e3);
  }
  unpack2(rs01, rs[0], rs[1]);
  unpack2(rs23, rs[2], rs[3]);
}
__device__ __forceinline__ void fwd_special_chunk_keep(uint32_t (&v)[32], float kq, float nshift, const float* __restrict__ qc,
                                                       bool same_mod, float diag_term, int dq, int nv, int gi, int partner,
                                                       float* __restrict__ stats, float (&rs)[4]) {
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const float x = fmaf(__uint_as_float(v[q]) * kq, qc[q], nshift);
    float e = fast_exp2(x);
    if (q == dq) {
      if (same_mod) e = diag_term;
      else { e = 0.f; stats[2 * (int64_t)gi + 1] = x; stats[2 * (int64_t)partner + 1] = x; }   // the mirrored tile is skipped
    }
    if (q >= nv) e = 0.f;
    rs[q & 3] += e;
    v[q] = __float_as_uint(e);
  }
}

// Column sums of a warp's 32 rows x 64 columns (a: columns 0..31, b: 32..63 of the slice; thread = row) by a butterfly
// transpose-reduce: five exchange steps halve the columns a lane is responsible for, 62 shuffles in all; lane l ends up
// with the sums of columns 2 l and 2 l + 1.
__device__ __forceinline__ void warp_column_sums(const uint32_t (&a)[32], const uint32_t (&b)[32], int lane, float& c0, float& c1) {
  float y[32];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float lo = __uint_as_float(a[i]), hi = __uint_as_float(b[i]);
    const float recv = __shfl_xor_sync(0xffffffffu, b4 ? lo : hi, 16);
    y[i] = (b4 ? hi : lo) + recv;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, b3 ? y[i] : y[i + 16], 8);
    y[i] = (b3 ? y[i + 16] : y[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, b2 ? y[i] : y[i + 8], 4);
    y[i] = (b2 ? y[i + 8] : y[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, b1 ? y[i] : y[i + 4], 2);
    y[i] = (b1 ? y[i + 4] : y[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, b0 ? y[i] : y[i + 2], 1);
    y[i] = (b0 ? y[i + 2] : y[i]) + recv;
  }
  c0 = y[0]; c1 = y[1];
}

template <bool kResident, bool kSym, int kOrder>
__global__ void __launch_bounds__(FWD2_THREADS, 1)
fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ feat, Geometry g,
               float* __restrict__ stats, int tiles_total, int ncb, int nrbp, int nk, int num_stages, int exp_flags,
               FwdFinalize fin) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_region = base;
  const uint32_t ring_base = a_region + (kResident ? nk * CHUNK_BYTES : 0);
  const uint32_t stage_bytes = (kResident ? 1 : 2) * CHUNK_BYTES;   // [own A chunk] + own half of the B chunk
  const uint32_t bar_base = ring_base + num_stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_SLOTS + s); };
  const uint32_t a_full = bar_base + 8u * (2 * MAX_SLOTS);
  const uint32_t a_empty = a_full + 8;
  auto tfull_bar = [&](int b) { return a_full + 16u + 8u * b; };
  auto tempty_bar = [&](int b) { return a_full + 32u + 8u * b; };
  const uint32_t tmem_slot = a_full + 48u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // column scales q_j of the tiles in flight: ring of FWD_QSTAGES x [256 columns], filled by the TMA producer warp one tile
  // ahead (a gather from the rows' tails), read by the epilogue warps with broadcast 16-byte loads
  float* q_ring = reinterpret_cast<float*>(smem_raw + (bar_base + kBarBytes - smem_u32(smem_raw)));
  auto qfull_bar = [&](int s) { return a_full + 64u + 8u * s; };
  auto qempty_bar = [&](int s) { return a_full + 96u + 8u * s; };
  const uint32_t idesc_s = make_idesc_f16(256, 256, 0, 0, 0, 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkd = g.dim / KC;            // nk counts the K chunks of the product (3 nkd for split rows, streamed only)

  if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < num_stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 16); }   // 8 warps of a group x 2 CTAs
    for (int s = 0; s < FWD_QSTAGES; ++s) { mbar_init(qfull_bar(s), 1); mbar_init(qempty_bar(s), 8); }   // the 8 warps of the tile's group
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers exist before any remote arrival / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // TMA producer (both CTAs): own A rows, own half of every B chunk; bytes counted on the leader's barriers
    Ring ring(num_stages);
    int cur_ib = -1;
    uint32_t a_cnt = 0;
    const uint32_t a_full_ldr = mapa_cluster(a_full, 0);
    FwdTileWalk<kSym, kOrder> w(pair, npairs, tiles_total, ncb, nrbp);
    // column scales of tile t + 1 are gathered (8 per lane) while tile t's chunks are issued, and published afterwards
    float qn[8];
    if (w.valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) qn[i] = row_q(feat, g, w.jb * FWD_TN + 32 * i + lane);
    }
    uint32_t qt = 0;
    auto publish_q = [&]() {
      const uint32_t qs = qt % FWD_QSTAGES;
      mbar_wait(qempty_bar(qs), ((qt / FWD_QSTAGES) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) q_ring[qs * FWD_TN + 32 * i + lane] = qn[i];
      __syncwarp();
      if (lane == 0) mbar_arrive(qfull_bar(qs));
      ++qt;
    };
    if (w.valid) publish_q();
    while (w.valid) {
      const int ib = w.ib, jb = w.jb;
      const int row0 = g.row_begin + (2 * ib + (int)rank) * TM, col0 = jb * FWD_TN + (int)rank * TM;
      w.advance();                                   // w is now the NEXT tile (if any)
      if (w.valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) qn[i] = row_q(feat, g, w.jb * FWD_TN + 32 * i + lane);
      }
      if (kResident && ib != cur_ib) {
        mbar_wait(a_empty, (a_cnt & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(a_full, 2 * nk * CHUNK_BYTES);
          for (int kc = 0; kc < nk; ++kc) tma_load_2d_2sm(a_region + kc * CHUNK_BYTES, &tmap, a_full_ldr, kc * KC, row0);
        }
        __syncwarp();
        cur_ib = ib; ++a_cnt;
      }
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
        if (elect_one()) {
          uint32_t st = ring_base + ring.stage * stage_bytes;
          const uint32_t full_ldr = mapa_cluster(full_bar(ring.stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(ring.stage), 2 * stage_bytes);
          if (!kResident) { tma_load_2d_2sm(st, &tmap, full_ldr, a_kcol(kc, nkd), row0); st += CHUNK_BYTES; }
          tma_load_2d_2sm(st, &tmap, full_ldr, b_kcol(kc, nkd), col0);
        }
        __syncwarp();
        ring.advance();
      }
      if (w.valid) publish_q();
    }
  } else if (warp == 1 && rank == 0) {
    // MMA issuer (leader only)
    Ring ring(num_stages);
    int cur_ib = -1;
    uint32_t a_cnt = 0, iter = 0;
    FwdTileWalk<kSym, kOrder> w(pair, npairs, tiles_total, ncb, nrbp);
    for (; w.valid; ++iter) {
      const int ib = w.ib;
      const uint32_t buf = iter & 1;
      mbar_wait_cluster(tempty_bar(buf), ((iter >> 1) & 1) ^ 1);
      if (kResident && ib != cur_ib) {
        mbar_wait(a_full, a_cnt & 1);
        cur_ib = ib; ++a_cnt;
      }
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * FWD_TN;
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(full_bar(ring.stage), ring.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = ring_base + ring.stage * stage_bytes;
          const uint64_t ad = kmajor_desc(kResident ? a_region + kc * CHUNK_BYTES : st);
          const uint64_t bd = kmajor_desc(kResident ? st : st + CHUNK_BYTES);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            if (!(exp_flags & 2))
              umma_ss_2sm(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s, (kc == 0 && k == 0) ? 0u : 1u);
          umma_commit_2sm(empty_bar(ring.stage), (uint16_t)3);
        }
        __syncwarp();
        ring.advance();
      }
      w.advance();
      const bool last_of_block = !w.valid || w.ib != ib;
      if (elect_one()) {
        umma_commit_2sm(tfull_bar(buf), (uint16_t)3);
        if (kResident && last_of_block) umma_commit_2sm(a_empty, (uint16_t)3);
      }
      __syncwarp();
    }
  } else if (warp >= EPI_WARP0) {
    // Sixteen epilogue warps in two ping-pong groups: group gsel takes the tiles of parity gsel, i.e. TMEM buffer gsel.  Inside
    // a group, warp (quadrant q, half h) owns rows [32 q, +32) x columns [128 h, +128) of the tile, as two 64-column chunks:
    // read-out of chunk 0 -> math -> read-out of chunk 1 -> buffer handed back -> math.  The TMEM read-out of a tile (128 KiB
    // at 64 B/clk: ~2k cycles) therefore overlaps the other group's arithmetic instead of preceding every warp's own.
    const int quad = warp & 3, half = ((warp - EPI_WARP0) >> 2) & 1, gsel = (warp - EPI_WARP0) >> 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + gsel * FWD_TN + half * TM;
    int cur_ib = -1, gi = 0;
    BlockSeg bi{0, 0};
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    float q_i = 1.f;
    const float k_diag_term = fast_exp2(-g.shift);
    float nshift = -g.shift;                       // -inf for a zero-padding row: all its exponentials are 0
    const uint32_t tempty_ldr = mapa_cluster(tempty_bar(gsel), 0);
    // tile coordinates: row-block pair ib, column block jb, and the segment / offset of this warp's half of the column block
    // (kept incrementally along a row sweep of the linear order: no divisions there)
    FwdTileWalk<kSym, kOrder> w(pair, npairs, tiles_total, ncb, nrbp);
    int ib = w.ib, jb = w.jb;
    int jseg = (jb * FWD_TN + half * TM) / g.bseg, joff = (jb * FWD_TN + half * TM) - jseg * g.bseg;
    uint32_t iter = 0;
    for (; w.valid; ++iter) {
      const bool mine = (iter & 1) == (uint32_t)gsel;
      if (mine && ib != cur_ib) {
        if (cur_ib >= 0) red_add_f32(&stats[2 * (int64_t)gi], (rs[0] + rs[1]) + (rs[2] + rs[3]));
        rs[0] = rs[1] = rs[2] = rs[3] = 0.f;
        cur_ib = ib;
        const int row0 = g.row_begin + (2 * ib + (int)rank) * TM;
        gi = row0 + r;
        bi = block_seg(row0, g.bseg);
        q_i = row_q(feat, g, gi);
        nshift = (gi % g.bseg) < g.bvalid ? -g.shift : -INFINITY;
      }
      const int ncv = g.bvalid - joff;                               // real columns in this half (>= TM: all of them)
      const bool same_mod = ((jseg & 1) == bi.mod);
      const bool diag_tile = ((jseg >> 1) * g.bseg + joff == bi.samp0);
      const float k = (same_mod ? g.k_intra : g.k_inter) * q_i;      // row part of the logit scale
      const bool col_sums = kSym && jb != ib;                        // off-diagonal tile of the symmetric walk
      const int col_row0 = jb * FWD_TN + half * TM;                  // stacked row of the warp's first column
      const uint32_t qs = iter % FWD_QSTAGES;
      const float* const qv = q_ring + qs * FWD_TN + half * TM;      // this half's column scales
      w.advance();                                                   // (ib, jb, jseg, joff) of the next tile; locals above are this tile's
      if (w.valid) {
        if (w.ib != ib || w.jb != jb + 1) { ib = w.ib; jb = w.jb; jseg = (jb * FWD_TN + half * TM) / g.bseg; joff = (jb * FWD_TN + half * TM) - jseg * g.bseg; }
        else { jb = w.jb; joff += FWD_TN; while (joff >= g.bseg) { joff -= g.bseg; ++jseg; } }
      }
      if (!mine) continue;
      mbar_wait(qfull_bar(qs), (iter / FWD_QSTAGES) & 1);
      mbar_wait(tfull_bar(gsel), (iter >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {                                   // 64-column chunks of the half
        uint32_t va[32], vb[32];
        tmem_ld32(lane_base + c * 64, va);
        tmem_ld32(lane_base + c * 64 + 32, vb);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();                                         // this warp's part of the tile is in registers
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty_ldr);   // TMEM hand-back: no memory to order
        }
        if (exp_flags & 1) continue;
        // the same-sample column r of a diagonal half sits in its 32-column chunk r >> 5 = quad
        const float* const qc = qv + c * 64;
        const int nva = ncv - c * 64, nvb = nva - 32;
        const bool dga = diag_tile && quad == 2 * c, dgb = diag_tile && quad == 2 * c + 1;
        if (!col_sums) {
          if (!dga && nva >= 32) fwd_sum_chunk(va, k, nshift, reinterpret_cast<const float4*>(qc), rs);
          else fwd_special_chunk(va, k, nshift, qc, same_mod, k_diag_term, dga ? (r & 31) : -1, nva, gi, stats, rs);
          if (!dgb && nvb >= 32) fwd_sum_chunk(vb, k, nshift, reinterpret_cast<const float4*>(qc + 32), rs);
          else fwd_special_chunk(vb, k, nshift, qc + 32, same_mod, k_diag_term, dgb ? (r & 31) : -1, nvb, gi, stats, rs);
        } else {
          const int partner = row_partner(gi, g.bseg);
          if (!dga && nva >= 32) fwd_sum_chunk_keep(va, k, nshift, reinterpret_cast<const float4*>(qc), rs);
          else fwd_special_chunk_keep(va, k, nshift, qc, same_mod, k_diag_term, dga ? (r & 31) : -1, nva, gi, partner, stats, rs);
          if (!dgb && nvb >= 32) fwd_sum_chunk_keep(vb, k, nshift, reinterpret_cast<const float4*>(qc + 32), rs);
          else fwd_special_chunk_keep(vb, k, nshift, qc + 32, same_mod, k_diag_term, dgb ? (r & 31) : -1, nvb, gi, partner, stats, rs);
          // the mirrored tile (jb, ib) is never computed: its row sums are this tile's column sums
          if (!(exp_flags & 16)) {
            float c0, c1;
            warp_column_sums(va, vb, lane, c0, c1);
            red_add_f32(&stats[2 * (int64_t)(col_row0 + c * 64 + 2 * lane)], c0);
            red_add_f32(&stats[2 * (int64_t)(col_row0 + c * 64 + 2 * lane + 1)], c1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(qempty_bar(qs));                      // this warp is done with the tile's column scales
    }
    if (cur_ib >= 0) atomicAdd(&stats[2 * (int64_t)gi], (rs[0] + rs[1]) + (rs[2] + rs[3]));
  }

  tc_fence_before();
  __syncthreads();
  if (fin.ticket != nullptr) {
    // Fused finalize (single rank): the last CTA of the grid to get here turns the complete row statistics into the loss
    // and the backward coefficients (finalize.cuh).  Scratch: the start of the operand area -- every MMA of the pair has
    // completed (the epilogue warps have consumed the last tile), and no peer writes there.
    double* s_sum = reinterpret_cast<double*>(smem_raw + (a_region - smem_u32(smem_raw)));
    float* s_rho = reinterpret_cast<float*>(s_sum + 32);
    int* s_last = reinterpret_cast<int*>(s_rho + 32);
    if (threadIdx.x == 0) {
      __threadfence();                                              // this CTA's atomics on `stats` before its ticket
      *s_last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      finalize_block<true>(g, stats, fin.coef, fin.loss, fin.scal, s_sum, s_rho);
      if (threadIdx.x == 0) *fin.ticket = 0u;
    }
  }
  cluster_sync_all();                    // neither CTA leaves (or frees TMEM) while the pair's MMAs / arrivals are in flight
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ================================================================================================
// Backward.  Work unit = (128-row block ib, 256-wide slab sb of D, 128-column block j); a CTA owns a contiguous,
// balanced range of units and walks it segment by segment (segment = the units of one (ib, sb) item).  Per j:
//   S(j)  = A_ib * Fhat_j^T            -> TMEM buffer b = tile & 1 (128 fp32 columns)
//   P(j)  = sigma * 2^(k S - shift) * (1/Z_g + 1/Z_j) * kappa   (same-sample column zeroed), fp16, written by the
//           epilogue warps IN PLACE over the first 64 columns of buffer b (tcgen05.st) -- no shared-memory copy
//   dF   += P(j) * Fhat_j[:, slab]     -> TMEM slab accumulator; A operand from TMEM, B read MN-major from the
//           same TMA boxes the S product uses
// MMA issue order: S(j0), S(j0+1), [dF(j), S(j+2)] ...; tcgen05.mma executes in issue order, so S(j+2) may reuse
// buffer b right behind dF(j), and the epilogue of tile j+1 overlaps dF(j) + S(j+2).
// Ring slots are 16 KiB boxes; an operand that needs two boxes (streamed A+B chunk, or a 128-wide dF operand)
// takes two consecutive slots under the first slot's barriers (the second slot's barriers are cycled in step).
// A segment that covers its whole item stores the slab; partial segments add into the zeroed dfhat (red.add).
// ================================================================================================
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct BwdSeg {
  int ib, sb, j0, j1;
  bool last_of_ib;      // no later segment of this CTA uses the same row block
};
struct BwdWalk {
  int u, u_end, ncb, n_slabs;
  __device__ BwdWalk(int u0, int u1, int ncb_, int ns_) : u(u0), u_end(u1), ncb(ncb_), n_slabs(ns_) {}
  __device__ __forceinline__ bool next(BwdSeg& s) {
    if (u >= u_end) return false;
    const int item = u / ncb;
    s.j0 = u - item * ncb;
    s.j1 = min(ncb, s.j0 + (u_end - u));
    s.ib = item / n_slabs;
    s.sb = item - s.ib * n_slabs;
    u += s.j1 - s.j0;
    s.last_of_ib = (u >= u_end) || ((u / ncb) / n_slabs != s.ib);
    return true;
  }
};

template <bool kResident>
__global__ void __launch_bounds__(NUM_THREADS, 1)
bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ feat, Geometry g,
              const float* __restrict__ coef, const float* __restrict__ scal, float* __restrict__ dfhat, int n_units,
              int n_slabs, int ncb, int nk, int num_slots) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_region = base;
  const uint32_t ring_base = a_region + (kResident ? nk * CHUNK_BYTES : 0);
  const uint32_t cvec_base = ring_base + num_slots * CHUNK_BYTES;                 // float2 [2][128]: (q_j, q_j kappa sigma / Z_j)
  const uint32_t bar_base = cvec_base + 2 * BWD_TN * 8;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_SLOTS + s); };
  const uint32_t a_full = bar_base + 8u * (2 * MAX_SLOTS);
  const uint32_t a_empty = a_full + 8;
  auto sfull_bar = [&](int b) { return a_full + 16u + 8u * b; };
  auto pfull_bar = [&](int b) { return a_full + 32u + 8u * b; };
  const uint32_t acc_full = a_full + 48u, acc_empty = a_full + 56u;
  const uint32_t tmem_slot = a_full + 64u;
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw_u32));
  float2* cvec = reinterpret_cast<float2*>(smem_raw + (cvec_base - raw_u32));
  const uint32_t idesc_s128 = make_idesc_f16(128, 128, 0, 0, 0, 0);   // S = A(K-major) * B(K-major)^T
  const uint32_t idesc_g128 = make_idesc_f16(128, 128, 0, 0, 0, 1);             // dF += P(f16) * F_J(MN-major)
  const uint32_t idesc_g64 = make_idesc_f16(128, 64, 0, 0, 0, 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u_begin = (int)((long long)blockIdx.x * n_units / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * n_units / gridDim.x);

  if (warp == 0 && lane == 0) prefetch_tmap(&tmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < num_slots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(pfull_bar(b), EPI_THREADS); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem_base + 2 * BWD_TN;     // slab accumulator: columns [256, 512)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    Ring ring(num_slots);
    int cur_ib = -1;
    uint32_t a_cnt = 0;
    BwdWalk walk(u_begin, u_end, ncb, n_slabs);
    BwdSeg sg;
    while (walk.next(sg)) {
      const int row0 = g.row_begin + sg.ib * TM;
      const int d0 = sg.sb * SLAB;
      const int nsc = min(SLAB, g.dim - d0) / KC;
      // a 128-wide dF operand spans two consecutive slots; slots stay pair-aligned only if every ring user moves in
      // pairs (streamed A+B chunks do; resident single-slot S chunks do when nk is even)
      const bool dpair = (nsc & 1) == 0 && (!kResident || (nk & 1) == 0);
      const bool dadv2 = dpair || !kResident;
      if (kResident && sg.ib != cur_ib) {
        mbar_wait(a_empty, (a_cnt & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full, nk * CHUNK_BYTES);
          for (int kc = 0; kc < nk; ++kc) tma_load_2d(a_region + kc * CHUNK_BYTES, &tmap, a_full, kc * KC, row0);
        }
        __syncwarp();
        cur_ib = sg.ib; ++a_cnt;
      }
      auto load_S = [&](int j) {           // operands of S(j)
        const int col0 = j * BWD_TN;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
          if (!kResident) mbar_wait(empty_bar(ring.stage + 1), ring.phase ^ 1);
          if (elect_one()) {
            const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
            if (kResident) {
              mbar_arrive_expect_tx(full_bar(ring.stage), CHUNK_BYTES);
              tma_load_2d(st, &tmap, full_bar(ring.stage), kc * KC, col0);
            } else {
              mbar_arrive_expect_tx(full_bar(ring.stage), 2 * CHUNK_BYTES);
              tma_load_2d(st, &tmap, full_bar(ring.stage), kc * KC, row0);
              tma_load_2d(st + CHUNK_BYTES, &tmap, full_bar(ring.stage), kc * KC, col0);
              mbar_arrive(full_bar(ring.stage + 1));
            }
          }
          __syncwarp();
          ring.advance();
          if (!kResident) ring.advance();
        }
      };
      auto load_dF = [&](int j) {          // B operand of dF(j): Fhat_j[:, slab] as [128 j][64 d] boxes
        const int col0 = j * BWD_TN;
        for (int c = 0; c < nsc; c += (dpair ? 2 : 1)) {
          mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
          if (dadv2) mbar_wait(empty_bar(ring.stage + 1), ring.phase ^ 1);
          if (elect_one()) {
            const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
            mbar_arrive_expect_tx(full_bar(ring.stage), dpair ? 2 * CHUNK_BYTES : CHUNK_BYTES);
            tma_load_2d(st, &tmap, full_bar(ring.stage), d0 + c * KC, col0);
            if (dpair) tma_load_2d(st + CHUNK_BYTES, &tmap, full_bar(ring.stage), d0 + (c + 1) * KC, col0);
            if (dadv2) mbar_arrive(full_bar(ring.stage + 1));   // keep the skipped slot's barriers in phase
          }
          __syncwarp();
          ring.advance();
          if (dadv2) ring.advance();
        }
      };
      load_S(sg.j0);
      if (sg.j0 + 1 < sg.j1) load_S(sg.j0 + 1);
      for (int j = sg.j0; j < sg.j1; ++j) {
        load_dF(j);
        if (j + 2 < sg.j1) load_S(j + 2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    Ring ring(num_slots);
    int cur_ib = -1;
    uint32_t a_cnt = 0, seg_iter = 0, s_issued = 0, p_cnt = 0;
    auto issue_S = [&]() {
      const uint32_t buf = s_issued & 1;
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(full_bar(ring.stage), ring.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
          const uint32_t a_addr = kResident ? a_region + kc * CHUNK_BYTES : st;
          const uint32_t b_addr = kResident ? st : st + CHUNK_BYTES;
          issue_s_chunk(tmem_base + buf * BWD_TN, a_addr, b_addr, idesc_s128, kc == 0);
          umma_commit(empty_bar(ring.stage));
          if (!kResident) umma_commit(empty_bar(ring.stage + 1));
        }
        __syncwarp();
        ring.advance();
        if (!kResident) ring.advance();
      }
      if (elect_one()) umma_commit(sfull_bar(buf));
      __syncwarp();
      ++s_issued;
    };
    BwdWalk walk(u_begin, u_end, ncb, n_slabs);
    BwdSeg sg;
    while (walk.next(sg)) {
      const int d0 = sg.sb * SLAB;
      const int nsc = min(SLAB, g.dim - d0) / KC;
      const bool dpair = (nsc & 1) == 0 && (!kResident || (nk & 1) == 0);
      const bool dadv2 = dpair || !kResident;
      if (kResident && sg.ib != cur_ib) {
        mbar_wait(a_full, a_cnt & 1);
        tc_fence_after();
        cur_ib = sg.ib; ++a_cnt;
      }
      issue_S();
      if (sg.j0 + 1 < sg.j1) issue_S();
      mbar_wait(acc_empty, (seg_iter & 1) ^ 1);       // the epilogue has drained the previous segment's slab
      tc_fence_after();
      for (int j = sg.j0; j < sg.j1; ++j, ++p_cnt) {
        const uint32_t buf = p_cnt & 1;
        mbar_wait(pfull_bar(buf), (p_cnt >> 1) & 1);
        tc_fence_after();
        const uint32_t p_tmem = tmem_base + buf * BWD_TN;
        for (int c = 0; c < nsc; c += (dpair ? 2 : 1)) {
          mbar_wait(full_bar(ring.stage), ring.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
#pragma unroll
            for (int k16 = 0; k16 < BWD_TN / 16; ++k16) {
              // A = P[:, 16 k16 .. +16) from TMEM: 8 columns of packed fp16 pairs per K = 16 step
              // B = Fhat_j[16 k16 .. +16, 64 or 128 d]: MN-major view of the TMA boxes; 16 K rows = 2048 bytes,
              // the second 64-wide MN atom is the next slot (LBO = 16 KiB)
              const uint64_t bd = make_smem_desc_sw128(st + k16 * 2048, 1024, CHUNK_BYTES);
              umma_ts(tmem_acc + c * KC, p_tmem + k16 * 8, bd, dpair ? idesc_g128 : idesc_g64,
                      (j > sg.j0 || k16 > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(ring.stage));
            if (dadv2) umma_commit(empty_bar(ring.stage + 1));
          }
          __syncwarp();
          ring.advance();
          if (dadv2) ring.advance();
        }
        if (j + 2 < sg.j1) issue_S();                  // reuses buffer `buf` right behind dF(j)
      }
      if (elect_one()) {
        umma_commit(acc_full);
        if (kResident && sg.last_of_ib) umma_commit(a_empty);
      }
      __syncwarp();
      ++seg_iter;
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - EPI_WARP0;
    const int r = ew * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
    const float sigma = scal[0];
    const float nshift = -g.shift;
    uint32_t seg_iter = 0, p_cnt = 0;
    BwdWalk walk(u_begin, u_end, ncb, n_slabs);
    BwdSeg sg;
    while (walk.next(sg)) {
      const int row0 = g.row_begin + sg.ib * TM;
      const int d0 = sg.sb * SLAB;
      const int slab_w = min(SLAB, g.dim - d0);
      const int gi = row0 + r;
      const BlockSeg bi = block_seg(row0, g.bseg);
      const float iz_i = coef[2 * (int64_t)gi];
      const float q_i = row_q(feat, g, gi);
      const float nshift_i = nshift + log2f(q_i);      // the tile is P q_g q_j (symmetric): q_g rides in the exponent
      for (int j = sg.j0; j < sg.j1; ++j, ++p_cnt) {
        const int col0 = j * BWD_TN;
        const BlockSeg bj = block_seg(col0, g.bseg);
        const bool same_mod = (bj.mod == bi.mod);
        const bool diag_tile = (bj.samp0 == bi.samp0);
        const float k = (same_mod ? g.k_intra : g.k_inter) * q_i;
        const float ks = (same_mod ? g.w : 1.0f) * sigma;
        const uint32_t buf = p_cnt & 1;
        const uint32_t tbuf = lane_base + buf * BWD_TN;
        float* cv = reinterpret_cast<float*>(cvec + buf * BWD_TN);   // per column pair: (q_j, q_j+1, w_j, w_j+1)
        {
          const float qj = row_q(feat, g, col0 + r);
          cv[(r >> 1) * 4 + (r & 1)] = qj;                                          // q_j
          cv[(r >> 1) * 4 + (r & 1) + 2] = coef[2 * (int64_t)(col0 + r)] * ks * qj;  // w_j = q_j kappa sigma / Z_j
        }
        const float a_i = iz_i * ks;
        named_bar_sync(1, EPI_THREADS);
        mbar_wait(sfull_bar(buf), (p_cnt >> 1) & 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        tmem_ld32(tbuf, va);
#pragma unroll
        for (int c = 0; c < BWD_TN / 32; ++c) {
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c + 1 < BWD_TN / 32) tmem_ld32(tbuf + (c + 1) * 32, (c & 1) ? va : vb);
          uint32_t packed[16];
          const float4* cv4 = reinterpret_cast<const float4*>(cv) + c * 16;
          const uint64_t k2 = pack2(k, k), ns2 = pack2(nshift_i, nshift_i), a2 = pack2(a_i, a_i);
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float4 cc = cv4[q >> 1];                       // (q_j, q_j+1, w_j, w_j+1) of columns q, q + 1
            // P~ = 2^x (1/Z_g + 1/Z_j) kappa sigma q_g q_j,  x = (f_g . f_j) (k q_g) q_j - shift; packed fp32 pairs
            const uint64_t q01 = pack2(cc.x, cc.y);
            const uint64_t x01 = ffma2(fmul2(pack2(__uint_as_float(v[q + 0]), __uint_as_float(v[q + 1])), k2), q01, ns2);
            float x0, x1, e0, e1;
            unpack2(x01, x0, x1);
            unpack2(fmul2(pack2(ex2_of(x0), ex2_of(x1)), ffma2(a2, q01, pack2(cc.z, cc.w))), e0, e1);
            if (diag_tile) {                                     // same-sample pair handled in grad_finish
              if (c * 32 + q + 0 == r) e0 = 0.f;
              if (c * 32 + q + 1 == r) e1 = 0.f;
            }
            __half2 h0 = __floats2half2_rn(e0, e1);
            packed[q >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          }
          // P(j) columns [32c, 32c+32) -> packed fp16 pairs in TMEM columns [16c, 16c+16) of the same buffer
          // (S columns < 32(c+1) are already in registers, so nothing unread is overwritten)
          tmem_st16(tbuf + c * 16, packed);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(pfull_bar(buf));
      }
      // slab accumulator -> dfhat (fp32, still scaled by sigma; grad_finish divides it out)
      mbar_wait(acc_full, seg_iter & 1);
      tc_fence_after();
      float* out = dfhat + (int64_t)(gi - g.row_begin) * g.dim + d0;
      const bool whole = (sg.j0 == 0 && sg.j1 == ncb);
#pragma unroll 1
      for (int c = 0; c < slab_w / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_base + 2 * BWD_TN + c * 32, v);
        tmem_ld_wait();
        if (whole) {
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<float4*>(out + c * 32 + q) =
                make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                            __uint_as_float(v[q + 3]));
        } else {
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            red_add_f32x4(out + c * 32 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]),
                          __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
      ++seg_iter;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// Backward for D <= 512: a CLUSTER OF TWO CTAs per work range, specialised by role, so that S is formed once per
// (row block, column block) instead of once per 256-wide slab of D.
//   rank 0, "S-CTA": S = A_ib Fhat_J^T for a 256-column block J (N = 256 MMAs, two TMEM buffers) -> two epilogue
//                    warpgroups ping-pong over the tiles and turn each into two 128-column fp16 probability tiles
//                    P(j), written row-major into a small per-pair ring of scratch tiles in global memory (8 x 32 KiB,
//                    L2-resident; distributed-shared-memory stores top out near 21 B/clk and stalled the exp warps);
//                    a signaller warp publishes each tile (gpu-scope fence) with an arrival on the peer's mbarrier
//   rank 1, "G-CTA": dF[128 x D] += P(j) Fhat_j   -- accumulator = up to all 512 TMEM columns; A = P(j) from its
//                    shared memory, B = [64 j][64 d] TMA boxes of Fhat_j read MN-major, four boxes (N = 256) per
//                    MMA group, four groups in flight; a loader warp TMA-loads each published P tile into one of 3
//                    shared-memory buffers; tcgen05.commit hands the scratch slot back with a multicast arrival on
//                    the S-CTA's mbarrier
// Both CTAs walk the same balanced range of (row block, 256-column block) units; per unit each spends the same
// 128*256*D MACs on its tensor core.  Shared-memory map (identical barrier block in both CTAs):
//   [0, 1 KiB) mbarriers + TMEM slot | [1, 3 KiB) column coefficient vectors (S-CTA) |
//   S-CTA: A row block (nk x 16 KiB) + ring of 32 KiB stages      G-CTA: 3 P tiles (32 KiB) + 4 groups of 4 x 8 KiB
// ================================================================================================
constexpr int PAIR_THREADS = 384;        // warps 0-3: producer / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int PAIR_EPI_THREADS = 256;
constexpr int PAIR_TN = 256;             // S tile columns in the S-CTA
constexpr int PAIR_PBUF = 3;             // 128-column P tiles in flight in the G-CTA's shared memory
constexpr int PAIR_NSLOT = 8;            // scratch P tiles per pair in global memory (ring)
constexpr int PAIR_GGROUPS = 4;          // G-CTA ring: groups of four [64 j][64 d] boxes
constexpr int PAIR_HDR = 5120;           // barriers (1 KiB) | column coefficient pairs [2][256] float2 (4 KiB)

struct PairSeg { int ib, j0, j1; bool last_of_ib; };
struct PairWalk {
  int u, u_end, ncb;
  __device__ PairWalk(int u0, int u1, int ncb_) : u(u0), u_end(u1), ncb(ncb_) {}
  __device__ __forceinline__ bool next(PairSeg& s) {
    if (u >= u_end) return false;
    s.ib = u / ncb;
    s.j0 = u - s.ib * ncb;
    s.j1 = min(ncb, s.j0 + (u_end - u));
    u += s.j1 - s.j0;
    s.last_of_ib = (u >= u_end) || (u / ncb != s.ib);
    return true;
  }
};

template <bool kResident>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
bwd_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap64,
                const uint8_t* __restrict__ feat, Geometry g, const float* __restrict__ coef,
                const float* __restrict__ scal, float* __restrict__ dfhat, uint8_t* __restrict__ scratch, int n_units,
                int ncb, int nk, int s_stages, int exp_flags, unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  // debug timeline (CROSSCLR_PAIR_TRACE): pair 0 stamps clock64 per role / tile / event
  auto TR = [&](int role, int tile, int ev) {
    if (trace != nullptr && blockIdx.x < 2 && tile < 64) trace[(role * 64 + tile) * 4 + ev] = clock64();
  };
  auto full_bar = [&](int s) { return base + 8u * s; };
  auto empty_bar = [&](int s) { return base + 96u + 8u * s; };
  const uint32_t a_full = base + 192u, a_empty = base + 200u;
  auto sfull_bar = [&](int b) { return base + 208u + 8u * b; };
  auto sempty_bar = [&](int b) { return base + 224u + 8u * b; };
  auto staged_bar = [&](int b) { return base + 240u + 8u * b; };     // S-CTA: a P tile's 128 rows are in global memory
  auto pready_bar = [&](int b) { return base + 304u + 8u * b; };     // G-CTA: that tile is published (remote arrival)
  auto pempty_bar = [&](int b) { return base + 368u + 8u * b; };     // S-CTA: scratch slot consumed (multicast commit)
  auto pbfull_bar = [&](int b) { return base + 432u + 8u * b; };     // G-CTA: P tile landed in a smem buffer (TMA)
  auto pbempty_bar = [&](int b) { return base + 456u + 8u * b; };    // G-CTA: smem buffer consumed
  const uint32_t acc_full = base + 480u, acc_empty = base + 488u;
  const uint32_t tmem_slot = base + 496u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + 496);
  float2* cvec = reinterpret_cast<float2*>(smem_raw + 1024);        // [2][256]: (q_j, q_j kappa sigma / Z_j)
  const uint32_t data = base + (cluster_ctarank() == 0 ? PAIR_HDR : 1024);   // G-CTAs need all of the rest for P tiles + boxes
  const uint32_t idesc_s256 = make_idesc_f16(128, 256, 0, 0, 0, 0);

  const uint32_t rank = cluster_ctarank();
  const uint32_t csize = cluster_nctarank();          // 1 S-CTA + (csize - 1) G-CTAs, one per 512-wide slab of D
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x / csize, npairs = gridDim.x / csize;
  const int u_begin = (int)((long long)pair * n_units / npairs);
  const int u_end = (int)((long long)(pair + 1) * n_units / npairs);

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap); prefetch_tmap(&tmap64); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sempty_bar(b), EPI_THREADS); }
    for (int b = 0; b < PAIR_NSLOT; ++b) {
      mbar_init(staged_bar(b), EPI_THREADS); mbar_init(pready_bar(b), 1); mbar_init(pempty_bar(b), csize - 1);
    }
    for (int b = 0; b < PAIR_PBUF; ++b) { mbar_init(pbfull_bar(b), 1); mbar_init(pbempty_bar(b), 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, PAIR_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (rank == 0 && (exp_flags & 8)) {
    // perf experiment: S-CTA idle
  } else if (rank >= 1 && (exp_flags & 4)) {
    // perf experiment: G-CTA idle
  } else if (rank == 0) {
    // =========================================================================== S-CTA
    const uint32_t a_region = data;
    const uint32_t ring_base = data + (kResident ? nk * CHUNK_BYTES : 0);
    const uint32_t stage_bytes = (kResident ? 2 : 3) * CHUNK_BYTES;     // [A chunk, streamed] + 256-row B chunk
    if (warp == 0) {
      Ring ring(s_stages);
      int cur_ib = -1;
      uint32_t a_cnt = 0;
      PairWalk walk(u_begin, u_end, ncb);
      PairSeg sg;
      while (walk.next(sg)) {
        const int row0 = g.row_begin + sg.ib * TM;
        if (kResident && sg.ib != cur_ib) {
          mbar_wait(a_empty, (a_cnt & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(a_full, nk * CHUNK_BYTES);
            for (int kc = 0; kc < nk; ++kc) tma_load_2d(a_region + kc * CHUNK_BYTES, &tmap, a_full, kc * KC, row0);
          }
          __syncwarp();
          cur_ib = sg.ib; ++a_cnt;
        }
        for (int j = sg.j0; j < sg.j1; ++j) {
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
            if (elect_one()) {
              uint32_t st = ring_base + ring.stage * stage_bytes;
              mbar_arrive_expect_tx(full_bar(ring.stage), stage_bytes);
              if (!kResident) { tma_load_2d(st, &tmap, full_bar(ring.stage), kc * KC, row0); st += CHUNK_BYTES; }
              tma_load_2d(st, &tmap, full_bar(ring.stage), kc * KC, j * PAIR_TN);
              tma_load_2d(st + CHUNK_BYTES, &tmap, full_bar(ring.stage), kc * KC, j * PAIR_TN + TM);
            }
            __syncwarp();
            ring.advance();
          }
        }
      }
    } else if (warp == 1) {
      Ring ring(s_stages);
      int cur_ib = -1;
      uint32_t a_cnt = 0, t = 0;
      PairWalk walk(u_begin, u_end, ncb);
      PairSeg sg;
      while (walk.next(sg)) {
        if (kResident && sg.ib != cur_ib) {
          mbar_wait(a_full, a_cnt & 1);
          cur_ib = sg.ib; ++a_cnt;
        }
        for (int j = sg.j0; j < sg.j1; ++j, ++t) {
          const uint32_t buf = t & 1;
          if (lane == 0) TR(0, t, 0);
          mbar_wait(sempty_bar(buf), ((t >> 1) & 1) ^ 1);
          if (lane == 0) TR(0, t, 1);
          tc_fence_after();
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(full_bar(ring.stage), ring.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t st = ring_base + ring.stage * stage_bytes;
              issue_s_chunk(tmem_base + buf * PAIR_TN, kResident ? a_region + kc * CHUNK_BYTES : st,
                            kResident ? st : st + CHUNK_BYTES, idesc_s256, kc == 0);
              umma_commit(empty_bar(ring.stage));
            }
            __syncwarp();
            ring.advance();
          }
          if (elect_one()) {
            umma_commit(sfull_bar(buf));
            if (kResident && sg.last_of_ib && j + 1 == sg.j1) umma_commit(a_empty);
          }
          if (lane == 0) TR(0, t, 2);
          __syncwarp();
        }
      }
    } else if (warp == 3) {
      // signaller: once the 128 rows of P tile th are in global memory (staged), make them visible GPU-wide and tell
      // the G-CTA's loader warp
      const uint32_t n_ptiles = 2u * (uint32_t)(u_end - u_begin);
      for (uint32_t th = 0; th < n_ptiles; ++th) {
        const uint32_t slot = th % PAIR_NSLOT, use = th / PAIR_NSLOT;
        mbar_wait(staged_bar(slot), use & 1);
        if (elect_one()) {
          fence_acq_rel_gpu();
          for (uint32_t gr = 1; gr < csize; ++gr) mbar_arrive_cluster_relaxed(mapa_cluster(pready_bar(slot), gr));
        }
        __syncwarp();
      }
    } else if (warp >= EPI_WARP0) {
      // Two epilogue warpgroups ping-pong over the S tiles (wg handles tiles t with (t & 1) == wg, i.e. TMEM buffer
      // wg), so one group's load / barrier latencies overlap the other group's exp work.  Tile t is unit
      // u_begin + t of this pair; it yields the 128-column P tiles 2t and 2t + 1.
      const int quad = warp & 3, wg = (warp - EPI_WARP0) >> 2;
      const int r = quad * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * PAIR_TN;
      const float sigma = scal[0];
      const float nshift = -g.shift;
      // scratch P tile layout = the no-swizzle K-major operand layout: [16 column chunks of 8][128 rows][16 bytes], so
      // one warp store instruction (32 rows x 16 B) writes 512 contiguous bytes
      uint8_t* const p_scratch = scratch + (size_t)pair * PAIR_NSLOT * PTILE_BYTES + (size_t)r * 16;
      const int n_tiles = u_end - u_begin;
      int cur_ib = -1, gi = 0;
      BlockSeg bi{0, 0};
      float iz_i = 0.f, q_i = 1.f, nshift_i = nshift;
      float* cv = reinterpret_cast<float*>(cvec + wg * PAIR_TN);     // per column pair: (q_j, q_j+1, w_j, w_j+1)
      // column coefficients of this thread's two columns for the group's next tile, fetched one tile ahead
      float izj0 = 0.f, izj1 = 0.f, qj0 = 1.f, qj1 = 1.f;
      if (wg < n_tiles) {
        const int jn = (u_begin + wg) % ncb;
        izj0 = coef[2 * (int64_t)(jn * PAIR_TN + r)];
        izj1 = coef[2 * (int64_t)(jn * PAIR_TN + TM + r)];
        qj0 = row_q(feat, g, jn * PAIR_TN + r);
        qj1 = row_q(feat, g, jn * PAIR_TN + TM + r);
      }
      for (int t = wg; t < n_tiles; t += 2) {
        const int u = u_begin + t;
        const int ib = u / ncb, j = u - ib * ncb;
        if (ib != cur_ib) {
          cur_ib = ib;
          const int row0 = g.row_begin + ib * TM;
          gi = row0 + r;
          bi = block_seg(row0, g.bseg);
          iz_i = coef[2 * (int64_t)gi];
          q_i = row_q(feat, g, gi);
          nshift_i = nshift + log2f(q_i);                // the tile is P q_g q_j (symmetric): q_g rides in the exponent
        }
        const BlockSeg bj0 = block_seg(j * PAIR_TN, g.bseg), bj1 = block_seg(j * PAIR_TN + TM, g.bseg);
        const float ks0 = ((bj0.mod == bi.mod) ? g.w : 1.0f) * sigma, ks1 = ((bj1.mod == bi.mod) ? g.w : 1.0f) * sigma;
        cv[(r >> 1) * 4 + (r & 1)] = qj0;                          // q_j, w_j = q_j kappa sigma / Z_j of the tile's 256 columns
        cv[(r >> 1) * 4 + (r & 1) + 2] = izj0 * ks0 * qj0;
        cv[((TM + r) >> 1) * 4 + (r & 1)] = qj1;
        cv[((TM + r) >> 1) * 4 + (r & 1) + 2] = izj1 * ks1 * qj1;
        if (t + 2 < n_tiles) {
          const int jn = (u + 2) % ncb;
          izj0 = coef[2 * (int64_t)(jn * PAIR_TN + r)];
          izj1 = coef[2 * (int64_t)(jn * PAIR_TN + TM + r)];
          qj0 = row_q(feat, g, jn * PAIR_TN + r);
          qj1 = row_q(feat, g, jn * PAIR_TN + TM + r);
        }
        named_bar_sync(1 + wg, EPI_THREADS);
        if (r == 0) TR(1 + wg, t >> 1, 0);
        mbar_wait(sfull_bar(wg), ((uint32_t)t >> 1) & 1);
        if (r == 0) TR(1 + wg, t >> 1, 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        tmem_ld32(lane_base, va);
#pragma unroll
        for (int h = 0; h < 2; ++h) {                              // the two 128-column P tiles of this S tile
          const BlockSeg bj = h ? bj1 : bj0;
          const bool same_mod = (bj.mod == bi.mod);
          const bool diag_tile = (bj.samp0 == bi.samp0);
          const float k = (same_mod ? g.k_intra : g.k_inter) * q_i;
          const float a_i = iz_i * (h ? ks1 : ks0);
          const uint32_t th = 2u * (uint32_t)t + h;                // P tile index of this pair
          const uint32_t slot = th % PAIR_NSLOT, use = th / PAIR_NSLOT;
          if (!(exp_flags & 2)) mbar_wait(pempty_bar(slot), (use & 1) ^ 1);   // the dF that read this scratch slot is done
          uint8_t* const prow = p_scratch + (size_t)slot * PTILE_BYTES;       // this thread's 16 bytes of column chunk 0
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int c = h * 4 + c4;                               // 32-column chunk of the S tile (0..7)
            uint32_t (&v)[32] = (c & 1) ? vb : va;
            tmem_ld_wait();
            if (c + 1 < 8) tmem_ld32(lane_base + (c + 1) * 32, (c & 1) ? va : vb);
            else { tc_fence_before(); mbar_arrive(sempty_bar(wg)); }   // whole S tile is in registers: free the buffer
            uint32_t packed[16];
            const float4* cv4 = reinterpret_cast<const float4*>(cv) + c * 16;
            const uint64_t k2 = pack2(k, k), ns2 = pack2(nshift_i, nshift_i), a2 = pack2(a_i, a_i);
#pragma unroll
            for (int q = 0; q < 32; q += 2) {
              const float4 cc = cv4[q >> 1];                         // (q_j, q_j+1, w_j, w_j+1) of columns q, q + 1
              const uint64_t q01 = pack2(cc.x, cc.y);
              const uint64_t x01 = ffma2(fmul2(pack2(__uint_as_float(v[q + 0]), __uint_as_float(v[q + 1])), k2), q01, ns2);
              float x0, x1, e0, e1;
              unpack2(x01, x0, x1);
              unpack2(fmul2(pack2(ex2_of(x0), ex2_of(x1)), ffma2(a2, q01, pack2(cc.z, cc.w))), e0, e1);
              if (diag_tile) {                                     // same-sample pair handled in grad_finish
                const int cbase = c4 * 32 + q;
                if (cbase + 0 == r) e0 = 0.f;
                if (cbase + 1 == r) e1 = 0.f;
              }
              __half2 h0 = __floats2half2_rn(e0, e1);
              packed[q >> 1] = *reinterpret_cast<uint32_t*>(&h0);
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              if (!(exp_flags & 1) || packed[ch * 4] == 0x7fff7fffu)    // exp bit0: perf experiment without the stores
                st_global_v4(prow + (c4 * 4 + ch) * (TM * 16), packed[ch * 4 + 0], packed[ch * 4 + 1], packed[ch * 4 + 2],
                             packed[ch * 4 + 3]);
          }
          // No proxy fence here: the arrive releases these generic-proxy stores (cta scope), the signaller's gpu-scope
          // fence is cumulative over them, and the reading side fences generic -> async before its TMA load.  (A
          // fence.proxy.async.global per epilogue thread is a MEMBAR.GPU each: it cost as much as the tile's math.)
          mbar_arrive(staged_bar(slot));
          if (r == 0) TR(1 + wg, t >> 1, 2 + h);
        }
        // every warp of the group is done reading this tile's column coefficients before any of them writes the next
        // tile's (cv is one buffer per group; without this a warp running ahead could overwrite its slice under a
        // slower warp's last chunks)
        named_bar_sync(1 + wg, EPI_THREADS);
      }
    }
  } else {
    // =========================================================================== G-CTA
    const uint32_t p_tiles = data;
    const uint32_t ring_base = data + PAIR_PBUF * 2 * CHUNK_BYTES;
    const uint32_t group_bytes = 4 * GBOX_BYTES;
    const int kc0 = (int)(rank - 1) * 8;              // this G-CTA's 512-wide slab of D starts at chunk kc0
    const int nkg = min(8, nk - kc0);                 // 64-wide chunks in the slab
    const int ndg = (nkg + 3) / 4;                    // d-groups of up to four 64-wide boxes (N = 64 * boxes <= 256)
    if (warp == 0) {
      Ring ring(PAIR_GGROUPS);
      PairWalk walk(u_begin, u_end, ncb);
      PairSeg sg;
      while (walk.next(sg)) {
        for (int jh = 2 * sg.j0; jh < 2 * sg.j1; ++jh) {           // 128-column P tiles
          for (int dg = 0; dg < ndg; ++dg) {
            const int nb = min(4, nkg - dg * 4);
            for (int kh = 0; kh < 2; ++kh) {                        // 64-row halves of the K = 128 j rows
              mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
              if (elect_one()) {
                const uint32_t st = ring_base + ring.stage * group_bytes;
                mbar_arrive_expect_tx(full_bar(ring.stage), nb * GBOX_BYTES);
                for (int q = 0; q < nb; ++q)
                  tma_load_2d(st + q * GBOX_BYTES, &tmap64, full_bar(ring.stage), (kc0 + dg * 4 + q) * KC,
                              jh * TM + kh * 64);
              }
              __syncwarp();
              ring.advance();
            }
          }
        }
      }
    } else if (warp == 3) {
      // P loader: published scratch tile -> one of the 3 shared-memory P buffers (two [128][64] swizzled boxes)
      const uint32_t n_ptiles = 2u * (uint32_t)(u_end - u_begin);
      for (uint32_t th = 0; th < n_ptiles; ++th) {
        const uint32_t slot = th % PAIR_NSLOT, use = th / PAIR_NSLOT;
        const uint32_t pb = th % PAIR_PBUF, puse = th / PAIR_PBUF;
        if (lane == 0) TR(4, th, 0);
        if (!(exp_flags & 2)) mbar_wait_cluster(pready_bar(slot), use & 1);
        if (lane == 0) TR(4, th, 1);
        fence_proxy_async_global();
        mbar_wait(pbempty_bar(pb), (puse & 1) ^ 1);
        if (lane == 0) TR(4, th, 2);
        if (elect_one()) {
          mbar_arrive_expect_tx(pbfull_bar(pb), PTILE_BYTES);
          bulk_load_1d(p_tiles + pb * PTILE_BYTES, scratch + ((size_t)pair * PAIR_NSLOT + slot) * PTILE_BYTES, PTILE_BYTES,
                       pbfull_bar(pb));
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      Ring ring(PAIR_GGROUPS);
      uint32_t th = 0, seg_iter = 0;
      PairWalk walk(u_begin, u_end, ncb);
      PairSeg sg;
      while (walk.next(sg)) {
        mbar_wait(acc_empty, (seg_iter & 1) ^ 1);       // the drain warps have emptied the previous segment's slab
        tc_fence_after();
        for (int jh = 2 * sg.j0; jh < 2 * sg.j1; ++jh, ++th) {
          const uint32_t pb = th % PAIR_PBUF, puse = th / PAIR_PBUF;
          const uint32_t slot = th % PAIR_NSLOT;
          if (lane == 0) TR(5, th, 0);
          mbar_wait(pbfull_bar(pb), puse & 1);          // the P tile has landed in shared memory (TMA)
          if (lane == 0) TR(5, th, 1);
          tc_fence_after();
          const uint32_t p_tile = p_tiles + pb * PTILE_BYTES;
          for (int dg = 0; dg < ndg; ++dg) {
            const int nb = min(4, nkg - dg * 4);
            const uint32_t idesc = make_idesc_f16(128, 64 * nb, 0, 0, 0, 1);
            for (int kh = 0; kh < 2; ++kh) {
              mbar_wait(full_bar(ring.stage), ring.phase);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t st = ring_base + ring.stage * group_bytes;
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                  // A = P[:, 64 kh + 16 k16 .. +16): no-swizzle K-major, column chunks 8 kh + 2 k16 and the next one
                  // (2048 bytes apart), 8-row groups 128 bytes apart; B = Fhat rows 64 kh + 16 k16 .. +16 of the
                  // [64 j][64 d] boxes, MN-major: 16 K rows = 2048 bytes, next 64-wide MN atom = next box
                  const uint64_t ad = make_smem_desc_nosw(p_tile + (kh * 8 + k16 * 2) * (TM * 16), TM * 16, 128);
                  const uint64_t bd = make_smem_desc_sw128(st + k16 * 2048, 1024, GBOX_BYTES);
                  umma_ss(tmem_base + dg * 256, ad, bd, idesc, (jh > 2 * sg.j0 || kh > 0 || k16 > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar(ring.stage));
              }
              __syncwarp();
              ring.advance();
            }
          }
          if (elect_one()) {
            umma_commit(pbempty_bar(pb));
            umma_commit_multicast(pempty_bar(slot), (uint16_t)1);   // scratch slot free -> the S-CTA (cluster rank 0)
          }
          if (lane == 0) TR(5, th, 2);
          __syncwarp();
        }
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
        ++seg_iter;
      }
    } else if (warp >= EPI_WARP0) {
      const int quad = warp & 3, wg = (warp - EPI_WARP0) >> 2;      // wg: which 256-column half of the slab
      const int r = quad * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
      uint32_t seg_iter = 0;
      PairWalk walk(u_begin, u_end, ncb);
      PairSeg sg;
      while (walk.next(sg)) {
        const int gi = g.row_begin + sg.ib * TM + r;
        mbar_wait(acc_full, seg_iter & 1);
        tc_fence_after();
        float* out = dfhat + (int64_t)(gi - g.row_begin) * g.dim + kc0 * KC;
        const bool whole = (sg.j0 == 0 && sg.j1 == ncb);
        const int c_end = min(nkg * KC, wg * 256 + 256) / 32;
#pragma unroll 1
        for (int c = wg * 8; c < c_end; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_base + c * 32, v);
          tmem_ld_wait();
          if (whole) {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              *reinterpret_cast<float4*>(out + c * 32 + q) =
                  make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                              __uint_as_float(v[q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              red_add_f32x4(out + c * 32 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]),
                            __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive(acc_empty);
        ++seg_iter;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // neither CTA leaves while its peer may still touch its shared memory
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// Self-test kernel: one CTA, D[128][n] = A[128][k] * B^T with the operand forms the real kernels use.
//   variant 0: A, B K-major from TMA tiles (the S product)
//   variant 1: A K-major written by threads with the swizzle formula (the P tile), B MN-major TMA tiles
//              (b_host is [k][n], n in {64, 128, 256}: n/64 boxes of [k rows][64], LBO = box size)  (the dF product)
//   variant 2: A from TMEM (tcgen05.st packed fp16 pairs), B K-major  (the single-CTA backward's P operand)
//   variant 3: A K-major WITHOUT swizzle, written by threads as [k/8][128 rows][16 B] (the paired backward's P), B K-major
// ================================================================================================
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int variant,
                int n, int k, const uint16_t* __restrict__ a_gmem, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nk = k / KC;
  const uint32_t a_s = base;                              // nk chunks of [128][64]
  const uint32_t b_s = a_s + nk * CHUNK_BYTES;            // variant 0/2: nk chunks of [n rows][64]; variant 1: n/64 boxes of [k rows][64]
  const uint32_t bchunk = (uint32_t)n * 128u;             // bytes of one [n rows][64] K-major chunk
  const uint32_t bar = b_s + (uint32_t)n * (uint32_t)k * 2u;
  const uint32_t done_bar = bar + 8;
  const uint32_t tmem_slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  (void)lane;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done_bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int r = threadIdx.x;

  if (variant == 1) {
    // write A with the P-tile formula: row r, columns in 16-byte chunks
    for (int atom = 0; atom < nk; ++atom)
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 val = *reinterpret_cast<const uint4*>(a_gmem + (size_t)r * k + atom * KC + ch * 8);
        const uint32_t addr = a_s + atom * CHUNK_BYTES + r * 128 + (((uint32_t)ch ^ (uint32_t)(r & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
      }
    fence_proxy_async_smem();
  }
  if (variant == 3) {
    // A in the no-swizzle K-major layout the paired backward uses for P: [k/8 column chunks][128 rows][16 bytes]
    for (int ch = 0; ch < k / 8; ++ch) {
      const uint4 val = *reinterpret_cast<const uint4*>(a_gmem + (size_t)r * k + ch * 8);
      const uint32_t addr = a_s + ch * 2048 + r * 16;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
    }
    fence_proxy_async_smem();
  }
  if (variant == 5) {
    // A in the layout the dataflow backward reads TRANSPOSED: element (m, kk) at (m/8) * (k*16) + kk*16 + (m%8)*2, i.e.
    // 16-byte chunks of 8 consecutive m, K rows 16 bytes apart: an MN-major operand without swizzle
    for (int kk = 0; kk < k; ++kk) {
      const uint16_t val = a_gmem[(size_t)r * k + kk];
      const uint32_t addr = a_s + (r >> 3) * (k * 16) + kk * 16 + (r & 7) * 2;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(val) : "memory");
    }
    fence_proxy_async_smem();
  }
  if (variant == 2) {
    // A -> TMEM columns [256, 256 + k/2): lane = row, column c holds elements (2c, 2c+1)
    for (int c0 = 0; c0 < k / 2; c0 += 16) {
      uint32_t v[16];
      for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<const uint32_t*>(a_gmem + (size_t)r * k + 2 * (c0 + i));
      tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (variant == 0 || variant == 6) {
      for (int kc = 0; kc < nk; ++kc) { tma_load_2d(a_s + kc * CHUNK_BYTES, &tmap_a, bar, kc * KC, 0); bytes += CHUNK_BYTES; }
    }
    if (variant == 1) {
      for (int a = 0; a < n / 64; ++a) {                  // boxes {64 n, k rows}
        tma_load_2d(b_s + a * k * 128, &tmap_b, bar, a * 64, 0);
        bytes += k * 128;
      }
    } else {
      for (int kc = 0; kc < nk; ++kc) { tma_load_2d(b_s + kc * bchunk, &tmap_b, bar, kc * KC, 0); bytes += bchunk; }
    }
    mbar_arrive_expect_tx(bar, bytes);
    mbar_wait(bar, 0);
    tc_fence_after();
    if (variant == 0 || variant == 6) {
      const uint32_t idesc = make_idesc_f16(128, n, 0, variant == 6 ? 1 : 0, 0, 0);      // variant 6: A fp16, B bf16
      for (int kc = 0; kc < nk; ++kc)
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem_base, kmajor_desc(a_s + kc * CHUNK_BYTES) + kk * 2, kmajor_desc(b_s + kc * bchunk) + kk * 2,
                  idesc, (kc | kk) ? 1u : 0u);
    } else if (variant == 1) {
      const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 0, 1);
      for (int k16 = 0; k16 < k / 16; ++k16)
        umma_ss(tmem_base, kmajor_desc(a_s + (k16 >> 2) * CHUNK_BYTES + (k16 & 3) * 32),
                make_smem_desc_sw128(b_s + k16 * 2048, 1024, (uint32_t)k * 128u), idesc, k16 ? 1u : 0u);
    } else if (variant == 3) {
      const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 0, 0);
      for (int k16 = 0; k16 < k / 16; ++k16)
        umma_ss(tmem_base, make_smem_desc_nosw(a_s + k16 * 2 * 2048, 2048, 128),
                kmajor_desc(b_s + (k16 >> 2) * bchunk) + (k16 & 3) * 2, idesc, k16 ? 1u : 0u);
    } else if (variant == 5) {
      const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 1, 0);      // A MN-major: LBO = 8-deep K groups, SBO = MN chunks
      for (int k16 = 0; k16 < k / 16; ++k16)
        umma_ss(tmem_base, make_smem_desc_nosw(a_s + k16 * 256, 128, (uint32_t)k * 16u),
                kmajor_desc(b_s + (k16 >> 2) * bchunk) + (k16 & 3) * 2, idesc, k16 ? 1u : 0u);
    } else {
      const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 0, 0);
      for (int kc = 0; kc < nk; ++kc)
        for (int kk = 0; kk < 4; ++kk)
          umma_ts(tmem_base, tmem_base + 256 + (kc * 4 + kk) * 8, kmajor_desc(b_s + kc * bchunk) + kk * 2, idesc,
                  (kc | kk) ? 1u : 0u);
    }
    umma_commit(done_bar);
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  for (int c = 0; c < n / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[(size_t)r * n + c * 32 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// variant 4: cta_group::2.  A cluster of two CTAs computes D[256][n] = A[256][k] * B[n][k]^T with ONE M = 256 MMA
// stream issued by the leader: CTA r stages rows [128 r, +128) of A and rows [r n/2, +n/2) of B (all TMA loads count
// on the leader's mbarrier), and reads its 128 accumulator rows back from its own TMEM.
__global__ void __launch_bounds__(128, 1)
selftest_2sm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int n, int k,
                    float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nk = k / KC;
  const uint32_t rank = cluster_ctarank();
  const uint32_t a_s = base;                                   // nk chunks of [128][64]
  const uint32_t b_s = a_s + nk * CHUNK_BYTES;                 // nk chunks of [n/2 rows][64]
  const uint32_t bchunk = (uint32_t)(n / 2) * 128u;
  const uint32_t bar = b_s + nk * bchunk;
  const uint32_t done_bar = bar + 8;
  const uint32_t tmem_slot = bar + 16;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done_bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int r = threadIdx.x;

  if (threadIdx.x == 0) {
    const uint32_t leader_bar = mapa_cluster(bar, 0);
    const uint32_t my_bytes = nk * CHUNK_BYTES + nk * bchunk;
    if (rank == 0) mbar_arrive_expect_tx(bar, 2 * my_bytes);
    for (int kc = 0; kc < nk; ++kc) {
      tma_load_2d_2sm(a_s + kc * CHUNK_BYTES, &tmap_a, leader_bar, kc * KC, (int)rank * TM);
      tma_load_2d_2sm(b_s + kc * bchunk, &tmap_b, leader_bar, kc * KC, (int)rank * (n / 2));
    }
    if (rank == 0) {
      mbar_wait(bar, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(256, n, 0, 0, 0, 0);
      for (int kc = 0; kc < nk; ++kc)
        for (int kk = 0; kk < 4; ++kk)
          umma_ss_2sm(tmem_base, kmajor_desc(a_s + kc * CHUNK_BYTES) + kk * 2, kmajor_desc(b_s + kc * bchunk) + kk * 2,
                      idesc, (kc | kk) ? 1u : 0u);
      umma_commit_2sm(done_bar, (uint16_t)3);
    }
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  for (int c = 0; c < n / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[((size_t)rank * TM + r) * n + c * 32 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2sm(tmem_base, 512);
}



// CROSSCLR_BWD_VARIANT (A/B measurements, tests): 1 forces the single-CTA slab kernel, 2 the 1 S-CTA + G-CTA(s)
// cluster kernel, 4 the dataflow kernel (flow_kernels.cu) below its size threshold; default 0 = chosen by shape.
int bwd_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CROSSCLR_BWD_VARIANT");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return v;
}

}  // namespace

// CROSSCLR_FWD_SYM=0 disables the symmetric (upper-triangle) walk of the single-rank forward (A/B measurements).
static bool fwd_sym_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CROSSCLR_FWD_SYM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

constexpr int kFwdQvBytes = FWD_QSTAGES * FWD_TN * 4;   // ring of per-tile column scales

template <bool kResident, bool kSym, int kOrder>
static int launch_fwd_tc2_t(const CUtensorMap& tmap, const void* feat, const Geometry& g, float* stats, cudaStream_t st,
                            const FwdFinalize& fin) {
  const int nk = s_chunks(g), ncb = g.rows / FWD_TN, nrbp = g.row_count / (2 * TM);
  const int tiles = kSym ? ncb * (ncb + 1) / 2 : nrbp * ncb;
  const size_t a_bytes = kResident ? (size_t)nk * CHUNK_BYTES : 0;
  const size_t stage_bytes = (size_t)(kResident ? 1 : 2) * CHUNK_BYTES;
  const size_t fixed = 1024 + kBarBytes + kFwdQvBytes + a_bytes;
  const int stages = (int)std::min<size_t>(MAX_SLOTS, (kMaxSmem - fixed) / stage_bytes);
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  CC_CHECK_CUDA(cudaFuncSetAttribute(fwd_tc2_kernel<kResident, kSym, kOrder>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * std::min(tiles, sm_count() / 2));
  cfg.blockDim = dim3(FWD2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  static const int exp_flags = getenv("CROSSCLR_FWD_EXP") ? atoi(getenv("CROSSCLR_FWD_EXP")) : 0;   // perf experiments only
  TimedLaunch timed(CROSSCLR_K_FWD, st);
  CC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fwd_tc2_kernel<kResident, kSym, kOrder>, tmap, (const uint8_t*)feat, g, stats, tiles, ncb,
                                   nrbp, nk, stages, exp_flags, fin));
  return check_launch("fwd_tc2_kernel");
}

bool fwd_tc_can_finalize(const Geometry& g) { return g.row_count == g.rows && g.row_begin == 0; }

int launch_fwd_tc(const Geometry& g, const void* feat, float* stats, cudaStream_t st, const FwdFinalize* fin_req) {
  const FwdFinalize fin = (fin_req != nullptr && fwd_tc_can_finalize(g)) ? *fin_req : FwdFinalize{nullptr, nullptr, nullptr, nullptr};
  if (fin_req != nullptr && fin.ticket == nullptr) { set_error("fused finalize needs a single-rank problem"); return CROSSCLR_EINVAL; }
  CUtensorMap tmap;
  int rc = CC_FEAT_TMAP(&tmap, feat, g, TM);
  if (rc) return rc;
  // Super-tile order: every pair inside the same 8 x 8 tiles at a time, so that row / column blocks come from HBM once per
  // super-tile instead of once per tile -- for problems whose stacked matrix is well beyond the 126 MB L2 AND whose row block
  // cannot stay resident anyway (D > 512: c5).  Measured on B200: c5 79 -> 63 ms on one GPU, 35 -> 28 ms per rank at N = 4;
  // with a resident row block (c4, D = 512) the linear order is as good or better (8.5 vs 9.0 ms).
  // Row-interleaved order: large problems with a resident row block (D <= 512: c4) -- pairs sweep whole rows in step, the row
  // block stays in shared memory and the column blocks are shared through the L2.  Measured (scripts/gpu_fwd_blocked.sh):
  // c4 8.2 -> 6.7 ms on one GPU, 3.0 -> 2.7 ms per rank at N = 4; B = 32768, D = 512 (75 MB) 1.72 -> 1.65 ms.
  // CROSSCLR_FWD_BLOCKED = 0 / 1 / 2 forces linear / super-tile / row-interleaved order.
  static const int force_order = getenv("CROSSCLR_FWD_BLOCKED") ? atoi(getenv("CROSSCLR_FWD_BLOCKED")) : -1;
  const bool can_reside = s_chunks(g) <= MAX_RES_CHUNKS && !g.split;    // split rows: A's K chunks are not B's, stream both
  const size_t bytes = (size_t)g.rows * g.pitch * 2;
  int order = force_order >= 0 ? force_order
                               : (can_reside ? (bytes > ((size_t)64 << 20) ? FWD_ORDER_ROWS : FWD_ORDER_LINEAR)
                                             : (bytes > ((size_t)100 << 20) ? FWD_ORDER_BLOCKED : FWD_ORDER_LINEAR));
  if (order == FWD_ORDER_ROWS && !can_reside) order = FWD_ORDER_BLOCKED;
  // single rank: all rows are owned, the Gram matrix is square and symmetric -> upper triangle of pair tiles only
  const bool sym = g.row_count == g.rows && g.row_begin == 0 && fwd_sym_enabled();
  if (order == FWD_ORDER_BLOCKED)
    return sym ? launch_fwd_tc2_t<false, true, FWD_ORDER_BLOCKED>(tmap, feat, g, stats, st, fin)
               : launch_fwd_tc2_t<false, false, FWD_ORDER_BLOCKED>(tmap, feat, g, stats, st, fin);
  if (order == FWD_ORDER_ROWS)
    return sym ? launch_fwd_tc2_t<true, true, FWD_ORDER_ROWS>(tmap, feat, g, stats, st, fin)
               : launch_fwd_tc2_t<true, false, FWD_ORDER_ROWS>(tmap, feat, g, stats, st, fin);
  if (sym) return can_reside ? launch_fwd_tc2_t<true, true, FWD_ORDER_LINEAR>(tmap, feat, g, stats, st, fin)
                             : launch_fwd_tc2_t<false, true, FWD_ORDER_LINEAR>(tmap, feat, g, stats, st, fin);
  return can_reside ? launch_fwd_tc2_t<true, false, FWD_ORDER_LINEAR>(tmap, feat, g, stats, st, fin)
                    : launch_fwd_tc2_t<false, false, FWD_ORDER_LINEAR>(tmap, feat, g, stats, st, fin);
}

// How many clusters of bwd_pair_kernel (1 S-CTA + csize-1 G-CTAs) can be resident at once.  Queried once per
// process and cluster size; 0 if the device cannot schedule that cluster shape.
template <bool kResident>
static int pair_clusters_resident(int csize) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  static bool done[9] = {false, false, false, false, false, false, false, false, false};
  if (!done[csize]) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize * (sm_count() / csize));
    cfg.blockDim = dim3(PAIR_THREADS);
    cfg.dynamicSmemBytes = kMaxSmem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int c = 0;
    cudaFuncSetAttribute(bwd_pair_kernel<kResident>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (cudaOccupancyMaxActiveClusters(&c, bwd_pair_kernel<kResident>, &cfg) != cudaSuccess || c < 0) {
      (void)cudaGetLastError();
      c = 0;
    }
    cache[csize] = std::min(c, sm_count() / csize);
    done[csize] = true;
  }
  return cache[csize];
}

size_t bwd_pair_scratch_bytes() { return (size_t)(sm_count() / 2) * PAIR_NSLOT * PTILE_BYTES; }

// cluster size for embedding dim `dim`: one G-CTA per 512 columns; 0 = use the single-CTA kernel
static int pair_cluster_size(int dim) {
  if (bwd_variant() == 1 || dim <= 256) return 0;   // (variants 0, 2, 3 fall through to the shape rule)
  const int csize = 1 + (dim + 511) / 512;
  if (csize > 4) return 0;
  const int resident = dim <= 512 ? pair_clusters_resident<true>(csize) : pair_clusters_resident<false>(csize);
  return resident * csize * 10 >= sm_count() * 8 ? csize : 0;     // needs >= 80 % of the SMs in clusters
}

// debug timeline of the role-specialised backward kernels: CROSSCLR_PAIR_TRACE=<file> makes cluster 0 stamp clock64
// per role / tile / event; the launch then synchronises and writes the table (perf debugging only)
static unsigned long long* trace_buffer(cudaStream_t st) {
  static unsigned long long* trace = nullptr;
  static const bool want_trace = getenv("CROSSCLR_PAIR_TRACE") != nullptr;
  if (!want_trace) return nullptr;
  if (trace == nullptr) cudaMalloc(&trace, 6 * 64 * 4 * 8);
  cudaMemsetAsync(trace, 0, 6 * 64 * 4 * 8, st);
  return trace;
}
static void trace_dump(unsigned long long* trace, cudaStream_t st) {
  if (trace == nullptr) return;
  static unsigned long long host[6 * 64 * 4];
  cudaStreamSynchronize(st);
  cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost);
  FILE* f = fopen(getenv("CROSSCLR_PAIR_TRACE"), "w");
  if (f) {
    for (int role = 0; role < 6; ++role)
      for (int t = 0; t < 64; ++t)
        fprintf(f, "%d %d %llu %llu %llu %llu\n", role, t, host[(role * 64 + t) * 4], host[(role * 64 + t) * 4 + 1],
                host[(role * 64 + t) * 4 + 2], host[(role * 64 + t) * 4 + 3]);
    fclose(f);
  }
}

template <bool kResident>
static int launch_bwd_pair_t(const CUtensorMap& tmap, const void* feat, const Geometry& g, const float* coef,
                             const float* scal, float* dfhat, void* scratch, int csize, cudaStream_t st) {
  CUtensorMap tmap64;                     // [64 rows][64 cols] boxes for the G-CTAs' MN-major operand groups
  int rc = CC_FEAT_TMAP(&tmap64, feat, g, 64);
  if (rc) return rc;
  const int nk = g.dim / KC, ncb = g.rows / PAIR_TN, nrb = g.row_count / TM;
  const long long n_units_ll = (long long)nrb * ncb;
  if (n_units_ll > 0x7fffffffLL) { set_error("crossclr_bwd: problem too large (%lld work units)", n_units_ll); return CROSSCLR_EINVAL; }
  const int n_units = (int)n_units_ll;
  CC_CHECK_CUDA(cudaFuncSetAttribute(bwd_pair_kernel<kResident>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
  const size_t a_bytes = kResident ? (size_t)nk * CHUNK_BYTES : 0;
  const int s_stages = std::min((int)((kMaxSmem - PAIR_HDR - a_bytes) / ((kResident ? 2 : 3) * CHUNK_BYTES)), MAX_SLOTS);
  const int nclusters = std::min(n_units, pair_clusters_resident<kResident>(csize));
  TimedLaunch timed(CROSSCLR_K_BWD, st);               // after the host-side preparation: the bracket holds device work only
  CC_CHECK_CUDA(cudaMemsetAsync(dfhat, 0, (size_t)g.row_count * g.dim * sizeof(float), st));
  static const int exp_flags = getenv("CROSSCLR_PAIR_EXP") ? atoi(getenv("CROSSCLR_PAIR_EXP")) : 0;
  unsigned long long* trace = trace_buffer(st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * nclusters);
  cfg.blockDim = dim3(PAIR_THREADS);
  cfg.dynamicSmemBytes = kMaxSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  CC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, bwd_pair_kernel<kResident>, tmap, tmap64, (const uint8_t*)feat, g, coef, scal, dfhat,
                                   (uint8_t*)scratch, n_units, ncb, nk, s_stages, exp_flags, trace));
  trace_dump(trace, st);
  return check_launch("bwd_pair_kernel");
}

const char* bwd_tc_kernel_name(const Geometry& g) {
  if (bwd_flow_applies(g)) return "bwd_flow_kernel";
  if (pair_cluster_size(g.dim)) return "bwd_pair_kernel";
  return "bwd_tc_kernel";
}

int launch_bwd_tc(const Geometry& g, const void* feat, const float* coef, const float* scal, float* dfhat,
                  float* dfhat_late, bool* two_partials, void* scratch, cudaStream_t st) {
  *two_partials = false;
  CUtensorMap tmap;
  int rc = CC_FEAT_TMAP(&tmap, feat, g, TM);
  if (rc) return rc;
  if (g.split && !bwd_flow_applies(g)) {
    set_error("CROSSCLR_PATH_TC_SPLIT needs the dataflow backward: >= 2048 stacked rows and dim <= 1024 (got %d rows, dim %d)",
              g.rows, g.dvalid);
    return CROSSCLR_EINVAL;
  }
  if (bwd_flow_applies(g))                             // producer pairs -> P-tile pool -> consumer pairs (flow_kernels.cu)
    return launch_bwd_flow(g, feat, coef, scal, dfhat, dfhat_late, two_partials, scratch, st);
  if (const int csize = pair_cluster_size(g.dim)) {    // > 1 slab would recompute S: role-specialised CTA clusters
    return g.dim <= 512 ? launch_bwd_pair_t<true>(tmap, feat, g, coef, scal, dfhat, scratch, csize, st)
                        : launch_bwd_pair_t<false>(tmap, feat, g, coef, scal, dfhat, scratch, csize, st);
  }
  const int nk = g.dim / KC, ncb = g.rows / BWD_TN, nrb = g.row_count / TM;
  const int n_slabs = (g.dim + SLAB - 1) / SLAB;
  const long long n_units_ll = (long long)nrb * n_slabs * ncb;
  if (n_units_ll > 0x7fffffffLL) { set_error("crossclr_bwd: problem too large (%lld work units)", n_units_ll); return CROSSCLR_EINVAL; }
  const int n_units = (int)n_units_ll;
  TimedLaunch timed(CROSSCLR_K_BWD, st);
  const bool resident = nk <= MAX_RES_CHUNKS;
  const size_t a_bytes = resident ? (size_t)nk * CHUNK_BYTES : 0;
  const size_t avail = kMaxSmem - 1024 - 2 * BWD_TN * 8 - kBarBytes - a_bytes;
  const int slots = std::min((int)(avail / CHUNK_BYTES) & ~1, MAX_SLOTS);
  const size_t smem = 1024 + a_bytes + (size_t)slots * CHUNK_BYTES + 2 * BWD_TN * 8 + kBarBytes;
  const int grid = std::min(n_units, sm_count());
  // CTAs own balanced unit ranges; a range that cuts an item adds its partial slab into dfhat
  CC_CHECK_CUDA(cudaMemsetAsync(dfhat, 0, (size_t)g.row_count * g.dim * sizeof(float), st));
  if (resident) {
    CC_CHECK_CUDA(cudaFuncSetAttribute(bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_tc_kernel<true><<<grid, NUM_THREADS, smem, st>>>(tmap, (const uint8_t*)feat, g, coef, scal, dfhat, n_units, n_slabs, ncb, nk, slots);
  } else {
    CC_CHECK_CUDA(cudaFuncSetAttribute(bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_tc_kernel<false><<<grid, NUM_THREADS, smem, st>>>(tmap, (const uint8_t*)feat, g, coef, scal, dfhat, n_units, n_slabs, ncb, nk, slots);
  }
  return check_launch("bwd_tc_kernel");
}

int run_selftest(int variant, const uint16_t* a, const uint16_t* b, float* out, int n, int k) {
  if (variant < 0 || variant > 6 || k % KC != 0 || k < KC || k > 256 || n % 32 != 0 || n < 32 || n > 256 ||
      (variant == 1 && n % 64 != 0) || (variant == 4 && n % 64 != 0)) {
    set_error("crossclr_selftest: unsupported variant/shape (variant %d n %d k %d)", variant, n, k);
    return CROSSCLR_EINVAL;
  }
  const int m = (variant == 4) ? 256 : 128;       // variant 4 (cta_group::2): a is [256][k], out [256][n]
  uint16_t *da = nullptr, *db = nullptr;
  float* dout = nullptr;
  CC_CHECK_CUDA(cudaMalloc(&da, (size_t)m * k * 2));
  CC_CHECK_CUDA(cudaMalloc(&db, (size_t)n * k * 2));
  CC_CHECK_CUDA(cudaMalloc(&dout, (size_t)m * n * 4));
  CC_CHECK_CUDA(cudaMemcpy(da, a, (size_t)m * k * 2, cudaMemcpyHostToDevice));
  CC_CHECK_CUDA(cudaMemcpy(db, b, (size_t)n * k * 2, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  int rc = make_tmap_f16(&ta, da, (uint64_t)m, (uint64_t)k, 128);
  if (!rc) {
    if (variant == 1) rc = make_tmap_f16(&tb, db, (uint64_t)k, (uint64_t)n, (uint32_t)k);        // b is [k][n], box {64, k}
    else if (variant == 4) rc = make_tmap_f16(&tb, db, (uint64_t)n, (uint64_t)k, (uint32_t)(n / 2));
    else rc = make_tmap_f16(&tb, db, (uint64_t)n, (uint64_t)k, (uint32_t)n);                      // b is [n][k]
  }
  if (!rc) {
    const size_t smem = 1024 + (size_t)(k / KC) * CHUNK_BYTES + (size_t)n * k * 2 + 64;
    if (variant == 4) {
      cudaFuncSetAttribute(selftest_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = nullptr;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, selftest_2sm_kernel, ta, tb, n, k, dout);
      if (e != cudaSuccess) { (void)cudaGetLastError(); set_error("selftest_2sm launch failed: %s", cudaGetErrorString(e)); rc = CROSSCLR_ECUDA; }
      else rc = check_launch("selftest_2sm_kernel");
    } else {
      cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      selftest_kernel<<<1, 128, smem>>>(ta, tb, variant, n, k, da, dout);
      rc = check_launch("selftest_kernel");
    }
    if (!rc) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { set_error("selftest kernel failed: %s", cudaGetErrorString(e)); rc = CROSSCLR_ECUDA; }
    }
    if (!rc) {
      cudaError_t e = cudaMemcpy(out, dout, (size_t)m * n * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { set_error("selftest copy-back failed: %s", cudaGetErrorString(e)); rc = CROSSCLR_ECUDA; }
    }
  }
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return rc;
}

}  // namespace crossclr
