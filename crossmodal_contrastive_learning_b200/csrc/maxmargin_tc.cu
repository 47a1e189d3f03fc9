// MaxMargin_coot (trainer/loss.py:17-41 of the reference) and the retrieval ranks of the same score matrix on the
// tensor cores: tcgen05 score tiles with a hinge epilogue.  SURVEY.md section 8 rows f1 / f4.
//
//   scores = im s^T (`cosine_sim`, :7-15: a plain mm, no normalisation),  d_i = scores_ii
//   forward   loss = (1/B^2) sum_{i != j} [ max(0, m + s_ij - d_i) + max(0, m + s_ij - d_j) ]                  (:34-41)
//             cnt_i = #{j != i : m + s_ij - d_i > 0} + #{j != i : m + s_ji - d_i > 0}   (what the backward's diagonal needs)
//             with m = 0 the two counts are the retrieval ranks of the matching item (im -> s and s -> im)
//   backward  G_ij = 1[m + s_ij - d_i > 0] + 1[m + s_ij - d_j > 0]  (i != j),  G_ii = -cnt_i
//             dL/dim = (c/B^2) G s,  dL/ds = (c/B^2) G^T im: the tile formula is symmetric under swapping the two operands,
//             so ONE kernel serves both (second launch with the tensor maps exchanged)
//
// One persistent CTA per SM walks a balanced range of work units (128-row block, 256-wide slab of D, 128-column block):
// TMA producer warp -> shared-memory ring (128-byte-swizzled [128][64] boxes straight out of the CALLER's tensors: 16-bit
// inputs need no staging copy, rows / columns past the edge are zero-filled by TMA), one MMA-issuing lane, score tiles
// double-buffered in TMEM, eight epilogue warps (two per TMEM lane quadrant, one per 64-column half of the tile).  Forward: hinge sums in registers, row counts in registers, column counts by
// warp ballots -> shared-memory atomics -> one global atomic per column and tile.  Backward: the indicator tile G (0 / 1 / 2:
// exact in fp16 and bf16) is written IN PLACE over the score tile in TMEM (tcgen05.st) and is the A operand of
// dA[128 x 256] += G(j) B_j[:, slab], B read MN-major from the same kind of boxes the score product uses; the diagonal
// entry is applied by the finishing pass (-cnt_i B_i), where it is exact whatever the count.
// fp32 inputs (the reference multiplies them in fp32) are staged as fp16 hi + lo pairs after a power-of-two scale per tensor and
// the score product runs over K = 3 D (hi.hi + lo.hi + hi.lo, a column map in the TMA producer): scores to ~2^-22 relative,
// the dropped lo.lo term included -- the precision of an fp32 dot product.  Small problems stay on the CUDA cores (maxmargin.cu).
#include "tc_common.cuh"

namespace crossclr {

namespace {

constexpr int MM_TN = 128;                    // score tile columns
constexpr int MM_THREADS = 384;                // TMA warp, MMA warp, TMEM warp, one idle, eight epilogue warps
constexpr int MM_EPI_THREADS = 256;
constexpr int MM_VEC_BYTES = 2 * MM_TN * 8;   // per-tile column vector (margin - d_j) and column counts, double-buffered

struct MmSeg { int ib, sb, j0, j1; bool last_of_ib; };
struct MmWalk {
  int u, u_end, ncb, n_slabs;
  __device__ MmWalk(int u0, int u1, int ncb_, int ns_) : u(u0), u_end(u1), ncb(ncb_), n_slabs(ns_) {}
  __device__ __forceinline__ bool next(MmSeg& s) {
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

template <int kFmt>
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  if (kFmt == 0) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kFmt: 0 = fp16, 1 = bf16 operands.  kGrad: false = forward (hinge sums, counts), true = one direction of the backward.
// kVar: 0 = row block streamed beside the column blocks, 1 = row block resident in shared memory (D <= 512), 2 = split rows
// (fp32 inputs as fp16 [hi | lo], streamed) -- compile-time, so that the 16-bit kernels carry none of the split bookkeeping.
template <int kFmt, bool kGrad, int kVar>
__global__ void __launch_bounds__(MM_THREADS, 1)
mm_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, int B, float margin,
             const float* __restrict__ diag, float* __restrict__ cnt, double* __restrict__ acc, int* __restrict__ rank_row,
             int* __restrict__ rank_col, float* __restrict__ dacc, int dpad, int n_units, int n_slabs, int ncb, int nk,
             int num_slots, const float* __restrict__ scales) {
  constexpr bool kResident = kVar == 1;
  constexpr bool split = kVar == 2;
  // split (fp32 inputs staged as fp16 [hi | lo] rows, lo at column dpad): nk = 3 dpad / 64 score chunks with the column maps
  // A = [hi | lo | hi], B = [hi | hi | lo] (tc_common.cuh), the gradient operand is hi + lo (two MMAs per box position), and
  // the staged rows carry a power-of-two scale per tensor: true score = accumulator * scales[0] * scales[1]
  const int nkd = dpad / KC;
  constexpr int nparts = split ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_region = base;
  const uint32_t ring_base = a_region + (kResident ? nk * CHUNK_BYTES : 0);
  const uint32_t vec_base = ring_base + num_slots * CHUNK_BYTES;
  const uint32_t bar_base = vec_base + MM_VEC_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_SLOTS + s); };
  const uint32_t a_full = bar_base + 8u * (2 * MAX_SLOTS);
  const uint32_t a_empty = a_full + 8;
  auto sfull_bar = [&](int b) { return a_full + 16u + 8u * b; };
  auto pfull_bar = [&](int b) { return a_full + 32u + 8u * b; };     // forward: "epilogue has read the tile"
  const uint32_t acc_full = a_full + 48u, acc_empty = a_full + 56u;
  const uint32_t tmem_slot = a_full + 64u;
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw_u32));
  float* cvec = reinterpret_cast<float*>(smem_raw + (vec_base - raw_u32));                 // [2][128] margin - d_j
  int* ccnt = reinterpret_cast<int*>(smem_raw + (vec_base - raw_u32) + 2 * MM_TN * 4);     // [2][128] column counts
  const uint32_t idesc_s = make_idesc_f16(128, MM_TN, kFmt, kFmt, 0, 0);     // S = A(K-major) B(K-major)^T
  const uint32_t idesc_g128 = make_idesc_f16(128, 128, kFmt, kFmt, 0, 1);    // dA += G(TMEM) B_j(MN-major)
  const uint32_t idesc_g64 = make_idesc_f16(128, 64, kFmt, kFmt, 0, 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u_begin = (int)((long long)blockIdx.x * n_units / gridDim.x);
  const int u_end = (int)((long long)(blockIdx.x + 1) * n_units / gridDim.x);

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmapA); prefetch_tmap(&tmapB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < num_slots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(pfull_bar(b), MM_EPI_THREADS); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, MM_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem_base + 2 * MM_TN;      // slab accumulator: columns [256, 512)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    Ring ring(num_slots);
    int cur_ib = -1;
    uint32_t a_cnt = 0;
    MmWalk walk(u_begin, u_end, ncb, n_slabs);
    MmSeg sg;
    while (walk.next(sg)) {
      const int row0 = sg.ib * TM;
      const int d0 = sg.sb * SLAB;
      const int nsc = min(SLAB, dpad - d0) / KC;
      // a 128-wide dA operand spans two consecutive slots; slots stay pair-aligned only if every ring user moves in pairs
      const bool dpair = (nsc & 1) == 0 && (!kResident || (nk & 1) == 0);
      const bool dadv2 = dpair || !kResident;
      if (kResident && sg.ib != cur_ib) {
        mbar_wait(a_empty, (a_cnt & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full, nk * CHUNK_BYTES);
          for (int kc = 0; kc < nk; ++kc) tma_load_2d(a_region + kc * CHUNK_BYTES, &tmapA, a_full, kc * KC, row0);
        }
        __syncwarp();
        cur_ib = sg.ib; ++a_cnt;
      }
      auto load_S = [&](int j) {
        const int col0 = j * MM_TN;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
          if (!kResident) mbar_wait(empty_bar(ring.stage + 1), ring.phase ^ 1);
          if (elect_one()) {
            const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
            if (kResident) {
              mbar_arrive_expect_tx(full_bar(ring.stage), CHUNK_BYTES);
              tma_load_2d(st, &tmapB, full_bar(ring.stage), kc * KC, col0);
            } else {
              mbar_arrive_expect_tx(full_bar(ring.stage), 2 * CHUNK_BYTES);
              tma_load_2d(st, &tmapA, full_bar(ring.stage), split ? a_kcol(kc, nkd) : kc * KC, row0);
              tma_load_2d(st + CHUNK_BYTES, &tmapB, full_bar(ring.stage), split ? b_kcol(kc, nkd) : kc * KC, col0);
              mbar_arrive(full_bar(ring.stage + 1));
            }
          }
          __syncwarp();
          ring.advance();
          if (!kResident) ring.advance();
        }
      };
      auto load_dA = [&](int j) {          // B operand of dA(j): B_j[:, slab] as [128 j][64 d] boxes
        const int col0 = j * MM_TN;
        for (int c = 0; c < nsc; c += (dpair ? 2 : 1)) {
          for (int part = 0; part < nparts; ++part) {          // hi, then lo (split rows)
            const int dc = part * dpad + d0 + c * KC;
            mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
            if (dadv2) mbar_wait(empty_bar(ring.stage + 1), ring.phase ^ 1);
            if (elect_one()) {
              const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
              mbar_arrive_expect_tx(full_bar(ring.stage), dpair ? 2 * CHUNK_BYTES : CHUNK_BYTES);
              tma_load_2d(st, &tmapB, full_bar(ring.stage), dc, col0);
              if (dpair) tma_load_2d(st + CHUNK_BYTES, &tmapB, full_bar(ring.stage), dc + KC, col0);
              if (dadv2) mbar_arrive(full_bar(ring.stage + 1));   // keep the skipped slot's barriers in phase
            }
            __syncwarp();
            ring.advance();
            if (dadv2) ring.advance();
          }
        }
      };
      load_S(sg.j0);
      if (sg.j0 + 1 < sg.j1) load_S(sg.j0 + 1);
      for (int j = sg.j0; j < sg.j1; ++j) {
        if (kGrad) load_dA(j);
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
          issue_s_chunk(tmem_base + buf * MM_TN, a_addr, b_addr, idesc_s, kc == 0);
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
    MmWalk walk(u_begin, u_end, ncb, n_slabs);
    MmSeg sg;
    while (walk.next(sg)) {
      const int d0 = sg.sb * SLAB;
      const int nsc = min(SLAB, dpad - d0) / KC;
      const bool dpair = (nsc & 1) == 0 && (!kResident || (nk & 1) == 0);
      const bool dadv2 = dpair || !kResident;
      if (kResident && sg.ib != cur_ib) {
        mbar_wait(a_full, a_cnt & 1);
        tc_fence_after();
        cur_ib = sg.ib; ++a_cnt;
      }
      issue_S();
      if (sg.j0 + 1 < sg.j1) issue_S();
      if (kGrad) {
        mbar_wait(acc_empty, (seg_iter & 1) ^ 1);     // the epilogue has drained the previous segment's slab
        tc_fence_after();
      }
      for (int j = sg.j0; j < sg.j1; ++j, ++p_cnt) {
        const uint32_t buf = p_cnt & 1;
        mbar_wait(pfull_bar(buf), (p_cnt >> 1) & 1);
        tc_fence_after();
        if (kGrad) {
          const uint32_t p_tmem = tmem_base + buf * MM_TN;
          for (int c = 0; c < nsc; c += (dpair ? 2 : 1))
          for (int part = 0; part < nparts; ++part) {
            mbar_wait(full_bar(ring.stage), ring.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t st = ring_base + ring.stage * CHUNK_BYTES;
#pragma unroll
              for (int k16 = 0; k16 < MM_TN / 16; ++k16) {
                // A = G[:, 16 k16 .. +16) from TMEM (8 columns of packed 16-bit pairs per K = 16 step; each 64-column half of
                // the tile is packed into the front of its own half); B = B_j[16 k16 .. +16, 64 or 128 d]: MN-major view of
                // the TMA boxes, 16 K rows = 2048 bytes, second 64-wide atom = next slot
                const uint64_t bd = make_smem_desc_sw128(st + k16 * 2048, 1024, CHUNK_BYTES);
                const uint32_t g_tmem = p_tmem + (k16 < 4 ? k16 * 8 : 64 + (k16 - 4) * 8);
                umma_ts(tmem_acc + c * KC, g_tmem, bd, dpair ? idesc_g128 : idesc_g64,
                        (j > sg.j0 || k16 > 0 || part > 0) ? 1u : 0u);
              }
              umma_commit(empty_bar(ring.stage));
              if (dadv2) umma_commit(empty_bar(ring.stage + 1));
            }
            __syncwarp();
            ring.advance();
            if (dadv2) ring.advance();
          }
        }
        if (j + 2 < sg.j1) issue_S();                  // reuses buffer `buf` (behind dA(j) in issue order)
      }
      if (elect_one()) {
        if (kGrad) umma_commit(acc_full);
        if (kResident && sg.last_of_ib) umma_commit(a_empty);
      }
      __syncwarp();
      ++seg_iter;
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue: 8 warps, two per TMEM lane quadrant
    // warp (quadrant qd = warp % 4, half h): rows [32 qd, +32) of the tile, columns [64 h, +64) as two chunks of 32
    const int qd = warp & 3, h = (warp - EPI_WARP0) >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qd * 32) << 16);
    const float ninf = __int_as_float(0xff800000);
    const float sc = split ? scales[0] * scales[1] : 1.0f;      // both powers of two: exact (16-bit inputs: folds away)
    uint32_t seg_iter = 0, p_cnt = 0;
    double tot = 0.0;
    int pend_col0 = -1;                      // forward: column block whose counts sit in ccnt[(p_cnt - 1) & 1]
    MmWalk walk(u_begin, u_end, ncb, n_slabs);
    MmSeg sg;
    while (walk.next(sg)) {
      const int row0 = sg.ib * TM;
      const int gi = row0 + r;
      const float mi = gi < B ? margin - diag[gi] : ninf;      // rows past the edge: every comparison false
      int rowc = 0;
      // d_j of the tile's columns, fetched one tile ahead (a global load right in front of the barrier costs its latency per tile)
      float dj_next = (h == 0 && sg.j0 * MM_TN + r < B) ? diag[sg.j0 * MM_TN + r] : 0.f;
      for (int j = sg.j0; j < sg.j1; ++j, ++p_cnt) {
        const int col0 = j * MM_TN;
        const uint32_t buf = p_cnt & 1;
        const uint32_t tbuf = lane_base + buf * MM_TN + h * 64;
        float* cv = cvec + buf * MM_TN;
        if (h == 0) {
          cv[r] = (col0 + r < B) ? margin - dj_next : ninf;
          if (!kGrad) ccnt[buf * MM_TN + r] = 0;
          if (j + 1 < sg.j1 && col0 + MM_TN + r < B) dj_next = diag[col0 + MM_TN + r];
        }
        named_bar_sync(1, MM_EPI_THREADS);
        if (!kGrad && h == 0 && pend_col0 >= 0) {      // every warp is past the previous tile: flush its column counts
          const int v = ccnt[(buf ^ 1) * MM_TN + r];
          if (v != 0 && pend_col0 + r < B) {
            atomicAdd(&cnt[pend_col0 + r], (float)v);
            if (rank_col != nullptr) atomicAdd(&rank_col[pend_col0 + r], v);
          }
        }
        // the diagonal tile and tiles on the lower / right edge mask single elements; all others run the plain loop
        const bool special = (j == sg.ib) || (col0 + MM_TN > B) || (row0 + TM > B);
        mbar_wait(sfull_bar(buf), (p_cnt >> 1) & 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        tmem_ld32(tbuf, va);
        float loc[4] = {0.f, 0.f, 0.f, 0.f};
        int rc4[4] = {0, 0, 0, 0};
        int colc[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          colc[c] = 0;
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          tmem_ld_wait();
          if (c == 0) tmem_ld32(tbuf + 32, vb);
          const int cb = h * 64 + c * 32;                 // first tile column of this chunk
          const float4* cv4 = reinterpret_cast<const float4*>(cv + cb);
          if (kGrad) {
            uint32_t packed[16];
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              const float4 cc = cv4[q >> 2];
              const float x0 = __uint_as_float(v[q]), x1 = __uint_as_float(v[q + 1]);
              const float x2 = __uint_as_float(v[q + 2]), x3 = __uint_as_float(v[q + 3]);
              float g0 = (fmaf(x0, sc, mi) > 0.f ? 1.f : 0.f) + (fmaf(x0, sc, cc.x) > 0.f ? 1.f : 0.f);
              float g1 = (fmaf(x1, sc, mi) > 0.f ? 1.f : 0.f) + (fmaf(x1, sc, cc.y) > 0.f ? 1.f : 0.f);
              float g2 = (fmaf(x2, sc, mi) > 0.f ? 1.f : 0.f) + (fmaf(x2, sc, cc.z) > 0.f ? 1.f : 0.f);
              float g3 = (fmaf(x3, sc, mi) > 0.f ? 1.f : 0.f) + (fmaf(x3, sc, cc.w) > 0.f ? 1.f : 0.f);
              if (j == sg.ib) {                        // G_ii is applied by the finishing pass
                const int cq = cb + q;
                if (cq + 0 == r) g0 = 0.f;
                if (cq + 1 == r) g1 = 0.f;
                if (cq + 2 == r) g2 = 0.f;
                if (cq + 3 == r) g3 = 0.f;
              }
              packed[q >> 1] = pack_pair<kFmt>(g0, g1);
              packed[(q >> 1) + 1] = pack_pair<kFmt>(g2, g3);
            }
            // G(j) columns [64h + 32c, +32) -> packed pairs in TMEM columns [64h + 16c, +16) of the same buffer: inside the
            // first chunk of this warp's own column half, which is in registers by now -- nothing unread is overwritten
            tmem_st16(tbuf + c * 16, packed);
          } else if (!special) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              const float4 cc = cv4[q >> 2];
              const float cj[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {            // four independent accumulation chains
                const float x = __uint_as_float(v[q + e]);
                const float hs = fmaf(x, sc, mi), hc = fmaf(x, sc, cj[e]);   // trainer/loss.py:34, :35
                loc[e] += fmaxf(hs, 0.f) + fmaxf(hc, 0.f);
                rc4[e] += hs > 0.f ? 1 : 0;
                const unsigned m = __ballot_sync(0xffffffffu, hc > 0.f);
                if (lane == q + e) colc[c] = __popc(m);
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int gj = col0 + cb + q;
              const bool valid = gi < B && gj < B && gi != gj;       // :36-40
              const float x = __uint_as_float(v[q]);
              const float hs = fmaf(x, sc, mi), hc = fmaf(x, sc, cv[cb + q]);
              if (valid) {
                loc[q & 3] += fmaxf(hs, 0.f) + fmaxf(hc, 0.f);
                rc4[q & 3] += hs > 0.f ? 1 : 0;
              }
              const unsigned m = __ballot_sync(0xffffffffu, valid && hc > 0.f);
              if (lane == q) colc[c] = __popc(m);
            }
          }
        }
        if (kGrad) tmem_st_wait();
        tc_fence_before();
        mbar_arrive(pfull_bar(buf));
        if (!kGrad) {
          tot += (double)((loc[0] + loc[1]) + (loc[2] + loc[3]));
          rowc += (rc4[0] + rc4[1]) + (rc4[2] + rc4[3]);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (colc[c] != 0) atomicAdd(&ccnt[buf * MM_TN + h * 64 + c * 32 + lane], colc[c]);
          pend_col0 = col0;
        }
      }
      if (!kGrad) {
        if (rowc != 0 && gi < B) {
          atomicAdd(&cnt[gi], (float)rowc);
          if (rank_row != nullptr) atomicAdd(&rank_row[gi], rowc);
        }
      } else {
        // slab accumulator -> dacc (fp32); a segment that covers its whole item stores, partial ones add.  The two warps of a
        // lane quadrant take alternate 32-column chunks.
        mbar_wait(acc_full, seg_iter & 1);
        tc_fence_after();
        const int d0 = sg.sb * SLAB;
        const int slab_w = min(SLAB, dpad - d0);
        float* out = dacc + (int64_t)gi * dpad + d0;
        const bool whole = (sg.j0 == 0 && sg.j1 == ncb);
#pragma unroll 1
        for (int c = h; c < slab_w / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(lane_base + 2 * MM_TN + c * 32, v);
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
              red_add4(out + c * 32 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                       __uint_as_float(v[q + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive(acc_empty);
      }
      ++seg_iter;
    }
    if (!kGrad) {
      named_bar_sync(1, MM_EPI_THREADS);
      if (h == 0 && pend_col0 >= 0) {
        const int v = ccnt[((p_cnt - 1) & 1) * MM_TN + r];
        if (v != 0 && pend_col0 + r < B) {
          atomicAdd(&cnt[pend_col0 + r], (float)v);
          if (rank_col != nullptr) atomicAdd(&rank_col[pend_col0 + r], v);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0 && tot != 0.0) atomicAdd(acc, tot);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// d_i = im_i . s_i in fp32 (one warp per row, the caller's own tensors), and the per-call state zeroed: counts, ranks, the
// loss sum.  kVec: rows are 16-byte aligned (16-byte loads); fp32 rows of the split path may have any alignment.
template <typename T, bool kVec>
__global__ void __launch_bounds__(256) mm_tc_diag_kernel(const T* __restrict__ im, int64_t im_stride, const T* __restrict__ s,
                                                        int64_t s_stride, int B, int D, float* __restrict__ diag,
                                                        float* __restrict__ cnt, int* __restrict__ rank_row,
                                                        int* __restrict__ rank_col, double* __restrict__ acc) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) acc[0] = 0.0;
  if (row >= B) return;
  const T* a = im + (int64_t)row * im_stride;
  const T* b = s + (int64_t)row * s_stride;
  constexpr int E = 16 / (int)sizeof(T);
  float dot = 0.f;
  const int dv = kVec ? (D / E) * E : 0;
  if (kVec) {
    for (int d = lane * E; d < dv; d += 32 * E) {
      const uint4 ua = *reinterpret_cast<const uint4*>(a + d), ub = *reinterpret_cast<const uint4*>(b + d);
      const T* pa = reinterpret_cast<const T*>(&ua);
      const T* pb = reinterpret_cast<const T*>(&ub);
#pragma unroll
      for (int e = 0; e < E; ++e) dot = fmaf(to_float<T>(pa[e]), to_float<T>(pb[e]), dot);
    }
  }
  for (int d = dv + lane; d < D; d += 32) dot = fmaf(to_float<T>(a[d]), to_float<T>(b[d]), dot);
  dot = warp_sum(dot);
  if (lane == 0) {
    diag[row] = dot;
    cnt[row] = 0.f;
    if (rank_row != nullptr) rank_row[row] = 0;
    if (rank_col != nullptr) rank_col[row] = 0;
  }
}

// ---- fp32 inputs: rows staged as fp16 [hi | lo] pairs after a power-of-two scale per tensor ----------------------------------
// max |x| of a tensor as the bit pattern of a non-negative float (ordered like unsigned integers); out[0] zeroed by the caller
__global__ void __launch_bounds__(256) mm_absmax_kernel(const float* __restrict__ x, int64_t stride, int B, int D,
                                                       unsigned int* __restrict__ out) {
  float m = 0.f;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < B; row += gridDim.x * (blockDim.x >> 5))
    for (int d = threadIdx.x & 31; d < D; d += 32) m = fmaxf(m, fabsf(x[(int64_t)row * stride + d]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// x' = x 2^-e with max |x'| in [2^13, 2^14) (the top of the fp16 range, so that hi keeps 11 bits for 27 binades below the
// maximum and hi + lo 22 bits for 16); hi = fp16(x'), lo = fp16(x' - hi); out row = [hi (dpad) | lo (dpad)], zero padded.
// scale_out[0] = 2^e: true score = (staged score) * scale_a * scale_b.
__global__ void __launch_bounds__(256) mm_split_kernel(const float* __restrict__ x, int64_t stride, int B, int D, int dpad,
                                                      const unsigned int* __restrict__ absmax, __half* __restrict__ out,
                                                      float* __restrict__ scale_out) {
  const float mx = __uint_as_float(absmax[0]);
  int e = 0;
  if (mx > 0.f && mx < __int_as_float(0x7f800000)) e = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127 - 13;
  e = max(-126, min(126, e));                     // 2^e and 2^-e both normal floats, built from their bit patterns
  const float down = __int_as_float((127 - e) << 23);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[0] = __int_as_float((127 + e) << 23);
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  __half* o = out + (int64_t)row * 2 * dpad;
  for (int d = lane * 2; d < dpad; d += 64) {
    const float v0 = d < D ? x[(int64_t)row * stride + d] * down : 0.f;
    const float v1 = d + 1 < D ? x[(int64_t)row * stride + d + 1] * down : 0.f;
    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
    *reinterpret_cast<__half2*>(o + d) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(o + dpad + d) =
        __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
  }
}

__global__ void mm_tc_loss_kernel(const double* __restrict__ acc, int B, double* __restrict__ loss) {
  loss[0] = acc[0] / ((double)B * (double)B);                              // trainer/loss.py:41
}

// out[a, :] = (c / B^2) (dacc[a, :] acc_scale - cnt_a Bm[a, :]); one warp per row.  acc_scale: the power-of-two scale of the
// staged gradient operand (split path), Bm the caller's own rows.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) mm_tc_finish_kernel(const float* __restrict__ dacc, int dpad, const T* __restrict__ Bm,
                                                          int64_t b_stride, const float* __restrict__ cnt,
                                                          const double* __restrict__ grad_out, int B, int D,
                                                          TO* __restrict__ out, int64_t out_stride,
                                                          const float* __restrict__ acc_scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  double c = 1.0 / ((double)B * (double)B);
  if (grad_out != nullptr) c *= grad_out[0];
  const float cf = (float)c, ca = cnt[row];
  const float as = acc_scale != nullptr ? acc_scale[0] : 1.0f;
  for (int d = lane; d < D; d += 32) {
    const float v = dacc[(int64_t)row * dpad + d] * as - ca * to_float<T>(Bm[(int64_t)row * b_stride + d]);
    out[(int64_t)row * out_stride + d] = from_float<TO>(cf * v);
  }
}

struct MmPlan { int nk, dpad, ncb, nrb, n_slabs, slots, split; bool resident; size_t smem; };

MmPlan mm_plan(int B, int D, bool split) {
  MmPlan p;
  p.split = split ? 1 : 0;
  p.dpad = (D + KC - 1) / KC * KC;
  p.nk = p.dpad / KC * (split ? 3 : 1);             // score chunks: hi.hi + lo.hi + hi.lo on split rows
  p.ncb = (B + MM_TN - 1) / MM_TN;
  p.nrb = (B + TM - 1) / TM;
  p.n_slabs = (p.dpad + SLAB - 1) / SLAB;
  p.resident = !split && p.nk <= MAX_RES_CHUNKS;    // split rows always stream A beside B
  const size_t a_bytes = p.resident ? (size_t)p.nk * CHUNK_BYTES : 0;
  const size_t avail = kMaxSmem - 1024 - MM_VEC_BYTES - kBarBytes - a_bytes;
  p.slots = std::min((int)(avail / CHUNK_BYTES) & ~1, MAX_SLOTS);
  p.smem = 1024 + a_bytes + (size_t)p.slots * CHUNK_BYTES + MM_VEC_BYTES + kBarBytes;
  return p;
}

template <int kFmt, bool kGrad, int kVar>
int mm_launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const MmPlan& p, int B, float margin, const float* diag,
                float* cnt, double* acc, int* rank_row, int* rank_col, float* dacc, const float* scales, cudaStream_t st) {
  const long long n_units_ll = (long long)p.nrb * (kGrad ? p.n_slabs : 1) * p.ncb;
  if (n_units_ll > 0x7fffffffLL) { set_error("maxmargin: problem too large (%lld work units)", n_units_ll); return CROSSCLR_EINVAL; }
  const int n_units = (int)n_units_ll;
  const int grid = std::min(n_units, sm_count());
  auto kern = mm_tc_kernel<kFmt, kGrad, kVar>;
  // opt-in shared memory: once per device and instantiation (the largest plan's size covers every other)
  static std::atomic<int> smem_set[64];
  const int slot = current_device_slot();
  if (smem_set[slot].load(std::memory_order_relaxed) < (int)p.smem) {
    CC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    smem_set[slot].store((int)p.smem, std::memory_order_relaxed);
  }
  kern<<<grid, MM_THREADS, p.smem, st>>>(ta, tb, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, p.dpad, n_units,
                                          kGrad ? p.n_slabs : 1, p.ncb, p.nk, p.slots, scales);
  return check_launch(kGrad ? "mm_tc_kernel<grad>" : "mm_tc_kernel<fwd>");
}

template <bool kGrad>
int mm_launch(int dtype, const CUtensorMap& ta, const CUtensorMap& tb, const MmPlan& p, int B, float margin, const float* diag,
              float* cnt, double* acc, int* rank_row, int* rank_col, float* dacc, const float* scales, cudaStream_t st) {
  if (p.split)          // fp32 inputs: staged fp16 [hi | lo] rows
    return mm_launch_t<0, kGrad, 2>(ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, scales, st);
  if (dtype == CROSSCLR_BF16)
    return p.resident ? mm_launch_t<1, kGrad, 1>(ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, scales, st)
                      : mm_launch_t<1, kGrad, 0>(ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, scales, st);
  return p.resident ? mm_launch_t<0, kGrad, 1>(ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, scales, st)
                    : mm_launch_t<0, kGrad, 0>(ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, dacc, scales, st);
}

int mm_tmaps(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D, CUtensorMap* ta,
             CUtensorMap* tb) {
  const bool bf16 = dtype == CROSSCLR_BF16;
  int rc = make_tmap_f16(ta, im, (uint64_t)B, (uint64_t)D, TM, true, (uint64_t)im_stride, bf16);
  if (!rc) rc = make_tmap_f16(tb, s, (uint64_t)B, (uint64_t)D, TM, true, (uint64_t)s_stride, bf16);
  return rc;
}

template <typename T, typename TO>
int mm_finish_t(const float* dacc, int dpad, const void* Bm, int64_t bs, const float* cnt, const double* go, int B, int D,
                void* out, int64_t os, const float* acc_scale, cudaStream_t st) {
  mm_tc_finish_kernel<T, TO><<<(B + 7) / 8, 256, 0, st>>>(dacc, dpad, (const T*)Bm, bs, cnt, go, B, D, (TO*)out, os, acc_scale);
  return check_launch("mm_tc_finish_kernel");
}

template <typename T>
int mm_finish(const float* dacc, int dpad, const void* Bm, int64_t bs, const float* cnt, const double* go, int B, int D,
              void* out, int64_t os, int out_dtype, const float* acc_scale, cudaStream_t st) {
  switch (out_dtype) {
    case CROSSCLR_F32: return mm_finish_t<T, float>(dacc, dpad, Bm, bs, cnt, go, B, D, out, os, acc_scale, st);
    case CROSSCLR_F16: return mm_finish_t<T, __half>(dacc, dpad, Bm, bs, cnt, go, B, D, out, os, acc_scale, st);
    case CROSSCLR_BF16: return mm_finish_t<T, __nv_bfloat16>(dacc, dpad, Bm, bs, cnt, go, B, D, out, os, acc_scale, st);
    default: set_error("crossclr_maxmargin_bwd: unsupported output dtype %d", out_dtype); return CROSSCLR_EINVAL;
  }
}

// Stage block of the split path inside the workspace: unsigned absmax[2] | float scale[2] | pad to 256 | half im[B][2 dpad] |
// half s[B][2 dpad]
struct MmStage { unsigned int* absmax; float* scales; __half* a; __half* b; };
MmStage mm_stage(void* stage, int B, int D) {
  const int dpad = (D + KC - 1) / KC * KC;
  MmStage m;
  m.absmax = (unsigned int*)stage;
  m.scales = (float*)((char*)stage + 8);
  m.a = (__half*)((char*)stage + 256);
  m.b = m.a + (size_t)B * 2 * dpad;
  return m;
}

}  // namespace

// 16-bit inputs: the tensor maps address the caller's tensors directly (16-byte aligned rows).  fp32 inputs: staged as fp16
// [hi | lo] rows in the workspace (any alignment).  Below one tile of rows or one K chunk of columns the CUDA-core kernels
// are the better fit.
bool maxmargin_tc_applies(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D) {
  if (B < 2 * TM || D < KC) return false;
  if (dtype == CROSSCLR_F32) return true;
  if (dtype != CROSSCLR_F16 && dtype != CROSSCLR_BF16) return false;
  if ((reinterpret_cast<uintptr_t>(im) | reinterpret_cast<uintptr_t>(s)) & 15u) return false;
  if ((im_stride | s_stride) & 7) return false;
  return true;
}

size_t maxmargin_tc_dacc_bytes(int B, int D) {
  const MmPlan p = mm_plan(B, D, false);
  return (size_t)p.nrb * TM * p.dpad * sizeof(float);
}

size_t maxmargin_tc_stage_bytes(int B, int D, int dtype) {
  if (dtype != CROSSCLR_F32) return 0;
  const int dpad = (D + KC - 1) / KC * KC;
  return 256 + 2 * (size_t)B * 2 * dpad * sizeof(__half);
}

int launch_maxmargin_tc_fwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                            float margin, float* diag, float* cnt, double* acc, void* stage, int* rank_row, int* rank_col,
                            double* loss, cudaStream_t st) {
  CUtensorMap ta, tb;
  const bool split = dtype == CROSSCLR_F32;
  const MmPlan p = mm_plan(B, D, split);
  const float* scales = nullptr;
  int rc;
  if (split) {
    const MmStage m = mm_stage(stage, B, D);
    CC_CHECK_CUDA(cudaMemsetAsync(m.absmax, 0, 8, st));
    const int grid = std::min((B + 7) / 8, 4 * sm_count());
    mm_absmax_kernel<<<grid, 256, 0, st>>>((const float*)im, im_stride, B, D, m.absmax);
    mm_absmax_kernel<<<grid, 256, 0, st>>>((const float*)s, s_stride, B, D, m.absmax + 1);
    mm_split_kernel<<<(B + 7) / 8, 256, 0, st>>>((const float*)im, im_stride, B, D, p.dpad, m.absmax, m.a, m.scales);
    mm_split_kernel<<<(B + 7) / 8, 256, 0, st>>>((const float*)s, s_stride, B, D, p.dpad, m.absmax + 1, m.b, m.scales + 1);
    rc = check_launch("mm_split_kernel");
    g_launches.fetch_add(3, std::memory_order_relaxed);
    if (rc) return rc;
    rc = mm_tmaps(m.a, 2 * p.dpad, m.b, 2 * p.dpad, CROSSCLR_F16, B, 2 * p.dpad, &ta, &tb);
    scales = m.scales;
    if (rc) return rc;
    mm_tc_diag_kernel<float, false><<<(B + 7) / 8, 256, 0, st>>>((const float*)im, im_stride, (const float*)s, s_stride, B, D,
                                                                 diag, cnt, rank_row, rank_col, acc);
  } else {
    rc = mm_tmaps(im, im_stride, s, s_stride, dtype, B, D, &ta, &tb);
    if (rc) return rc;
    if (dtype == CROSSCLR_F16)
      mm_tc_diag_kernel<__half, true><<<(B + 7) / 8, 256, 0, st>>>((const __half*)im, im_stride, (const __half*)s, s_stride, B,
                                                                   D, diag, cnt, rank_row, rank_col, acc);
    else
      mm_tc_diag_kernel<__nv_bfloat16, true><<<(B + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)im, im_stride,
                                                                          (const __nv_bfloat16*)s, s_stride, B, D, diag, cnt,
                                                                          rank_row, rank_col, acc);
  }
  rc = check_launch("mm_tc_diag_kernel");
  if (rc) return rc;
  rc = mm_launch<false>(dtype, ta, tb, p, B, margin, diag, cnt, acc, rank_row, rank_col, nullptr, scales, st);
  if (rc || loss == nullptr) return rc;
  mm_tc_loss_kernel<<<1, 1, 0, st>>>(acc, B, loss);
  return check_launch("mm_tc_loss_kernel");
}

int launch_maxmargin_tc_bwd(const void* im, int64_t im_stride, const void* s, int64_t s_stride, int dtype, int B, int D,
                            float margin, const float* diag, const float* cnt, float* dacc, void* stage,
                            const double* grad_out, void* d_im, int64_t d_im_stride, void* d_s, int64_t d_s_stride,
                            int out_dtype, cudaStream_t st) {
  CUtensorMap ta, tb;
  const bool split = dtype == CROSSCLR_F32;
  const MmPlan p = mm_plan(B, D, split);
  const float* scales = nullptr;
  int rc;
  if (split) {                     // the forward staged the rows (same inputs, same workspace)
    const MmStage m = mm_stage(stage, B, D);
    rc = mm_tmaps(m.a, 2 * p.dpad, m.b, 2 * p.dpad, CROSSCLR_F16, B, 2 * p.dpad, &ta, &tb);
    scales = m.scales;
  } else {
    rc = mm_tmaps(im, im_stride, s, s_stride, dtype, B, D, &ta, &tb);
  }
  if (rc) return rc;
  const size_t dacc_bytes = maxmargin_tc_dacc_bytes(B, D);
  for (int dir = 0; dir < 2 && !rc; ++dir) {
    // dir 0: dL/dim = G s (rows of im against rows of s); dir 1: dL/ds = G^T im -- the same tile formula, operands exchanged
    CC_CHECK_CUDA(cudaMemsetAsync(dacc, 0, dacc_bytes, st));
    rc = mm_launch<true>(dtype, dir ? tb : ta, dir ? ta : tb, p, B, margin, diag, const_cast<float*>(cnt), nullptr, nullptr,
                         nullptr, dacc, scales, st);
    if (rc) break;
    const void* Bm = dir ? im : s;
    const int64_t bs = dir ? im_stride : s_stride;
    void* out = dir ? d_s : d_im;
    const int64_t os = dir ? d_s_stride : d_im_stride;
    const float* as = split ? scales + (dir ? 0 : 1) : nullptr;        // the scale of the gradient operand's tensor
    if (dtype == CROSSCLR_F32) rc = mm_finish<float>(dacc, p.dpad, Bm, bs, cnt, grad_out, B, D, out, os, out_dtype, as, st);
    else if (dtype == CROSSCLR_F16) rc = mm_finish<__half>(dacc, p.dpad, Bm, bs, cnt, grad_out, B, D, out, os, out_dtype, as, st);
    else rc = mm_finish<__nv_bfloat16>(dacc, p.dpad, Bm, bs, cnt, grad_out, B, D, out, os, out_dtype, as, st);
  }
  return rc;
}

}  // namespace crossclr
