// Dataflow backward of libcrossclr_b200 for D <= 512 (sm_100a only).
//
//   dFhat_g = sum_j P_gj Fhat_j,   P_gj = sigma 2^(L_gj) (1/Z_g + 1/Z_j) kappa_gj       (autograd of trainer/loss.py:79-114)
//
// dF[256 rows x D] in fp32 fills all 512 TMEM columns of a CTA pair, so one SM cannot hold the similarity tile S and the
// gradient accumulator at once: the two products live on different SMs.  Here the chip is split into PRODUCER pairs and
// CONSUMER pairs (clusters of two CTAs, each pair one cta_group::2 MMA stream, M = 256) that meet in global memory:
//
//   producer pair s : S[256 x 256] = F_I F_J^T (rows of I resident, J streamed) -> epilogue warps turn the accumulators into
//                     the fp16 probability tile P(I, J) and store it into one of the pair's FLOW_NSLOT pool slots (L2-resident,
//                     4 sub-tiles [row half][column half] of 32 KiB in the un-swizzled UMMA operand layout) -> a publisher warp
//                     appends a descriptor {producer, slot, other block, transposed} to the ring of every consumer of the tile
//   consumer pair c : owns dF of one 256-row block I (or of a column subset of it, `parts` > 1) in TMEM for a whole sweep;
//                     takes descriptors in arrival order (the sum does not care), TMA-loads its two P sub-tiles and the F
//                     boxes of the other block, and accumulates dF_I += P F_J; a release counter hands the slot back
//
// Two schedules:
//   symmetric (one rank owns every row, nb = R / 256 <= 48):  P is symmetric, so only tiles (I, I + d), d = 0 .. nb/2 (cyclic)
//       are formed and every off-diagonal tile is used TWICE: by the consumer of I as P (A operand K-major) and by the
//       consumer of J as P^T (the same bytes read as an MN-major A operand).  S work halves: 12 B^2 D executed instead of 16.
//   row band (multi-rank, or large R): consumers sweep the owned row blocks in waves of n_g; every tile has one consumer.
// The producer : consumer ratio is free (an S tile costs ~5.4k cycles of epilogue-paced work, a dF tile ~4.1k of MMA).
//
// All CTAs must be co-resident (they spin on each other through global memory); the launcher sizes the grid with
// cudaOccupancyMaxActiveClusters and every spin has a poll budget that traps instead of hanging the GPU.
#include "tc_common.cuh"

namespace crossclr {
using namespace ptx;

namespace {

constexpr int FLOW_THREADS = 384;        // warps 0-3: TMA / MMA / TMEM+credit|dispatch / publisher; warps 4-11: epilogue | drain
constexpr int FLOW_NSLOT = 8;            // pool tiles per producer pair
constexpr int FLOW_PBUF = 3;             // 32 KiB P sub-tiles in flight in a consumer CTA's shared memory
constexpr int FLOW_GGROUPS = 4;          // consumer ring: groups of dim/128 [64 j][64 d] boxes
constexpr int FLOW_DQ = 8;               // descriptors in flight inside a consumer CTA
constexpr int FLOW_TN = 256;             // S tile columns
constexpr int FLOW_HDR = 1024 + 8 * 2 * 128 * 8 + 1024;   // barriers | per-warp column coefficient pairs | pad: 18 KiB
constexpr uint16_t kMaskPair = 0x3;

struct FlowParams {
  int n_g, n_s;          // consumer pairs, producer pairs
  int nb;                // 256-column blocks of the stacked matrix
  int nrb;               // owned 256-row blocks
  int parts;             // consumers per row block: consumer (I, part) takes the other-blocks x with x % parts == part
  int cnt;               // tile products per consumer ring = nb / parts
  int sym;               // symmetric schedule
  int d0;                // symmetric schedule: tiles of cyclic distance >= d0 go to the LATE rings (consumed by producer pairs
                         // once their own production is over); nb + 1 = no late rings
  int cnt_main, cnt_late;   // tile products per main ring (= cnt without late rings) / per late ring
  int ds, dw;            // consumers split D into ds slabs of dw columns (D <= 512: 1 x D; else 2 x D/2): physical ring =
                         // logical ring * ds + slab, every tile is consumed once per slab
  int a_res;             // producer keeps its 128 rows of A resident (nk <= 8); else A streams beside B
  int jmajor;            // row-band schedule: a wave's tile list runs column block by column block and producer s takes every
                         // n_s-th tile (all producers and consumers inside the same few column blocks at a time: one HBM read
                         // per block and wave), instead of a contiguous range of one row block's sweep
  int nk;                // K chunks of the similarity product: dim / 64, or 3 dim / 64 for split rows
  int s_stages;          // producer ring depth
  uint32_t* ring;        // [nrb * parts (+ nb late)][cnt] descriptors, zero = not yet published
  uint32_t* tail;        // [nrb * parts (+ nb late)]
  uint32_t* release;     // [n_s][FLOW_NSLOT] consumer releases per pool slot (monotonic)
  uint8_t* pool;         // [n_s][FLOW_NSLOT][4][32 KiB]
  float* dfhat_late;     // [row_count][dim]: partial dF of the late consumers (plain stores; grad_finish adds it)
  unsigned long long* trace;   // debug timeline (CROSSCLR_FLOW_TRACE), or nullptr
  int exp;               // perf experiments (CROSSCLR_FLOW_EXP; results are wrong): 1 consumers run on made-up descriptors,
                         // 2 producers neither wait for credits nor publish, 4 producers idle, 8 consumers idle
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// No release semantics on purpose (a release RMW is a MEMBAR.GPU in front of the RED): what must precede the producer's
// reuse of the slot are this CTA pair's TMA reads of it, and those have completed (mbarrier complete_tx observed by the
// issuing thread) before this instruction is issued.
__device__ __forceinline__ void red_relaxed_gpu_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint32_t kFlowPollBudget = 4000000u;   // x >= ~0.5 us per poll: seconds; a protocol bug traps instead of hanging

__device__ __forceinline__ uint32_t poll_descriptor(const uint32_t* p) {
  uint32_t v = ld_acquire_gpu_u32(p), n = 0;
  while (!(v & 0x80000000u)) {
    __nanosleep(40);
    v = ld_acquire_gpu_u32(p);
    if (++n > kFlowPollBudget) __trap();
  }
  return v;
}
__device__ __forceinline__ void poll_counter_ge(const uint32_t* p, uint32_t want) {
  uint32_t n = 0;
  while ((int32_t)(ld_acquire_gpu_u32(p) - want) < 0) {
    __nanosleep(40);
    if (++n > kFlowPollBudget) __trap();
  }
}

// descriptor word: bit 31 valid | bit 30 transposed | [23, 30) producer | [20, 23) slot | [0, 20) other block
static_assert(FLOW_NSLOT == 8, "the descriptor packs producer * FLOW_NSLOT + slot into bits [20, 30)");
__device__ __forceinline__ uint32_t flow_desc(int prod, int slot, int other, bool trans) {
  return 0x80000000u | (trans ? 0x40000000u : 0u) | ((uint32_t)prod << 23) | ((uint32_t)slot << 20) | (uint32_t)other;
}

struct FlowTile {
  int I, J;              // owned 256-row block (local index), 256-column block (global index)
  int ring_d, ring_t;    // consumer ring of the direct product, of the transposed product (-1: none)
};

// The ordered list of S tiles of producer `s` over all waves (every producer warp walks its own copy).  Symmetric schedule
// with late rings: the producer's chunk of the (row block, cyclic distance) list is walked twice, first the tiles of the
// main rings (d < d0), then its late tiles (d >= d0) -- fewer than FLOW_NSLOT, so the slots they hold until the late
// consumers start are never needed again by this producer.
struct FlowProdWalk {
  int n_g, n_s, nb, parts, cnt, sym, s, nrings, nwaves, d0, jmajor;
  int wave, k, k0, k_end, c0, pass, kstep, nr_wave;
  int I, d;                                   // symmetric schedule: row block, cyclic distance
  __device__ FlowProdWalk(const FlowParams& P, int s_)
      : n_g(P.n_g), n_s(P.n_s), nb(P.nb), parts(P.parts), cnt(P.cnt), sym(P.sym), s(s_), nrings(P.nrb * P.parts),
        wave(-1), k(0), k0(0), k_end(0), c0(0), pass(1), kstep(1), nr_wave(1), I(0), d(0) {
    jmajor = (!P.sym && P.jmajor) ? 1 : 0;
    n_g = P.n_g / P.ds;                        // logical rings per wave (n_g is a multiple of ds)
    nwaves = (nrings + n_g - 1) / n_g;
    d0 = P.d0;
  }
  __device__ __forceinline__ int nd(int i) const { return (nb & 1) ? (nb + 1) / 2 : (i < nb / 2 ? nb / 2 + 1 : nb / 2); }
  __device__ __forceinline__ void seek(int kk) {      // symmetric schedule: (I, d) of linear tile index kk
    I = 0;
    int rem = kk;
    while (I < nb && rem >= nd(I)) { rem -= nd(I); ++I; }
    d = rem;
  }
  __device__ bool next(FlowTile& t) {
    for (;;) {
      if (k >= k_end) {
        if (sym && pass == 0 && d0 <= nb) {             // second pass over the chunk: the late tiles
          pass = 1; k = k0; seek(k);
          continue;
        }
        if (++wave >= nwaves) return false;
        c0 = wave * n_g;
        const int nr = min(n_g, nrings - c0);
        const long long T = sym ? (long long)nb * (nb + 1) / 2 : (long long)nr * cnt;
        if (jmajor) { k0 = k = s; k_end = (int)T; kstep = n_s; nr_wave = nr; }
        else {
          k0 = k = (int)((long long)s * T / n_s);
          k_end = (int)((long long)(s + 1) * T / n_s);
        }
        pass = 0;
        if (sym) seek(k);
        continue;
      }
      if (!sym) {
        const int jdx = jmajor ? k / nr_wave : k % cnt;        // which of the ring's cnt column blocks
        const int c = c0 + (jmajor ? k - jdx * nr_wave : k / cnt);
        t.I = c / parts;
        t.J = (c - t.I * parts) + parts * jdx;
        t.ring_d = c;
        t.ring_t = -1;
        k += kstep;
        return true;
      }
      const int Ic = I, dc = d;
      if (++d == nd(I)) { d = 0; ++I; }
      ++k;
      if ((dc >= d0) != (pass == 1)) continue;          // not this pass's kind
      int J = Ic + dc;
      if (J >= nb) J -= nb;
      t.I = Ic; t.J = J;
      if (dc >= d0) {                                   // late tile: both products go to the rows' late rings (parts == 1)
        t.ring_d = nrings + Ic;
        t.ring_t = nrings + J;
      } else {
        t.ring_d = Ic * parts + (J % parts);
        t.ring_t = dc ? J * parts + (Ic % parts) : -1;
      }
      return true;
    }
  }
};

__global__ void __launch_bounds__(FLOW_THREADS, 1)
bwd_flow_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap64,
                const __grid_constant__ CUtensorMap tmap_p, const uint8_t* __restrict__ feat, Geometry g,
                const float* __restrict__ coef, const float* __restrict__ scal, float* __restrict__ dfhat, FlowParams P) {
  float* const dfhat_late = P.dfhat_late;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  auto full_bar = [&](int s) { return base + 8u * s; };
  auto empty_bar = [&](int s) { return base + 96u + 8u * s; };
  const uint32_t a_full = base + 192u, a_empty = base + 200u;
  auto sfull_bar = [&](int b) { return base + 208u + 8u * b; };
  auto sempty_bar = [&](int b) { return base + 224u + 8u * b; };
  auto staged_bar = [&](int b) { return base + 240u + 8u * b; };     // producer: the CTA's two sub-tiles of a pool tile are stored
  auto pempty_bar = [&](int b) { return base + 304u + 8u * b; };     // producer: pool slot released by its consumer(s)
  auto pub_bar = [&](int b) { return base + 368u + 8u * b; };        // producer leader: the peer's half is fenced
  const uint32_t tmem_slot = base + 432u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + 432);
  // consumer roles: their own barrier region, untouched by the producer roles (a producer pair turns consumer at the end)
  auto cfull_bar = [&](int s) { return base + 640u + 8u * s; };      // dF operand ring (FLOW_GGROUPS)
  auto cempty_bar = [&](int s) { return base + 672u + 8u * s; };
  auto pbfull_bar = [&](int b) { return base + 704u + 8u * b; };     // consumer leader: both CTAs' P sub-tiles landed (TMA)
  auto pbempty_bar = [&](int b) { return base + 728u + 8u * b; };    // consumer: smem P buffer consumed (multicast commit)
  auto dqfull_bar = [&](int b) { return base + 752u + 8u * b; };     // consumer: descriptor b of the local queue is valid
  const uint32_t acc_full = base + 816u, acc_empty = base + 824u;
  volatile uint32_t* dq = reinterpret_cast<volatile uint32_t*>(smem_raw + 832);   // consumer: [FLOW_DQ] descriptor words
  float2* cvw = reinterpret_cast<float2*>(smem_raw + 1024);         // producer: [8 epilogue warps][2 tile parities][128] (q_j, w_j)
  const uint32_t idesc_s = make_idesc_f16(256, 256, 0, 0, 0, 0);

  // debug timeline: producer 0 and consumer 0 (leader CTAs) stamp %globaltimer per tile / event
  auto TR = [&](int role, uint32_t tile, int ev) {
    if (P.trace != nullptr && tile < 64) P.trace[(role * 64 + tile) * 8 + ev] = global_ns();
  };
  unsigned long long* const pair_stamp = P.trace != nullptr ? P.trace + 3 * 64 * 8 + (blockIdx.x >> 1) * 4 : nullptr;
  const uint32_t sub = cluster_ctarank();              // position in the pair = which 128 rows of a 256-row block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x >> 1;
  const bool is_prod = pair >= P.n_g;
  const int prod = pair - P.n_g;
  const uint32_t data = base + (is_prod ? FLOW_HDR : 1024);
  const int nk = P.nk;                   // K chunks of the similarity product (3 dim / 64 for split rows, streamed only)
  const int nkd = g.dim / KC;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap); prefetch_tmap(&tmap64); prefetch_tmap(&tmap_p); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sempty_bar(b), 16); }    // 8 warps x 2 CTAs
    for (int b = 0; b < FLOW_NSLOT; ++b) { mbar_init(staged_bar(b), 8); mbar_init(pempty_bar(b), 1); mbar_init(pub_bar(b), 1); }
    for (int s = 0; s < FLOW_GGROUPS; ++s) { mbar_init(cfull_bar(s), 1); mbar_init(cempty_bar(s), 1); }
    for (int b = 0; b < FLOW_PBUF; ++b) { mbar_init(pbfull_bar(b), 1); mbar_init(pbempty_bar(b), 1); }
    for (int b = 0; b < FLOW_DQ; ++b) mbar_init(dqfull_bar(b), 1);
    mbar_init(acc_full, 1); mbar_init(acc_empty, 16);                                             // 8 warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int nrings_main = P.nrb * P.parts;
  // The consumer roles of a pair for the rings c_first, c_first + c_stride, ... < c_end (late: rings of the late set).
  auto consumer_roles = [&](int c_first, int c_stride, int c_end, bool late) {
    const uint32_t p_tiles = base + 1024;
    const uint32_t ring_base = p_tiles + FLOW_PBUF * PTILE_BYTES;
    const int dw = P.dw;                               // columns of D this consumer accumulates (its slab)
    const int nbox = dw / 128;                         // own [64 j][64 d] boxes per 64-row group: half of each MMA's N
    const uint32_t group_bytes = (uint32_t)nbox * GBOX_BYTES;
    const int nmma = (dw + 255) / 256;                 // MMAs per K = 16 step: N = 256 each (last one 128 if dw % 256)
    const int cnt_c = late ? P.cnt_late : P.cnt_main;
    if (warp == 0) {
      // TMA producer of the dF B operand: rows of the other block as [64 j][64 d] boxes, this CTA's half of every MMA's N
      Ring ring(FLOW_GGROUPS);
      uint32_t i = 0;
      for (int c = c_first; c < c_end; c += c_stride) {
        for (int p = 0; p < cnt_c; ++p, ++i) {
          mbar_wait(dqfull_bar(i % FLOW_DQ), (i / FLOW_DQ) & 1);
          const uint32_t e = dq[i % FLOW_DQ];
          const int other = (int)(e & 0xFFFFFu);
          for (int a = 0; a < 2; ++a) {
            for (int kh = 0; kh < 2; ++kh) {                          // 64-row halves of the K = 128 rows of a sub-tile
              mbar_wait(cempty_bar(ring.stage), ring.phase ^ 1);
              if (elect_one()) {
                const uint32_t st = ring_base + ring.stage * group_bytes;
                const uint32_t full_ldr = mapa_cluster(cfull_bar(ring.stage), 0);
                if (sub == 0) mbar_arrive_expect_tx(cfull_bar(ring.stage), 2 * group_bytes);
                for (int q = 0; q < nbox; ++q) {
                  const int m = q >> 1;                               // MMA index; its N = nm, this CTA's half = nm / 2
                  const int nm = min(256, dw - m * 256);
                  const int d0 = (c % P.ds) * dw + m * 256 + (int)sub * (nm >> 1) + (q & 1) * 64;
                  tma_load_2d_2sm(st + q * GBOX_BYTES, &tmap64, full_ldr, d0, other * FLOW_TN + a * TM + kh * 64);
                }
              }
              __syncwarp();
              ring.advance();
            }
          }
        }
      }
    } else if (warp == 2) {
      // dispatcher + P loader: next descriptor of the ring (arrival order) -> local queue -> this CTA's two sub-tiles
      uint32_t i = 0, th = 0;
      unsigned long long waited = 0;
      for (int c = c_first; c < c_end; c += c_stride) {
        for (int p = 0; p < cnt_c; ++p, ++i) {
          if (lane == 0) {
            const unsigned long long tw0 = (pair_stamp && !late) ? global_ns() : 0;
            if ((pair == 0 && !late) && sub == 0) TR(1, i, 0);
            const uint32_t e = (P.exp & 1) ? flow_desc((int)((i + pair) % (uint32_t)P.n_s), (int)(i & 7), (int)((i + pair) % (uint32_t)P.nb), (i & 1) != 0)
                                           : poll_descriptor(P.ring + (size_t)c * P.cnt + p);
            if ((pair == 0 && !late) && sub == 0) TR(1, i, 1);
            if (pair_stamp && !late && sub == 0) {
              const unsigned long long tw1 = global_ns();
              if (i == 0) pair_stamp[0] = tw1; else waited += tw1 - tw0;
              pair_stamp[2] = waited;
            }
            dq[i % FLOW_DQ] = e;
            mbar_arrive(dqfull_bar(i % FLOW_DQ));
            fence_proxy_async_global();                              // the tile's generic-proxy stores -> our TMA reads
            const bool trans = (e & 0x40000000u) != 0;
            const int tile = (int)((e >> 20) & 0x3FFu);              // producer * FLOW_NSLOT + slot
            for (int a = 0; a < 2; ++a, ++th) {
              const uint32_t pb = th % FLOW_PBUF, puse = th / FLOW_PBUF;
              mbar_wait(pbempty_bar(pb), (puse & 1) ^ 1);
              if (sub == 0) mbar_arrive_expect_tx(pbfull_bar(pb), 2 * PTILE_BYTES);
              const int st_idx = trans ? (a * 2 + (int)sub) : ((int)sub * 2 + a);
              tma_load_2d_2sm(p_tiles + pb * PTILE_BYTES, &tmap_p, mapa_cluster(pbfull_bar(pb), 0), 0,
                              (tile * 4 + st_idx) * 256);
            }
            if ((pair == 0 && !late) && sub == 0) TR(1, i, 2);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1 && sub == 0) {
      Ring ring(FLOW_GGROUPS);
      uint32_t i = 0, th = 0, seg_iter = 0;
      for (int c = c_first; c < c_end; c += c_stride) {
        mbar_wait_cluster(acc_empty, (seg_iter & 1) ^ 1);   // both CTAs' drain warps have emptied the previous sweep
        tc_fence_after();
        for (int p = 0; p < cnt_c; ++p, ++i) {
          mbar_wait(dqfull_bar(i % FLOW_DQ), (i / FLOW_DQ) & 1);
          const uint32_t e = dq[i % FLOW_DQ];
          const bool trans = (e & 0x40000000u) != 0;
          for (int a = 0; a < 2; ++a, ++th) {
            const uint32_t pb = th % FLOW_PBUF, puse = th / FLOW_PBUF;
            if ((pair == 0 && !late) && lane == 0 && a == 0) TR(1, i, 3);
            mbar_wait(pbfull_bar(pb), puse & 1);              // both CTAs' P sub-tiles have landed in shared memory (TMA)
            if ((pair == 0 && !late) && lane == 0) TR(1, i, 4 + a);
            tc_fence_after();
            if (a == 1 && elect_one())                        // both sub-tiles of the pool tile are on chip: slot released
              red_relaxed_gpu_add_u32(P.release + ((e >> 20) & 0x3FFu), 1u);
            const uint32_t p_tile = p_tiles + pb * PTILE_BYTES;
            for (int kh = 0; kh < 2; ++kh) {
              mbar_wait(cfull_bar(ring.stage), ring.phase);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t st = ring_base + ring.stage * group_bytes;
                for (int m = 0; m < nmma; ++m) {
                  const int nm = min(256, dw - m * 256);
                  const uint32_t idesc = make_idesc_f16(256, nm, 0, 0, trans ? 1 : 0, 1);
#pragma unroll
                  for (int k16 = 0; k16 < 4; ++k16) {
                    // direct:     A = P[:, 64 kh + 16 k16 .. +16) K-major: column chunks 8 kh + 2 k16 and the next one (2048 B
                    //             apart), 8-row groups 128 B apart
                    // transposed: A = P^T: M = the sub-tile's 128 columns (16-byte chunks 2048 B apart), K = its rows
                    //             64 kh + 16 k16 .. +16 (8-row groups 128 B apart): the same bytes read MN-major
                    // B = Fhat rows 64 kh + 16 k16 .. +16 of this CTA's boxes 2m, 2m + 1, MN-major
                    const uint64_t ad = trans ? make_smem_desc_nosw(p_tile + (kh * 8 + k16 * 2) * 128, 128, TM * 16)
                                              : make_smem_desc_nosw(p_tile + (kh * 8 + k16 * 2) * (TM * 16), TM * 16, 128);
                    const uint64_t bd = make_smem_desc_sw128(st + 2 * m * GBOX_BYTES + k16 * 2048, 1024, GBOX_BYTES);
                    umma_ss_2sm(tmem_base + m * 256, ad, bd, idesc, (p > 0 || a > 0 || kh > 0 || k16 > 0) ? 1u : 0u);
                  }
                }
                umma_commit_2sm(cempty_bar(ring.stage), kMaskPair);
              }
              __syncwarp();
              ring.advance();
            }
            if (elect_one()) umma_commit_2sm(pbempty_bar(pb), kMaskPair);
            if ((pair == 0 && !late) && lane == 0 && a == 1) TR(1, i, 6);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit_2sm(acc_full, kMaskPair);
        __syncwarp();
        ++seg_iter;
        if (pair_stamp && lane == 0) pair_stamp[late ? 3 : 1] = global_ns();
      }
    } else if (warp >= EPI_WARP0) {
      // drain: dF of the sweep -> dfhat (fp32, still scaled by sigma; grad_finish divides it out)
      const int quadw = warp & 3, wg = (warp - EPI_WARP0) >> 2;     // wg: which 256-column half of D
      const int r = quadw * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quadw * 32) << 16);
      const uint32_t acc_empty_ldr = mapa_cluster(acc_empty, 0);
      const bool whole = P.parts == 1;      // one main (and one late) consumer per row block: plain stores
      uint32_t seg_iter = 0;
      for (int c = c_first; c < c_end; c += c_stride, ++seg_iter) {
        const int cl = c / P.ds;                         // logical ring, slab c % ds
        const int I = late ? cl - nrings_main : cl / P.parts;
        mbar_wait(acc_full, seg_iter & 1);
        tc_fence_after();
        float* out = (late ? dfhat_late : dfhat) + ((int64_t)I * FLOW_TN + (int)sub * TM + r) * g.dim + (c % P.ds) * dw;
        const int cc_end = min(dw, wg * 256 + 256) / 32;
#pragma unroll 1
        for (int cc = wg * 8; cc < cc_end; ++cc) {
          uint32_t v[32];
          tmem_ld32(lane_base + cc * 32, v);
          tmem_ld_wait();
          if (whole) {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              *reinterpret_cast<float4*>(out + cc * 32 + q) =
                  make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                              __uint_as_float(v[q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + cc * 32 + q), "f"(__uint_as_float(v[q])),
                           "f"(__uint_as_float(v[q + 1])), "f"(__uint_as_float(v[q + 2])), "f"(__uint_as_float(v[q + 3]))
                           : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_ldr);   // TMEM hand-back: no memory to order (see tc_ptx.cuh)
      }
    }
  };

  if (is_prod && !(P.exp & 4)) {
    // =========================================================================== producer pair
    const bool a_res = P.a_res != 0;
    const uint32_t a_region = data;
    const uint32_t ring_base = data + (a_res ? nk * CHUNK_BYTES : 0);
    const uint32_t stage_bytes = a_res ? CHUNK_BYTES : 2 * CHUNK_BYTES;      // [own A chunk, streamed] + own half of the B chunk
    if (warp == 0) {
      // TMA: own 128 rows of A (resident per row block, or streamed), own half of each 256-row B chunk; bytes counted on the leader
      Ring ring(P.s_stages);
      int cur_I = -1;
      uint32_t a_cnt = 0;
      const uint32_t a_full_ldr = mapa_cluster(a_full, 0);
      FlowProdWalk walk(P, prod);
      FlowTile tl;
      while (walk.next(tl)) {
        const int row0 = g.row_begin + tl.I * FLOW_TN + (int)sub * TM;
        if (a_res && tl.I != cur_I) {
          mbar_wait(a_empty, (a_cnt & 1) ^ 1);
          if (elect_one()) {
            if (sub == 0) mbar_arrive_expect_tx(a_full, 2 * nk * CHUNK_BYTES);
            for (int kc = 0; kc < nk; ++kc) tma_load_2d_2sm(a_region + kc * CHUNK_BYTES, &tmap, a_full_ldr, kc * KC, row0);
          }
          __syncwarp();
          cur_I = tl.I; ++a_cnt;
        }
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(empty_bar(ring.stage), ring.phase ^ 1);
          if (elect_one()) {
            uint32_t st = ring_base + ring.stage * stage_bytes;
            const uint32_t full_ldr = mapa_cluster(full_bar(ring.stage), 0);
            if (sub == 0) mbar_arrive_expect_tx(full_bar(ring.stage), 2 * stage_bytes);
            if (!a_res) { tma_load_2d_2sm(st, &tmap, full_ldr, a_kcol(kc, nkd), row0); st += CHUNK_BYTES; }
            tma_load_2d_2sm(st, &tmap, full_ldr, b_kcol(kc, nkd), tl.J * FLOW_TN + (int)sub * TM);
          }
          __syncwarp();
          ring.advance();
        }
      }
    } else if (warp == 1 && sub == 0) {
      // MMA issuer (leader): S tile t -> TMEM buffer t & 1 of both CTAs
      Ring ring(P.s_stages);
      int cur_I = -1;
      uint32_t a_cnt = 0, t = 0;
      FlowProdWalk walk(P, prod);
      FlowTile tl, nx;
      bool have = walk.next(tl);
      while (have) {
        const bool have_next = walk.next(nx);
        if (a_res && tl.I != cur_I) {
          mbar_wait(a_full, a_cnt & 1);
          cur_I = tl.I; ++a_cnt;
        }
        const uint32_t buf = t & 1;
        if (prod == 0 && lane == 0) TR(0, t, 0);
        mbar_wait_cluster(sempty_bar(buf), ((t >> 1) & 1) ^ 1);
        if (prod == 0 && lane == 0) TR(0, t, 1);
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(full_bar(ring.stage), ring.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = ring_base + ring.stage * stage_bytes;
            const uint64_t ad = kmajor_desc(a_res ? a_region + kc * CHUNK_BYTES : st);
            const uint64_t bd = kmajor_desc(a_res ? st : st + CHUNK_BYTES);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_ss_2sm(tmem_base + buf * FLOW_TN, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s,
                          (kc == 0 && k == 0) ? 0u : 1u);
            umma_commit_2sm(empty_bar(ring.stage), kMaskPair);
          }
          __syncwarp();
          ring.advance();
        }
        if (elect_one()) {
          umma_commit_2sm(sfull_bar(buf), kMaskPair);
          if (a_res && (!have_next || nx.I != tl.I)) umma_commit_2sm(a_empty, kMaskPair);   // last tile of this row block
        }
        if (prod == 0 && lane == 0) TR(0, t, 2);
        __syncwarp();
        ++t;
        tl = nx;
        have = have_next;
      }
    } else if (warp == 1 && sub == 1) {
      // credit poller (the non-leader's MMA warp has nothing else to do): a pool slot may be overwritten once every consumer
      // of the tile it held has loaded it; arrives on both CTAs' pempty barriers
      uint32_t cum[FLOW_NSLOT] = {0, 0, 0, 0, 0, 0, 0, 0};
      uint32_t t = 0;
      FlowProdWalk walk(P, prod);
      FlowTile tl;
      const uint32_t pempty_ldr = mapa_cluster(pempty_bar(0), 0);
      while (walk.next(tl)) {
        const uint32_t slot = t % FLOW_NSLOT;
        if (t >= FLOW_NSLOT) {
          if (lane == 0) {
            if (!(P.exp & 2)) poll_counter_ge(P.release + prod * FLOW_NSLOT + slot, cum[slot]);
            mbar_arrive(pempty_bar(slot));
            mbar_arrive_cluster_relaxed(pempty_ldr + 8u * slot);   // ordered behind the acquire load of the counter
          }
          __syncwarp();
        }
        cum[slot] += ((tl.ring_t >= 0) ? 2u : 1u) * (uint32_t)P.ds;
        ++t;
      }
    } else if (warp == 2 || warp == 3) {
      // two publishers (tiles of even / odd index): the CTA's half of pool tile t is in global memory -> fence it GPU-wide;
      // the leader, once both halves are fenced, appends the tile's descriptor to its consumers' rings
      uint32_t t = 0;
      FlowProdWalk walk(P, prod);
      FlowTile tl;
      const uint32_t pub_ldr = mapa_cluster(pub_bar(0), 0);
      while (walk.next(tl)) {
        if ((t & 1) == (uint32_t)(warp - 2)) {
          const uint32_t slot = t % FLOW_NSLOT, use = t / FLOW_NSLOT;
          mbar_wait(staged_bar(slot), use & 1);
          if (lane == 0) {
            if (prod == 0 && sub == 0) TR(0, t, 4);
            if (sub == 1) {
              fence_acq_rel_gpu();                         // this CTA's half (observed through staged_bar) is visible GPU-wide
              mbar_arrive_cluster(pub_ldr + 8u * slot);
            } else if (!(P.exp & 2)) {
              // ring positions first (two independent atomics in flight), then the peer's half, then ONE gpu-scope fence that
              // is cumulative over both halves (own: staged_bar, peer: cluster-scope acquire), then plain flag stores
              uint32_t pos_d[2], pos_t[2] = {0u, 0u};
              for (int h = 0; h < P.ds; ++h) {
                pos_d[h] = atomicAdd(P.tail + tl.ring_d * P.ds + h, 1u);
                if (tl.ring_t >= 0) pos_t[h] = atomicAdd(P.tail + tl.ring_t * P.ds + h, 1u);
              }
              mbar_wait_cluster(pub_bar(slot), use & 1);
              fence_acq_rel_gpu();
              for (int h = 0; h < P.ds; ++h) {
                st_relaxed_gpu_u32(P.ring + (size_t)(tl.ring_d * P.ds + h) * P.cnt + pos_d[h], flow_desc(prod, (int)slot, tl.J, false));
                if (tl.ring_t >= 0)
                  st_relaxed_gpu_u32(P.ring + (size_t)(tl.ring_t * P.ds + h) * P.cnt + pos_t[h], flow_desc(prod, (int)slot, tl.I, true));
              }
              if (prod == 0) TR(0, t, 5);
            }
          }
          __syncwarp();
        }
        ++t;
      }
    } else if (warp >= EPI_WARP0) {
      // Eight epilogue warps, two per TMEM lane quadrant: warp (quadrant q, half wg) owns rows [32 q, +32) x columns
      // [128 wg, +128) of every S tile = sub-tile [sub][wg] of the pool tile, as four 32-column chunks run through a
      // two-register-buffer software pipeline that does not stop at tile boundaries (chunk 0 of tile t + 1 leaves TMEM
      // under chunk 3 of tile t).
      const int quadw = warp & 3, wg = (warp - EPI_WARP0) >> 2;
      const int r = quadw * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quadw * 32) << 16) + wg * TM;
      const float sigma = scal[0];
      const float nshift = -g.shift;
      const uint32_t sempty_ldr0 = mapa_cluster(sempty_bar(0), 0), sempty_ldr1 = mapa_cluster(sempty_bar(1), 0);
      // un-swizzled K-major operand layout of a sub-tile: [16 column chunks of 8][128 rows][16 bytes]
      uint8_t* const p_pool = P.pool + ((size_t)prod * FLOW_NSLOT * 4 + sub * 2 + wg) * PTILE_BYTES + (size_t)r * 16;
      float* const cv_warp = reinterpret_cast<float*>(cvw + (warp - EPI_WARP0) * 256);   // [tile parity][64 column pairs][4]
      const float* const coef_col = coef + 2 * (int64_t)(wg * TM + lane);
      FlowProdWalk walk(P, prod);
      FlowTile tl, nx;
      bool have = walk.next(tl);
      int cur_I = -1, gi = 0;
      BlockSeg bi{0, 0};
      float iz_i = 0.f, q_i = 1.f, nshift_i = nshift;
      float izj[4] = {0.f, 0.f, 0.f, 0.f}, qj[4] = {1.f, 1.f, 1.f, 1.f};
      uint32_t va[32], vb[32];
      if (have) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          izj[q] = coef_col[2 * ((int64_t)tl.J * FLOW_TN + 32 * q)];
          qj[q] = row_q(feat, g, (int64_t)tl.J * FLOW_TN + wg * TM + lane + 32 * q);
        }
        mbar_wait(sfull_bar(0), 0);                                  // prologue of the pipeline: chunk 0 of tile 0
        tc_fence_after();
        tmem_ld32(lane_base, va);
      }
      uint32_t t = 0;
      const bool tracer = prod == 0 && sub == 0 && r == 0 && wg == 0;
      while (have) {
        if (tracer) TR(2, t, 0);
        const bool have_next = walk.next(nx);
        if (tl.I != cur_I) {
          cur_I = tl.I;
          const int row0 = g.row_begin + tl.I * FLOW_TN + (int)sub * TM;
          gi = row0 + r;
          bi = block_seg(row0, g.bseg);
          iz_i = coef[2 * (int64_t)gi];
          q_i = row_q(feat, g, gi);
          nshift_i = nshift + log2f(q_i);                // the tile is P q_g q_j -- symmetric, so that its transposed read is
        }                                                // valid; q_g rides in the exponent, grad_finish divides it out
        const int jrow0 = tl.J * FLOW_TN + wg * TM;
        const BlockSeg bj = block_seg(jrow0, g.bseg);
        const bool same_mod = (bj.mod == bi.mod);
        const bool diag_tile = (bj.samp0 == bi.samp0);
        const float k = (same_mod ? g.k_intra : g.k_inter) * q_i;   // row part of the logit scale
        const float ks = (same_mod ? g.w : 1.0f) * sigma;
        float* cv = cv_warp + (t & 1) * (2 * TM);                  // per column pair: (q_j, q_j+1, w_j, w_j+1)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = 32 * q + lane;
          cv[(col >> 1) * 4 + (col & 1)] = qj[q];                    // q_j
          cv[(col >> 1) * 4 + (col & 1) + 2] = izj[q] * ks * qj[q];  // w_j = q_j kappa sigma / Z_j
        }
        if (have_next) {                                             // next tile's column coefficients, a whole tile ahead
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            izj[q] = coef_col[2 * ((int64_t)nx.J * FLOW_TN + 32 * q)];
            qj[q] = row_q(feat, g, (int64_t)nx.J * FLOW_TN + wg * TM + lane + 32 * q);
          }
        }
        __syncwarp();
        if (tracer) TR(2, t, 1);
        const float a_i = iz_i * ks;
        const uint32_t buf = t & 1;
        const uint32_t tb = lane_base + buf * FLOW_TN;
        const uint32_t tb_next = lane_base + (buf ^ 1) * FLOW_TN;
        const uint32_t slot = t % FLOW_NSLOT, use = t / FLOW_NSLOT;
        if (prod == 0 && sub == 0 && r == 0 && wg == 0) TR(0, t, 3);
        const unsigned long long tw0 = (pair_stamp && sub == 0 && r == 0 && wg == 0) ? global_ns() : 0;
        mbar_wait(pempty_bar(slot), (use & 1) ^ 1);                  // every consumer of the slot's previous tile has loaded it
        if (prod == 0 && sub == 0 && r == 0 && wg == 0) TR(0, t, 6);
        if (pair_stamp && sub == 0 && r == 0 && wg == 0) {
          const unsigned long long tw1 = global_ns();
          if (t == 0) pair_stamp[0] = tw1;
          pair_stamp[2] += tw1 - tw0;
          pair_stamp[1] = tw1;
        }
        uint8_t* const prow = p_pool + (size_t)slot * (4 * PTILE_BYTES);
        auto p_chunk = [&](const uint32_t (&v)[32], int c) {         // c: 32-column chunk of the warp's half (0..3)
          uint32_t packed[16];
          const float4* cv4 = reinterpret_cast<const float4*>(cv) + c * 16;
          const uint64_t k2 = pack2(k, k), ns2 = pack2(nshift_i, nshift_i), a2 = pack2(a_i, a_i);
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float4 cc = cv4[q >> 1];                         // (q_j, q_j+1, w_j, w_j+1) of columns q, q + 1
            // P~ = 2^x (1/Z_g + 1/Z_j) kappa sigma q_g q_j,  x = (f_g . f_j) (k q_g) q_j - shift     (q_g dF_g = sum_j P~ f_j)
            // on packed fp32 pairs: FMUL2, FFMA2 -> 2 x ex2 -> FFMA2, FMUL2 -> one fp16 pair
            const uint64_t q01 = pack2(cc.x, cc.y);
            const uint64_t x01 = ffma2(fmul2(pack2(__uint_as_float(v[q + 0]), __uint_as_float(v[q + 1])), k2), q01, ns2);
            float x0, x1, e0, e1;
            unpack2(x01, x0, x1);
            unpack2(fmul2(pack2(fast_exp2(x0), fast_exp2(x1)), ffma2(a2, q01, pack2(cc.z, cc.w))), e0, e1);
            __half2 h0 = __floats2half2_rn(e0, e1);
            packed[q >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          }
          if (diag_tile && c == quadw) {                              // warp-uniform: this chunk holds column r of the sub-tile
            const int pi = (r & 31) >> 1;                             // same-sample pair: handled in grad_finish
            const uint32_t keep = (r & 1) ? 0x0000ffffu : 0xffff0000u;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == pi) packed[i] &= keep;
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            st_global_v4(prow + (c * 4 + ch) * (TM * 16), packed[ch * 4 + 0], packed[ch * 4 + 1], packed[ch * 4 + 2],
                         packed[ch * 4 + 3]);
        };
        tmem_ld_wait();                          // chunk 0 (va)
        if (tracer) TR(2, t, 2);
        tmem_ld32(tb + 32, vb);
        p_chunk(va, 0);
        if (tracer) TR(2, t, 3);
        tmem_ld_wait();                          // chunk 1 (vb)
        tmem_ld32(tb + 64, va);
        p_chunk(vb, 1);
        if (tracer) TR(2, t, 4);
        tmem_ld_wait();                          // chunk 2 (va)
        tmem_ld32(tb + 96, vb);
        p_chunk(va, 2);
        if (tracer) TR(2, t, 5);
        tmem_ld_wait();                          // chunk 3 (vb): this warp's part of the S tile is in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(buf ? sempty_ldr1 : sempty_ldr0);   // TMEM hand-back: relaxed (a release
                                                                              // arrive is a MEMBAR behind this warp's P stores)
        if (tracer) TR(2, t, 6);
        bool prefetched = false;
        if (have_next && mbar_try_wait(sfull_bar(buf ^ 1), ((t + 1) >> 1) & 1)) {
          tc_fence_after();
          tmem_ld32(tb_next, va);
          prefetched = true;
        }
        p_chunk(vb, 3);
        // The arrive releases these generic-proxy stores (cta scope); the publisher's gpu-scope fence is cumulative over
        // them, and the consumer fences generic -> async before its TMA load.
        __syncwarp();
        if (lane == 0) mbar_arrive(staged_bar(slot));
        if (prod == 0 && sub == 0 && r == 0 && wg == 0) TR(0, t, 7);
        if (have_next && !prefetched) {
          mbar_wait(sfull_bar(buf ^ 1), ((t + 1) >> 1) & 1);
          tc_fence_after();
          tmem_ld32(tb_next, va);
        }
        if (tracer) TR(2, t, 7);
        ++t;
        tl = nx;
        have = have_next;
      }
    }
  }

  if (!is_prod && !(P.exp & 8)) consumer_roles(pair, P.n_g, nrings_main * P.ds, false);
  if (is_prod && P.d0 <= P.nb && !(P.exp & 4)) {
    // production is over for this pair: every tile it made is stored and published, its TMEM and shared memory are free.
    // It now consumes the late rings of rows prod, prod + n_s, ... (symmetric schedule only).
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    consumer_roles(nrings_main + prod, P.n_s, nrings_main + P.nb, true);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // neither CTA leaves while its peer may still touch its shared memory / TMEM pair
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// resident clusters (pairs) of bwd_flow_kernel on the current device; queried once per device
int flow_pairs_resident() {
  static std::atomic<int> cache[64];
  const int dev = current_device_slot();
  int c = cache[dev].load(std::memory_order_relaxed);
  if (c == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (sm_count() / 2));
    cfg.blockDim = dim3(FLOW_THREADS);
    cfg.dynamicSmemBytes = kMaxSmem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    cudaFuncSetAttribute(bwd_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (cudaOccupancyMaxActiveClusters(&n, bwd_flow_kernel, &cfg) != cudaSuccess || n < 0) {
      (void)cudaGetLastError();
      n = 0;
    }
    c = std::min(n, sm_count() / 2);
    cache[dev].store(c > 0 ? c : -1, std::memory_order_relaxed);
  }
  return c > 0 ? c : 0;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

struct FlowPlan {
  bool ok;
  int n_g, n_s, parts, cnt, sym, nb, nrb, ds;
  int d0, cnt_main, cnt_late;       // late rings of the symmetric schedule (d0 = nb + 1: none)
};

// Schedule for a problem, or ok = false when the dataflow kernel does not apply (the caller falls back).
FlowPlan flow_plan(const Geometry& g, int pairs) {
  FlowPlan f = {};
  static const int variant = env_int("CROSSCLR_BWD_VARIANT", 0);
  if (variant != 0 && variant != 4) return f;
  if (g.dim > 1024 || g.dim % 128 != 0 || (g.dim > 512 && g.dim % 256 != 0) || g.rows % FLOW_TN != 0 ||
      g.row_count % FLOW_TN != 0 || pairs < 32)
    return f;
  if (variant == 0 && g.rows < 2048) return f;                      // tiny problems: latency-bound either way
  f.ds = g.dim > 512 ? 2 : 1;                                       // consumers take 512-column (or D/2) slabs of D
  f.nb = g.rows / FLOW_TN;
  f.nrb = g.row_count / FLOW_TN;
  if (f.nb >= (1 << 20) || pairs > 127) return f;
  static const int sym_max = env_int("CROSSCLR_FLOW_SYM_MAX", 48);
  static const int force_ns = env_int("CROSSCLR_FLOW_NS", 0);
  static const int force_ng = env_int("CROSSCLR_FLOW_NG", 0);
  f.sym = (g.row_count == g.rows && g.row_begin == 0 && f.nb * f.ds <= std::min(sym_max, pairs - 8)) ? 1 : 0;
  // consumers per row block: as many as fit in ~3/5 of the pairs (symmetric) or in the 32 consumer pairs of a wave
  const int g_cap = force_ng > 0 ? force_ng : (f.sym ? std::max(f.nb * f.ds, pairs * 3 / 5) : (pairs * 32) / 74);
  f.parts = 1;
  while (f.nrb * f.parts * 2 * f.ds <= g_cap && f.nb % (f.parts * 2) == 0) f.parts *= 2;
  f.cnt = f.nb / f.parts;
  const int nrings = f.nrb * f.parts * f.ds;                         // physical rings: (row block, part, slab of D)
  if (f.sym) {
    f.n_g = nrings;                                                  // single wave by construction (a multiple of ds)
    const long long tiles = (long long)f.nb * (f.nb + 1) / 2;
    // every remaining pair produces: measured at B = 4096, D = 512 (32 consumer pairs): 25 / 32 / 42 producer pairs ->
    // 129 / 127 / 121 us; the consumers are MMA-bound from their first tile on, more producers get them there sooner
    f.n_s = (int)std::max<long long>(1, std::min<long long>(tiles, pairs - f.n_g));
  } else {
    f.n_g = std::min(nrings, std::max(f.ds, std::min(g_cap, pairs - 1) / f.ds * f.ds));   // a multiple of ds
    f.n_s = pairs - f.n_g;
    const long long tiles = (long long)f.nrb * f.nb;
    f.n_s = (int)std::max<long long>(1, std::min<long long>(f.n_s, tiles));
  }
  if (force_ns > 0) f.n_s = std::max(1, std::min(force_ns, pairs - f.n_g));
  // Late rings (symmetric schedule, one consumer per row block): the producers finish long before the consumers (an S tile
  // costs ~1.3 dF tile products, and every S tile feeds two products), so the products of the tiles with the largest
  // cyclic distance, d >= d0, are left to the producer pairs, which turn consumer when their production is over.  Balance:
  // (nb - L) products on a main consumer  =  production (tiles / n_s S tiles) + L products on a late consumer.
  f.d0 = f.nb + 1; f.cnt_main = f.cnt; f.cnt_late = 0;
  static const int force_d0 = env_int("CROSSCLR_FLOW_D0", -1);
  if (f.sym && f.parts == 1 && f.ds == 1 && force_d0 != 0) {
    const double tiles = 0.5 * f.nb * (f.nb + 1);
    const double l_target = 0.5 * (f.nb - 1.3 * tiles / f.n_s);
    int d0 = (f.nb & 1) ? (f.nb - 1) / 2 + 1 - (int)(l_target / 2.0) : (int)std::ceil(f.nb / 2.0 - (l_target - 1.0) / 2.0);
    if (force_d0 > 0) d0 = force_d0;
    const int nd_max = (f.nb & 1) ? (f.nb + 1) / 2 : f.nb / 2 + 1;      // distances 0 .. nd_max - 1
    const int late_tiles_per_row = nd_max - d0;
    // a producer's chunk crosses at most one row end (chunk < row), so it holds at most late_tiles_per_row late tiles in
    // its FLOW_NSLOT slots while it goes on producing: keep two slots free
    if (d0 >= 2 && late_tiles_per_row >= 1 && late_tiles_per_row <= FLOW_NSLOT - 3 && tiles / f.n_s < nd_max - 1 && f.n_s >= f.nb / 2) {
      f.d0 = d0;
      f.cnt_main = 2 * d0 - 1;
      f.cnt_late = f.nb - f.cnt_main;
    }
  }
  f.ok = f.n_g >= 1 && f.n_s >= 1 && f.n_g + f.n_s <= pairs;
  return f;
}

// CROSSCLR_FLOW_TRACE=<file>: producer 0 / consumer 0 stamp %globaltimer per tile and event; the launch then synchronises
// and writes the table (perf debugging only)
unsigned long long* flow_trace_buffer(cudaStream_t st) {
  static unsigned long long* trace = nullptr;
  static const bool want = getenv("CROSSCLR_FLOW_TRACE") != nullptr;
  if (!want) return nullptr;
  if (trace == nullptr) cudaMalloc(&trace, (3 * 64 * 8 + 128 * 4) * 8);
  cudaMemsetAsync(trace, 0, (3 * 64 * 8 + 128 * 4) * 8, st);
  return trace;
}
void flow_trace_dump(unsigned long long* trace, cudaStream_t st, const FlowPlan& f) {
  if (trace == nullptr) return;
  static unsigned long long host[3 * 64 * 8 + 128 * 4];
  cudaStreamSynchronize(st);
  cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost);
  FILE* fp = fopen(getenv("CROSSCLR_FLOW_TRACE"), "w");
  if (!fp) return;
  unsigned long long t0 = ~0ull;
  for (int i = 0; i < 3 * 64 * 8; ++i) if (host[i] && host[i] < t0) t0 = host[i];
  fprintf(fp, "# n_g %d n_s %d parts %d cnt %d sym %d nb %d nrb %d ds %d d0 %d cnt_main %d cnt_late %d\n", f.n_g, f.n_s, f.parts, f.cnt, f.sym, f.nb, f.nrb, f.ds, f.d0, f.cnt_main, f.cnt_late);
  fprintf(fp, "# role 0 = producer 0: mma_wait_sempty mma_start mma_issued epi_start staged_seen published epi_slot_ok epi_done\n");
  fprintf(fp, "# role 1 = consumer 0: poll_start desc_seen p_issued mma_wait_p p0_landed p1_landed mma_issued -   (ns since first stamp)\n");
  fprintf(fp, "# role 2 = producer 0, epilogue warp 0: loop_top coef_staged chunk0_in_regs chunk0_done chunk1_done chunk2_done tmem_released next_tile_ready\n");
  for (int role = 0; role < 3; ++role)
    for (int t = 0; t < 64; ++t) {
      fprintf(fp, "%d %2d", role, t);
      for (int e = 0; e < 8; ++e) {
        const unsigned long long x = host[(role * 64 + t) * 8 + e];
        fprintf(fp, " %8lld", x ? (long long)(x - t0) : -1ll);
      }
      fprintf(fp, "\n");
    }
  fprintf(fp, "# per pair (consumers first, then producers): first_ns last_ns waited_ns late_consumer_done_ns\n");
  for (int p = 0; p < f.n_g + f.n_s && p < 128; ++p) {
    const unsigned long long* q = host + 3 * 64 * 8 + p * 4;
    fprintf(fp, "P %3d %s %8lld %8lld %8lld %4lld\n", p, p < f.n_g ? "cons" : "prod", q[0] ? (long long)(q[0] - t0) : -1ll,
            q[1] ? (long long)(q[1] - t0) : -1ll, (long long)q[2], q[3] ? (long long)(q[3] - t0) : -1ll);
  }
  fclose(fp);
}

// upper bound of the control words of any plan for this problem: rings (main + late, or two slabs), tails (parts <= 64),
// releases
size_t flow_control_bytes(int nrb, int nb, int pairs) {
  const size_t words = 2 * (size_t)nrb * nb + 2 * (size_t)nrb * 64 + nb + (size_t)pairs * FLOW_NSLOT;
  return (words * 4 + 1023) & ~(size_t)1023;
}

}  // namespace

// scratch for the dataflow backward of a problem with `row_count` owned of `rows` stacked rows: control words + tile pool
size_t bwd_flow_scratch_bytes(int rows, int row_count) {
  const int pairs = sm_count() / 2;
  return flow_control_bytes((row_count + FLOW_TN - 1) / FLOW_TN, (rows + FLOW_TN - 1) / FLOW_TN, pairs) +
         (size_t)pairs * FLOW_NSLOT * 4 * PTILE_BYTES;
}

bool bwd_flow_applies(const Geometry& g) {
  const int pairs = flow_pairs_resident();
  return flow_plan(g, pairs).ok;
}

int launch_bwd_flow(const Geometry& g, const void* feat, const float* coef, const float* scal, float* dfhat,
                    float* dfhat_late, bool* used_late, void* scratch, cudaStream_t st) {
  const int pairs = flow_pairs_resident();
  const FlowPlan f = flow_plan(g, pairs);
  if (!f.ok) { set_error("crossclr_bwd: the dataflow kernel does not apply to this problem"); return CROSSCLR_EINVAL; }
  CUtensorMap tmap, tmap64, tmap_p;
  int rc = CC_FEAT_TMAP(&tmap, feat, g, TM);
  if (rc) return rc;
  rc = CC_FEAT_TMAP(&tmap64, feat, g, 64);
  if (rc) return rc;
  const size_t ctl_bytes = flow_control_bytes(f.nrb, f.nb, sm_count() / 2);
  uint8_t* pool = (uint8_t*)scratch + ctl_bytes;
  rc = make_tmap_f16(&tmap_p, pool, (uint64_t)f.n_s * FLOW_NSLOT * 4 * 256, 64, 256, /*swizzle=*/false);
  if (rc) return rc;
  FlowParams P;
  P.n_g = f.n_g; P.n_s = f.n_s; P.nb = f.nb; P.nrb = f.nrb; P.parts = f.parts; P.cnt = f.cnt; P.sym = f.sym;
  P.d0 = f.d0; P.cnt_main = f.cnt_main; P.cnt_late = f.cnt_late;
  const bool late_on = f.d0 <= f.nb;
  const size_t n_rings = (size_t)f.nrb * f.parts * f.ds + (late_on ? f.nb : 0);
  P.nk = s_chunks(g);
  P.ds = f.ds; P.dw = g.dim / f.ds;
  // Column-block-major producer order for row bands whose stacked matrix does not sit comfortably in the 126 MB L2: with
  // contiguous sweeps every producer and consumer walks the column blocks at its own pace and each block comes from HBM
  // once per tile (c5: ~1 TB per launch); interleaved, all of them are inside the same few column blocks.  Measured on B200
  // (scripts/gpu_flow_jmajor.sh): c3 3.42 -> 3.17 ms, c4 34.6 -> 31.8 ms, c5 298 -> 236 ms; B = 16384, D = 512 (38 MB) unchanged.
  // CROSSCLR_FLOW_JMAJOR = 0 / 1 forces the order.
  static const int force_jmajor = env_int("CROSSCLR_FLOW_JMAJOR", -1);
  P.jmajor = f.sym ? 0 : (force_jmajor >= 0 ? (force_jmajor != 0) : ((size_t)g.rows * g.pitch * 2 > ((size_t)48 << 20) ? 1 : 0));
  P.a_res = (P.nk <= MAX_RES_CHUNKS && !g.split && !P.jmajor) ? 1 : 0;
  P.s_stages = P.a_res ? std::min((int)((kMaxSmem - FLOW_HDR - (size_t)P.nk * CHUNK_BYTES) / CHUNK_BYTES), MAX_SLOTS)
                       : std::min((int)((kMaxSmem - FLOW_HDR) / (2 * CHUNK_BYTES)), MAX_SLOTS);
  P.ring = (uint32_t*)scratch;
  P.tail = P.ring + n_rings * f.cnt;
  P.release = P.tail + n_rings;
  P.pool = pool;
  P.dfhat_late = dfhat_late;
  *used_late = late_on;
  const size_t ctl_used = ((n_rings * f.cnt + n_rings + (size_t)f.n_s * FLOW_NSLOT) * 4 + 255) & ~(size_t)255;
  P.trace = flow_trace_buffer(st);
  static const int exp_flags = env_int("CROSSCLR_FLOW_EXP", 0);
  P.exp = exp_flags;
  CC_CHECK_CUDA(cudaFuncSetAttribute(bwd_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
  TimedLaunch timed(CROSSCLR_K_BWD, st);               // after the host-side preparation: the bracket holds device work only
  CC_CHECK_CUDA(cudaMemsetAsync(scratch, 0, ctl_used, st));
  if (f.parts > 1) CC_CHECK_CUDA(cudaMemsetAsync(dfhat, 0, (size_t)g.row_count * g.dim * sizeof(float), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (f.n_g + f.n_s));
  cfg.blockDim = dim3(FLOW_THREADS);
  cfg.dynamicSmemBytes = kMaxSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  CC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, bwd_flow_kernel, tmap, tmap64, tmap_p, (const uint8_t*)feat, g, coef, scal, dfhat, P));
  flow_trace_dump(P.trace, st, f);
  return check_launch("bwd_flow_kernel");
}

}  // namespace crossclr
