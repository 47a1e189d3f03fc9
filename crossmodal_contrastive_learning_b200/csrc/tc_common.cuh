// Pieces shared by the tensor-core translation units (tc_kernels.cu, flow_kernels.cu): tile constants, ring bookkeeping,
// operand descriptors of the stacked-matrix boxes, and the host-side tensor-map / device queries.
#pragma once

#include "common.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace crossclr {
using namespace ptx;

namespace {

constexpr int TM = 128;                 // tile rows (TMEM lanes)
constexpr int KC = 64;                  // K chunk: 64 fp16 = one 128-byte swizzle row
constexpr int CHUNK_BYTES = TM * KC * 2;   // 16 KiB: a [128 rows][64 elems] box
constexpr int MAX_RES_CHUNKS = 8;       // A row block stays resident for D <= 512
constexpr int FWD_TN = 256;             // forward similarity tile columns (one N=256 MMA)
constexpr int BWD_TN = 128;             // backward similarity / probability tile columns
constexpr int SLAB = 256;               // dFhat columns per backward work item (TMEM columns)
constexpr int MAX_SLOTS = 12;
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 128;


__device__ __forceinline__ uint64_t kmajor_desc(uint32_t addr) { return make_smem_desc_sw128(addr, 1024, 0); }

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ __forceinline__ void advance() {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

// The 4 K=16 MMAs of one 64-wide K chunk into a 128 x N accumulator (A, B K-major 128-byte-swizzled boxes).
__device__ __forceinline__ void issue_s_chunk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc,
                                              bool first_chunk) {
  const uint64_t ad = kmajor_desc(a_addr), bd = kmajor_desc(b_addr);
#pragma unroll
  for (int k = 0; k < KC / 16; ++k)
    umma_ss(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (first_chunk && k == 0) ? 0u : 1u);
}

// segment bookkeeping of a 128-aligned block of stacked rows starting at `r0`
struct BlockSeg {
  int mod;      // modality of the block (0 video / 1 text)
  int samp0;    // sample index of its first row
};
__device__ __forceinline__ BlockSeg block_seg(int r0, int bseg) {
  const int seg = r0 / bseg;
  return BlockSeg{seg & 1, (seg >> 1) * bseg + (r0 - seg * bseg)};
}

// residual scale q_j of stacked row j (normalised row = q_j * stored row; include/crossclr_b200.h), first word of the row tail
__device__ __forceinline__ float row_q(const uint8_t* __restrict__ feat, const Geometry& g, int64_t j) {
  return __ldg(reinterpret_cast<const float*>(feat + (j * g.pitch + (g.pitch - CROSSCLR_ROW_TAIL)) * 2));
}

// K chunks of the similarity product and the stacked-row column each one reads.  Plain rows: chunk kc of A and of B is
// column 64 kc.  Split rows [hi | lo]: K = 3 dim, A = [hi | lo | hi], B = [hi | hi | lo] (hi.hi + lo.hi + hi.lo); with
// nk = dim / 64 both maps below reduce to 64 kc in the plain case (kc < nk).
__host__ __device__ __forceinline__ int s_chunks(const Geometry& g) { return (g.dim / KC) * (g.split ? 3 : 1); }
__device__ __forceinline__ int a_kcol(int kc, int nk) { return (kc < 2 * nk ? kc : kc - 2 * nk) * KC; }
__device__ __forceinline__ int b_kcol(int kc, int nk) { return (kc < nk ? kc : kc - nk) * KC; }

// stacked-matrix tensor map of a geometry: {64, box_rows} boxes of the [rows][dim] operand inside the pitched rows
#define CC_FEAT_TMAP(m, feat, g, box_rows) \
  make_tmap_f16(m, feat, (uint64_t)(g).rows, (uint64_t)((g).pitch - CROSSCLR_ROW_TAIL), box_rows, true, (uint64_t)(g).pitch, false)

constexpr int kBarBytes = 8 * (2 * MAX_SLOTS) + 128;   // mbarrier block of the forward / single-CTA backward kernels
constexpr int PTILE_BYTES = TM * 128 * 2;   // a [128 rows][128 columns] fp16 probability tile: 32 KiB
constexpr int GBOX_BYTES = 64 * KC * 2;     // a [64 rows][64 columns] box of the dF operand: 8 KiB
constexpr size_t kMaxSmem = 232448;         // 227 KiB opt-in dynamic shared memory per CTA

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 16-bit matrix [rows][cols] (row-major, row pitch `pitch` elements; 0 = cols) tiled into {64 cols, box_rows} boxes,
// 128-byte swizzle (or none)
inline int make_tmap_f16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool swizzle = true,
                         uint64_t pitch = 0, bool bf16 = false) {
  // The driver-API encode needs a current context on THIS thread.  torch's autograd worker threads only get
  // one lazily (first runtime call), so bind the primary context here; cudaFree(nullptr) is the documented no-op
  // that does it.
  static thread_local int bound_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_dev) {
    cudaFree(nullptr);
    bound_dev = dev;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available"); return CROSSCLR_ECUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {(pitch ? pitch : cols) * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return CROSSCLR_ECUDA; }
  return CROSSCLR_OK;
}

// SM count of the CURRENT device (cached per device: one process may drive several)
inline int sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

}  // namespace
}  // namespace crossclr
