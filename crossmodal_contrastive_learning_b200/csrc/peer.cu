// Single-node exchange of row shards through NVLink peer memory (SURVEY.md section 8 rows e / f2): the replacement of the two
// NCCL all-gathers of the row-sharded step.  Every rank owns buffers of one size that all ranks of the node have mapped
// (CUDA IPC); crossclr_peer_exchange is ONE kernel that
//   0. (entry barrier, optional) signals every peer and waits until every peer has ENTERED the same exchange -- so that all
//      of them are done reading what the stores below overwrite (the stacked rows the previous step's backward reads),
//   1. stores the caller's slice [offset, offset + bytes) of the local buffer into the same place of every peer's buffer
//      (16-byte stores over NVLink; the slice is already in place locally: `pack` wrote it there), and
//   2. in its last block -- after every block has fenced its stores at system scope -- signals every peer (flag = epoch) and
//      waits for every peer's signal: when the kernel ends, all slices of all ranks are in this rank's buffer.
// No host synchronisation, no separate barrier launch, capturable in a CUDA graph (the epoch lives in device memory).
// Ranks must call it in lock-step (the same sequence of exchanges on every rank), like a collective.
#include "common.cuh"

namespace crossclr {

namespace {

constexpr int kMaxPeers = 16;
struct PeerArgs {
  uint8_t* base[kMaxPeers];
  uint32_t* flags[kMaxPeers];     // flags[p][q]: the last epoch at which rank q signalled rank p (exit barrier);
};                                //   flags[p][n + q]: likewise for the entry barrier

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void wait_flag(const uint32_t* f, uint32_t epoch) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
    if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a rank that never arrives is a protocol bug, not a hang
  }
}

// state: [0] exit epoch, [1] block ticket, [2] entry epoch
__global__ void __launch_bounds__(256) peer_exchange_kernel(PeerArgs pa, int n, int rank, size_t offset, size_t nvec,
                                                            int entry_barrier, uint32_t* __restrict__ state) {
  if (entry_barrier) {
    const uint32_t e_in = state[2] + 1;                // rewritten only by the last block, after every block's ticket
    const int p = threadIdx.x;
    if (p < n && p != rank) {
      if (blockIdx.x == 0) st_release_sys(pa.flags[p] + n + rank, e_in);
      wait_flag(pa.flags[rank] + n + p, e_in);
    }
    __syncthreads();
  }
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(pa.base[rank] + offset);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = src[i];
    for (int k = 1; k < n; ++k) {                      // peers in a rank-staggered order: no two ranks start on the same link
      const int p = rank + k < n ? rank + k : rank + k - n;
      reinterpret_cast<uint4*>(pa.base[p] + offset)[i] = v;
    }
  }
  __threadfence_system();                              // this thread's peer stores are performed before the ticket below
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&state[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  const uint32_t epoch = state[0] + 1;
  const int p = threadIdx.x;
  if (p < n && p != rank) {
    __threadfence_system();
    st_release_sys(pa.flags[p] + rank, epoch);
    wait_flag(pa.flags[rank] + p, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    state[0] = epoch; state[1] = 0;
    if (entry_barrier) state[2] = state[2] + 1;
  }
}

}  // namespace

int launch_peer_exchange(void* const* bases, uint32_t* const* flags, int n, int rank, size_t offset, size_t bytes,
                         int entry_barrier, uint32_t* state, cudaStream_t st) {
  PeerArgs pa;
  for (int i = 0; i < kMaxPeers; ++i) { pa.base[i] = nullptr; pa.flags[i] = nullptr; }
  for (int i = 0; i < n; ++i) { pa.base[i] = (uint8_t*)bases[i]; pa.flags[i] = flags[i]; }
  const size_t nvec = bytes / 16;
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t want = (nvec + 255) / 256;
  const int grid = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sms * 4));
  peer_exchange_kernel<<<grid, 256, 0, st>>>(pa, n, rank, offset, nvec, entry_barrier, state);
  return check_launch("peer_exchange_kernel");
}

int peer_max_ranks() { return kMaxPeers; }

}  // namespace crossclr
