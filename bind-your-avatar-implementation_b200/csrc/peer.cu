// Exchanges of the sequence-parallel step over NVLink peer memory (SURVEY.md §8e "target form"): no NCCL on the data path.
//
// Every rank maps every other rank's exchange buffers (symmetric allocations, torch.distributed._symmetric_memory hands
// out the peer pointers).  Two mechanisms:
//   * PUSH: the producing kernel itself stores into the consumer's memory — the fused QKV GEMM's epilogue TMA-stores each
//     destination rank's [q|k|v] column block straight into that rank's attention input (gemm_tcgen05.cu: peer_out), the
//     attention epilogue stores each query row straight into the K-blocked A operand of its owner's out-projection
//     (fa_tcgen05.cu: out_peers).  The exchange overlaps the math tile by tile; what remains is a barrier.
//   * COPY: one kernel moves strided segments between a local buffer and the peers' buffers in the layout the consumer
//     wants — the router's spatial-attention exchanges (the position gather / scatter permutations are folded into the
//     segment strides), the face queries of a rank's router positions, the routing result.  Pushed (posted writes) by
//     default; a pull variant (remote reads) exists for comparison.
// Ordering: a device-side epoch barrier (bya_peer_barrier): each rank release-stores its epoch into every peer's flag
// array after its producer kernel has finished (stream order) and acquire-spins on its own flags.  Buffers need no
// double buffering: between a consumer's read and the next overwrite of the same buffer there is always another barrier
// of the same stream-ordered sequence.
#include "common.cuh"
#include "../../include/bya.h"

#include <type_traits>

namespace bya {

BYA_DEVICE void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
BYA_DEVICE int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One block of >= n_ranks threads.  counter: this rank's epoch (device memory, advanced here: CUDA-graph replays keep
// counting).  peer_flags[r]: rank r's flag array [n_ranks] (peer-mapped); local flag slot s is written by rank s.
__global__ void peer_barrier_kernel(int* counter, int* const* peer_flags, int my_rank, int n_ranks) {
  __shared__ int epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = *counter + 1;
    *counter = epoch_s;
  }
  __syncthreads();
  const int epoch = epoch_s;
  const int r = threadIdx.x;
  if (r < n_ranks) {
    __threadfence_system();
    st_release_sys(peer_flags[r] + my_rank, epoch);           // "rank my_rank has reached epoch" -> rank r
    const int* mine = peer_flags[my_rank] + r;               // ... and wait for rank r to reach it
    unsigned spins = 0;
    while (ld_acquire_sys(mine) - epoch < 0) {
      if (++spins > (1u << 27)) __trap();                    // a lost peer must not hang the box (seconds, then abort)
      __nanosleep(64);
    }
  }
}

struct PullSeg {      // one strided 3-D block copied from a peer buffer into a local buffer; all sizes in BYTES
  long long src_off, dst_off;
  long long src_outer_stride, dst_outer_stride, src_row_stride, dst_row_stride;
  int peer, outer, rows, row_bytes;
};
static_assert(sizeof(PullSeg) == sizeof(ByaPullSeg), "ByaPullSeg layout");

// grid (blocks per segment, n_segs).  PUSH: src = local + src_off, dst = peers[seg.peer] + dst_off (posted NVLink writes:
// latency-tolerant, the default); PULL: src = peers[seg.peer] + src_off, dst = local + dst_off (remote reads need
// ~2 MB in flight to cover the NVLink round trip: measured 16 GB/s with 4 096 threads of one load each).
// VEC = bytes per access (16 / 8 / 4); four independent accesses in flight per thread.
template <int VEC, bool PUSH>
__global__ void __launch_bounds__(256) peer_copy_kernel(const PullSeg* __restrict__ segs, char* const* __restrict__ peers,
                                                        char* __restrict__ local) {
  using V = typename std::conditional<VEC == 16, uint4, typename std::conditional<VEC == 8, uint2, uint32_t>::type>::type;
  const PullSeg s = segs[blockIdx.y];
  const char* src = (PUSH ? local : peers[s.peer]) + s.src_off;
  char* d = (PUSH ? peers[s.peer] : local) + s.dst_off;
  const int per_row = s.row_bytes / VEC;
  const long long total = (long long)s.outer * s.rows * per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto addr = [&](long long i, const char*& sp, char*& dp) {
    const int v = int(i % per_row);
    const long long rr = i / per_row;
    const int r = int(rr % s.rows), o = int(rr / s.rows);
    sp = src + o * s.src_outer_stride + r * s.src_row_stride + (long long)v * VEC;
    dp = d + o * s.dst_outer_stride + r * s.dst_row_stride + (long long)v * VEC;
  };
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {
    const char* sp[4];
    char* dp[4];
    V val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) addr(i + u * stride, sp[u], dp[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) val[u] = *reinterpret_cast<const V*>(sp[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<V*>(dp[u]) = val[u];
  }
  for (; i < total; i += stride) {
    const char* sp;
    char* dp;
    addr(i, sp, dp);
    *reinterpret_cast<V*>(dp) = *reinterpret_cast<const V*>(sp);
  }
}

}  // namespace bya

using namespace bya;

extern "C" int bya_peer_barrier(void* stream, int* counter, int* const* peer_flags, int my_rank, int n_ranks) {
  if (!counter || !peer_flags || n_ranks < 1 || n_ranks > BYA_MAX_PEERS || my_rank < 0 || my_rank >= n_ranks) return BYA_ERR_SHAPE;
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(counter, peer_flags, my_rank, n_ranks);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_peer_copy(void* stream, const ByaPullSeg* segs, int n_segs, void* const* peers, void* local, int push,
                             int vec_bytes, int blocks_per_seg) {
  if (!segs || !peers || !local || n_segs <= 0 || n_segs > 65535 || blocks_per_seg <= 0) return BYA_ERR_SHAPE;
  dim3 grid(blocks_per_seg, n_segs);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const PullSeg* sg = reinterpret_cast<const PullSeg*>(segs);
  char* const* pp = reinterpret_cast<char* const*>(peers);
  char* lc = static_cast<char*>(local);
#define BYA_PEER_COPY(V)                                                        \
  if (push) peer_copy_kernel<V, true><<<grid, 256, 0, s>>>(sg, pp, lc);         \
  else peer_copy_kernel<V, false><<<grid, 256, 0, s>>>(sg, pp, lc)
  switch (vec_bytes) {
    case 16: BYA_PEER_COPY(16); break;
    case 8: BYA_PEER_COPY(8); break;
    case 4: BYA_PEER_COPY(4); break;
    default: return BYA_ERR_ALIGN;
  }
#undef BYA_PEER_COPY
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
