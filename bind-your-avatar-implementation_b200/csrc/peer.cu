// Exchanges of the sequence-parallel step over NVLink peer memory (SURVEY.md §8e "target form"): no NCCL on the data path.
//
// Every rank maps every other rank's exchange buffers (symmetric allocations, torch.distributed._symmetric_memory hands
// out the peer pointers).  Two mechanisms:
//   * PUSH: the producing kernel itself stores into the consumer's memory — the fused QKV GEMM's epilogue TMA-stores each
//     destination rank's [q|k|v] column block straight into that rank's attention input (gemm_tcgen05.cu: peer_out), the
//     attention epilogue stores each query row straight into the K-blocked A operand of its owner's out-projection
//     (fa_tcgen05.cu: out_peers).  The exchange overlaps the math tile by tile; what remains is a barrier.
//   * PULL: one kernel gathers strided segments from the peers' buffers into the local layout the consumer wants — the
//     router's spatial-attention exchanges (the position gather / scatter permutations are folded into the segment
//     strides), the face queries of a rank's router positions, the routing result.
// Ordering: a device-side epoch barrier (bya_peer_barrier): each rank release-stores its epoch into every peer's flag
// array after its producer kernel has finished (stream order) and acquire-spins on its own flags.  Buffers need no
// double buffering: between a consumer's read and the next overwrite of the same buffer there is always another barrier
// of the same stream-ordered sequence.
#include "common.cuh"
#include "../../include/bya.h"

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

// grid (blocks per segment, n_segs).  src bases: peer_src[seg.peer]; VEC = bytes per access (16 / 8 / 4).
template <int VEC>
__global__ void __launch_bounds__(256) peer_pull_kernel(const PullSeg* __restrict__ segs, const char* const* __restrict__ peer_src,
                                                        char* __restrict__ dst) {
  const PullSeg s = segs[blockIdx.y];
  const char* src = peer_src[s.peer] + s.src_off;
  char* d = dst + s.dst_off;
  const int per_row = s.row_bytes / VEC;
  const long long total = (long long)s.outer * s.rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = int(i % per_row);
    const long long rr = i / per_row;
    const int r = int(rr % s.rows), o = int(rr / s.rows);
    const char* sp = src + o * s.src_outer_stride + r * s.src_row_stride + (long long)v * VEC;
    char* dp = d + o * s.dst_outer_stride + r * s.dst_row_stride + (long long)v * VEC;
    if (VEC == 16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
    else if (VEC == 8) *reinterpret_cast<uint2*>(dp) = *reinterpret_cast<const uint2*>(sp);
    else *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
  }
}

}  // namespace bya

using namespace bya;

extern "C" int bya_peer_barrier(void* stream, int* counter, int* const* peer_flags, int my_rank, int n_ranks) {
  if (!counter || !peer_flags || n_ranks < 1 || n_ranks > BYA_MAX_PEERS || my_rank < 0 || my_rank >= n_ranks) return BYA_ERR_SHAPE;
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(counter, peer_flags, my_rank, n_ranks);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_peer_pull(void* stream, const ByaPullSeg* segs, int n_segs, const void* const* peer_src, void* dst, int vec_bytes,
                             int blocks_per_seg) {
  if (!segs || !peer_src || !dst || n_segs <= 0 || n_segs > 65535 || blocks_per_seg <= 0) return BYA_ERR_SHAPE;
  dim3 grid(blocks_per_seg, n_segs);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const PullSeg* sg = reinterpret_cast<const PullSeg*>(segs);
  const char* const* ps = reinterpret_cast<const char* const*>(peer_src);
  switch (vec_bytes) {
    case 16: peer_pull_kernel<16><<<grid, 256, 0, s>>>(sg, ps, static_cast<char*>(dst)); break;
    case 8: peer_pull_kernel<8><<<grid, 256, 0, s>>>(sg, ps, static_cast<char*>(dst)); break;
    case 4: peer_pull_kernel<4><<<grid, 256, 0, s>>>(sg, ps, static_cast<char*>(dst)); break;
    default: return BYA_ERR_ALIGN;
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
