// Routed per-character cross-attention with 32 keys per (character, frame), and the router's tiny strided
// self-attentions.   sm_100a build; warp-level mma.sync tiles (the problem is 32 keys wide — far below a tcgen05
// tile — and HBM/latency bound: q is read once, the blended result written once).
//
//   out[n, h*d : (h+1)*d] = sum_c w[n,c] * softmax_k(scale * q[n,h,:] . K[g(c,n)][h][k][:]) @ V[g(c,n)][h]
//   g(c, n) = c * kv_frames + n / tokens_per_frame
// i.e. the per-character attention AND the routed blend of models/transformer.py:821-822 / :925-926 in one pass:
// each token gathers only the K/V of its own frame, and the per-character results are never materialised
// (blend-then-project).  Replaces the attention core of
// PerceiverCrossAttention.forward (models/router.py:256-273; softmax in fp32) and of the audio cross-attention
// (models/audio_model.py:253-256 -> diffusers AttnProcessor2_0).
#include "common.cuh"
#include "../../include/bya.h"

#include <cstdlib>

namespace bya {

BYA_DEVICE void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

BYA_DEVICE void ldmatrix_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

constexpr int XA_WARPS = 8;
constexpr int XA_TOK = XA_WARPS * 16;  // tokens per block

// K  : [G][H][32][D]   (keys x head-dim, head-dim contiguous)
// Vt : [G][H][D][32]   (V transposed: head-dim x keys, keys contiguous)  — both prepared once per generation
template <int D, int C>
__global__ void __launch_bounds__(XA_WARPS * 32, D == 64 ? 3 : 2)
xattn_kv32_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ K,
                  const __nv_bfloat16* __restrict__ Vt, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                  int ldo, int heads, int tokens_per_frame, int kv_frames, float scale_log2, long long tok_begin,
                  long long tok_count) {
  constexpr int KP = D + 8;    // padded row of the K tile (bank-conflict-free fragment loads)
  constexpr int VP = 32 + 8;   // padded row of the V^T tile
  extern __shared__ __align__(16) uint8_t xa_smem[];
  typedef __nv_bfloat16 (*KTile)[32][KP];
  typedef __nv_bfloat16 (*VTile)[D][VP];
  // two staging buffers: K / V^T of head h+1 stream in (cp.async) while head h is being computed
  constexpr size_t kBufBytes = size_t(C) * 32 * KP * 2 + size_t(C) * D * VP * 2;
  auto kbuf = [&](int i) { return reinterpret_cast<KTile>(xa_smem + i * kBufBytes); };
  auto vbuf = [&](int i) { return reinterpret_cast<VTile>(xa_smem + i * kBufBytes + size_t(C) * 32 * KP * 2); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int frame = blockIdx.y;
  const int tok0 = blockIdx.x * XA_TOK + warp * 16;               // within the frame
  const int r0 = tok0 + g, r1 = tok0 + g + 8;                     // this thread's two rows (within the frame)
  // global token index -> local row of q / w / out (this rank owns tokens [tok_begin, tok_begin + tok_count))
  const long long blk_lo = (long long)frame * tokens_per_frame + blockIdx.x * XA_TOK;
  if (blk_lo >= tok_begin + tok_count || blk_lo + XA_TOK <= tok_begin) return;  // no owned token in this block
  const long long gl0 = (long long)frame * tokens_per_frame + r0 - tok_begin;
  const long long gl1 = (long long)frame * tokens_per_frame + r1 - tok_begin;
  const bool ok0 = r0 < tokens_per_frame && gl0 >= 0 && gl0 < tok_count;
  const bool ok1 = r1 < tokens_per_frame && gl1 >= 0 && gl1 < tok_count;
  const size_t n0 = size_t(ok0 ? gl0 : 0), n1 = size_t(ok1 ? gl1 : 0);

  float wt0[C], wt1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    wt0[c] = (w && ok0) ? w[n0 * C + c] : 1.f;
    wt1[c] = (w && ok1) ? w[n1 * C + c] : 1.f;
  }

  // blockIdx.z = head group: one block per SM (143 blocks at the c2 size) left the kernel latency-bound (224 us
  // against a 33 us HBM roofline); 4 head groups give every SM 3-4 resident blocks to overlap staging and math
  const int hpg = (heads + gridDim.z - 1) / gridDim.z;
  const int h_begin = blockIdx.z * hpg, h_end = min(heads, int(blockIdx.z + 1) * hpg);
  // ---- stage K_h and V^T_h of every character (this block's frame) into buffer `b` with 16-byte cp.async
  auto stage = [&](int h, int b) {
    KTile dK = kbuf(b);
    VTile dV = vbuf(b);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const size_t grp = (size_t(c) * kv_frames + frame) * heads + h;
      const __nv_bfloat16* gk = K + grp * 32 * D;
      for (int i = threadIdx.x; i < 32 * D / 8; i += blockDim.x) {
        const int row = i / (D / 8), cc = i % (D / 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&dK[c][row][cc * 8])), "l"(gk + i * 8) : "memory");
      }
      const __nv_bfloat16* gv = Vt + grp * 32 * D;
      for (int i = threadIdx.x; i < D * 32 / 8; i += blockDim.x) {
        const int row = i / 4, cc = i % 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&dV[c][row][cc * 8])), "l"(gv + i * 8) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (h_begin < h_end) stage(h_begin, 0);
  for (int h = h_begin; h < h_end; ++h) {
    const int cur = (h - h_begin) & 1;
    if (h + 1 < h_end) {
      stage(h + 1, cur ^ 1);     // buffer cur^1 was last read in iteration h-1: every warp passed that iteration's trailing barrier
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    KTile sK = kbuf(cur);
    VTile sV = vbuf(cur);

    // ---- Q fragments of this warp's 16 tokens for head h (straight from global; zero for rows past the frame)
    uint32_t qa[D / 16][4];
    const __nv_bfloat16* q0 = q + n0 * ldq + h * D;
    const __nv_bfloat16* q1 = q + n1 * ldq + h * D;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      qa[kk][0] = ok0 ? *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * t) : 0u;
      qa[kk][1] = ok1 ? *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 2 * t) : 0u;
      qa[kk][2] = ok0 ? *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8 + 2 * t) : 0u;
      qa[kk][3] = ok1 ? *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8 + 2 * t) : 0u;
    }

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

    // B fragments come from shared memory with ldmatrix.x4: four 8x8 blocks per instruction, i.e. both halves of TWO
    // k-steps — a quarter of the shared-memory instructions of per-fragment 32-bit loads (ncu, round 1: this kernel sat
    // on the shared-memory pipe: mio / short-scoreboard stalls)
    const int lrow = lane & 7, lcol = (lane >> 3) * 8;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      // characters none of this warp's 16 tokens is routed to contribute exactly zero: skip them (stage-2 hard masks:
      // every token has one character, transformer.py:821-822 / :925-926 multiply the others by 0)
      if (w != nullptr && __all_sync(0xffffffffu, wt0[c] == 0.f && wt1[c] == 0.f)) continue;
      // S = Q K_c^T : 16 x 32
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < D / 32; ++k2) {
          uint32_t bk[4];
          ldmatrix_x4(bk, smem_u32(&sK[c][nt * 8 + lrow][k2 * 32 + lcol]));
          mma_bf16_16816(s[nt], qa[2 * k2], bk[0], bk[1]);
          mma_bf16_16816(s[nt], qa[2 * k2 + 1], bk[2], bk[3]);
        }
      }
      // softmax over the 32 keys of character c (rows g and g+8), fp32
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = exp2f((s[nt][0] - m0) * scale_log2);
        s[nt][1] = exp2f((s[nt][1] - m0) * scale_log2);
        s[nt][2] = exp2f((s[nt][2] - m1) * scale_log2);
        s[nt][3] = exp2f((s[nt][3] - m1) * scale_log2);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float f0 = wt0[c] / l0, f1 = wt1[c] / l1;
      // O += (w_c * P_c) V_c
      uint32_t pa[2][4];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        pa[kk][0] = pack_bf16x2(s[2 * kk][0] * f0, s[2 * kk][1] * f0);
        pa[kk][1] = pack_bf16x2(s[2 * kk][2] * f1, s[2 * kk][3] * f1);
        pa[kk][2] = pack_bf16x2(s[2 * kk + 1][0] * f0, s[2 * kk + 1][1] * f0);
        pa[kk][3] = pack_bf16x2(s[2 * kk + 1][2] * f1, s[2 * kk + 1][3] * f1);
      }
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        uint32_t bv[4];   // V^T rows nt*8.. (head-dims), keys 0-7 | 8-15 | 16-23 | 24-31
        ldmatrix_x4(bv, smem_u32(&sV[c][nt * 8 + lrow][lcol]));
        mma_bf16_16816(o[nt], pa[0], bv[0], bv[1]);
        mma_bf16_16816(o[nt], pa[1], bv[2], bv[3]);
      }
    }
    // ---- store
    __nv_bfloat16* d0 = out + n0 * ldo + h * D;
    __nv_bfloat16* d1 = out + n1 * ldo + h * D;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
      if (ok0) *reinterpret_cast<uint32_t*>(d0 + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0], o[nt][1]);
      if (ok1) *reinterpret_cast<uint32_t*>(d1 + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2], o[nt][3]);
    }
    __syncthreads();   // everyone is done with buffer `cur` before the next iteration refills it
  }
}

// Router temporal / multi-ID self-attention (models/router.py:478-488): many independent tiny sequences (L = 13
// frames, or L = C characters) of rows of the [rows, 3*heads*64] qkv matrix; rows of sequence s are tok_stride apart,
// its first row is  base(s) = (s / inner) * outer_stride + (s % inner).
// One WARP per (group of consecutive sequences filling a 16- or 32-row tile, head): their q, k and v rows (128 B each) are read ONCE
// with 16-byte coalesced loads into a padded shared tile (short sequences share a tile and are kept apart by a
// block-diagonal mask), S = Q K^T and O = P V run as m16n8k16 mma.sync tiles fed by ldmatrix, the softmax stays
// in the accumulator registers, and the L output rows leave through the same tile with 16-byte stores.
// HBM-bound: qkv is read once (108 MB per call at the c2 size) and out written once (36 MB).  The first version
// (one thread per query row) re-read every K/V row L times from L2 and took 143-180 us per call; the roofline of
// the call is 22 us.
constexpr int SA_WARPS = 4;
constexpr int SA_PITCH = 72;   // bf16 elements per shared row: 144 B keeps ldmatrix and 16 B accesses conflict-free

BYA_DEVICE void ldmatrix_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int MT>   // MT 16-row tiles per warp: sequences of up to 16 * MT rows
__global__ void __launch_bounds__(SA_WARPS * 32 / MT)
small_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, __nv_bfloat16* __restrict__ out, int ldo,
                       int n_seq, int L, int heads, int inner, long long outer_stride, long long tok_stride,
                       float scale_log2) {
  constexpr int R = 16 * MT;        // tile rows
  constexpr int NW = SA_WARPS / MT; // warps per block (static shared memory stays under 48 KB)
  __shared__ __align__(16) __nv_bfloat16 tile[NW][3][R][SA_PITCH];   // q | k | v, padding rows zero
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spw = R / L;                                                     // sequences per warp tile
  const int rows = spw * L;
  const long long unit = (long long)blockIdx.x * NW + warp;                  // (sequence group, head), heads fastest
  if (unit >= (long long)((n_seq + spw - 1) / spw) * heads) return;
  const int h = int(unit % heads);
  const int s0 = int(unit / heads) * spw;
  const int HD = heads * 64;
  // tile row r = (sequence s0 + r / L, position r % L) -> row of qkv / out (-1: padding)
  auto grow = [&](int r) -> long long {
    const int si = r / L, s = s0 + si;
    if (r >= rows || s >= n_seq) return -1;
    return (long long)(s / inner) * outer_stride + (s % inner) + (long long)(r - si * L) * tok_stride;
  };
  __nv_bfloat16 (*T)[R][SA_PITCH] = tile[warp];

  // ---- load: lane -> (row = lane / 8 + 4 i, 16-byte chunk = lane % 8): 4 rows x 128 B per instruction
  long long gr[4 * MT];
  {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 4 * MT; ++i) gr[i] = grow((lane >> 3) + 4 * i);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
      for (int i = 0; i < 4 * MT; ++i) {
        const int r = (lane >> 3) + 4 * i;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (gr[i] >= 0) v = *reinterpret_cast<const uint4*>(qkv + size_t(gr[i]) * ld + m * HD + h * 64 + ch * 8);
        *reinterpret_cast<uint4*>(&T[m][r][ch * 8]) = v;
      }
    }
  }
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
  // ---- S = Q K^T  (R x R: MT m-tiles x 2 MT n-tiles of 8 keys), k-dim = 64 in 4 steps
  float sc[MT][2 * MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int n = 0; n < 2 * MT; ++n) sc[mt][n][0] = sc[mt][n][1] = sc[mt][n][2] = sc[mt][n][3] = 0.f;
  {
    // ldmatrix.x4 address pattern: lanes 0-15 -> rows 0-15 of the left 8 columns, lanes 16-31 -> the right 8 columns
    const uint32_t qa = smem_u32(&T[0][lane & 15][(lane >> 4) * 8]);
    // K as the col-major B operand: matrices (keys 0-7, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 0-7), (keys 8-15, d 8-15)
    const uint32_t ka = smem_u32(&T[1][(lane & 7) + ((lane >> 4) << 3)][((lane >> 3) & 1) * 8]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) ldmatrix_x4(a[mt], qa + mt * 16 * SA_PITCH * 2 + k * 32);
#pragma unroll
      for (int kt = 0; kt < MT; ++kt) {   // 16 keys per ldmatrix.x4
        uint32_t b[4];
        ldmatrix_x4(b, ka + kt * 16 * SA_PITCH * 2 + k * 32);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(sc[mt][2 * kt], a[mt], b[0], b[1]);
          mma_bf16_16816(sc[mt][2 * kt + 1], a[mt], b[2], b[3]);
        }
      }
    }
  }
  // ---- softmax over the keys of the query's own sequence (rows g and g + 8 of each m-tile; a row lives in one quad)
  uint32_t pa[MT][MT][4];   // P as the A operand of O = P V: accumulator layout == A-fragment layout
  float inv[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
    const int rs0 = (mt * 16 + g) / L, rs1 = (mt * 16 + g + 8) / L;   // which sequence of the tile the rows belong to
#pragma unroll
    for (int n = 0; n < 2 * MT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n * 8 + 2 * t + e;         // key column: only keys of the query's own sequence count
        const int cs = c / L;
        sc[mt][n][e] = (c < rows && cs == rs0) ? sc[mt][n][e] * scale_log2 : -INFINITY;
        sc[mt][n][2 + e] = (c < rows && cs == rs1) ? sc[mt][n][2 + e] * scale_log2 : -INFINITY;
        mx0 = fmaxf(mx0, sc[mt][n][e]);
        mx1 = fmaxf(mx1, sc[mt][n][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    mx0 = (mx0 == -INFINITY) ? 0.f : mx0;       // padding rows: no valid key
    mx1 = (mx1 == -INFINITY) ? 0.f : mx1;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < 2 * MT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[mt][n][e] = exp2f(sc[mt][n][e] - mx0);
        sc[mt][n][2 + e] = exp2f(sc[mt][n][2 + e] - mx1);
        l0 += sc[mt][n][e];
        l1 += sc[mt][n][2 + e];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    inv[mt][0] = 1.f / l0;
    inv[mt][1] = 1.f / l1;
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
      pa[mt][kt][0] = pack_bf16x2(sc[mt][2 * kt][0], sc[mt][2 * kt][1]);
      pa[mt][kt][1] = pack_bf16x2(sc[mt][2 * kt][2], sc[mt][2 * kt][3]);
      pa[mt][kt][2] = pack_bf16x2(sc[mt][2 * kt + 1][0], sc[mt][2 * kt + 1][1]);
      pa[mt][kt][3] = pack_bf16x2(sc[mt][2 * kt + 1][2], sc[mt][2 * kt + 1][3]);
    }
  }
  // ---- O = P V : 8 n-tiles of 8 head-dims; V[key][d] is the row-major [k][n] operand -> ldmatrix.trans
  float o[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int n = 0; n < 8; ++n) o[mt][n][0] = o[mt][n][1] = o[mt][n][2] = o[mt][n][3] = 0.f;
  {
    // matrices: (keys 0-7, d 0-7), (keys 8-15, d 0-7), (keys 0-7, d 8-15), (keys 8-15, d 8-15)
    const uint32_t va = smem_u32(&T[2][(lane & 7) + (((lane >> 3) & 1) << 3)][(lane >> 4) * 8]);
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, va + kt * 16 * SA_PITCH * 2 + n2 * 32);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(o[mt][2 * n2], pa[mt][kt], b[0], b[1]);
          mma_bf16_16816(o[mt][2 * n2 + 1], pa[mt][kt], b[2], b[3]);
        }
      }
    }
  }
  // ---- normalise, stage through the (now dead) q tile, store the valid rows with 16-byte accesses
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(&T[0][mt * 16 + g][n * 8 + 2 * t]) =
          pack_bf16x2(o[mt][n][0] * inv[mt][0], o[mt][n][1] * inv[mt][0]);
      *reinterpret_cast<uint32_t*>(&T[0][mt * 16 + g + 8][n * 8 + 2 * t]) =
          pack_bf16x2(o[mt][n][2] * inv[mt][1], o[mt][n][3] * inv[mt][1]);
    }
  }
  __syncwarp();
  {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 4 * MT; ++i) {
      const int r = (lane >> 3) + 4 * i;
      if (gr[i] >= 0)
        *reinterpret_cast<uint4*>(out + size_t(gr[i]) * ldo + h * 64 + ch * 8) =
            *reinterpret_cast<const uint4*>(&T[0][r][ch * 8]);
    }
  }
}

}  // namespace bya

using namespace bya;

extern "C" int bya_xattn_kv32(void* stream, const void* q, int ldq, const void* K, const void* Vt, const float* w,
                              void* out, int ldo, int tokens, int heads, int head_dim, int chars, int kv_frames,
                              float scale, long long tok_begin, long long total_tokens) {
  if (total_tokens <= 0) { total_tokens = tokens; tok_begin = 0; }
  if (!q || !K || !Vt || !out || tokens <= 0 || heads <= 0 || kv_frames <= 0 || total_tokens % kv_frames ||
      tok_begin < 0 || tok_begin + tokens > total_tokens)
    return BYA_ERR_SHAPE;
  if (ldq % 2 || ldo % 2) return BYA_ERR_ALIGN;
  const int tpf = int(total_tokens / kv_frames);
  static int xa_hpg = 0;   // heads per block (tuning knob BYA_XA_HPG)
  if (!xa_hpg) { const char* e = std::getenv("BYA_XA_HPG"); xa_hpg = e ? std::atoi(e) : 4; if (xa_hpg < 1) xa_hpg = 4; }
  dim3 grid((tpf + XA_TOK - 1) / XA_TOK, kv_frames, (heads + xa_hpg - 1) / xa_hpg);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float sl2 = scale * 1.4426950408889634f;
#define BYA_XA(D_, C_)                                                                                            \
  do {                                                                                                            \
    constexpr int smem = 2 * (C_ * 32 * (D_ + 8) * 2 + C_ * D_ * 40 * 2);   /* two staging buffers */             \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      if (cudaFuncSetAttribute(xattn_kv32_kernel<D_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=   \
          cudaSuccess)                                                                                            \
        return BYA_ERR_CUDA;                                                                                      \
      attr = true;                                                                                                \
    }                                                                                                             \
    xattn_kv32_kernel<D_, C_><<<grid, XA_WARPS * 32, smem, s>>>(                                                  \
        (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)K, (const __nv_bfloat16*)Vt, w, (__nv_bfloat16*)out,  \
        ldo, heads, tpf, kv_frames, sl2, tok_begin, tokens);                                                      \
  } while (0)
  if (head_dim == 64 && chars == 1) BYA_XA(64, 1);
  else if (head_dim == 64 && chars == 2) BYA_XA(64, 2);
  else if (head_dim == 64 && chars == 3) BYA_XA(64, 3);
  else if (head_dim == 128 && chars == 1) BYA_XA(128, 1);
  else if (head_dim == 128 && chars == 2) BYA_XA(128, 2);
  else if (head_dim == 128 && chars == 3) BYA_XA(128, 3);
  else return BYA_ERR_SHAPE;
#undef BYA_XA
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_small_attention(void* stream, const void* qkv, int ld, void* out, int ldo, int n_seq, int seq_len,
                                   int heads, int inner, long long outer_stride, long long tok_stride, float scale) {
  if (!qkv || !out || n_seq <= 0 || seq_len <= 0 || seq_len > 32 || heads <= 0 || inner <= 0) return BYA_ERR_SHAPE;
  if (ld % 8 || ldo % 8) return BYA_ERR_ALIGN;
  const int mt = seq_len <= 16 ? 1 : 2;          // 97-frame clips have 25 latent frames: two 16-row tiles
  const int spw = 16 * mt / seq_len;
  const int nw = SA_WARPS / mt;
  const long long units = (long long)((n_seq + spw - 1) / spw) * heads;
  const int blocks = int((units + nw - 1) / nw);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float sl2 = scale * 1.4426950408889634f;
  if (mt == 1)
    small_attention_kernel<1><<<blocks, nw * 32, 0, st>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)out, ldo, n_seq,
                                                          seq_len, heads, inner, outer_stride, tok_stride, sl2);
  else
    small_attention_kernel<2><<<blocks, nw * 32, 0, st>>>((const __nv_bfloat16*)qkv, ld, (__nv_bfloat16*)out, ldo, n_seq,
                                                          seq_len, heads, inner, outer_stride, tok_stride, sl2);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
