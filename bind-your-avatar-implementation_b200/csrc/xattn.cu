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

constexpr int XA_WARPS = 4;

BYA_DEVICE float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BYA_DEVICE float xa_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (one issue slot for two lanes of an accumulator register pair)
BYA_DEVICE void xa_fma2(float& a0, float& a1, float b, float c) {   // a = a * b + c
  asm("{\n\t.reg .b64 va, vb, vc;\n\tmov.b64 va, {%0, %1};\n\tmov.b64 vb, {%2, %2};\n\tmov.b64 vc, {%3, %3};\n\t"
      "fma.rn.f32x2 va, va, vb, vc;\n\tmov.b64 {%0, %1}, va;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b), "f"(c));
}
BYA_DEVICE void xa_mul2(float& a0, float& a1, float b) {            // a = a * b
  asm("{\n\t.reg .b64 va, vb;\n\tmov.b64 va, {%0, %1};\n\tmov.b64 vb, {%2, %2};\n\t"
      "mul.rn.f32x2 va, va, vb;\n\tmov.b64 {%0, %1}, va;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b));
}
BYA_DEVICE void xa_add2(float& a0, float& a1, float b0, float b1) { // a = a + b
  asm("{\n\t.reg .b64 va, vb;\n\tmov.b64 va, {%0, %1};\n\tmov.b64 vb, {%2, %3};\n\t"
      "add.rn.f32x2 va, va, vb;\n\tmov.b64 {%0, %1}, va;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}

// K  : [G][H][32][D]   (keys x head-dim, head-dim contiguous)
// Vt : [G][H][D][32]   (V transposed: head-dim x keys, keys contiguous)  — both prepared once per generation
//
// Round-2 rewrite (the first version — K / V^T staged in shared memory per block, 4-byte q loads and 4-byte output
// stores in the mma fragment layout — ran at 0.27 of the HBM roofline: mio / short-scoreboard / barrier stalls, and
// every warp instruction touched 8 half-used sectors).  Now:
//   * no block-level synchronisation at all: a WARP owns 16*MT tokens x the block's heads;
//   * every operand moves in 16-byte pieces.  The contraction index of an MMA may be enumerated in any order as long as
//     A and B agree, and so may the column index of B as long as the consumer of C knows it.  With
//         k-slot (2t, 2t+1, 2t+8, 2t+9) of k-steps 2kk / 2kk+1   <->   elements 32kk + 8t + (0..3) / (4..7)
//     thread (g, t) feeds both k-steps of Q K^T from ONE 16-byte piece of its q row and ONE of a K row; with
//         column g of score tile nt   <->   key 8(g>>1) + 2nt + (g&1)
//     the 8 scores a thread holds per row are the keys 8t..8t+7 — exactly one 16-byte piece of a V^T row for P V; with
//         column g of output tile nt  <->   dim 32(nt>>2) + 8(g>>1) + 2(nt&3) + (g&1)
//     a thread ends up with 8 CONSECUTIVE output dims per group of four tiles: one 16-byte store.
//     K / V^T (10 MB in all) are read through L1 in their natural layouts — no staging, no ldmatrix — and the block
//     prefetches the NEXT head's K / V^T lines into L1 (one line per thread) while it works on the current one: without
//     it every first touch of a line is an exposed L2 round trip in front of an MMA (ncu: 23 % of the samples);
//   * q rows stream in with cp.async (16 B, zero-fill for rows this rank does not own) one head ahead into a per-warp,
//     XOR-swizzled landing buffer, after an L2 prefetch of the warp's whole q range (all heads of the block);
//   * MT = 2 token tiles per warp at d = 64 share every K / V^T fragment (L1 wavefronts per token halved);
//   * the softmax runs on ex2.approx / rcp.approx and packed f32x2 arithmetic (the kernel is issue-bound: ~1 600 warp
//     instructions per 32 tokens x head before, the MMAs are 128 of them).
// pf_mode: 0 no K / V^T prefetch, 1 prefetch.global.L1 (default; measured 97.7 -> 90.0 us at the audio shape).
// STAGE: K / V^T of the block's current and next head staged in shared memory by cp.async instead (one block barrier per
// head): 82 us at the audio shape — kept as BYA_XA_VAR=1 for the 2-character shapes; the default path for 1-2 characters is
// the tensor-memory kernel (xattn_tc.cu), this one serves 3 characters and BYA_XA_TC=0.
template <int D, int C, int MT, int XA_BLOCKS, bool STAGE>
__global__ void __launch_bounds__(XA_WARPS * 32, XA_BLOCKS)
xattn_kv32_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ K,
                  const __nv_bfloat16* __restrict__ Vt, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                  int ldo, int heads, int tokens_per_frame, int kv_frames, float scale_log2, long long tok_begin,
                  long long tok_count, int pf_mode) {
  constexpr int ROWS = 16 * MT;   // tokens per warp
  constexpr int CPR = D / 8;      // 16-byte chunks per q row of one head
  constexpr int KK = D / 32;      // 16-byte pieces per thread and row (two k-steps each)
  constexpr int NJ = ROWS * CPR / 32;   // cp.async pieces per lane and head
  // dynamic shared memory: per-warp q landing buffers [XA_WARPS][2][ROWS * CPR], then (STAGE) the block's K / V^T
  // buffers [2][C][K: 32 rows x CPR chunks, chunk ^ 4 on odd rows | V^T: D rows x 4 chunks, linear]
  extern __shared__ __align__(128) uint4 xa_dyn[];
  uint4 (*qs)[2][ROWS * CPR] = reinterpret_cast<uint4 (*)[2][ROWS * CPR]>(xa_dyn);
  constexpr int KVC = 4 * D;                       // 16-byte chunks of one K (or V^T) tile of one character
  uint4* kvs = xa_dyn + XA_WARPS * 2 * ROWS * CPR;   // [2][C][2][KVC]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t zr = uint32_t(pf_mode) >> 8;   // 0 (pf_mode is 0..2), opaque to the compiler
  const int frame = blockIdx.y;
  const int tokw = (blockIdx.x * XA_WARPS + warp) * ROWS;          // this warp's first token within the frame
  // global token index -> local row of q / w / out (this rank owns tokens [tok_begin, tok_begin + tok_count))
  const long long base = (long long)frame * tokens_per_frame + tokw - tok_begin;
  auto row_ok = [&](int rr) { return tokw + rr < tokens_per_frame && base + rr >= 0 && base + rr < tok_count; };
  const int hpg = (heads + gridDim.z - 1) / gridDim.z;
  const int h_begin = blockIdx.z * hpg, h_end = min(heads, int(blockIdx.z + 1) * hpg);
  if (h_begin >= h_end) return;
  // a warp without an owned token still helps with the block's K / V^T prefetches; it just has no rows
  const bool warp_has_rows = tokw < tokens_per_frame && base + ROWS > 0 && base < tok_count;

  // validity of the rows this thread computes (bit mt*2+j: row mt*16 + g + 8j) and copies (bit 8+j: row (lane + 32j) / CPR)
  uint32_t vmask = 0;
  float wt[MT][2][C];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int rr = mt * 16 + g + 8 * j;
      const bool okr = warp_has_rows && row_ok(rr);
      vmask |= uint32_t(okr) << (mt * 2 + j);
#pragma unroll
      for (int c = 0; c < C; ++c) wt[mt][j][c] = okr ? (w ? w[size_t(base + rr) * C + c] : 1.f) : 0.f;
    }
#pragma unroll
  for (int j = 0; j < NJ; ++j) vmask |= uint32_t(warp_has_rows && row_ok((lane + 32 * j) / CPR)) << (8 + j);

  // ---- L2 prefetch of the warp's q rows for all heads of this block (contiguous per row)
  if (warp_has_rows) {
    const int lines = ((h_end - h_begin) * D * 2 + 127) / 128;
    for (int rr = lane / 8; rr < ROWS; rr += 4) {
      if (!row_ok(rr)) continue;
      const char* p = reinterpret_cast<const char*>(q + size_t(base + rr) * ldq + h_begin * D);
      for (int ln = lane & 7; ln < lines; ln += 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + ln * 128));
    }
  }
  // ---- K / V^T of head h (every character) -> L1, one 128-byte line per thread and round
  auto prefetch_kv = [&](int h) {
    if (pf_mode == 0) return;
    constexpr int LPT = 32 * D * 2 / 128;          // lines per (character, K or V^T) tile
    for (int i = threadIdx.x; i < C * 2 * LPT; i += XA_WARPS * 32) {
      const int which = i / LPT, ln = i - which * LPT;
      const size_t grp = (size_t(which >> 1) * kv_frames + frame) * heads + h;
      const char* p = reinterpret_cast<const char*>(((which & 1) ? Vt : K) + grp * 32 * D) + ln * 128;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
  };
  // ---- q rows of head h -> landing buffer b (16-byte cp.async, chunk ^ 4 on odd rows keeps the fragment reads conflict-free)
  const __nv_bfloat16* qlane = q + (lane % CPR) * 8;
  auto issue = [&](int h, int b) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int rr = (lane + 32 * j) / CPR, ch = lane % CPR;
      const bool v = (vmask >> (8 + j)) & 1u;
      const __nv_bfloat16* src = v ? qlane + size_t(base + rr) * ldq + h * D : q;
      const uint32_t dst = smem_u32(&qs[warp][b][rr * CPR + (ch ^ ((rr & 1) << 2))]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(v ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- (STAGE) K / V^T of head h, every character -> shared buffer b, by the whole block
  auto stage_kv = [&](int h, int b) {
    for (int i = threadIdx.x; i < C * 2 * KVC; i += XA_WARPS * 32) {
      const int which = i / KVC, idx = i - which * KVC;   // which = 2 c + (0: K, 1: V^T)
      const size_t grp = (size_t(which >> 1) * kv_frames + frame) * heads + h;
      const __nv_bfloat16* src = ((which & 1) ? Vt : K) + grp * 32 * D + idx * 8;
      int d = idx;
      if (!(which & 1)) { const int row = idx / CPR, ch = idx % CPR; d = row * CPR + (ch ^ ((row & 1) << 2)); }
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&kvs[(b * C * 2 + which) * KVC + d])), "l"(src) : "memory");
    }
  };
  if constexpr (STAGE) {
    stage_kv(h_begin, 0);
    if (warp_has_rows) issue(h_begin, 0);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
    if (warp_has_rows) issue(h_begin, 0);
  }
  for (int h = h_begin; h < h_end; ++h) {
    const int cur = (h - h_begin) & 1;
    if constexpr (STAGE) {
      // one group per head holds this thread's q pieces AND its share of the block's K / V^T; one barrier per head:
      // head h has landed for everyone, and everyone is done with head h-1 (whose buffers the next copies overwrite)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (h + 1 < h_end) {
        stage_kv(h + 1, cur ^ 1);
        if (warp_has_rows) issue(h + 1, cur ^ 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (!warp_has_rows) continue;
    } else {
      if (h + 1 < h_end) prefetch_kv(h + 1);
      if (!warp_has_rows) continue;
      __syncwarp();                 // every lane has taken its fragments of head h-1 out of buffer cur^1
      if (h + 1 < h_end) {
        issue(h + 1, cur ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();                 // ... and sees the pieces the other lanes fetched for head h
    }

    // A fragments of both k-steps of every 32-element slice: (row g | row g+8) x (elements 0-1 | 2-3) and (4-5 | 6-7)
    uint32_t qa[MT][2 * KK][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        const uint4 r0 = qs[warp][cur][(mt * 16 + g) * CPR + ((kk * 4 + t) ^ ((g & 1) << 2))];
        const uint4 r1 = qs[warp][cur][(mt * 16 + g + 8) * CPR + ((kk * 4 + t) ^ ((g & 1) << 2))];
        // `^ zr` (a zero ptxas cannot see through): without it the fragment words keep living in the two 16-byte load
        // results, whose register adjacency is not the MMA's, and every MMA gets its A quad copied together first
        // (SASS: 4 moves in front of each of the 128 HMMAs)
        qa[mt][2 * kk][0] = r0.x ^ zr, qa[mt][2 * kk][1] = r1.x ^ zr, qa[mt][2 * kk][2] = r0.y ^ zr, qa[mt][2 * kk][3] = r1.y ^ zr;
        qa[mt][2 * kk + 1][0] = r0.z ^ zr, qa[mt][2 * kk + 1][1] = r1.z ^ zr, qa[mt][2 * kk + 1][2] = r0.w ^ zr,
        qa[mt][2 * kk + 1][3] = r1.w ^ zr;
      }

    float o[MT][D / 8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < D / 8; ++i) o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.f;

#pragma unroll
    for (int c = 0; c < C; ++c) {
      // characters none of this warp's tokens is routed to contribute exactly zero: skip them (stage-2 hard masks:
      // every token has one character, transformer.py:821-822 / :925-926 multiply the others by 0)
      bool none = true;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) none = none && wt[mt][0][c] == 0.f && wt[mt][1][c] == 0.f;
      if (__all_sync(0xffffffffu, none)) continue;
      const size_t grp = (size_t(c) * kv_frames + frame) * heads + h;
      const __nv_bfloat16* Kc = K + grp * 32 * D + 8 * t + (8 * (g >> 1) + (g & 1)) * D;
      const __nv_bfloat16* Vc = Vt + grp * 32 * D + 8 * t + (8 * (g >> 1) + (g & 1)) * 32;
      // S = Q K_c^T : (16 MT) x 32; column g of tile nt is key 8(g>>1) + 2nt + (g&1)
      float s[MT][4][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
      // (k-slice outermost, score tile innermost: an A quad is put together once and feeds four MMAs)
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        uint4 b[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if constexpr (STAGE)
            b[nt] = kvs[(cur * C * 2 + 2 * c) * KVC + (8 * (g >> 1) + 2 * nt + (g & 1)) * CPR + ((kk * 4 + t) ^ ((g & 1) << 2))];
          else
            b[nt] = __ldg(reinterpret_cast<const uint4*>(Kc + 2 * nt * D + kk * 32));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(s[mt][nt], qa[mt][2 * kk], b[nt].x, b[nt].y);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(s[mt][nt], qa[mt][2 * kk + 1], b[nt].z, b[nt].w);
        }
      }
      // softmax over the 32 keys of character c (rows g and g+8 of every token tile), fp32; w_c / sum folded into P
      uint32_t pa[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        float f[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {          // j = 0: row g (accumulator lanes 0, 1); j = 1: row g + 8 (lanes 2, 3)
          float m = fmaxf(fmaxf(s[mt][0][2 * j], s[mt][0][2 * j + 1]), fmaxf(s[mt][1][2 * j], s[mt][1][2 * j + 1]));
          m = fmaxf(m, fmaxf(fmaxf(s[mt][2][2 * j], s[mt][2][2 * j + 1]), fmaxf(s[mt][3][2 * j], s[mt][3][2 * j + 1])));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          const float nm = -m * scale_log2;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            xa_fma2(s[mt][nt][2 * j], s[mt][nt][2 * j + 1], scale_log2, nm);
            s[mt][nt][2 * j] = xa_ex2(s[mt][nt][2 * j]);
            s[mt][nt][2 * j + 1] = xa_ex2(s[mt][nt][2 * j + 1]);
          }
          float l0 = s[mt][0][2 * j], l1 = s[mt][0][2 * j + 1];
#pragma unroll
          for (int nt = 1; nt < 4; ++nt) xa_add2(l0, l1, s[mt][nt][2 * j], s[mt][nt][2 * j + 1]);
          float l = l0 + l1;
          l += __shfl_xor_sync(0xffffffffu, l, 1);
          l += __shfl_xor_sync(0xffffffffu, l, 2);
          f[j] = wt[mt][j][c] * xa_rcp(l);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) xa_mul2(s[mt][nt][2 * j], s[mt][nt][2 * j + 1], f[j]);
        }
        // k-slots (2t, 2t+1 | 2t+8, 2t+9) of k-step kk hold the keys 8t + 4kk + (0, 1 | 2, 3): tiles 2kk and 2kk+1
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          pa[mt][kk][0] = pack_bf16x2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
          pa[mt][kk][1] = pack_bf16x2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
          pa[mt][kk][2] = pack_bf16x2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
          pa[mt][kk][3] = pack_bf16x2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
        }
      }
      // O += (w_c * P_c) V_c; column g of tile nt is head-dim 32(nt>>2) + 8(g>>1) + 2(nt&3) + (g&1)
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        uint4 b;   // keys 8t .. 8t+7 of that head-dim
        if constexpr (STAGE)
          b = kvs[(cur * C * 2 + 2 * c + 1) * KVC + (32 * (nt >> 2) + 8 * (g >> 1) + 2 * (nt & 3) + (g & 1)) * 4 + t];
        else
          b = __ldg(reinterpret_cast<const uint4*>(Vc + (32 * (nt >> 2) + 2 * (nt & 3)) * 32));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(o[mt][nt], pa[mt][0], b.x, b.y);
          mma_bf16_16816(o[mt][nt], pa[mt][1], b.z, b.w);
        }
      }
    }
    // ---- store: 8 consecutive dims per thread and group of four tiles
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!((vmask >> (mt * 2 + j)) & 1u)) continue;
        __nv_bfloat16* d = out + size_t(base + mt * 16 + g + 8 * j) * ldo + h * D + 8 * t;
#pragma unroll
        for (int G = 0; G < D / 32; ++G) {
          uint4 v;
          v.x = pack_bf16x2(o[mt][4 * G][2 * j], o[mt][4 * G][2 * j + 1]);
          v.y = pack_bf16x2(o[mt][4 * G + 1][2 * j], o[mt][4 * G + 1][2 * j + 1]);
          v.z = pack_bf16x2(o[mt][4 * G + 2][2 * j], o[mt][4 * G + 2][2 * j + 1]);
          v.w = pack_bf16x2(o[mt][4 * G + 3][2 * j], o[mt][4 * G + 3][2 * j + 1]);
          __stcs(reinterpret_cast<uint4*>(d + 32 * G), v);
        }
      }
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

namespace bya {
int xattn_tc_dispatch(cudaStream_t s, const void* q, int ldq, const void* K, const void* Vt, const float* w, void* out,
                      int ldo, int tokens, int heads, int head_dim, int chars, int kv_frames, float scale,
                      long long tok_begin, long long total_tokens);   // xattn_tc.cu
}

using namespace bya;

extern "C" int bya_xattn_kv32(void* stream, const void* q, int ldq, const void* K, const void* Vt, const float* w,
                              void* out, int ldo, int tokens, int heads, int head_dim, int chars, int kv_frames,
                              float scale, long long tok_begin, long long total_tokens) {
  if (total_tokens <= 0) { total_tokens = tokens; tok_begin = 0; }
  if (!q || !K || !Vt || !out || tokens <= 0 || heads <= 0 || kv_frames <= 0 || total_tokens % kv_frames ||
      tok_begin < 0 || tok_begin + tokens > total_tokens)
    return BYA_ERR_SHAPE;
  // every q / out / K / V^T access is a 16-byte piece
  if (ldq % 8 || ldo % 8) return BYA_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(K) |
       reinterpret_cast<uintptr_t>(Vt)) & 15)
    return BYA_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {   // 1-2 characters: tensor-memory form (xattn_tc.cu)
    const int r = xattn_tc_dispatch(s, q, ldq, K, Vt, w, out, ldo, tokens, heads, head_dim, chars, kv_frames, scale,
                                    tok_begin, total_tokens);
    if (r != 1) return r;
  }
  const int tpf = int(total_tokens / kv_frames);
  static int xa_hpg = -1;   // heads per block (tuning knob BYA_XA_HPG; default 4 at d = 64, 2 at d = 128)
  if (xa_hpg < 0) { const char* e = std::getenv("BYA_XA_HPG"); xa_hpg = e ? std::atoi(e) : 0; if (xa_hpg < 0) xa_hpg = 0; }
  const int hpg = xa_hpg ? xa_hpg : (head_dim == 64 ? 4 : 2);
  static int xa_pf = -1;    // K / V^T L1 prefetch of the next head (BYA_XA_PF: 0 off, 1 prefetch.global.L1)
  if (xa_pf < 0) { const char* e = std::getenv("BYA_XA_PF"); xa_pf = e ? std::atoi(e) : 1; if (xa_pf < 0 || xa_pf > 1) xa_pf = 1; }
  const float sl2 = scale * 1.4426950408889634f;
#define BYA_XA(D_, C_, MT_, BL_, ST_)                                                                             \
  do {                                                                                                            \
    dim3 grid((tpf + XA_WARPS * 16 * MT_ - 1) / (XA_WARPS * 16 * MT_), kv_frames, (heads + hpg - 1) / hpg);       \
    constexpr int smem = XA_WARPS * 2 * 16 * MT_ * (D_ / 8) * 16 + (ST_ ? 2 * C_ * 2 * 4 * D_ * 16 : 0);          \
    auto kern = xattn_kv32_kernel<D_, C_, MT_, BL_, ST_>;                                                         \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)           \
        return BYA_ERR_CUDA;                                                                                      \
      attr = true;                                                                                                \
    }                                                                                                             \
    kern<<<grid, XA_WARPS * 32, smem, s>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)K,                 \
                                           (const __nv_bfloat16*)Vt, w, (__nv_bfloat16*)out, ldo, heads, tpf,     \
                                           kv_frames, sl2, tok_begin, tokens, xa_pf);                             \
  } while (0)
  static int xa_var = -1;   // tuning knob BYA_XA_VAR: 0 K / V^T through L1, 1 staged in shared memory
  if (xa_var < 0) { const char* e = std::getenv("BYA_XA_VAR"); xa_var = e ? std::atoi(e) : 0; }
  if (head_dim == 64 && chars == 1) BYA_XA(64, 1, 2, 3, false);
  else if (head_dim == 64 && chars == 2) {
    if (xa_var == 1) BYA_XA(64, 2, 2, 3, true);
    else BYA_XA(64, 2, 2, 3, false);
  }
  else if (head_dim == 64 && chars == 3) BYA_XA(64, 3, 2, 3, false);
  else if (head_dim == 128 && chars == 1) BYA_XA(128, 1, 1, 3, false);
  else if (head_dim == 128 && chars == 2) {
    if (xa_var == 1) BYA_XA(128, 2, 1, 2, true);
    else BYA_XA(128, 2, 1, 3, false);
  }
  else if (head_dim == 128 && chars == 3) BYA_XA(128, 3, 1, 3, false);
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
