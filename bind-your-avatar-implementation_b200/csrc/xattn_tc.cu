// Routed 32-key cross-attention on the tcgen05 tensor cores (round 2, second form of `bya_xattn_kv32`).
//
//   out[n, h*d : (h+1)*d] = sum_c w[n,c] * softmax_k(scale * q[n,h,:] . K[g(c,n)][h][k][:]) @ V[g(c,n)][h]
//
// Why a second kernel: the mma.sync form (xattn.cu) needs ~1 100 warp instructions per 32 tokens x head — fragment
// loads, quad shuffles for every row maximum / row sum, 128 HMMAs — and plateaus at 0.35-0.4 of the HBM roofline however
// its operands are staged (ncu: no pipe above 35 %, the warps wait on each other's fixed latencies).  Here a CTA owns a
// 128-token tile of one frame and walks the heads of its group:
//   * one elected thread moves everything with TMA: the q tile (128 x d; rows outside this rank's rows are zero-filled,
//     rows past the end of the frame belong to the next frame and are computed but never stored), the 32 keys of EVERY
//     character stacked into one K operand (32 C rows) and each character's V^T (d rows of 32 keys), 2-3 stages ahead;
//   * S = Q K^T is ONE tcgen05.mma chain (M = 128, N = 32 C) into tensor memory; a softmax thread owns a token ROW:
//     it pulls its 32 C scores with tcgen05.ld, takes max / exp2 / sum per character without a single shuffle, folds
//     w_c / sum_c into the probabilities and writes them back to tensor memory as packed bf16;
//   * O = sum_c (w_c P_c) V_c runs as tcgen05.mma with A = P from tensor memory and B = V^T_c (K-major, 64-byte rows,
//     64B swizzle) accumulating over the characters — the routed blend of transformer.py:821-822 / :925-926 costs nothing;
//   * the same four warps drain O (tcgen05.ld), round to bf16 and leave through a 64B-swizzled staging tile + TMA
//     store; warps whose 32 rows are not all valid (end of a frame, edges of a sequence-parallel shard) store per row.
// The q / out tensor maps cover the rows THIS RANK holds; a tile's first row may be negative or run past the end (TMA
// zero-fills / clips), so the same kernel — and the same per-row arithmetic, bit for bit — serves the whole clip and
// sequence-parallel shards that start in the middle of a frame.  3 characters stay on the mma.sync kernel.
// Same contract, same operand layouts (include/bya.h).
#include "common.cuh"
#include "../../include/bya.h"

#include <cstdlib>

namespace bya {

// NWG softmax warpgroups per CTA take alternating heads (two independent score -> probability -> output chains keep the
// tensor pipe and the TMA engine busy while a chain waits on the other units); NST operand stages.
//   d = 64 : NWG 1, NST 2, 128 TMEM columns, 73 KB  -> three CTAs per SM
//   d = 128: NWG 2, NST 3, 512 TMEM columns, 209 KB -> one CTA per SM
template <int D>
struct XtCfg {
  static constexpr int NWG = D == 64 ? 1 : 2;
  static constexpr int NST = D == 64 ? 2 : 3;
  static constexpr int kThreads = 128 * NWG + 32;   // warps [0, 4 NWG): softmax + epilogue (one token row per thread); last warp: TMA + MMA issue
  static constexpr int kCtas = D == 64 ? 3 : 1;
  static constexpr int kWgCols = 64 + D;            // per warpgroup: S (32 C fp32, P written over it as packed bf16) | O (d fp32)
  static constexpr int kTmemCols = NWG * kWgCols <= 128 ? 128 : (NWG * kWgCols <= 256 ? 256 : 512);
};

template <int D, int C>
struct XtSmem {
  static constexpr int KD = D / 64;                     // 64-column (128-byte) operand boxes per q / K row
  static constexpr int kQBox = 128 * 128;               // [128 tokens][128 B]
  static constexpr int kKBox = 32 * C * 128;            // [32 C keys][128 B], characters stacked
  static constexpr int kQ = KD * kQBox;
  static constexpr int kK = KD * kKBox;
  static constexpr int kVTile = D * 64;                 // [d][32 keys] of one character
  static constexpr int kV = C * kVTile;
  static constexpr int kStage = kQ + kK + kV;
  static constexpr int kStgOffset = XtCfg<D>::NST * kStage;         // per softmax warp [32 rows][64 B]
  static constexpr int kBarOffset = kStgOffset + 4 * XtCfg<D>::NWG * 2048;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // + barriers / TMEM slot + alignment slack
  static_assert(kStage % 1024 == 0, "operand tiles must stay 1024-byte aligned");
};

struct XtArgs {
  const float* w;
  __nv_bfloat16* out;      // for the guarded stores of warps whose 32 rows are not all valid (frame end, shard edges)
  long long tok_begin;     // this rank holds the tokens [tok_begin, tok_begin + tok_count) of the clip; q / w / out rows are local
  int tok_count, ldo;
  int heads, tpf, kv_frames, hpg;
  float scale_log2;
};

// K-major operand with 64-byte rows (32 bf16) under the 64B swizzle: 8-row groups are 512 B apart
BYA_DEVICE uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;                 // leading byte offset (unused for swizzled K-major): 16 B
  d |= uint64_t(512 >> 4) << 32;          // stride byte offset
  d |= uint64_t(1) << 46;                 // descriptor version (Blackwell)
  d |= uint64_t(4) << 61;                 // layout = SWIZZLE_64B
  return d;
}

BYA_DEVICE float xt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BYA_DEVICE float xt_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D, int C>
__global__ void __launch_bounds__(XtCfg<D>::kThreads, XtCfg<D>::kCtas)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmo, const XtArgs p) {
  using L = XtSmem<D, C>;
  using Cfg = XtCfg<D>;
  constexpr int NWG = Cfg::NWG, NST = Cfg::NST;
  extern __shared__ uint8_t xt_raw[];
  uint8_t* smem = xt_raw + ((1024u - (smem_u32(xt_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* full = bars;                  // [NST] operands of a head have landed
  uint64_t* empty = full + NST;           // [NST] both MMAs of the head that used the stage are complete
  uint64_t* s_full = empty + NST;         // [NWG] scores in tensor memory
  uint64_t* p_full = s_full + NWG;        // [NWG] probabilities written (4 warps)
  uint64_t* o_full = p_full + NWG;        // [NWG] output accumulator complete
  uint64_t* o_free = o_full + NWG;        // [NWG] output accumulator drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + NWG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x, frame = blockIdx.y;
  const int h_begin = blockIdx.z * p.hpg, h_end = min(p.heads, h_begin + p.hpg);
  const int nh = h_end - h_begin;
  constexpr int WI = 4 * NWG;   // the issuing warp
  // local row of the tile's first token: the q / out tensor maps cover THIS RANK's rows, and TMA zero-fills / clips rows
  // outside them (negative or past the end), so a tile may hang over either edge of the shard
  const long long local0_ll = (long long)frame * p.tpf + (long long)tile * 128 - p.tok_begin;
  if (local0_ll + 128 <= 0 || local0_ll >= p.tok_count) return;   // no owned token in this tile (whole CTA)
  const int local0 = int(local0_ll);

  if (warp == WI && lane == 0) {
    tma_prefetch_desc(&tmq);
    tma_prefetch_desc(&tmk);
    tma_prefetch_desc(&tmv);
    tma_prefetch_desc(&tmo);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < NWG; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WI) {
    if (lane == 0 && nh > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 32 * C, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 0);
      auto load = [&](int i) {
        const int st = i % NST, h = h_begin + i;
        uint8_t* sQ = smem + st * L::kStage;
        uint8_t* sK = sQ + L::kQ;
        uint8_t* sV = sK + L::kK;
        mbar_arrive_expect_tx(&full[st], L::kStage);
#pragma unroll
        for (int kd = 0; kd < L::KD; ++kd) {
          tma_load_2d(sQ + kd * L::kQBox, &tmq, &full[st], h * D + kd * 64, local0, kEvictFirst);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int grp = (c * p.kv_frames + frame) * p.heads + h;
#pragma unroll
          for (int kd = 0; kd < L::KD; ++kd)
            tma_load_2d(sK + kd * L::kKBox + c * 32 * 128, &tmk, &full[st], kd * 64, grp * 32, kEvictLast);
          tma_load_2d(sV + c * L::kVTile, &tmv, &full[st], 0, grp * D, kEvictLast);
        }
      };
      for (int j = 0; j < NST && j < nh; ++j) load(j);
      // issue order: P V of head i - NWG (its warpgroup's score / probability columns are then free), Q K^T of head i
      for (int i = 0; i < nh + NWG; ++i) {
        const int j = i - NWG;
        if (j >= 0) {
          const int g = j % NWG, u = j / NWG, st = j % NST;
          const uint32_t sV = smem_u32(smem + st * L::kStage) + L::kQ + L::kK;
          const uint32_t tW = tmem_base + g * Cfg::kWgCols;
          mbar_wait(&p_full[g], uint32_t(u) & 1u);
          if (u > 0) mbar_wait(&o_free[g], uint32_t(u - 1) & 1u);
          tc_fence_after();
          // O[128, d] = sum_c P_c V_c   (P already carries w_c / rowsum_c)
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const uint64_t dv = make_smem_desc_sw64(sV + c * L::kVTile);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_ts(tW + 64, tW + c * 16 + ks * 8, dv + uint64_t(2 * ks), idesc_pv, (c | ks) != 0);
          }
          umma_commit(&o_full[g]);
          umma_commit(&empty[st]);
        }
        if (i < nh) {
          const int st = i % NST;
          const uint32_t sQ = smem_u32(smem + st * L::kStage);
          const uint32_t sK = sQ + L::kQ;
          mbar_wait(&full[st], uint32_t(i / NST) & 1u);
          tc_fence_after();
          // S[128, 32 C] = Q K^T
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(sQ + (ks >> 2) * L::kQBox, 16, 1024) + uint64_t(2 * (ks & 3));
            const uint64_t db = make_smem_desc_sw128(sK + (ks >> 2) * L::kKBox, 16, 1024) + uint64_t(2 * (ks & 3));
            umma_ss(tmem_base + (i % NWG) * Cfg::kWgCols, da, db, idesc_qk, ks != 0);
          }
          umma_commit(&s_full[i % NWG]);
        }
        if (j >= 0 && j + NST < nh) {   // the stage head j used serves head j + NST once both of j's MMAs are complete
          mbar_wait(&empty[j % NST], uint32_t((j + NST) / NST - 1) & 1u);
          load(j + NST);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: thread = token row
    const int g = warp >> 2, q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int tok = tile * 128 + r;                               // within the frame
    const int lrow = local0 + r;                                  // local row of q / w / out
    // rows past the end of the frame hold the NEXT frame's tokens (wrong K / V), rows outside the shard are zero-filled
    const bool valid = tok < p.tpf && lrow >= 0 && lrow < p.tok_count;
    const bool warp_full = __all_sync(0xffffffffu, valid);        // all 32 rows valid: the warp's tile leaves by TMA
    float wt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) wt[c] = valid ? (p.w ? p.w[size_t(lrow) * C + c] : 1.f) : 0.f;
    const uint32_t tS = tmem_base + g * Cfg::kWgCols + (uint32_t(q4 * 32) << 16);
    const uint32_t tO = tS + 64;
    uint8_t* stg = smem + L::kStgOffset + warp * 2048;
    const uint32_t stg_row = smem_u32(stg) + lane * 64;
    const int row0 = local0 + q4 * 32;
    const float sl2 = p.scale_log2;
    for (int i = g; i < nh; i += NWG) {
      const int h = h_begin + i;
      const uint32_t par = uint32_t(i / NWG) & 1u;
      mbar_wait(&s_full[g], par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < C; ++c) {
        uint32_t sr[32];
        tmem_ld_x32(tS + c * 32, sr);
        tmem_ld_wait();
        float m0 = fmaxf(__uint_as_float(sr[0]), __uint_as_float(sr[1])), m1 = fmaxf(__uint_as_float(sr[2]), __uint_as_float(sr[3]));
#pragma unroll
        for (int k = 4; k < 32; k += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(sr[k]), __uint_as_float(sr[k + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(sr[k + 2]), __uint_as_float(sr[k + 3])));
        }
        const float nm = -fmaxf(m0, m1) * sl2;
        float e[32];
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          e[k] = xt_ex2(fmaf(__uint_as_float(sr[k]), sl2, nm));
          e[k + 1] = xt_ex2(fmaf(__uint_as_float(sr[k + 1]), sl2, nm));
          l0 += e[k];
          l1 += e[k + 1];
        }
        const float f = wt[c] * xt_rcp(l0 + l1);
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) pk[k] = pack_bf16x2(e[2 * k] * f, e[2 * k + 1] * f);
        // P_c goes over the first score columns: [16 c, 16 c + 16) lies inside S_0 .. S_c, all of which this thread
        // already holds in registers (thread-private TMEM lane); S_{c+1} starts at column 32 (c + 1)
        tmem_st_x16(tS + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);

      mbar_wait(&o_full[g], par);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t orr[32];
        tmem_ld_x32(tO + ch * 32, orr);
        tmem_ld_wait();
        if (ch == D / 32 - 1) {   // the accumulator sits in registers: the tensor pipe may start this warpgroup's next P V
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[g]);
        }
        if (!warp_full) {   // frame end / shard edge: every valid row stores its own 64 bytes
          if (valid) {
            uint4* d = reinterpret_cast<uint4*>(p.out + size_t(lrow) * p.ldo + h * D + ch * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d[j] = make_uint4(pack_bf16x2(__uint_as_float(orr[8 * j]), __uint_as_float(orr[8 * j + 1])),
                                pack_bf16x2(__uint_as_float(orr[8 * j + 2]), __uint_as_float(orr[8 * j + 3])),
                                pack_bf16x2(__uint_as_float(orr[8 * j + 4]), __uint_as_float(orr[8 * j + 5])),
                                pack_bf16x2(__uint_as_float(orr[8 * j + 6]), __uint_as_float(orr[8 * j + 7])));
          }
          continue;
        }
        if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(stg_row + (((j ^ (lane >> 1)) & 3) << 4),
                       pack_bf16x2(__uint_as_float(orr[8 * j]), __uint_as_float(orr[8 * j + 1])),
                       pack_bf16x2(__uint_as_float(orr[8 * j + 2]), __uint_as_float(orr[8 * j + 3])),
                       pack_bf16x2(__uint_as_float(orr[8 * j + 4]), __uint_as_float(orr[8 * j + 5])),
                       pack_bf16x2(__uint_as_float(orr[8 * j + 6]), __uint_as_float(orr[8 * j + 7])));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmo, stg, h * D + ch * 32, row0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int D, int C>
static int launch_xattn_tc(cudaStream_t s, const void* q, int ldq, const void* K, const void* Vt, const float* w, void* out,
                           int ldo, int tokens, int heads, int kv_frames, float scale, long long tok_begin,
                           long long total_tokens, int hpg) {
  using L = XtSmem<D, C>;
  const int tpf = int(total_tokens / kv_frames);
  const int G = C * kv_frames;
  CUtensorMap tmq, tmk, tmv, tmo;
  int rc = bya_host::encode_tmap_bf16(&tmq, q, uint64_t(heads) * D, tokens, uint64_t(ldq) * 2, 64, 128);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tmo, out, uint64_t(heads) * D, tokens, uint64_t(ldo) * 2, 32, 32);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tmk, K, D, uint64_t(G) * heads * 32, uint64_t(D) * 2, 64, 32);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tmv, Vt, 32, uint64_t(G) * heads * D, 64, 32, D);
  if (rc) return rc;
  auto kern = xattn_tc_kernel<D, C>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess) return BYA_ERR_CUDA;
    attr = true;
  }
  XtArgs a;
  a.w = w;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.tok_begin = tok_begin, a.tok_count = tokens, a.ldo = ldo;
  a.heads = heads, a.tpf = tpf, a.kv_frames = kv_frames, a.hpg = hpg;
  a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((tpf + 127) / 128, kv_frames, (heads + hpg - 1) / hpg);
  kern<<<grid, XtCfg<D>::kThreads, L::kTotal, s>>>(tmq, tmk, tmv, tmo, a);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

// Called by bya_xattn_kv32 (xattn.cu) for calls with 1 or 2 characters; returns 1 when the call is not for this kernel
// (the mma.sync form then runs), BYA_OK / an error code otherwise.
int xattn_tc_dispatch(cudaStream_t s, const void* q, int ldq, const void* K, const void* Vt, const float* w, void* out,
                      int ldo, int tokens, int heads, int head_dim, int chars, int kv_frames, float scale,
                      long long tok_begin, long long total_tokens) {
  static int on = -1, hpg_env = 0;
  if (on < 0) {
    const char* e = std::getenv("BYA_XA_TC");
    on = e ? std::atoi(e) : 1;
    const char* g = std::getenv("BYA_XA_TC_HPG");
    hpg_env = g ? std::atoi(g) : 0;
  }
  if (!on || chars > 2) return 1;
  const int hpg = hpg_env > 0 ? hpg_env : (head_dim == 64 ? 8 : 4);
#define BYA_XT(D_, C_) \
  return launch_xattn_tc<D_, C_>(s, q, ldq, K, Vt, w, out, ldo, tokens, heads, kv_frames, scale, tok_begin, total_tokens, hpg)
  if (head_dim == 64 && chars == 1) BYA_XT(64, 1);
  if (head_dim == 64 && chars == 2) BYA_XT(64, 2);
  if (head_dim == 128 && chars == 1) BYA_XT(128, 1);
  if (head_dim == 128 && chars == 2) BYA_XT(128, 2);
#undef BYA_XT
  return 1;
}

}  // namespace bya
