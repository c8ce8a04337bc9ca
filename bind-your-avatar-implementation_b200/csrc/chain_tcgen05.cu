// One "link" of the router's SpatialTemporalAttentionBlock chain as ONE kernel (sm_100a only):
//
//     X    = resid + A1 · W1^T + b1                      (attention out-projection / MLP down-projection + residual)
//     out2 = act( LayerNorm_512(X) · W2^T + b2 )         (the NEXT sub-block's norm + fused q|k|v / MLP up-projection)
//
// i.e. models/router.py:474-476 -> :478-479, :481-483 -> :485-486, :488 -> :490-491 and :491 -> the next block's :474
// (`x = x + attn(...)`, `x_norm = norm(x)`, `to_q/to_k/to_v` or `mlp[0]`), which the unfused path runs as GEMM ->
// LayerNorm kernel -> GEMM with two round trips of the [C*Nv, 512] activations through L2 / HBM in between.
//
// Why it pays: with K = 512 every 128 x 256 output tile of the plain GEMM re-streams its operands for 4 096 cycles of
// MMA, and ncu shows those GEMMs sitting on the L2 -> SM bandwidth (12.7 TB/s: 736 MB in 58 us for the 1536-wide one),
// not on the tensor pipe.  Here a CTA PAIR owns 256 rows: each CTA keeps ITS 128 rows of X in TENSOR MEMORY (bf16, 256
// columns) as the A operand of the second GEMM (tcgen05.mma with A from TMEM), so the second GEMM streams only weights
// — and only HALF of every weight tile per CTA (cta_group::2: the pair's tensor cores read both halves).
//
// LayerNorm is folded algebraically, so X is never rewritten:  LN(x)·W^T = rstd * (x·W'^T - mean * csum) + b',
// W' = W diag(gamma) (bf16), csum[n] = sum_k W'[n,k], b' = b + W·beta  (host side: `engine.RouterPack`).  The row
// statistics come from the bf16-rounded X values the first epilogue writes — the same numbers the reference's
// LayerNorm sees — accumulated in fp32 by the thread that owns the row (TMEM lane == row) and reused by the same
// thread in the second epilogue: no cross-thread traffic beyond one exchange between the two column halves.
//
// Per CTA: warps 0..7 epilogue, warp 8 TMA producer, warp 9 MMA issuer (leader CTA only), warp 10 TMEM allocator.  The
// issuing warps carry the HIGHEST warp ids: the sub-partition arbiter prefers higher ids, and an MMA issuer that shares
// its sub-partition with two busy epilogue warps must not wait behind them (measured: the same kernel with the issuer
// as warp 1 lost a third of its speed as soon as the epilogues did real work).
// TMEM (512 columns): X [0,256) as packed bf16 pairs, two 128-column fp32 accumulator stages [256,384) / [384,512).
// Shared memory: 7-stage ring of 24 KB (first GEMM: A1 k-block 16 KB + half W1 tile 8 KB; second GEMM: two half W2
// k-blocks of 8 KB), bias / csum vectors, statistics exchange, output staging tiles.
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

constexpr int CH_D = 512;         // router width: K of both GEMMs and N of the first
constexpr int CH_BM = 128;        // rows per CTA (256 per pair)
constexpr int CH_BK = 64;
constexpr int CH_BN = 128;        // accumulator chunk (columns)
constexpr int CH_HALF = CH_BN / 2;   // weight rows staged by ONE CTA per chunk
constexpr int CH_STAGES = 7;
constexpr int CH_A_BYTES = CH_BM * CH_BK * 2;        // 16 KB
constexpr int CH_W_BYTES = CH_HALF * CH_BK * 2;      // 8 KB
constexpr int CH_STAGE_BYTES = CH_A_BYTES + CH_W_BYTES;
constexpr int CH_THREADS = 384;
constexpr int CH_EPI = 256;
constexpr int CH_MAX_N2 = 1536;
constexpr int CH_VEC_OFF = CH_STAGES * CH_STAGE_BYTES;                    // b1 [512] | csum [1536] | b2 [1536] fp32
constexpr int CH_VEC_BYTES = (CH_D + 2 * CH_MAX_N2) * 4;
constexpr int CH_STAT_OFF = CH_VEC_OFF + CH_VEC_BYTES;                    // [4 chunks x 2 halves][128 rows] float2
constexpr int CH_STAT_BYTES = 2 * (CH_D / CH_BN) * CH_BM * 8;
constexpr int CH_STG_OFF = ((CH_STAT_OFF + CH_STAT_BYTES + 1023) / 1024) * 1024;   // [8 warps][32 rows x 64 B]
constexpr int CH_STG_BYTES = 8 * 4096;   // [8 warps][32 rows x 128 B]
constexpr int CH_BAR_OFF = CH_STG_OFF + CH_STG_BYTES;
constexpr int CH_SMEM = CH_BAR_OFF + 256 + 1024;
constexpr int CH_TMEM_X = 0, CH_TMEM_ACC = 256;

struct ChainParams {
  int M, N2, act, n_split;
  int ldr, ldx, ldc;
  long long col_block_stride;
  __nv_bfloat16* x_out;
  __nv_bfloat16* out2;
  int a_kblock, col_block;
  int store_x;
  float ln_eps;
  const __nv_bfloat16* b1;
  const __nv_bfloat16* resid;
  const float* csum;
  const float* b2;
};

// (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1) as ONE packed FFMA2 (each half rounds like fmaf)
BYA_DEVICE void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 va, vb, vc;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vb, {%4, %5};\n\t"
      "mov.b64 vc, {%6, %7};\n\t"
      "fma.rn.f32x2 va, va, vb, vc;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

// D[tmem of both CTAs] (+)= A[tmem of both CTAs] * B[smem halves of both]   (M = 256)
BYA_DEVICE void umma_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int ACT>   // GEMM_ACT_* of the second GEMM, compile-time: the unrolled epilogue holds one variant, not four
__global__ void __launch_bounds__(CH_THREADS, 1)
gemm_ln_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_w1,
                    const __grid_constant__ CUtensorMap tmap_w2, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the shared array (not an integer round trip) keeps the address space visible to the compiler:
  // plain C++ accesses below become LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t rank = cluster_ctarank();   // 0 = leader (issues the MMAs)
  const int pair = int(blockIdx.x >> 1);
  const int num_pairs = int(gridDim.x) >> 1;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CH_BAR_OFF);
  uint64_t* empty_bar = full_bar + CH_STAGES;
  uint64_t* tfull_bar = empty_bar + CH_STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained (leader's copy collects both CTAs)
  uint64_t* xfull_bar = tempty_bar + 2;          // [1] X of this unit is in TMEM (leader's copy collects both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + 1);
  float* sb1 = reinterpret_cast<float*>(smem + CH_VEC_OFF);
  float* scs = sb1 + CH_D;
  float* sb2 = scs + CH_MAX_N2;
  float2* sstat = reinterpret_cast<float2*>(smem + CH_STAT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_mp = (p.M + 2 * CH_BM - 1) / (2 * CH_BM);
  const int num_units = num_mp * p.n_split;
  const int n2_slice = p.N2 / p.n_split;        // columns of out2 per unit
  const int nch2 = n2_slice / CH_BN;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);   // one arrive per epilogue warp of both CTAs
    }
    mbar_init(xfull_bar, 16);   // one arrive per epilogue warp of both CTAs
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair; u < num_units; u += num_pairs) {
        const int mp = u / p.n_split, sl = u - mp * p.n_split;
        const int arow = mp * 2 * CH_BM + int(rank) * CH_BM;
        // first GEMM: 4 chunks of 128 columns, 8 k-blocks each
        for (int c1 = 0; c1 < CH_D / CH_BN; ++c1) {
          for (int kb = 0; kb < CH_D / CH_BK; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * CH_STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CH_STAGE_BYTES * 2);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (p.a_kblock) {
              const int k0 = kb * CH_BK;
              tma_load_3d_pair(sa, &tmap_a1, bar, k0 % p.a_kblock, arow, k0 / p.a_kblock, kEvictNormal);
            } else {
              tma_load_2d_pair(sa, &tmap_a1, bar, kb * CH_BK, arow, kEvictNormal);
            }
            tma_load_2d_pair(sa + CH_A_BYTES, &tmap_w1, bar, kb * CH_BK, c1 * CH_BN + int(rank) * CH_HALF, kEvictLast);
            if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        // second GEMM: weights only, two k-blocks per stage
        for (int c2 = 0; c2 < nch2; ++c2) {
          const int wrow = sl * n2_slice + c2 * CH_BN + int(rank) * CH_HALF;
          for (int kp = 0; kp < CH_D / CH_BK / 2; ++kp) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * CH_STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CH_W_BYTES * 4);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(sa, &tmap_w2, bar, (2 * kp) * CH_BK, wrow, kEvictLast);
            tma_load_2d_pair(sa + CH_W_BYTES, &tmap_w2, bar, (2 * kp + 1) * CH_BK, wrow, kEvictLast);
            if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    constexpr uint32_t idesc = make_idesc_bf16(2 * CH_BM, CH_BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0, xphase = 0;
    for (int u = pair; u < num_units; u += num_pairs) {
      for (int c1 = 0; c1 < CH_D / CH_BN; ++c1) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + CH_TMEM_ACC + as * CH_BN;
        for (int kb = 0; kb < CH_D / CH_BK; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * CH_STAGE_BYTES);
            const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(a_addr + CH_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < CH_BK / 16; ++k)
              umma_ss_pair(tacc, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == CH_D / CH_BK - 1) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      // X of both CTAs has been written to TMEM by the first epilogue
      mbar_wait(xfull_bar, xphase);
      xphase ^= 1;
      tc_fence_after();
      const uint32_t tx = tmem_base + CH_TMEM_X;
      for (int c2 = 0; c2 < nch2; ++c2) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + CH_TMEM_ACC + as * CH_BN;
        for (int kp = 0; kp < CH_D / CH_BK / 2; ++kp) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_addr = smem_u32(smem + stage * CH_STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t db = make_smem_desc_sw128(w_addr + h * CH_W_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < CH_BK / 16; ++k) {
                const int kk = (2 * kp + h) * (CH_BK / 16) + k;   // K step of 16 elements = 8 packed TMEM columns
                umma_ts_pair(tacc, tx + kk * 8, db + uint64_t(k * 2), idesc, kk > 0 ? 1u : 0u);
              }
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kp == CH_D / CH_BK / 2 - 1) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue
    // Two warps per TMEM lane quarter; each takes 64 of a chunk's 128 columns.  A warp pulls its 64 accumulator columns
    // into registers in one go and hands the stage back AT ONCE (one cheap arrive per warp, no fence: only tensor memory
    // changes hands), so the tensor pipe never waits for epilogue math or stores.
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = warp >> 2;          // which 64-column half of every 128-column chunk
    const int et = threadIdx.x;
    const int trow = q * 32 + lane;      // row within this CTA's 128 rows == TMEM lane
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const uint32_t xfull_leader = mapa_shared(smem_u32(xfull_bar), 0);
    // Output path: a thread owns an accumulator ROW, so it stages its 64 columns (128 B) into the warp's 32 x 64 tile
    // (16 B chunks XOR-swizzled by row: conflict-free both ways) and the warp writes the tile out with 16-byte stores,
    // 8 lanes per 128-byte row segment = 4 full lines per instruction.
    uint8_t* stg = smem + CH_STG_OFF + warp * 4096;
    const uint32_t stg_u32 = smem_u32(stg);
    const uint32_t sstat_u32 = smem_u32(sstat);
    auto stage32 = [&](const uint32_t* pk, int h) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(stg_u32 + lane * 128 + ((((h << 2) + j) ^ (lane & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2],
                     pk[4 * j + 3]);
    };
    // rows [wrow0, wrow0 + 32) x 64 columns starting at `g` (= address of [wrow0][first column]), row stride ld
    auto flush64 = [&](__nv_bfloat16* g, long long ld, int wrow0) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3), c = lane & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(stg_u32 + r * 128 + ((c ^ (r & 7)) << 4)));
        if (wrow0 + r < p.M) *reinterpret_cast<uint4*>(g + size_t(r) * ld + c * 8) = v;
      }
      __syncwarp();
    };
    // first GEMM's bias once per CTA
    for (int i = et; i < CH_D; i += CH_EPI) sb1[i] = p.b1 ? __bfloat162float(p.b1[i]) : 0.f;
    int as = 0;
    uint32_t aphase = 0;
    int cur_sl = -1;
    for (int u = pair; u < num_units; u += num_pairs) {
      const int mp = u / p.n_split, sl = u - mp * p.n_split;
      const int row0 = mp * 2 * CH_BM + int(rank) * CH_BM;   // first row of this CTA
      const int row = row0 + trow;
      const bool row_ok = row < p.M;
      const int wrow0 = row0 + q * 32;                       // first row of this warp's 32-row store tile
      if (sl != cur_sl) {   // second GEMM's folded vectors for this column slice
        asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI) : "memory");   // nobody still reads the previous slice's vectors
        for (int i = et; i < n2_slice; i += CH_EPI) {
          scs[i] = p.csum[sl * n2_slice + i];
          sb2[i] = p.b2[sl * n2_slice + i];
        }
        cur_sl = sl;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI) : "memory");
      // ---- first epilogue: X = resid + acc + b1 -> bf16 -> (global x_out) + TMEM; row statistics
      const __nv_bfloat16* rrow = p.resid + size_t(row_ok ? row : 0) * p.ldr + half * CH_HALF;
#pragma unroll 1
      for (int c1 = 0; c1 < CH_D / CH_BN; ++c1) {
        const int col0 = c1 * CH_BN + half * CH_HALF;   // first of my 64 columns of X
        // my row's residual, fetched before the accumulator is awaited.  (Re-fetching the next chunk's halves into the
        // same registers as soon as a piece has consumed them measured SLOWER: 77.1 vs 74.9 us per link.)
        uint4 rs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rs[i] = row_ok ? reinterpret_cast<const uint4*>(rrow + c1 * CH_BN)[i] : make_uint4(0, 0, 0, 0);
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        uint32_t r[64];
        const uint32_t tacc = tmem_base + CH_TMEM_ACC + as * CH_BN + lane_off + half * CH_HALF;
        tmem_ld_x32(tacc, r);
        tmem_ld_x32(tacc + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_nofence(tempty_leader + uint32_t(as) * 8u);
        float s1 = 0.f, s2 = 0.f;   // sum / sum of squares of my 64 X values of this chunk
        {
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 rv = rs[4 * pc + i];
              const uint32_t xs[4] = {rv.x, rv.y, rv.z, rv.w};
              const float4 ba = *reinterpret_cast<const float4*>(sb1 + col0 + 32 * pc + 8 * i);
              const float4 bb = *reinterpret_cast<const float4*>(sb1 + col0 + 32 * pc + 8 * i + 4);
              const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = 32 * pc + 8 * i + 2 * j;
                const float v0 = bf16_lo(xs[j]) + (__uint_as_float(r[e]) + bs[2 * j]);
                const float v1 = bf16_hi(xs[j]) + (__uint_as_float(r[e + 1]) + bs[2 * j + 1]);
                const uint32_t w = pack_bf16x2(v0, v1);
                pk[4 * i + j] = w;
                const float x0 = bf16_lo(w), x1 = bf16_hi(w);   // the rounded values are what LayerNorm sees
                s1 += x0 + x1;
                s2 = fmaf(x0, x0, fmaf(x1, x1, s2));
              }
            }
            tmem_st_x16(tmem_base + CH_TMEM_X + lane_off + ((col0 + 32 * pc) >> 1), pk);
            if (p.store_x && sl == 0) stage32(pk, pc);
          }
          if (p.store_x && sl == 0) flush64(p.x_out + size_t(wrow0) * p.ldx + col0, p.ldx, wrow0);
        }
        sts_f2(sstat_u32 + ((c1 * 2 + half) * CH_BM + trow) * 8, s1, s2);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_nofence(xfull_leader);
      // Row statistics: the eight 64-column partial sums of a row (two threads x four chunks) added in column order,
      // identically by both threads of the row
      asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI) : "memory");
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2 * CH_D / CH_BN; ++c) {
        const float2 o = lds_f2(sstat_u32 + (c * CH_BM + trow) * 8);
        t1 += o.x;
        t2 += o.y;
      }
      const float mean = t1 * (1.f / CH_D);
      const float var = fmaxf(t2 * (1.f / CH_D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.ln_eps);
      const float nmr = -mean * rstd;
      // ---- second epilogue: out2 = act(rstd * (acc - mean * csum) + b2')
#pragma unroll 1
      for (int c2 = 0; c2 < nch2; ++c2) {
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        uint32_t r[64];
        const uint32_t tacc = tmem_base + CH_TMEM_ACC + as * CH_BN + lane_off + half * CH_HALF;
        tmem_ld_x32(tacc, r);
        tmem_ld_x32(tacc + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_nofence(tempty_leader + uint32_t(as) * 8u);
        if (++as == 2) { as = 0; aphase ^= 1; }
        const int lc0 = c2 * CH_BN + half * CH_HALF;   // first of my 64 columns within the slice
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const int lc = lc0 + 32 * pc;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 cs = *reinterpret_cast<const float4*>(scs + lc + 4 * i);
            const float4 bb = *reinterpret_cast<const float4*>(sb2 + lc + 4 * i);
            float v0, v1, v2, v3;   // rstd * acc + (nmr * csum + b2), two columns per packed FMA
            ffma2(v0, v1, cs.x, cs.y, nmr, nmr, bb.x, bb.y);
            ffma2(v2, v3, cs.z, cs.w, nmr, nmr, bb.z, bb.w);
            ffma2(v0, v1, __uint_as_float(r[32 * pc + 4 * i]), __uint_as_float(r[32 * pc + 4 * i + 1]), rstd, rstd, v0, v1);
            ffma2(v2, v3, __uint_as_float(r[32 * pc + 4 * i + 2]), __uint_as_float(r[32 * pc + 4 * i + 3]), rstd, rstd, v2, v3);
            if (ACT == GEMM_ACT_GELU_ERF) {
              v0 = gelu_erf(v0); v1 = gelu_erf(v1); v2 = gelu_erf(v2); v3 = gelu_erf(v3);
            } else if (ACT == GEMM_ACT_GELU_TANH) {
              v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); v2 = gelu_tanh(v2); v3 = gelu_tanh(v3);
            } else if (ACT == GEMM_ACT_RELU) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            pk[2 * i] = pack_bf16x2(v0, v1);
            pk[2 * i + 1] = pack_bf16x2(v2, v3);
          }
          stage32(pk, pc);
        }
        {
          const int c0 = sl * n2_slice + lc0;   // first of the 64 staged columns (64 | col_block: never straddles a block)
          __nv_bfloat16* g = p.col_block
                                 ? p.out2 + size_t(c0 / p.col_block) * p.col_block_stride + size_t(wrow0) * p.ldc + c0 % p.col_block
                                 : p.out2 + size_t(wrow0) * p.ldc + c0;
          flush64(g, p.ldc, wrow0);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace bya

extern "C" int bya_gemm_ln_gemm_bf16(void* stream, const void* A1, int lda, const void* W1, int ldw1, const void* W2f,
                                     int ldw2, const ByaChainArgs* args) {
  using namespace bya;
  if (!A1 || !W1 || !W2f || !args) return BYA_ERR_SHAPE;
  const ByaChainArgs& a = *args;
  if (a.M <= 0 || a.N2 <= 0 || a.N2 > CH_MAX_N2 || a.N2 % CH_BN || !a.resid || !a.x_out || !a.out2 || !a.csum || !a.b2)
    return BYA_ERR_SHAPE;
  int n_split = a.n_split > 0 ? a.n_split : 1;
  if ((a.N2 / CH_BN) % n_split) return BYA_ERR_SHAPE;
  // several column slices of one row tile run in different CTAs: the residual they all read must not be the X one writes
  if (n_split > 1 && a.store_x && a.resid == a.x_out) return BYA_ERR_SHAPE;
  if (lda % 8 || ldw1 % 8 || ldw2 % 8 || a.ldr % 8 || a.ldx % 8 || a.ldc % 8) return BYA_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(A1) | reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(W2f) |
       reinterpret_cast<uintptr_t>(a.resid) | reinterpret_cast<uintptr_t>(a.x_out) | reinterpret_cast<uintptr_t>(a.out2)) & 15)
    return BYA_ERR_ALIGN;
  if (a.col_block && (a.col_block % 64 || a.N2 % a.col_block || a.col_block_stride % 8)) return BYA_ERR_SHAPE;
  int a_kblock = a.a_kblock == CH_D ? 0 : a.a_kblock;
  if (a_kblock && (a_kblock % CH_BK || CH_D % a_kblock || a.a_kblock_stride % 8)) return BYA_ERR_SHAPE;
  int col_block = a.col_block == a.N2 ? 0 : a.col_block;

  CUtensorMap ta, tw1, tw2;
  int rc = a_kblock ? bya_host::encode_tmap_bf16(&ta, A1, a_kblock, a.M, uint64_t(lda) * 2, CH_BK, CH_BM, CH_D / a_kblock,
                                                 uint64_t(a.a_kblock_stride) * 2)
                    : bya_host::encode_tmap_bf16(&ta, A1, CH_D, a.M, uint64_t(lda) * 2, CH_BK, CH_BM);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tw1, W1, CH_D, CH_D, uint64_t(ldw1) * 2, CH_BK, CH_HALF);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tw2, W2f, CH_D, a.N2, uint64_t(ldw2) * 2, CH_BK, CH_HALF);
  if (rc) return rc;

  ChainParams p;
  p.M = a.M;
  p.N2 = a.N2;
  p.act = a.act;
  p.n_split = n_split;
  p.ldr = a.ldr;
  p.ldx = a.ldx;
  p.ldc = a.ldc;
  p.col_block_stride = a.col_block_stride;
  p.x_out = a.x_out;
  p.out2 = a.out2;
  p.a_kblock = a_kblock;
  p.col_block = col_block;
  p.store_x = a.store_x;
  p.ln_eps = a.ln_eps;
  p.b1 = a.b1;
  p.resid = a.resid;
  p.csum = a.csum;
  p.b2 = a.b2;

  if (a.act < 0 || a.act > 3) return BYA_ERR_SHAPE;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const ChainParams);
  static const Kern kerns[4] = {gemm_ln_gemm_kernel<0>, gemm_ln_gemm_kernel<1>, gemm_ln_gemm_kernel<2>, gemm_ln_gemm_kernel<3>};
  Kern kern = kerns[a.act];
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[a.act]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM) != cudaSuccess) return BYA_ERR_CUDA;
    attr_set[a.act] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_pairs = 0;
  if (!max_pairs) {
    cfg.gridDim = dim3(bya_host::num_sms());
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = bya_host::num_sms() / 2;
    max_pairs = n;
  }
  const int units = ((a.M + 2 * CH_BM - 1) / (2 * CH_BM)) * n_split;
  const int pairs = units < max_pairs ? units : max_pairs;
  cfg.gridDim = dim3(2 * pairs);
  if (cudaLaunchKernelEx(&cfg, kern, ta, tw1, tw2, p) != cudaSuccess) return BYA_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
