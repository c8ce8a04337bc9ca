// Flash-style multi-head attention forward, head_dim 64, bf16, no mask, non-causal.   sm_100a only.
//   O[b, n, h*64 : h*64+64] = softmax(Q_h K_h^T * scale) V_h
// Replaces F.scaled_dot_product_attention as called by diffusers' CogVideoXAttnProcessor2_0 for the joint
// [text; video] self-attention (models/transformer.py:241-245) and by the router's spatial attention
// (models/router.py:474-476, K10).  Q/K/V are read in place from the (fused) projection output — row-major
// [rows, ld] with head h at columns h*64.. — so no head-major copy is ever made.
//
// One CTA = one (batch, head, 256 query rows) = two 128-row query tiles; KV tiles of 128 keys.  Warp roles:
//   warps 0-3    softmax for query tile 0   (one thread per query row; tcgen05.ld S -> exp2 -> tcgen05.st P)
//   warps 4-7    softmax for query tile 1
//   warp 8       TMA producer (one thread): Q once, then 4-deep rings of K and V tiles
//   warp 9       MMA issuer (one thread):   S_t = Q_t K^T (tcgen05.mma SS, fp32 in TMEM),
//                                           O_t += P_t V  (tcgen05.mma TS: P read from TMEM, V consumed MN-major
//                                           straight from its natural [key, d] layout)
//   warp 10      TMEM allocator
// The issuing warps carry the HIGHEST warp ids on purpose: the sub-partition arbiter prefers higher ids, and a
// starved MMA issuer (it shares a sub-partition with two softmax warps) was the critical path of the whole kernel
// (clock64 timeline: ~2000 cycles per KV tile spent in its ~270-instruction loop).
// TMEM (512 columns): S_t [128t, 128t+128)   P_t [256+64t, +64) (packed bf16)   O_t [384+64t, +64).
// S and P live in DIFFERENT columns, so the softmax warps hand S back (`s_free`) as soon as the scores are in
// registers and Q K^T of KV tile j+1 runs on the tensor pipe while the exponentials of tile j are being computed:
// in steady state neither side waits for the other.
// d=64 attention is exp-bound (MUFU.EX2: 8 cycles per warp instruction, measured; the MMAs of the same scores take
// 4): POLY16 of every 16 exponentials are evaluated on the FMA pipe instead (Cody-Waite split + degree-3 minimax
// polynomial, max rel. error 8.8e-5, far below the bf16 rounding of P).  Row sums are accumulated in registers.
// Two softmax variants:
//   general  (bya_attention_d64):         running max with lazy rescaling (O in TMEM is only rescaled when the max
//                                         grows by more than 2^8); the whole 128-score row is held in registers.
//   bounded  (bya_attention_d64_bounded): the caller guarantees |q.k| <= bound (log2 units, bound <= 64) — true for the
//                                         joint self-attention, whose q and k are LayerNorm-ed per head (|q|,|k| <=
//                                         8 max|gamma| + |beta|) — so softmax(s) = 2^s / sum 2^s needs no max, no
//                                         subtraction and no rescaling: the scores go straight from the TMEM
//                                         load into the exponentials.
#include "common.cuh"
#include "../../include/bya.h"

#include <cstdlib>
#include <type_traits>

namespace bya {

constexpr int FA_D = 64;
constexpr int FA_BM = 128;         // query rows per tile (2 tiles per CTA)
constexpr int FA_BN = 128;         // keys per KV tile
__host__ __device__ constexpr int fa_stages(int qt) { return qt == 1 ? 3 : 4; }   // K ring depth == V ring depth
constexpr int FA_TILE_BYTES = FA_BM * FA_D * 2;  // 16 KB
// qt query tiles per CTA, nt softmax threads per query row, + one issuing warpgroup
__host__ __device__ constexpr int fa_threads(int qt, int nt) { return 128 * qt * nt + 128; }
// V ring depth: with two CTAs per SM (qt == 1) each may use at most ~113 KB: 1 Q + 3 K + 2 V tiles
__host__ __device__ constexpr int fa_vstages(int qt) { return qt == 1 ? 2 : 4; }
__host__ __device__ constexpr int fa_smem(int qt) {
  return (qt + fa_stages(qt) + fa_vstages(qt)) * FA_TILE_BYTES + 512 + 1024;
}

struct FaArgs {
  int seq;        // rows per batch element (queries == keys)
  int seq_stride; // first row of batch element b is b * seq_stride (>= seq: padding rows between sequences are skipped)
  int heads;
  int batch;
  int ldo;        // row stride of O in elements
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  // PUSH exchange of sequence parallelism (rows_per_peer > 0): row n is stored to out_peers[n / rows_per_peer] at local
  // row n % rows_per_peer — the owner rank's out-projection operand, possibly over NVLink — instead of `out`
  __nv_bfloat16* out_peers[BYA_MAX_PEERS];
  int rows_per_peer;
  long long* trace;   // debug (BYA_FA_TRACE): [32 iterations][16 events] clock64 stamps of CTA (0,0,0), else NULL
};

// debug timeline: event e of iteration j (only CTA 0, only while j < 32)
#define FA_TRACE(e, j)                                                                              \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 32 && lane == 0)  \
      p.trace[(j) * 16 + (e)] = clock64();                                                          \
  } while (0)

template <int R>
BYA_DEVICE void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
BYA_DEVICE void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

BYA_DEVICE float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// (t0, t1) = (a0, a1) * c + d   as one packed FFMA2
BYA_DEVICE void fma2(float& t0, float& t1, float a0, float a1, float c, float d) {
  asm("{\n\t.reg .b64 va, vc, vd;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vc, {%4, %4};\n\t"
      "mov.b64 vd, {%5, %5};\n\t"
      "fma.rn.f32x2 va, va, vc, vd;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(t0), "=f"(t1)
      : "f"(a0), "f"(a1), "f"(c), "f"(d));
}
// (t0, t1) = (a0, a1) * (b0, b1) + c
BYA_DEVICE void fma2v(float& t0, float& t1, float a0, float a1, float b0, float b1, float c) {
  asm("{\n\t.reg .b64 va, vb, vc;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vb, {%4, %5};\n\t"
      "mov.b64 vc, {%6, %6};\n\t"
      "fma.rn.f32x2 va, va, vb, vc;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(t0), "=f"(t1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
// (a0, a1) += (b0, b1)
BYA_DEVICE void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 va, vb;\n\t"
      "mov.b64 va, {%0, %1};\n\t"
      "mov.b64 vb, {%2, %3};\n\t"
      "add.rn.f32x2 va, va, vb;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}
BYA_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for |x| < 126 on the FMA / ALU pipes (no MUFU): x = n + f, n = floor(x) via a round-down magic add,
// 2^f ~ degree-3 minimax polynomial on [0,1), exponent patched in with one integer shift-add.
template <bool CLAMP>   // CLAMP: inputs may be below -126 (the result then saturates at 2^-126 instead of wrapping)
BYA_DEVICE void ex2_poly2(float& y0, float& y1, float x0, float x1) {
  if (CLAMP) {
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
  }
  float t0, t1, n0, n1, f0, f1, p0, p1;
  asm("{\n\t.reg .b64 va, vb;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vb, {%4, %4};\n\t"
      "add.rm.ftz.f32x2 va, va, vb;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(t0), "=f"(t1)
      : "f"(x0), "f"(x1), "f"(12582912.0f));
  asm("{\n\t.reg .b64 va, vb;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vb, {%4, %4};\n\t"
      "add.rn.ftz.f32x2 va, va, vb;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(n0), "=f"(n1)
      : "f"(t0), "f"(t1), "f"(-12582912.0f));
  asm("{\n\t.reg .b64 va, vb;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vb, {%4, %5};\n\t"
      "sub.rn.ftz.f32x2 va, va, vb;\n\t"
      "mov.b64 {%0, %1}, va;\n\t}\n"
      : "=f"(f0), "=f"(f1)
      : "f"(x0), "f"(x1), "f"(n0), "f"(n1));
  fma2v(p0, p1, f0, f1, 0.0771190897f, 0.0771190897f, 0.2275643945f);
  fma2v(p0, p1, p0, p1, f0, f1, 0.6951461434f);
  fma2v(p0, p1, p0, p1, f0, f1, 1.0f);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// POLY16: exponentials per group of 16 evaluated by ex2_poly2 (0,4,8,12).  BOUNDED: see the header.
// NT: softmax threads per query row.  NT = 2 (bounded only: without a running max the two halves of a row share
// nothing but the final row sum) doubles the warps per SM sub-partition.  Measured: no gain (1066 vs 1055 TFLOP/s) —
// the exp phase is bound by dispatch/pipe occupancy (every f32x2 / F2FP instruction holds its pipe for 2 cycles), not by
// latency hiding — so NT = 1 is the default; the variant is kept for the next round of tuning.
// QT: 128-row query tiles per CTA.  QT = 2: one CTA per SM, the two tiles ping-pong on the tensor pipe but their softmax
// warps fall into lock-step (the MMA warp serves them in order, the arbiter prefers the higher warp id), so MUFU idles
// while BOTH load / wait / store.  QT = 1: two independent CTAs per SM (256 TMEM columns and half the registers each)
// whose phases drift apart freely.
template <int POLY16, bool BOUNDED, int NT, int QT>
__global__ void __launch_bounds__(fa_threads(QT, NT), QT == 1 ? 2 : 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ FaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  constexpr int FA_STAGES = fa_stages(QT);
  constexpr int FA_VSTAGES = fa_vstages(QT);
  uint8_t* sQ = smem;                                   // [QT][16 KB]
  uint8_t* sK = sQ + QT * FA_TILE_BYTES;                // [STAGES][16 KB]
  uint8_t* sV = sK + FA_STAGES * FA_TILE_BYTES;         // [STAGES][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_VSTAGES * FA_TILE_BYTES);
  uint64_t* q_full = bars;                    // [1]
  uint64_t* k_full = q_full + 1;              // [STAGES]
  uint64_t* k_empty = k_full + FA_STAGES;     // [STAGES]
  uint64_t* v_full = k_empty + FA_STAGES;     // [STAGES]
  uint64_t* v_empty = v_full + FA_VSTAGES;    // [VSTAGES]
  uint64_t* s_full = v_empty + FA_VSTAGES;     // [2]  S_t(j) = Q_t K(j)^T landed in TMEM
  uint64_t* s_free = s_full + 2;              // [2]  S_t(j) is in the softmax registers: S_t may be overwritten
  uint64_t* p_full = s_free + 2;              // [2]  P_t(j) written to TMEM
  uint64_t* pv_done = p_full + 2;             // [2]  O_t += P_t(j) V(j) complete: P_t free, O_t stable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: role/TMEM addresses
  const int lane = threadIdx.x & 31;                                // stay in uniform registers
  const int qblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.seq_stride;        // first row of this batch element in the [rows, ld] matrices
  const int q0 = qblk * (QT * FA_BM);           // first query row (within the batch element)
  const int n_kv = (p.seq + FA_BN - 1) / FA_BN;
  const int col = head * FA_D;

  constexpr int W0 = 4 * QT * NT;   // first warp of the issuing warpgroup
  if (warp == W0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == W0 + 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < FA_VSTAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < QT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4 * NT);    // one arrive per softmax warp
      mbar_init(&p_full[t], 4 * NT);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  constexpr int kTmemCols = 256 * QT;
  if (warp == W0 + 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColP = 128 * QT, kColO = 192 * QT;

  if (warp >= W0) {
    // register pool of the launch: (65536 / CTAs per SM / threads) rounded down to 8 per thread
    setmaxnreg_dec<((NT == 1 && QT == 2) ? 56 : 40)>();
    if (warp == W0) {
      // ---------------------------------------------------------------- TMA producer (one elected thread)
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, QT * FA_TILE_BYTES);
        tma_load_2d(sQ, &tmap_q, q_full, col, row_base + q0, kEvictFirst);
        if (QT == 2) tma_load_2d(sQ + FA_TILE_BYTES, &tmap_q, q_full, col, row_base + q0 + FA_BM, kEvictFirst);
        // order: K(0), then K(j+1), V(j): K runs one tile ahead because S(j+1) is computed before O += P(j) V(j)
        mbar_arrive_expect_tx(&k_full[0], FA_TILE_BYTES);
        tma_load_2d(sK, &tmap_k, &k_full[0], col, row_base, kEvictLast);
        int ks = 1, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) {
            mbar_wait(&k_empty[ks], kph ^ 1);
            mbar_arrive_expect_tx(&k_full[ks], FA_TILE_BYTES);
            tma_load_2d(sK + ks * FA_TILE_BYTES, &tmap_k, &k_full[ks], col, row_base + (j + 1) * FA_BN, kEvictLast);
            if (++ks == FA_STAGES) { ks = 0; kph ^= 1; }
          }
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], FA_TILE_BYTES);
          tma_load_2d(sV + vs * FA_TILE_BYTES, &tmap_v, &v_full[vs], col, row_base + j * FA_BN, kEvictLast);
          if (++vs == FA_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == W0 + 1) {
      // ---------------------------------------------------------------- MMA issuer
      // The whole warp runs the warp-uniform loop and polls the barriers; one elected lane issues (ptxas keeps
      // descriptors in uniform registers only when control flow is provably uniform).
      constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BM, FA_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BM, FA_D, 0, 1);  // B = V is MN-major (d contiguous)
      const uint64_t dq0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dq1 = make_smem_desc_sw128(smem_u32(sQ + (QT - 1) * FA_TILE_BYTES), 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV), 1024, 1024);
      constexpr uint64_t kStageStep = FA_TILE_BYTES >> 4;   // descriptor address field is in 16 B units
      const uint32_t tS0 = tmem_base + kColS, tS1 = tmem_base + kColS + 128;
      const uint32_t tP0 = tmem_base + kColP, tP1 = tmem_base + kColP + 64;
      const uint32_t tO0 = tmem_base + kColO, tO1 = tmem_base + kColO + 64;

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) umma_ss(tS0, dq0 + uint64_t(2 * k), dk0 + uint64_t(2 * k), idesc_qk, k != 0);
        umma_commit(&s_full[0]);
        if (QT == 2) {
#pragma unroll
          for (int k = 0; k < FA_D / 16; ++k) umma_ss(tS1, dq1 + uint64_t(2 * k), dk0 + uint64_t(2 * k), idesc_qk, k != 0);
          umma_commit(&s_full[1]);
        }
        umma_commit(&k_empty[0]);
      }
      __syncwarp();
      int ks = 1, vs = 0;          // ring slots of K(j+1) and V(j)
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t jp = j & 1;
        if (j + 1 < n_kv) {
          // S_t(j+1) = Q_t K(j+1)^T as soon as the softmax warps hold S_t(j) in registers
          const uint64_t dk = dk0 + uint64_t(ks) * kStageStep;
          mbar_wait(&k_full[ks], kph);
          mbar_wait(&s_free[0], jp);
          FA_TRACE(8, j);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < FA_D / 16; ++k)
              umma_ss(tS0, dq0 + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_qk, k != 0);
            umma_commit(&s_full[0]);
            if (QT == 1) umma_commit(&k_empty[ks]);
          }
          __syncwarp();
          if (QT == 2) {
            mbar_wait(&s_free[1], jp);
            FA_TRACE(9, j);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < FA_D / 16; ++k)
                umma_ss(tS1, dq1 + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_qk, k != 0);
              umma_commit(&s_full[1]);
              umma_commit(&k_empty[ks]);
            }
            __syncwarp();
          }
          if (++ks == FA_STAGES) { ks = 0; kph ^= 1; }
        }
        // O_t += P_t(j) V(j)
        const uint64_t dv = dv0 + uint64_t(vs) * kStageStep;
        const uint32_t acc = j > 0;
        mbar_wait(&v_full[vs], vph);
        mbar_wait(&p_full[0], jp);
        FA_TRACE(10, j);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < FA_BN / 16; ++k)
            umma_ts(tO0, tP0 + k * 8, dv + uint64_t(128 * k), idesc_pv, acc | (k != 0));
          umma_commit(&pv_done[0]);
          if (QT == 1) umma_commit(&v_empty[vs]);
        }
        __syncwarp();
        if (QT == 2) {
          mbar_wait(&p_full[1], jp);
          FA_TRACE(11, j);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < FA_BN / 16; ++k)
              umma_ts(tO1, tP1 + k * 8, dv + uint64_t(128 * k), idesc_pv, acc | (k != 0));
            umma_commit(&pv_done[1]);
            umma_commit(&v_empty[vs]);
          }
          __syncwarp();
        }
        FA_TRACE(12, j);
        if (++vs == FA_VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (one thread per query row)
    static_assert(NT == 1 || BOUNDED, "two threads per row need the max-free softmax");
    // register pool of the launch: 65536 / threads rounded down to 8 -> 168 (NT=1) / 96 (NT=2) per thread
    setmaxnreg_inc<(QT == 1 ? 216 : (NT == 1 ? 224 : 104))>();
    constexpr int NC = FA_BN / NT;   // score columns per thread
    constexpr int OC = FA_D / NT;    // O columns per thread (final normalisation + store)
    const int t = (warp / (4 * NT));   // query tile 0 / 1
    const int h = (warp >> 2) % NT;  // which part of the row
    const int q = warp & 3;          // TMEM lane quarter
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + kColS + t * 128 + h * NC + lane_off;
    const uint32_t tP = tmem_base + kColP + t * 64 + h * (NC / 2) + lane_off;
    const uint32_t tO = tmem_base + kColO + t * 64 + h * OC + lane_off;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;   // row-sum accumulators (two packed f32x2)

    if constexpr (BOUNDED) {
      // P = 2^s, no max, no scaling: exponentiate straight off the TMEM load, P stored back in 32-key chunks.
      auto exp_chunk = [&](const uint32_t* sc, uint32_t* pk, int col0, int valid, auto masked) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float a0 = __uint_as_float(sc[i]), a1 = __uint_as_float(sc[i + 1]);
          float a2 = __uint_as_float(sc[i + 2]), a3 = __uint_as_float(sc[i + 3]);
          if (decltype(masked)::value) {   // last KV tile only: keys past the sequence end get weight 2^-100
            a0 = (col0 + i < valid) ? a0 : -100.f;
            a1 = (col0 + i + 1 < valid) ? a1 : -100.f;
            a2 = (col0 + i + 2 < valid) ? a2 : -100.f;
            a3 = (col0 + i + 3 < valid) ? a3 : -100.f;
          }
          if ((i & 15) < POLY16) {   // compile-time after unrolling
            ex2_poly2<false>(a0, a1, a0, a1);
            ex2_poly2<false>(a2, a3, a2, a3);
          } else {
            a0 = ex2(a0);
            a1 = ex2(a1);
            a2 = ex2(a2);
            a3 = ex2(a3);
          }
          add2(l0, l1, a0, a1);
          add2(l2, l3, a2, a3);
          pk[i / 2] = pack_bf16x2(a0, a1);
          pk[i / 2 + 1] = pack_bf16x2(a2, a3);
        }
      };
      // one KV tile: S_t(j) -> P_t(j)
      auto tile = [&](int j, int valid, auto masked) {
        if (q == 0 && h == 0) FA_TRACE(t * 4 + 0, j);
        uint32_t s[NC];
#pragma unroll
        for (int i = 0; i < NC; i += 32) tmem_ld_x32(tS + i, s + i);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);   // hand S_t back at once: Q K^T of tile j+1 overlaps these exps
        if (q == 0 && h == 0) FA_TRACE(t * 4 + 1, j);
        uint32_t pk[NC / 2];
        exp_chunk(s, pk, 0, valid, masked);
        if (NC > 32) exp_chunk(s + 32, pk + 16, 32, valid, masked);
        // P_t(j-1) must have been consumed by O_t += P_t(j-1) V(j-1) before it is overwritten.  That MMA was issued
        // when the previous iteration ended (and queues behind the other tile's): waiting here, well into this
        // iteration, costs nothing; waiting before the first exponential stalled tile 1 for ~500 cycles per KV tile.
        if (q == 0 && h == 0) FA_TRACE(t * 4 + 2, j);
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
        }
        if (q == 0 && h == 0) FA_TRACE(t * 4 + 3, j);
        if (NC > 32) tmem_st_x32(tP, pk);
        else tmem_st_x16(tP, pk);
        if (NC > 64) {
          exp_chunk(s + 64, pk + 32, 64, valid, masked);
          tmem_st_x16(tP + 32, pk + 32);
          exp_chunk(s + 96, pk + 48, 96, valid, masked);
          tmem_st_x16(tP + 48, pk + 48);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      };
      for (int j = 0; j < n_kv; ++j) {
        const int valid = p.seq - j * FA_BN - h * NC;   // valid keys among this thread's NC columns
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        if (valid >= NC) tile(j, valid, std::false_type{});
        else tile(j, valid, std::true_type{});
      }
    } else {
      const float c = p.scale_log2;
      float m = -INFINITY;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[128];
        tmem_ld_x32(tS, s);
        tmem_ld_x32(tS + 32, s + 32);
        tmem_ld_x32(tS + 64, s + 64);
        tmem_ld_x32(tS + 96, s + 96);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);   // the tensor pipe may start S_t(j+1) now
        const int valid = p.seq - j * FA_BN;
        if (valid < FA_BN) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          mx[0] = max3(mx[0], __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx[1] = max3(mx[1], __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
          mx[2] = max3(mx[2], __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
          mx[3] = max3(mx[3], __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
        }
        const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (j == 0) {
          m = mt;
        } else {
          const bool grow = (mt - m) * c > 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // O_t may only be touched once P(j-1) V(j-1) has landed (issued a whole softmax ago)
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
            const float mn = fmaxf(m, mt);
            const float alpha = ex2((m - mn) * c);
            m = mn;
            uint32_t o[64];
            tmem_ld_x32(tO, o);
            tmem_ld_x32(tO + 32, o + 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO, o);
            tmem_st_x32(tO + 32, o + 32);
            l0 *= alpha;
            l1 *= alpha;
            l2 *= alpha;
            l3 *= alpha;
          }
        }
        const float nmc = -m * c;
        uint32_t pk[64];
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          float a0, a1, a2, a3;
          fma2(a0, a1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), c, nmc);
          fma2(a2, a3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), c, nmc);
          if ((i & 15) < POLY16) {   // compile-time after unrolling
            ex2_poly2<true>(a0, a1, a0, a1);
            ex2_poly2<true>(a2, a3, a2, a3);
          } else {
            a0 = ex2(a0);
            a1 = ex2(a1);
            a2 = ex2(a2);
            a3 = ex2(a3);
          }
          add2(l0, l1, a0, a1);
          add2(l2, l3, a2, a3);
          pk[i / 2] = pack_bf16x2(a0, a1);
          pk[i / 2 + 1] = pack_bf16x2(a2, a3);
        }
        if (j > 0) {   // P_t(j-1) must have been consumed by O_t += P_t(j-1) V(j-1) before it is overwritten
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_x32(tP, pk);
        tmem_st_x32(tP + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
    }
    // ---- epilogue: O / l -> bf16 -> global
    float l = (l0 + l1) + (l2 + l3);
    if (NT == 2) {   // the two threads of a row exchange their partial row sums through (now idle) K-ring shared memory
      mbar_wait(&pv_done[t], (n_kv - 1) & 1);   // every MMA of this tile has completed: no tile reads sK any more
      float* xch = reinterpret_cast<float*>(sK) + t * 256 + q * 32 + lane;   // [tile][half][row]
      xch[h * 128] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + t * 4 + q) : "memory");
      l += xch[(h ^ 1) * 128];
    }
    mbar_wait(&pv_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    uint32_t o[OC];
#pragma unroll
    for (int i = 0; i < OC; i += 32) tmem_ld_x32(tO + i, o + i);
    tmem_ld_wait();
    const int qrow = q0 + t * FA_BM + q * 32 + lane;
    if (qrow < p.seq) {
      const float inv = 1.0f / l;
      __nv_bfloat16* orow = p.rows_per_peer > 0
                                ? p.out_peers[qrow / p.rows_per_peer] + size_t(qrow % p.rows_per_peer) * p.ldo
                                : p.out + size_t(row_base + qrow) * p.ldo;
      uint4* dst = reinterpret_cast<uint4*>(orow + col + h * OC);
#pragma unroll
      for (int i = 0; i < OC / 8; ++i) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        dst[i] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W0 + 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

using FaKernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const FaArgs);

// Tuning knobs, read once: BYA_FA_POLY16 / BYA_FA_POLY16_BOUNDED (0, 4, 8, 12: exponentials per 16 moved from MUFU
// to the FMA pipe), BYA_FA_NT (bounded kernel only: 1 or 2 softmax threads per query row), BYA_FA_QT / BYA_FA_QT_BOUNDED
// (force 1 or 2 query tiles per CTA; default: by sequence length, see fa_launch).
template <bool BOUNDED>
static FaKernel fa_pick_kernel(int qt, int* threads) {
  int poly = 4, nt = 1;
  if (const char* e = std::getenv(BOUNDED ? "BYA_FA_POLY16_BOUNDED" : "BYA_FA_POLY16")) poly = std::atoi(e);
  if (const char* e = std::getenv("BYA_FA_NT")) nt = (BOUNDED && std::atoi(e) == 2) ? 2 : 1;
  if (qt == 1) nt = 1;
  *threads = fa_threads(qt, nt);
  if (qt == 1) {
    switch (poly) {
      case 0: return fa_fwd_kernel<0, BOUNDED, 1, 1>;
      case 8: return fa_fwd_kernel<8, BOUNDED, 1, 1>;
      default: return fa_fwd_kernel<4, BOUNDED, 1, 1>;
    }
  }
  if constexpr (BOUNDED) {
    if (nt == 2) {
      switch (poly) {
        case 0: return fa_fwd_kernel<0, true, 2, 2>;
        case 4: return fa_fwd_kernel<4, true, 2, 2>;
        case 12: return fa_fwd_kernel<12, true, 2, 2>;
        default: return fa_fwd_kernel<8, true, 2, 2>;
      }
    }
  }
  switch (poly) {
    case 0: return fa_fwd_kernel<0, BOUNDED, 1, 2>;
    case 4: return fa_fwd_kernel<4, BOUNDED, 1, 2>;
    case 12: return fa_fwd_kernel<12, BOUNDED, 1, 2>;
    default: return fa_fwd_kernel<8, BOUNDED, 1, 2>;
  }
}

template <bool BOUNDED>
static int fa_launch(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo, int batch,
                     int seq, int seq_stride, int heads, float scale, void* const* out_peers = nullptr, int n_peers = 0,
                     int rows_per_peer = 0) {
  if (!q || !k || !v || (!out && !out_peers) || batch <= 0 || seq <= 0 || heads <= 0 || seq_stride < seq) return BYA_ERR_SHAPE;
  if (out_peers && (n_peers < 1 || n_peers > BYA_MAX_PEERS || rows_per_peer <= 0 || batch != 1 ||
                    (long long)n_peers * rows_per_peer < seq))
    return BYA_ERR_SHAPE;
  if (ld % 8 || ldo % 8 || ld < heads * FA_D || ldo < heads * FA_D) return BYA_ERR_ALIGN;
  const uint64_t rows = uint64_t(batch - 1) * seq_stride + seq;
  // Short sequences (the router's 1 350-token frames): one 128-row query tile per CTA and two CTAs per SM — a CTA only
  // lives for ~11 KV tiles, so its fill / drain overlaps the other CTA's work (measured 152 vs 175 us per call).  Long
  // sequences: two tiles per CTA sharing every K/V tile (947 vs 890 TFLOP/s at 17 776 tokens).
  static int forced_qt = -1;
  if (forced_qt < 0) {
    const char* e = std::getenv(BOUNDED ? "BYA_FA_QT_BOUNDED" : "BYA_FA_QT");
    forced_qt = e ? (std::atoi(e) == 1 ? 1 : 2) : 0;
  }
  const int qt = forced_qt ? forced_qt : (seq < 4096 ? 1 : 2);
  CUtensorMap tq, tk, tv;
  int rc = bya_host::encode_tmap_bf16(&tq, q, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BM);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tk, k, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BN);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tv, v, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BN);
  if (rc) return rc;
  static FaKernel kern[3] = {nullptr, nullptr, nullptr};
  static int threads[3] = {0, 0, 0};
  if (!kern[qt]) {
    FaKernel kk = fa_pick_kernel<BOUNDED>(qt, &threads[qt]);
    if (cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, fa_smem(qt)) != cudaSuccess)
      return BYA_ERR_CUDA;
    kern[qt] = kk;
  }
  FaArgs a;
  a.seq = seq;
  a.seq_stride = seq_stride;
  a.heads = heads;
  a.batch = batch;
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.rows_per_peer = out_peers ? rows_per_peer : 0;
  for (int i = 0; i < BYA_MAX_PEERS; ++i)
    a.out_peers[i] = (out_peers && i < n_peers) ? reinterpret_cast<__nv_bfloat16*>(out_peers[i]) : nullptr;
  if (out_peers)
    for (int i = 0; i < n_peers; ++i)
      if (!out_peers[i] || (reinterpret_cast<uintptr_t>(out_peers[i]) & 15)) return BYA_ERR_ALIGN;
  static long long* trace = nullptr;   // debug timeline buffer (tools/gpu_fa_trace.py), looked up once
  static bool trace_looked_up = false;
  if (!trace_looked_up) {
    trace_looked_up = true;
    if (const char* e = std::getenv("BYA_FA_TRACE")) trace = reinterpret_cast<long long*>(std::strtoull(e, nullptr, 0));
  }
  a.trace = trace;
  dim3 grid((seq + qt * FA_BM - 1) / (qt * FA_BM), heads, batch);
  kern[qt]<<<grid, threads[qt], fa_smem(qt), reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, a);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

}  // namespace bya

extern "C" int bya_attention_d64(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                                 int batch, int seq, int heads, float scale) {
  return bya::fa_launch<false>(stream, q, k, v, ld, out, ldo, batch, seq, seq, heads, scale);
}

extern "C" int bya_attention_d64_strided(void* stream, const void* q, const void* k, const void* v, int ld, void* out,
                                         int ldo, int batch, int seq, int seq_stride, int heads, float scale) {
  return bya::fa_launch<false>(stream, q, k, v, ld, out, ldo, batch, seq, seq_stride, heads, scale);
}

extern "C" int bya_attention_d64_scatter(void* stream, const void* q, const void* k, const void* v, int ld, void* const* out_peers,
                                         int n_peers, int rows_per_peer, int ldo, int seq, int heads, float scale,
                                         float score_bound_log2) {
  if (score_bound_log2 > 0.f) {
    if (score_bound_log2 > 64.f) return BYA_ERR_SHAPE;
    return bya::fa_launch<true>(stream, q, k, v, ld, nullptr, ldo, 1, seq, seq, heads, 1.0f, out_peers, n_peers, rows_per_peer);
  }
  return bya::fa_launch<false>(stream, q, k, v, ld, nullptr, ldo, 1, seq, seq, heads, scale, out_peers, n_peers, rows_per_peer);
}

extern "C" int bya_attention_d64_bounded(void* stream, const void* q, const void* k, const void* v, int ld, void* out,
                                         int ldo, int batch, int seq, int heads, float score_bound_log2) {
  if (!(score_bound_log2 > 0.f) || score_bound_log2 > 64.f) return BYA_ERR_SHAPE;   // also rejects NaN
  return bya::fa_launch<true>(stream, q, k, v, ld, out, ldo, batch, seq, seq, heads, 1.0f);
}
