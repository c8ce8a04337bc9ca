// Flash-style multi-head attention forward, head_dim 64, bf16, no mask, non-causal.   sm_100a only.
//   O[b, n, h*64 : h*64+64] = softmax(Q_h K_h^T * scale) V_h
// Replaces F.scaled_dot_product_attention as called by diffusers' CogVideoXAttnProcessor2_0 for the joint
// [text; video] self-attention (models/transformer.py:241-245) and by the router's spatial attention
// (models/router.py:474-476, K10).  Q/K/V are read in place from the (fused) projection output — row-major
// [rows, ld] with head h at columns h*64.. — so no head-major copy is ever made.
//
// One CTA = one (batch, head, 256 query rows).  Warp roles:
//   warp 0       TMA producer: Q (2 x 128 rows) once, then a ring of K and V tiles (128 keys each)
//   warp 1       MMA issuer:   S_t = Q_t K^T (tcgen05.mma SS, fp32 in TMEM), O_t += P_t V (tcgen05.mma TS, P read
//                              from TMEM where the softmax warps wrote it over S_t)
//   warp 2       TMEM allocator
//   warps 4-7    softmax for query tile 0   (one thread per query row; tcgen05.ld S -> exp2 -> tcgen05.st P)
//   warps 8-11   softmax for query tile 1
// The two query tiles ping-pong on the tensor pipe: while one tile's softmax runs, the other's MMAs issue.
// Running max uses lazy rescaling (O in TMEM is only rescaled when the max grows by more than 2^8).
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

constexpr int FA_D = 64;
constexpr int FA_BM = 128;         // query rows per tile (2 tiles per CTA)
constexpr int FA_BN = 128;         // keys per KV tile
constexpr int FA_STAGES = 4;       // K ring depth == V ring depth
constexpr int FA_THREADS = 384;
constexpr int FA_TILE_BYTES = FA_BM * FA_D * 2;  // 16 KB
constexpr int FA_ONES_BYTES = 2048;             // [16 x 64] bf16 ones: B operand of the row-sum MMA
constexpr int FA_SMEM = (2 + 2 * FA_STAGES) * FA_TILE_BYTES + FA_ONES_BYTES + 512 + 1024;

struct FaArgs {
  int seq;        // rows per batch element (queries == keys)
  int heads;
  int batch;
  int ldo;        // row stride of O in elements
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
};

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
BYA_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FA_THREADS, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const FaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][16 KB]
  uint8_t* sK = smem + 2 * FA_TILE_BYTES;               // [STAGES][16 KB]
  uint8_t* sV = sK + FA_STAGES * FA_TILE_BYTES;         // [STAGES][16 KB]
  uint8_t* sOnes = sV + FA_STAGES * FA_TILE_BYTES;      // [2 KB] bf16 1.0 (any swizzle of all-ones is all-ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + FA_ONES_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [STAGES]
  uint64_t* k_empty = k_full + FA_STAGES;  // [STAGES]
  uint64_t* v_full = k_empty + FA_STAGES;
  uint64_t* v_empty = v_full + FA_STAGES;
  uint64_t* s_full = v_empty + FA_STAGES;  // [2]
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* o_full = p_full + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.seq;               // first row of this batch element in the [rows, ld] matrices
  const int q0 = qblk * (2 * FA_BM);            // first query row (within the batch element)
  const int n_kv = (p.seq + FA_BN - 1) / FA_BN;
  const int col = head * FA_D;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  for (int i = threadIdx.x; i < FA_ONES_BYTES / 4; i += FA_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384)  L0 [384,400)  L1 [400,416);
  // P_t overwrites S_t[0,64) as packed bf16; L_t = P_t * ones accumulates the softmax row sums on the tensor pipe
  constexpr uint32_t kColS = 0, kColO = 256, kColL = 384;

  if (warp < 4) {
    setmaxnreg_dec<56>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * FA_TILE_BYTES);
        tma_load_2d(sQ, &tmap_q, q_full, col, row_base + q0, kEvictFirst);
        tma_load_2d(sQ + FA_TILE_BYTES, &tmap_q, q_full, col, row_base + q0 + FA_BM, kEvictFirst);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&k_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&k_full[stage], FA_TILE_BYTES);
          tma_load_2d(sK + stage * FA_TILE_BYTES, &tmap_k, &k_full[stage], col, row_base + j * FA_BN, kEvictLast);
          mbar_wait(&v_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&v_full[stage], FA_TILE_BYTES);
          tma_load_2d(sV + stage * FA_TILE_BYTES, &tmap_v, &v_full[stage], col, row_base + j * FA_BN, kEvictLast);
          if (++stage == FA_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BM, FA_BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BM, FA_D, 0, 1);  // B = V is MN-major (d contiguous)
      constexpr uint32_t idesc_pl = make_idesc_bf16(FA_BM, 16, 0, 0);    // B = ones [16 x K], K-major
      const uint64_t d_ones = make_smem_desc_sw128(smem_u32(sOnes), 16, 1024);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int t, int stage) {
        const uint64_t da = make_smem_desc_sw128(q_addr + t * FA_TILE_BYTES, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sK + stage * FA_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k)
          umma_ss(tmem_base + kColS + t * 128, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_qk, k != 0);
      };
      auto issue_pv = [&](int t, int stage, bool acc) {
        const uint64_t db = make_smem_desc_sw128(smem_u32(sV + stage * FA_TILE_BYTES), 1024, 1024);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) {
          umma_ts(tmem_base + kColO + t * 64, tmem_base + kColS + t * 128 + k * 8, db + uint64_t(128 * k), idesc_pv,
                  acc || k != 0);
          umma_ts(tmem_base + kColL + t * 16, tmem_base + kColS + t * 128 + k * 8, d_ones, idesc_pl, acc || k != 0);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, 0);
        umma_commit(&s_full[0]);
        issue_qk(1, 0);
        umma_commit(&s_full[1]);
        umma_commit(&k_empty[0]);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == FA_STAGES) { nstage = 0; nphase ^= 1; }
        const bool more = (j + 1 < n_kv);
        mbar_wait(&v_full[stage], phase);
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
        if (elect_one()) issue_pv(0, stage, j > 0);
        __syncwarp();
        if (more) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          if (elect_one()) {
            issue_qk(0, nstage);
            umma_commit(&s_full[0]);
          }
          __syncwarp();
        }
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(1, stage, j > 0);
          umma_commit(&v_empty[stage]);
          if (more) {
            issue_qk(1, nstage);
            umma_commit(&s_full[1]);
            umma_commit(&k_empty[nstage]);
          } else {
            umma_commit(o_full);
          }
        }
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = (warp - 4) >> 2;   // query tile 0 / 1
    const int q = warp & 3;          // TMEM lane quarter
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + kColS + t * 128 + lane_off;
    const uint32_t tO = tmem_base + kColO + t * 64 + lane_off;
    const uint32_t tL = tmem_base + kColL + t * 16 + lane_off;
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
          const float mn = fmaxf(m, mt);
          const float alpha = ex2((m - mn) * c);
          m = mn;
          uint32_t o[64];
          tmem_ld_x32(tO, o);
          tmem_ld_x32(tO + 32, o + 32);
          uint32_t lsum;
          tmem_ld_x1(tL, &lsum);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 64; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          lsum = __float_as_uint(__uint_as_float(lsum) * alpha);
          tmem_st_x32(tO, o);
          tmem_st_x32(tO + 32, o + 32);
          tmem_st_x1(tL, lsum);
        }
      }
      const float nmc = -m * c;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        float t0, t1;
        fma2(t0, t1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), c, nmc);
        pk[i / 2] = pack_bf16x2(ex2(t0), ex2(t1));
      }
      tmem_st_x32(tS, pk);
      tmem_st_x32(tS + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o[64];
    tmem_ld_x32(tO, o);
    tmem_ld_x32(tO + 32, o + 32);
    tmem_ld_wait();
    uint32_t lbits;
    tmem_ld_x1(tL, &lbits);
    tmem_ld_wait();
    const int qrow = q0 + t * FA_BM + q * 32 + lane;
    if (qrow < p.seq) {
      const float inv = 1.0f / __uint_as_float(lbits);
      uint4* dst = reinterpret_cast<uint4*>(p.out + size_t(row_base + qrow) * p.ldo + col);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace bya

extern "C" int bya_attention_d64(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                                 int batch, int seq, int heads, float scale) {
  using namespace bya;
  if (!q || !k || !v || !out || batch <= 0 || seq <= 0 || heads <= 0) return BYA_ERR_SHAPE;
  if (ld % 8 || ldo % 8 || ld < heads * FA_D || ldo < heads * FA_D) return BYA_ERR_ALIGN;
  const uint64_t rows = uint64_t(batch) * seq;
  CUtensorMap tq, tk, tv;
  int rc = bya_host::encode_tmap_bf16(&tq, q, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BM);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tk, k, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BN);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tv, v, uint64_t(heads) * FA_D, rows, uint64_t(ld) * 2, FA_D, FA_BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM) != cudaSuccess)
      return BYA_ERR_CUDA;
    attr_set = true;
  }
  FaArgs a;
  a.seq = seq;
  a.heads = heads;
  a.batch = batch;
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  dim3 grid((seq + 2 * FA_BM - 1) / (2 * FA_BM), heads, batch);
  fa_fwd_kernel<<<grid, FA_THREADS, FA_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, a);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
