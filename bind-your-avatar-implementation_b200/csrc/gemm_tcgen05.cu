// C[M,N] = epilogue(A[M,K] · W[N,K]^T)   bf16 in, fp32 accumulate in TMEM, bf16 out.   sm_100a only.
//
// Two builds of one kernel template:
//   CTAS = 2 (default for the step's large GEMMs): a CTA PAIR (cluster of 2 = the two SMs of a TPC) owns a 256 x 256
//            output tile and runs tcgen05.mma.cta_group::2 with M = 256: each CTA stages its own 128 rows of A and only
//            HALF of the B tile (128 of the 256 weight rows) per k-block — 32 KB per stage instead of 48 KB, so 6
//            stages fit, the B operand crosses L2 -> SM once per pair, and shared-memory read traffic per MMA halves.
//            Only the even CTA issues MMAs; its barriers collect the peer's TMA bytes (.cta_group::2 loads) and the
//            peer's epilogue arrivals (remote mbarrier.arrive); tcgen05.commit multicasts to both CTAs.
//   CTAS = 1: one CTA per 128 x BN tile (small M, BN < 256, or BYA_GEMM_CTAS=1).
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled tiles, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       (tcgen05.ld -> registers -> fused epilogue -> global), double-buffered against the
//                              next tile's main loop through two TMEM accumulator stages.  Two warps per TMEM lane
//                              quarter (each takes half of the tile's columns); the tile's bias and gate vectors are
//                              staged in shared memory BEFORE the accumulator is awaited and the residual is
//                              prefetched one chunk ahead: the K=512 router GEMMs were bound by exactly these
//                              global-load latencies (ncu: epilogue warps in long_sb on the bias loads)
// Replaces, on the hot path, every nn.Linear of the reference's denoising step:
// models/transformer.py:200-221 (attn1.to_q/k/v/to_out, ff), models/router.py:226-228,301-302,430-466,
// models/audio_model.py:179-185.  The fused epilogues replace the elementwise ops the reference runs after them:
// bias, GELU, qk-LayerNorm + RoPE (diffusers CogVideoXAttnProcessor2_0), gate * x + residual (transformer.py:247-260),
// scale * x + residual (transformer.py:832, :936).
#include "common.cuh"
#include "gemm.h"

#include <cstdlib>

namespace bya {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;

template <int BN, int CTAS>
struct GemmSmem {
  static constexpr int kStages = CTAS == 2 ? 6 : 4;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = BN / CTAS;            // weight rows staged by ONE CTA per k-block
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = kStages * kStageBytes;          // [2 stages][bias | gate_a | gate_b][BN] fp32
  static constexpr int kEpiBytes = 2 * 3 * BN * 4;
  static constexpr int kStgOffset = kEpiOffset + ((kEpiBytes + 1023) / 1024) * 1024;   // [8 warps][32 rows x 64 B], 64B-swizzled
  static constexpr int kStgBytes = 8 * 2048;
  static constexpr int kBarOffset = kStgOffset + kStgBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + slack for 1024 B alignment
};

// debug timeline (BYA_GEMM_TRACE=<device pointer>): clock64 stamps of CTA 0, event e of its i-th tile (i < 16)
__device__ long long* g_gemm_trace = nullptr;
#define GEMM_TRACE(e, i)                                                                  \
  do {                                                                                    \
    if (g_gemm_trace && blockIdx.x == 0 && (i) < 16 && lane == 0) g_gemm_trace[(i) * 8 + (e)] = clock64(); \
  } while (0)

// store maps of the sequence-parallel PUSH exchange: one 2-D [M, col_block] map per destination rank (peer memory)
struct alignas(64) PeerMaps {
  CUtensorMap m[BYA_MAX_PEERS];
};

template <int BN, int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const GemmArgs p, const __grid_constant__ PeerMaps pmaps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  using L = GemmSmem<BN, CTAS>;
  constexpr int kStages = L::kStages;
  constexpr int TM = BM * CTAS;                         // output rows per tile (per CTA pair when CTAS = 2)
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int unit = CTAS == 2 ? int(blockIdx.x >> 1) : int(blockIdx.x);   // scheduler unit: CTA or CTA pair
  const int num_units = int(gridDim.x) / CTAS;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + TM - 1) / TM;
  const int num_n = p.N / BN;
  const int splits = p.split_k > 1 ? p.split_k : 1;   // GEMM_EPI_SPLITK_F32: k-ranges run as independent tiles
  const int num_tiles = num_m * num_n * splits;
  const int num_kb = p.K / BK;
  constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;   // power of 2

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8 * CTAS);   // one arrive per epilogue warp; the leader's copy also collects the peer's
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) tmem_alloc_pair<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();   // the peer's barriers exist before anything is signalled across
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (m_blk, n_blk): groups of `group_m` row-blocks sweep all column-blocks, so that concurrently running
  // CTAs share a small set of A and W tiles in L2.
  auto k_range = [&](int t, int& kb0, int& kb1) {
    const int ks = t % splits;
    kb0 = int((long long)ks * num_kb / splits);
    kb1 = int((long long)(ks + 1) * num_kb / splits);
  };
  auto tile_coord = [&](int t, int& mb, int& nb) {
    t /= splits;
    const int gm = p.group_m;
    const int per_group = gm * num_n;
    const int g = t / per_group;
    const int first_m = g * gm;
    const int rows = min(gm, num_m - first_m);
    const int r = t - g * per_group;
    mb = first_m + r % rows;
    nb = r / rows;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int ti = 0;
      for (int t = unit; t < num_tiles; t += num_units, ++ti) {
        int mb, nb;
        tile_coord(t, mb, nb);
        const int arow = mb * TM + int(rank) * BM;             // this CTA's 128 rows of A
        const int brow = nb * BN + int(rank) * L::kBRows;      // this CTA's share of the weight rows
        int kb0, kb1;
        k_range(t, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kb == kb0) GEMM_TRACE(0, ti);
          if (kb == kb1 - 1) GEMM_TRACE(1, ti);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if constexpr (CTAS == 2) {
            // the LEADER's barrier collects the bytes of both CTAs' loads of this stage
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes * 2);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (p.a_kblock) {
              const int k0 = kb * BK;
              tma_load_3d_pair(sa, &tmap_a, bar, k0 % p.a_kblock, arow, k0 / p.a_kblock, kEvictNormal);
            } else {
              tma_load_2d_pair(sa, &tmap_a, bar, kb * BK, arow, kEvictNormal);
            }
            tma_load_2d_pair(sb, &tmap_b, bar, kb * BK, brow, kEvictLast);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            if (p.a_kblock) {
              const int k0 = kb * BK;
              tma_load_3d(sa, &tmap_a, &full_bar[stage], k0 % p.a_kblock, arow, k0 / p.a_kblock, kEvictNormal);
            } else {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, arow, kEvictNormal);
            }
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, brow, kEvictLast);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (the leader CTA of a pair)
    constexpr uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int ti = 0;
    for (int t = unit; t < num_tiles; t += num_units, ++ti) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      GEMM_TRACE(2, ti);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * BN;
      int kb0, kb1;
      k_range(t, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (kb == kb0) GEMM_TRACE(3, ti);
        if (kb == kb1 - 1) GEMM_TRACE(4, ti);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
          const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            if constexpr (CTAS == 2) umma_ss_pair(tmem_acc, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, acc);
            else umma_ss(tmem_acc, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, acc);
          }
          if constexpr (CTAS == 2) {   // both CTAs' producers / epilogues are released by the same completion
            umma_commit_pair(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit_pair(&tfull_bar[as]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;    // which half of the tile's columns
    const int et = threadIdx.x - 128;    // 0..255
    float* epi = reinterpret_cast<float*>(smem + L::kEpiOffset);
    int as = 0;
    uint32_t aphase = 0;
    int ti = 0;
    const uint32_t tempty_leader = CTAS == 2 ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
    for (int t = unit; t < num_tiles; t += num_units, ++ti) {
      int mb, nb;
      tile_coord(t, mb, nb);
      const int col0 = nb * BN;
      const int trow0 = mb * TM + int(rank) * BM;   // first output row of this CTA's half of the tile
      if (warp == 4) GEMM_TRACE(5, ti);
      // stage this tile's column vectors while the main loop is still running
      float* sbias = epi + as * 3 * BN;
      float* sga = sbias + BN;
      float* sgb = sga + BN;
      for (int i = et; i < BN; i += kEpiThreads) {
        sbias[i] = p.bias ? __bfloat162float(p.bias[col0 + i]) : 0.f;
        if (p.mode == GEMM_EPI_RESIDUAL) {
          sga[i] = p.gate_a ? p.gate_a[col0 + i] : 1.f;
          sgb[i] = p.gate_b ? p.gate_b[col0 + i] : 1.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      mbar_wait(&tfull_bar[as], aphase);
      if (warp == 4) GEMM_TRACE(6, ti);
      tc_fence_after();
      const int row = trow0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + as * BN + (uint32_t(q * 32) << 16);
      // Output: 32 rows x 32 columns per warp and chunk -> shared memory (64B-swizzled, conflict-free 16 B stores)
      // -> ONE TMA store.  A thread owns an accumulator ROW, so direct global stores put 32 different rows (32
      // half-written sectors) into every store instruction: the clock64 timeline showed 10.8k cycles of epilogue
      // per 128x256 tile, 2.6x the K=512 main loop.  TMA clips rows >= M; with col_block (sequence-parallel send
      // buffer) the map is 3-D [dest][row][col].
      uint8_t* stg = smem + L::kStgOffset + (warp - 4) * 2048;
      const uint32_t stg_row = smem_u32(stg) + lane * 64;
      const int row0 = trow0 + q * 32;
      auto store32 = [&](const float* v, int col) {
        if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has finished reading the buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(stg_row + (((j ^ (lane >> 1)) & 3) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]),
                       pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                       pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.col_block) {
            if (p.peer_out[0]) tma_store_2d(&pmaps.m[col / p.col_block], stg, col % p.col_block, row0);   // NVLink push
            else tma_store_3d(&tmap_c, stg, col % p.col_block, row0, col / p.col_block);
          } else {
            tma_store_2d(&tmap_c, stg, col, row0);
          }
          tma_store_commit();
        }
      };

      // The accumulator stage goes back to the MMA warp as soon as this warp's LAST chunk of it sits in registers —
      // before that chunk's math and stores.  One arrive per warp, without a memory fence: only tensor memory changes
      // hands (tcgen05.fence), the output stores are tracked by their bulk groups.
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) mbar_arrive_remote_nofence(tempty_leader + uint32_t(as) * 8u);
          else mbar_arrive(&tempty_bar[as]);
        }
      };
      if (p.mode == GEMM_EPI_SPLITK_F32) {
        // partial product of one k-range -> its own fp32 slice [ks][M][ldc] of the workspace (plain stores: the
        // reduction over the slices, bias and activation happen in bya_splitk_finalize in a FIXED order, so the
        // result does not depend on which CTA finished first)
        float* orow = reinterpret_cast<float*>(p.out) + (size_t(t % splits) * p.M + size_t(row_ok ? row : 0)) * p.ldc + col0;
        constexpr int CH = (BN >= 64) ? 32 : BN / 2;
        constexpr int NCH = (BN / 2) / CH;
#pragma unroll 1
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = half * (BN / 2) + ci * CH;
          uint32_t r[CH];
          if (CH == 32) tmem_ld_x32(taddr + c, r);
          else tmem_ld_x16(taddr + c, r);
          tmem_ld_wait();
          if (ci == NCH - 1) release_acc();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < CH; i += 4)
              *reinterpret_cast<float4*>(orow + c + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                     __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          }
        }
      } else if (p.mode == GEMM_EPI_QKV) {
        // one 64-wide head per iteration: + bias, LayerNorm(64) on q/k heads, RoPE on video rows
        const bool is_video = row >= p.split_row;
        const float* cs = p.rope_cos + size_t(max(row - p.split_row, 0) + p.rope_row0) * 64;
        const float* sn = p.rope_sin + size_t(max(row - p.split_row, 0) + p.rope_row0) * 64;
        const int blk = p.qkv_block ? p.qkv_block : p.N;
        constexpr int HC = (BN / 2 >= 64) ? BN / 2 : 64;   // columns per warp (whole heads)
        // Rotary values.  ncu (source page): with the full tables the rotation's first multiplies carried 23 % of all
        // stall samples of the kernel — eight exposed L2 round trips per head, because the 32 scattered 16-byte loads per
        // head and thread are issued in batches between the arithmetic (next to no L1 beside 220 KB of shared memory).
        // With the packed table (each pair once) the 16 values of HALF a head are fetched before the accumulator is even
        // loaded, and the other half while the first half is stored: their latency hides under the LayerNorm math.
        // (Keeping the first half across the heads of a tile and fetching the second before the LayerNorm spilled: slower.)
        // (use_pk is uniform over the CTA — store32 below synchronises the warp —; do_rot is per row)
        const bool use_pk = p.rope_cs != nullptr && *p.rope_mismatch == 0;
        const bool do_rot = is_video && row_ok;
        const float4* rpk = reinterpret_cast<const float4*>(
            p.rope_cs + ((use_pk && do_rot) ? size_t(max(row - p.split_row, 0) + p.rope_row0) * 64 : 0));
#pragma unroll 1
        for (int c = half * HC; c < (half + 1) * HC && c < BN; c += 64) {
          float4 rc[4], rs[4];   // cos / sin of the pairs of columns [0, 32) of the head
          if (use_pk && do_rot) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { rc[i] = rpk[i]; rs[i] = rpk[8 + i]; }
          }
          uint32_t r[64];
          tmem_ld_x32(taddr + c, r);
          tmem_ld_x32(taddr + c + 32, r + 32);
          tmem_ld_wait();
          if (!(c + 64 < (half + 1) * HC && c + 64 < BN)) release_acc();   // last head of this warp
          const int col = col0 + c;
          float v[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]) + sbias[c + i];
          const int cin = col % blk;  // position inside the [q|k|v] group
          if (3 * cin < 2 * blk) {
            const bool is_k = 3 * cin >= blk;
            const __nv_bfloat16* gw = is_k ? p.nk_w : p.nq_w;
            const __nv_bfloat16* gb = is_k ? p.nk_b : p.nq_b;
            float mean = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) mean += v[i];
            mean *= (1.f / 64.f);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) { float d = v[i] - mean; var += d * d; }
            const float rstd = rsqrtf(var * (1.f / 64.f) + p.ln_eps);
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              uint32_t ww = *reinterpret_cast<const uint32_t*>(gw + i);
              uint32_t bb = *reinterpret_cast<const uint32_t*>(gb + i);
              v[i] = (v[i] - mean) * rstd * bf16_lo(ww) + bf16_lo(bb);
              v[i + 1] = (v[i + 1] - mean) * rstd * bf16_hi(ww) + bf16_hi(bb);
            }
            const float qs = (!is_k && p.q_premul != 0.f) ? p.q_premul : 1.f;   // softmax scale * log2(e) folded into q
            if (use_pk) {
              auto rot16 = [&](float* x, const float4* cc, const float4* ss) {   // 32 columns = 16 pairs
                if (!do_rot) return;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float cv[4] = {cc[i].x, cc[i].y, cc[i].z, cc[i].w};
                  const float sv[4] = {ss[i].x, ss[i].y, ss[i].z, ss[i].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float x0 = x[8 * i + 2 * j], x1 = x[8 * i + 2 * j + 1];
                    x[8 * i + 2 * j] = x0 * cv[j] - x1 * sv[j];
                    x[8 * i + 2 * j + 1] = x1 * cv[j] + x0 * sv[j];
                  }
                }
              };
              rot16(v, rc, rs);
              if (do_rot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { rc[i] = rpk[4 + i]; rs[i] = rpk[12 + i]; }   // second half: in flight during the store
              }
              if (qs != 1.f) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= qs;
              }
              store32(v, col);
              rot16(v + 32, rc, rs);
              if (qs != 1.f) {
#pragma unroll
                for (int i = 32; i < 64; ++i) v[i] *= qs;
              }
              store32(v + 32, col + 32);
              continue;
            }
            if (is_video && row_ok) {
#pragma unroll
              for (int i = 0; i < 64; i += 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(cs + i);
                const float4 s4 = *reinterpret_cast<const float4*>(sn + i);
                const float x0 = v[i], x1 = v[i + 1], x2 = v[i + 2], x3 = v[i + 3];
                v[i] = x0 * c4.x - x1 * s4.x;
                v[i + 1] = x1 * c4.y + x0 * s4.y;
                v[i + 2] = x2 * c4.z - x3 * s4.z;
                v[i + 3] = x3 * c4.w + x2 * s4.w;
              }
            }
            if (qs != 1.f) {
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] *= qs;
            }
          }
          store32(v, col);
          store32(v + 32, col + 32);
        }
        if (half * HC >= BN) release_acc();   // BN = 64: the second warp of a lane quarter has no head to drain
      } else {
        const float* gate = nullptr;
        float rscale = p.alpha;
        float bscale = 1.f;
        const bool resid_mode = p.mode == GEMM_EPI_RESIDUAL;
        if (resid_mode) {
          gate = (row < p.split_row) ? sga : sgb;
          if (p.row_bias_scale && row_ok) bscale = p.row_bias_scale[row];
        }
        constexpr int CH = (BN >= 64) ? 32 : BN / 2;   // columns per chunk; each warp owns BN/2 columns
        constexpr int NCH = (BN / 2) / CH;
        const int cbeg = half * (BN / 2);
        const __nv_bfloat16* rrow = resid_mode ? p.resid + size_t(row_ok ? row : 0) * p.ldr + col0 + cbeg : nullptr;
        uint4 rnext[CH / 8];
        if (resid_mode) {
#pragma unroll
          for (int i = 0; i < CH / 8; ++i) rnext[i] = reinterpret_cast<const uint4*>(rrow)[i];
        }
#pragma unroll 1
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = cbeg + ci * CH;
          uint32_t r[CH];
          if (CH == 32) tmem_ld_x32(taddr + c, r);
          else tmem_ld_x16(taddr + c, r);
          uint4 rcur[CH / 8];
          if (resid_mode) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) rcur[i] = rnext[i];
            if (ci + 1 < NCH) {   // prefetch the next chunk's residual under this chunk's math
#pragma unroll
              for (int i = 0; i < CH / 8; ++i) rnext[i] = reinterpret_cast<const uint4*>(rrow + (ci + 1) * CH)[i];
            }
          }
          tmem_ld_wait();
          if (ci == NCH - 1) release_acc();
          const int col = col0 + c;
          float v[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]) + sbias[c + i] * bscale;
          if (p.act == GEMM_ACT_GELU_TANH) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = gelu_tanh(v[i]);
          } else if (p.act == GEMM_ACT_GELU_ERF) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
          } else if (p.act == GEMM_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (resid_mode) {
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
              const uint32_t xs[4] = {rcur[i].x, rcur[i].y, rcur[i].z, rcur[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = 8 * i + 2 * j;
                v[e] = bf16_lo(xs[j]) + rscale * gate[c + e] * v[e];
                v[e + 1] = bf16_hi(xs[j]) + rscale * gate[c + e + 1] * v[e + 1];
              }
            }
          }
          store32(v, col);
        }
      }
      if (warp == 4) GEMM_TRACE(7, ti);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();   // every output tile has landed before the CTA retires
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();   // neither CTA's barriers / TMEM go away while the peer still signals them
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BN, int CTAS>
static int launch_gemm(const GemmArgs& a, const void* A, int lda, const void* W, int ldw, cudaStream_t stream) {
  using L = GemmSmem<BN, CTAS>;
  CUtensorMap ta, tb;
  int rc = a.a_kblock ? bya_host::encode_tmap_bf16(&ta, A, a.a_kblock, a.M, uint64_t(lda) * 2, BK, BM, a.K / a.a_kblock,
                                                   uint64_t(a.a_kblock_stride) * 2)
                      : bya_host::encode_tmap_bf16(&ta, A, a.K, a.M, uint64_t(lda) * 2, BK, BM);
  if (rc) return rc;
  rc = bya_host::encode_tmap_bf16(&tb, W, a.K, a.N, uint64_t(ldw) * 2, BK, L::kBRows);
  if (rc) return rc;
  CUtensorMap tc;   // output: 32 x 32 boxes; with col_block a 3-D [dest][row][col] view of the send buffer
  static PeerMaps pm_none = {};
  PeerMaps pm_local;
  const PeerMaps* pm = &pm_none;
  if (a.col_block && a.peer_out[0]) {
    const int nd = a.N / a.col_block;
    for (int d = 0; d < nd; ++d) {
      if (!a.peer_out[d]) return BYA_ERR_SHAPE;
      rc = bya_host::encode_tmap_bf16(&pm_local.m[d], a.peer_out[d], a.col_block, a.M, uint64_t(a.ldc) * 2, 32, 32);
      if (rc) return rc;
    }
    for (int d = nd; d < BYA_MAX_PEERS; ++d) pm_local.m[d] = pm_local.m[0];
    pm = &pm_local;
  }
  if (a.mode == GEMM_EPI_SPLITK_F32 || pm != &pm_none) tc = ta;   // no local store map (fp32 workspace / peer stores)
  else
    rc = a.col_block ? bya_host::encode_tmap_bf16(&tc, a.out, a.col_block, a.M, uint64_t(a.ldc) * 2, 32, 32, a.N / a.col_block,
                                                  uint64_t(a.col_block_stride) * 2)
                     : bya_host::encode_tmap_bf16(&tc, a.out, a.N, a.M, uint64_t(a.ldc) * 2, 32, 32);
  if (rc) return rc;
  auto kern = gemm_bf16_kernel<BN, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess)
      return BYA_ERR_CUDA;
    attr_set = true;
  }
  static bool trace_set = false;
  if (!trace_set) {
    trace_set = true;
    if (const char* e = std::getenv("BYA_GEMM_TRACE")) {
      long long* ptr = reinterpret_cast<long long*>(std::strtoull(e, nullptr, 0));
      cudaMemcpyToSymbol(g_gemm_trace, &ptr, sizeof(ptr));
    }
  }
  const int num_tiles = ((a.M + BM * CTAS - 1) / (BM * CTAS)) * (a.N / BN) * (a.split_k > 1 ? a.split_k : 1);
  if constexpr (CTAS == 1) {
    const int grid = num_tiles < bya_host::num_sms() ? num_tiles : bya_host::num_sms();
    kern<<<grid, kThreads, L::kTotal, stream>>>(ta, tb, tc, a, *pm);
  } else {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // CTA pairs that can be resident at once (a persistent grid must not exceed it): one per TPC
    static int max_pairs = 0;
    if (!max_pairs) {
      cfg.gridDim = dim3(bya_host::num_sms());
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = bya_host::num_sms() / 2;
      max_pairs = n;
    }
    const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
    cfg.gridDim = dim3(2 * pairs);
    if (cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, a, *pm) != cudaSuccess) return BYA_ERR_CUDA;
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

// CTA-pair policy: BYA_GEMM_CTAS=1 / 2 forces a build (A/B timing); default: pairs whenever the shape allows it
static int gemm_ctas_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("BYA_GEMM_CTAS");
    v = e ? std::atoi(e) : 0;
  }
  return v;
}

}  // namespace bya

extern "C" int bya_gemm_bf16(void* stream, const void* A, int lda, const void* W, int ldw, const ByaGemmArgs* args) {
  using namespace bya;
  if (!A || !W || !args || !args->out) return BYA_ERR_SHAPE;
  GemmArgs a = *args;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % BK != 0) return BYA_ERR_SHAPE;
  if (lda % 8 || ldw % 8 || a.ldc % 8 || (a.mode == GEMM_EPI_RESIDUAL && (a.ldr % 8 || !a.resid))) return BYA_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(a.out)) & 15)
    return BYA_ERR_ALIGN;
  if (a.group_m <= 0) a.group_m = 16;
  if (a.mode == GEMM_EPI_QKV) {
    const int blk = a.qkv_block ? a.qkv_block : a.N;
    if (a.N % 64 || blk % 192 || a.N % blk || !a.rope_cos || !a.rope_sin || !a.nq_w || !a.nq_b || !a.nk_w || !a.nk_b)
      return BYA_ERR_SHAPE;
  }
  if (a.col_block && (a.col_block % 64 || a.N % a.col_block || a.col_block_stride % 8 || a.mode == GEMM_EPI_RESIDUAL))
    return BYA_ERR_SHAPE;
  if (a.peer_out[0] && (!a.col_block || a.N / a.col_block > BYA_MAX_PEERS)) return BYA_ERR_SHAPE;
  if (a.mode == GEMM_EPI_SPLITK_F32) {
    if (a.col_block || a.ldc % 4 || (reinterpret_cast<uintptr_t>(a.out) & 15)) return BYA_ERR_SHAPE;
    if (a.split_k < 1) a.split_k = 1;
    if (a.split_k > a.K / BK) a.split_k = a.K / BK;
  } else {
    a.split_k = 0;
  }
  if (a.a_kblock == a.K) a.a_kblock = 0;
  if (a.col_block == a.N && !a.peer_out[0]) a.col_block = 0;   // a single column block is the plain layout
  if (a.a_kblock && (a.a_kblock % BK || a.K % a.a_kblock || a.a_kblock_stride % 8)) return BYA_ERR_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int force = gemm_ctas_override();
  if (a.N % 256 == 0) {
    // measured (gpurun_out/gemm_pair.log): pairs win 9-15 % on the K >= 3072 DiT shapes and lose 12 % on the K = 512
    // router GEMMs, whose 8 k-blocks are over before the deeper pipeline and the cluster launch pay off
    const bool pair = a.mode != GEMM_EPI_SPLITK_F32 && (force ? force == 2 : (a.M > 256 && a.K >= 1024));
    if (pair && !a.col_block && a.mode != GEMM_EPI_QKV && a.N % 192 == 0) {
      // Wave quantisation with few tiles (the per-rank M of 8-way sequence parallelism: 2 222 rows x 3 072 columns = 108
      // pair tiles = 1.46 waves of 74 pairs, i.e. 27 % of the second wave idle): 256 x 192 pair tiles make it 144 tiles
      // = 1.95 waves of 3/4-size tiles.  (256 x 128 pair tiles were measured SLOWER: 0.156 vs 0.134 ms at
      // 2222 x 3072 x 12288 — per-tile efficiency drops faster than the tail shrinks.)
      const long long pairs = bya_host::num_sms() / 2;
      const long long mt = (a.M + 255) / 256;
      const double c256 = double((mt * (a.N / 256) + pairs - 1) / pairs);
      const double c192 = double((mt * (a.N / 192) + pairs - 1) / pairs) * 0.75 / 0.96;
      if (c192 < 0.97 * c256 && std::getenv("BYA_GEMM_NO_PAIR192") == nullptr) return launch_gemm<192, 2>(a, A, lda, W, ldw, s);
    }
    return pair ? launch_gemm<256, 2>(a, A, lda, W, ldw, s) : launch_gemm<256, 1>(a, A, lda, W, ldw, s);
  }
  if (a.N % 128 == 0) return launch_gemm<128, 1>(a, A, lda, W, ldw, s);
  if (a.N % 64 == 0) return launch_gemm<64, 1>(a, A, lda, W, ldw, s);
  return BYA_ERR_SHAPE;
}
