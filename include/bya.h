/* bya.h — C ABI of libbya.so: the B200 (sm_100a) kernels behind the denoising hot path of Bind-Your-Avatar.
 *
 * The reference has no native layer at all: its hot path is Python calling torch / diffusers.
 * Each entry point below therefore cites the reference *Python* code it replaces (file:line under /root/reference).
 * Conventions ("C-ABI layer"): raw device pointers + sizes + a cudaStream_t passed as void*;
 * the caller owns every buffer; kernels never allocate, never synchronise, never touch another stream.
 * Return value: 0 on success, <0 on error (bad shape / alignment / arch / CUDA launch failure) — the Python host
 * (`bya_b200/ops.py`) turns non-zero into RuntimeError, the way the reference asserts / raises
 * (models/transformer.py:636, :928).  All matrices are row-major bf16 unless stated.
 */
#ifndef BYA_H_
#define BYA_H_
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_bf16.h>
typedef __nv_bfloat16 bya_bf16;
#else
typedef uint16_t bya_bf16;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BYA_ABI_VERSION 1
int bya_abi_version(void);
/* 0 if the current device is sm_100 (B200) and the driver exposes cuTensorMapEncodeTiled, else <0. */
int bya_check_device(void);

/* ---------------------------------------------------------------- GEMM  out = epilogue(A[M,K] · W[N,K]^T)
 * Replaces every nn.Linear on the path: transformer.py:200-221 (attn1 / ff), router.py:226-228, :301-302, :430-466,
 * audio_model.py:179-185, plus the elementwise ops that follow them in the reference (see epilogue modes). */
#define BYA_MAX_PEERS 8
enum { GEMM_EPI_STORE = 0, GEMM_EPI_RESIDUAL = 1, GEMM_EPI_QKV = 2,
       GEMM_EPI_SPLITK_F32 = 3 /* out is an fp32 [split_k][M][ldc] workspace: k-split s STORES its partial product into
                                  slice s; bya_splitk_finalize sums the slices in a fixed order (deterministic) and
                                  applies bias / act */ };
enum { GEMM_ACT_NONE = 0, GEMM_ACT_GELU_TANH = 1, GEMM_ACT_GELU_ERF = 2, GEMM_ACT_RELU = 3 };

typedef struct ByaGemmArgs {
  int M, N, K;            /* K % 64 == 0, N % 64 == 0 */
  int mode;               /* GEMM_EPI_* */
  int act;                /* GEMM_ACT_*  (applied to acc + bias) */
  int group_m;            /* row-blocks per rasterisation group (0 -> 16) */
  const bya_bf16* bias;   /* [N] or NULL */
  bya_bf16* out;          /* [M, ldc] */
  int ldc;
  /* GEMM_EPI_RESIDUAL: out = resid + alpha * gate[row-class] * (acc + bias * row_bias_scale[row])
   *   (transformer.py:247-248, :259-260 gated residuals; :832 face blend residual; :936 audio residual) */
  const bya_bf16* resid;  /* [M, ldr]; may alias out */
  int ldr;
  const float* gate_a;    /* [N] gate for rows <  split_row (text rows), NULL -> 1 */
  const float* gate_b;    /* [N] gate for rows >= split_row (video rows), NULL -> 1 */
  int split_row;
  float alpha;
  const float* row_bias_scale; /* [M] or NULL */
  /* GEMM_EPI_QKV: q and k heads (64 columns each): + bias, LayerNorm(64, ln_eps, affine), RoPE on rows
   *   >= split_row (rope row = row - split_row + rope_row0); v columns only get the bias (diffusers
   *   CogVideoXAttnProcessor2_0 as used at transformer.py:241-245). */
  int qkv_block;          /* width of one [q|k|v] column group (0 -> N); N may hold several groups (one per
                             sequence-parallel destination rank): within a group the first third is q, the second k */
  float ln_eps;
  const float* rope_cos;  /* [>= rope_row0 + M - split_row, 64] fp32 */
  const float* rope_sin;
  int rope_row0;
  const bya_bf16 *nq_w, *nq_b, *nk_w, *nk_b; /* [64] each */
  /* Sequence-parallel plumbing (no reference counterpart).  Output column-block scatter: column c is
   * written at out + (c / col_block) * col_block_stride + row * ldc + c % col_block (col_block % 64 == 0; 0 -> off),
   * i.e. the GEMM writes the all-to-all send buffer [dest][rows][cols_of_dest] directly.  K-blocked A: A is given as
   * K / a_kblock row-major [M, a_kblock] blocks a_kblock_stride elements apart (the all-to-all receive buffer
   * [src][rows][cols_of_src]); a_kblock % 64 == 0; 0 -> off. */
  int col_block;
  long long col_block_stride;
  int a_kblock;
  long long a_kblock_stride;
  /* GEMM_EPI_QKV: the q heads are multiplied by q_premul after LayerNorm + RoPE (0 -> 1).  With
   * q_premul = head_dim^-1/2 * log2(e) the scores Q K^T come out in log2 units, which is what
   * bya_attention_d64_bounded consumes (the softmax scale of diffusers' Attention folded into the projection). */
  float q_premul;
  /* GEMM_EPI_SPLITK_F32 only: the K range is cut into split_k parts that run as independent tiles (skinny,
   * weight-streaming GEMMs of the per-generation prologue — audio_model.py:78-114: M <= 128 rows against a 2.4 GB
   * weight — need more than N / 256 CTAs to pull HBM bandwidth).  0 / 1 -> off. */
  int split_k;
  /* Sequence-parallel PUSH exchange (with col_block): when peer_out[0] != NULL, column block d is TMA-stored to
   * peer_out[d] — the base of a row-major [M, col_block] block with row stride ldc that may live in ANOTHER GPU's memory
   * (NVLink peer mapping) — instead of out + d * col_block_stride.  N / col_block <= BYA_MAX_PEERS. */
  void* peer_out[8];
  /* GEMM_EPI_QKV, optional: the rotary table with every (cos, sin) pair stored ONCE per row — rope_cs [rows, 64] fp32 =
   * [cos[0], cos[2], .., cos[62] | sin[0], sin[2], .., sin[62]] — built by bya_rope_pack, which also writes *rope_mismatch
   * != 0 if some pair of the original tables differs (cos[2i] != cos[2i+1]; diffusers' tables repeat every value twice).
   * When rope_cs != NULL and *rope_mismatch == 0 the epilogue reads 64 values per head and row instead of 128, early
   * enough to hide their latency (same arithmetic, same bits); otherwise it reads rope_cos / rope_sin. */
  const float* rope_cs;
  const int* rope_mismatch;
} ByaGemmArgs;

int bya_gemm_bf16(void* stream, const void* A, int lda, const void* W, int ldw, const ByaGemmArgs* args);

/* ---------------------------------------------------------------- one link of the router's block chain, fused
 *   X    = resid + A1[M,512] · W1[512,512]^T + b1                      -> x_out (bf16; the LayerNorm input)
 *   out2 = act( LayerNorm_512(X; eps, gamma, beta) · W2[N2,512]^T + b2 )
 * Replaces `x = x + attn.to_out(...)` / `x = x + mlp[2](...)` followed by the NEXT sub-block's `normN(x)` and its
 * to_q|to_k|to_v (or mlp[0] + GELU) in SpatialTemporalAttentionBlock.forward (models/router.py:474-491): three launches
 * and two round trips of the [C*Nv, 512] activations become one kernel that keeps its 128 rows of X in tensor memory.
 * The LayerNorm is folded (exact in real arithmetic): the caller passes W2f = bf16(W2 · diag(gamma)),
 * csum[n] = sum_k float(W2f[n,k]) and b2[n] = bias2[n] + sum_k W2[n,k] * beta[k]; the kernel computes mean / rstd of the
 * bf16-rounded X rows and forms rstd * (X · W2f^T - mean * csum) + b2.
 * n_split > 1: the N2 columns are cut into n_split slices handled by different CTAs (the first GEMM is recomputed per
 * slice; for small M); resid must then not alias x_out.  a_kblock / col_block: as in ByaGemmArgs (sequence-parallel
 * receive / send buffers) for A1 and out2.  N2 % 128 == 0, N2 <= 1536. */
typedef struct ByaChainArgs {
  int M, N2;
  int act;                 /* GEMM_ACT_* on out2 */
  int n_split;             /* 0 / 1 -> off; (N2 / 128) % n_split == 0 */
  const bya_bf16* b1;      /* [512] or NULL */
  const bya_bf16* resid;   /* [M, ldr] */
  int ldr;
  bya_bf16* x_out;         /* [M, ldx]; may alias resid when n_split <= 1 */
  int ldx;
  int store_x;             /* 0: X is not written back (the caller only needs out2) */
  float ln_eps;
  const float* csum;       /* [N2] */
  const float* b2;         /* [N2] folded bias */
  bya_bf16* out2;          /* [M, ldc] (or the first column block, see col_block) */
  int ldc;
  int col_block;
  long long col_block_stride;
  int a_kblock;
  long long a_kblock_stride;
} ByaChainArgs;
int bya_gemm_ln_gemm_bf16(void* stream, const void* A1, int lda, const void* W1, int ldw1, const void* W2f, int ldw2,
                          const ByaChainArgs* args);

/* ---------------------------------------------------------------- multi-head attention, head_dim 64, no mask
 * out[b*seq+n, h*64..] = softmax(Q_h K_h^T * scale) V_h ; q/k/v/out are row-major [batch*seq, ld] views whose head h
 * lives at columns [h*64, h*64+64) (i.e. column slices of the fused projection output).
 * Replaces F.scaled_dot_product_attention inside diffusers CogVideoXAttnProcessor2_0 (transformer.py:241-245) and
 * inside the router's spatial attention (router.py:474-476). */
int bya_attention_d64(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                      int batch, int seq, int heads, float scale);
/* Same, with batch element b starting at row b * seq_stride (seq_stride >= seq; the rows in between are ignored).
 * Used by the sequence-parallel router, whose frames are padded to a multiple of the group size. */
int bya_attention_d64_strided(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                              int batch, int seq, int seq_stride, int heads, float scale);
/* Same attention for BOUNDED, pre-scaled scores: out = softmax_2(Q_h K_h^T) V_h with softmax_2(s) = 2^s / sum 2^s,
 * i.e. q already carries scale * log2(e) (ByaGemmArgs.q_premul), and the caller guarantees |q.k| <= score_bound_log2
 * <= 64 for every (query, key) pair — true for the joint self-attention because diffusers applies LayerNorm(64) to
 * every q and k head (CogVideoXAttnProcessor2_0 via transformer.py:241-245): |q| <= 8 max|gamma_q| + |beta_q|, same
 * for k, and RoPE is a rotation.  No running max, no rescaling: about 1.5x the speed of the general kernel.
 * Returns BYA_ERR_SHAPE if the bound is not in (0, 64]. */
int bya_attention_d64_bounded(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                              int batch, int seq, int heads, float score_bound_log2);

/* Sequence-parallel PUSH exchange: the attention for this rank's heads over ALL rows, each output row stored straight
 * into its owner's buffer: row n goes to out_peers[n / rows_per_peer] + (n % rows_per_peer) * ldo + h*64 (the K-blocked A
 * operand of the owner's out-projection; the pointers may be NVLink peer mappings).  batch 1.  score_bound_log2 > 0
 * selects the bounded kernel (see above), otherwise `scale` is used. */
int bya_attention_d64_scatter(void* stream, const void* q, const void* k, const void* v, int ld, void* const* out_peers,
                              int n_peers, int rows_per_peer, int ldo, int seq, int heads, float scale,
                              float score_bound_log2);

/* ---------------------------------------------------------------- routed small-KV cross-attention (32 keys)
 * out[n, h*d..] = sum_c w[n,c] * softmax_k(scale * q[n,h,:].K[g][h][k][:]) @ V[g][h],  g = c*kv_frames + n/(tokens/kv_frames)
 * K  : [chars*kv_frames][heads][32][head_dim],  Vt : [chars*kv_frames][heads][head_dim][32]  (V transposed), bf16.
 * w  : [tokens, chars] fp32 routing / audio weights (NULL -> 1).  head_dim in {64,128}, chars in {1,2,3}.
 * q, out, K, Vt 16-byte aligned, ldq / ldo multiples of 8 (every access is a 16-byte piece; BYA_ERR_ALIGN otherwise).
 * Replaces the attention core + routed blend of PerceiverCrossAttention (router.py:256-273 with transformer.py:821-822)
 * and of the audio cross-attention (audio_model.py:253-256 with transformer.py:925-926). */
/* Sequence-parallel use: q/w/out hold only the `tokens` video tokens [tok_begin, tok_begin+tokens) of `total_tokens`
 * (total_tokens <= 0 -> tokens are the whole clip, tok_begin ignored). */
int bya_xattn_kv32(void* stream, const void* q, int ldq, const void* K, const void* Vt, const float* w, void* out,
                   int ldo, int tokens, int heads, int head_dim, int chars, int kv_frames, float scale,
                   long long tok_begin, long long total_tokens);

/* Router temporal / multi-ID self-attention (router.py:478-488): n_seq sequences of seq_len <= 32 rows of the
 * [rows, 3*heads*64] qkv matrix, rows of one sequence `tok_stride` apart, first row (s/inner)*outer_stride + s%inner. */
int bya_small_attention(void* stream, const void* qkv, int ld, void* out, int ldo, int n_seq, int seq_len, int heads,
                        int inner, long long outer_stride, long long tok_stride, float scale);

/* ---------------------------------------------------------------- row kernels (HBM-bound)
 * out = (LN(x)*gamma+beta) * (1+scale[cls]) + shift[cls] (+ add[row % add_rows]); cls a: rows < split_row, b: others.
 * dim in {512,768,1024,2048,3072}.  gamma/beta/add bf16, scale/shift fp32; any of them may be NULL.
 * Replaces nn.LayerNorm + CogVideoXLayerNormZero / AdaLayerNorm modulation (transformer.py:233,:251,:944-948) and the
 * LayerNorms of router.py:247-248,:380-399,:475-491 and audio_model.py:249. */
int bya_layernorm_modulate(void* stream, const void* x, int ldx, void* out, int ldo, int rows, int dim, float eps,
                           const void* gamma, const void* beta, const float* scale_a, const float* shift_a,
                           const float* scale_b, const float* shift_b, int split_row, const void* add, int add_rows);

/* y[b,n] = out_act(bias[n] + sum_k W[n,k] * in_act(x[b,k]));  act: 0 none, 1 SiLU; batch <= 4; fp32 in/out.
 * One call evaluates every adaLN linear of a step (CogVideoXLayerNormZero.linear x 2L + AdaLayerNorm.linear,
 * transformer.py:198,212,420) on the shared temb; also time_embedding.linear_1/2 (transformer.py:686). */
int bya_gemv(void* stream, const void* W, const void* bias, const float* x, float* y, int batch, int N, int K,
             int in_act, int out_act);

/* packed[r, i] = cos[r, 2i], packed[r, 32 + i] = sin[r, 2i] (i < 32); *mismatch (device int, zeroed by the caller) is set to 1
 * if cos[r, 2i] != cos[r, 2i+1] or sin[r, 2i] != sin[r, 2i+1] anywhere.  See ByaGemmArgs.rope_cs. */
int bya_rope_pack(void* stream, const float* cos, const float* sin, float* packed, int* mismatch, int rows);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) (transformer.py:680): out [batch, dim] fp32 */
int bya_timestep_features(void* stream, const int64_t* t, float* out, int batch, int dim);

/* im2col of CogVideoXPatchEmbed's Conv2d(k=2,s=2) for one batch element: latents [F,C,H,W] -> [F*H/2*W/2, ldo] with
 * columns (c,dy,dx), zero padded to ldo (transformer.py:690); and the inverse scatter of transformer.py:955-957. */
int bya_patchify(void* stream, const void* latents, void* out, int frames, int channels, int height, int width, int ldo);
int bya_unpatchify(void* stream, const void* y, int ldy, void* out, int frames, int channels, int grid_h, int grid_w);

/* r[n,c] = sigmoid(w . x[c*rows+n,:] + b): MultiIPRouter.final_proj + permute (router.py:408-411); r fp32 [rows,chars] */
int bya_router_head(void* stream, const void* x, const void* w, const void* b, float* r, int rows, int chars, int dim);

/* ---------------------------------------------------------------- 3-D masks -> routing (bit-exact integer/0-1 path)
 * masks uint8 [chars,T,H,W] (>0 inside) -> index_mask int64 [frames*grid_h*grid_w] (-1 background, later character
 * wins; may be NULL) and one-hot logits fp32 [tokens, chars].  util/utils.py:871-936 (+ :481-514). */
int bya_masks_to_routing(void* stream, const uint8_t* masks, int chars, int T, int H, int W, int frames, int grid_h,
                         int grid_w, int64_t* index_mask, float* logits);
/* OR (max) over the frame axis broadcast back to all frames: transformer.py:815-818 */
int bya_routing_frame_or(void* stream, const float* logits, float* out, int frames, int tokens_per_frame, int chars);
/* w[n,c] = 1 - max_{c'!=c} (af @ r[n])[c'] (== 1 - swap(af@r) for two characters), wsum[n] = sum_c w[n,c] (may be
 * NULL): transformer.py:860-863,:899-900 */
int bya_audio_weights(void* stream, const float* af, const float* routing, float* w, float* wsum, int tokens, int chars);

/* ---------------------------------------------------------------- per-generation prologue helpers
 * The timestep-invariant sub-graphs (LocalFacialExtractor router.py:157-193, AudioProjModel audio_model.py:78-114, face /
 * router / audio K-V precompute router.py:247-254,:377-383, audio_model.py:241-256) run on bya_gemm_bf16,
 * bya_layernorm_*, bya_attention_d64 plus the data-movement kernels below; nothing of it goes through a library. */

/* out[r, c] = (bf16) src[r, c] for r < rows, c < cols; src is bf16 (src_f32 == 0) or fp32; row strides in elements.
 * torch.cat / repeat / unfold / .to(bf16) of the reference's prologue (router.py:166-176, audio_model.py:188-193). */
int bya_copy2d(void* stream, const void* src, long long lds, int src_f32, void* out, long long ldo, int rows, int cols);
/* Zero `bytes` bytes (a memset node, no kernel): split-K workspaces, padded operand tails. */
int bya_memset_zero(void* stream, void* ptr, long long bytes);
/* out[r, c] = act(sum_s ws[s][row0 + r][c] + bias[c]) as bf16 for r < rows: ws is the [splits][ws_rows][ldw] workspace a
 * GEMM_EPI_SPLITK_F32 call filled; (row0, rows) selects a row range of it.  act: GEMM_ACT_* */
int bya_splitk_finalize(void* stream, const float* ws, int splits, int ws_rows, long long ldw, int row0, const void* bias,
                        int act, void* out, long long ldo, int rows, int cols);
/* LayerNorm(dim 1024, affine) followed by LeakyReLU(0.01): the `Linear -> LayerNorm -> LeakyReLU` stages of
 * LocalFacialExtractor's mapping MLPs (router.py:118-154). */
int bya_layernorm_leakyrelu(void* stream, const void* x, int ldx, void* out, int ldo, int rows, int dim, float eps,
                            const void* gamma, const void* beta, float slope);
/* x [groups*32, ldx]: k of head h at columns k_off + h*head_dim.., v at v_off + h*head_dim..  ->  K [groups][heads][32]
 * [head_dim], Vt [groups][heads][head_dim][32]: the layouts bya_xattn_kv32 consumes (the reference's reshape_tensor /
 * transpose of router.py:250-254 and diffusers' head split of audio_model.py:179-185). */
int bya_kv_pack(void* stream, const void* x, long long ldx, int k_off, int v_off, void* K, void* Vt, int groups, int heads,
                int head_dim);
/* k [chars*32, ldk] (routed keys, natural head-major columns h*head_dim + d) -> block-structured score matrix
 * [chars*32*heads, heads*head_dim]: row (c, tok*heads + h) holds the key of head h in columns h*head_dim.., zeros
 * elsewhere — so that the router's per-head q.k^T (router.py:385-393) is one dense GEMM against it. */
int bya_router_keys_scatter(void* stream, const void* k, long long ldk, void* mat, int chars, int heads, int head_dim);

/* ---------------------------------------------------------------- exchanges over NVLink peer memory
 * No reference counterpart (the reference is single-GPU); oracle = the single-GPU result.
 * bya_peer_barrier: epoch barrier of the sequence-parallel group.  counter: this rank's epoch (device int, zero at start,
 * advanced by the kernel); peer_flags: device array of n_ranks pointers, peer_flags[r] = rank r's flag array [n_ranks]
 * (zero at start; peer-mapped).  Returns once every rank of the group has issued the same barrier. */
int bya_peer_barrier(void* stream, int* counter, int* const* peer_flags, int my_rank, int n_ranks);
/* bya_peer_copy: moves n_segs strided 3-D blocks between a local buffer and peer buffers.  push != 0: for each segment, for
 * o < outer, r < rows: copy row_bytes bytes from local + src_off + o*src_outer_stride + r*src_row_stride to
 * peers[peer] + dst_off + o*dst_outer_stride + r*dst_row_stride (posted NVLink writes); push == 0: the same with the roles
 * swapped (source peers[peer], destination local: remote reads).  All offsets / strides / row_bytes multiples of vec_bytes
 * (16, 8 or 4).  segs and peers are device arrays. */
typedef struct ByaPullSeg {
  long long src_off, dst_off;
  long long src_outer_stride, dst_outer_stride, src_row_stride, dst_row_stride;
  int peer, outer, rows, row_bytes;
} ByaPullSeg;
int bya_peer_copy(void* stream, const ByaPullSeg* segs, int n_segs, void* const* peers, void* local, int push, int vec_bytes,
                  int blocks_per_seg);

/* ---------------------------------------------------------------- the step either side of the path
 * Classifier-free-guidance combine + CogVideoXDPMScheduler.step + the write of x_{t-1} into the next step's model
 * input, as ONE elementwise kernel (models/pipeline_bindyouravatar.py:897-906 concat / scale_model_input, :923-931
 * guidance, :934-945 scheduler step and cast; the scheduler itself is diffusers 0.34.0.dev0
 * schedulers/scheduling_dpm_cogvideox.py, not vendored in the reference).  Every product is rounded separately, in the
 * dtype torch's promotion rules give the reference expression (0-dim coefficient x bf16 tensor -> bf16; x fp32
 * tensor -> fp32), so the result is bit-identical to the reference arithmetic on the same noise draws.
 *
 * Per step the kernel reads one row of `coef` (device memory, so the launch can sit inside a CUDA graph that is
 * replayed for every step): row = coef + BYA_DPM_NCOEF * (step_index ? *step_index : 0). */
enum { BYA_DPM_GUIDANCE = 0,   /* guidance scale (only read when cfg_batch == 2) */
       BYA_DPM_SQRT_ALPHA = 1, /* alphas_cumprod[t] ** 0.5 */
       BYA_DPM_SQRT_BETA = 2,  /* (1 - alphas_cumprod[t]) ** 0.5 */
       BYA_DPM_MULT0 = 3, BYA_DPM_MULT1 = 4, BYA_DPM_MULT2 = 5, BYA_DPM_MULT3 = 6, /* get_mult() */
       BYA_DPM_MULT_NOISE = 7,
       BYA_DPM_SECOND_ORDER = 8, /* != 0: old_pred_original_sample is not None and prev_timestep >= 0 */
       BYA_DPM_INV_SQRT_ALPHA = 9, /* fp32(1 / alphas_cumprod[t] ** 0.5), the reciprocal taken in the table's precision:
                                      what torch multiplies by for `tensor / cpu scalar` (epsilon prediction only) */
       BYA_DPM_NCOEF = 12 };
enum { BYA_PRED_EPSILON = 0, BYA_PRED_SAMPLE = 1, BYA_PRED_V = 2 };

typedef struct ByaDpmStepArgs {
  int frames, channels, hw;      /* latents are [frames, channels, hw] (one sample); channels * hw % 8 == 0 */
  int cfg_batch;                 /* 1: model output is the prediction; 2: [uncond | cond], combined with the guidance */
  int prediction_type;           /* BYA_PRED_* (CogVideoX: v_prediction) */
  const bya_bf16* model_out;     /* bf16 [cfg_batch, frames, channels, hw] (the transformer's output), or NULL */
  const float* model_out_f32;    /* fp32 [frames, channels, hw] when the caller already combined (cfg_batch == 1) */
  const bya_bf16* sample;        /* x_t bf16 */
  bya_bf16* prev_sample;         /* x_{t-1} bf16 (the reference casts to the prompt dtype, :945); may alias sample */
  const float* old_pred;         /* previous step's pred_original_sample fp32; read only on second-order steps */
  float* pred_out;               /* this step's pred_original_sample fp32; may alias old_pred */
  const bya_bf16* noise;         /* bf16 [steps][2][frames*channels*hw]: the step's first / second randn draw */
  bya_bf16* model_input;         /* optional bf16 [in_batch, frames, in_channels, hw]: channels [0, channels) of every
                                    batch entry receive x_{t-1} (:897-906; scale_model_input is the identity) */
  int in_batch, in_channels;
  const float* coef;             /* device fp32 [steps][BYA_DPM_NCOEF] */
  const int* step_index;         /* device, or NULL (row 0, noise pair 0) */
} ByaDpmStepArgs;
int bya_cfg_dpm_step(void* stream, const ByaDpmStepArgs* args);

/* One thread: i = *counter; timestep_out[0..batch) = timesteps[i]; *step_index = i; *counter = i + 1.  Lets a captured
 * graph [select, transformer step, bya_cfg_dpm_step] be replayed for the whole loop (:893-896, :908) with no host
 * work between steps. */
int bya_denoise_select_step(void* stream, const int64_t* timesteps, int n_steps, int64_t* timestep_out, int batch,
                            int* counter, int* step_index);

#ifdef __cplusplus
}
#endif
#endif /* BYA_H_ */
