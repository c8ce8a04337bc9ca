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
enum { GEMM_EPI_STORE = 0, GEMM_EPI_RESIDUAL = 1, GEMM_EPI_QKV = 2 };
enum { GEMM_ACT_NONE = 0, GEMM_ACT_GELU_TANH = 1, GEMM_ACT_GELU_ERF = 2 };

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
  /* GEMM_EPI_QKV: columns [0, qk_cols) are q|k heads of 64: + bias, LayerNorm(64, ln_eps, affine), RoPE on rows
   *   >= split_row; columns >= qk_cols (v) only get the bias (diffusers CogVideoXAttnProcessor2_0 as used at
   *   transformer.py:241-245). */
  int qk_cols;
  float ln_eps;
  const float* rope_cos;  /* [M - split_row, 64] fp32 */
  const float* rope_sin;
  const bya_bf16 *nq_w, *nq_b, *nk_w, *nk_b; /* [64] each */
} ByaGemmArgs;

int bya_gemm_bf16(void* stream, const void* A, int lda, const void* W, int ldw, const ByaGemmArgs* args);

/* ---------------------------------------------------------------- multi-head attention, head_dim 64, no mask
 * out[b*seq+n, h*64..] = softmax(Q_h K_h^T * scale) V_h ; q/k/v/out are row-major [batch*seq, ld] views whose head h
 * lives at columns [h*64, h*64+64) (i.e. column slices of the fused projection output).
 * Replaces F.scaled_dot_product_attention inside diffusers CogVideoXAttnProcessor2_0 (transformer.py:241-245) and
 * inside the router's spatial attention (router.py:474-476). */
int bya_attention_d64(void* stream, const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                      int batch, int seq, int heads, float scale);

#ifdef __cplusplus
}
#endif
#endif /* BYA_H_ */
