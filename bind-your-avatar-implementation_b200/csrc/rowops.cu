// HBM-bound row kernels of the denoising step: LayerNorm(+adaLN modulation), the adaLN/time-embedding GEMV,
// sinusoidal timestep features, patchify / unpatchify, router output head.  One warp per row, 16-byte vector
// loads, fp32 statistics — sized so every kernel is a single pass over its input.
#include <cstdlib>
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

// ---------------------------------------------------------------------------------------------------------------
// out[r,:] = (LN(x[r,:]) * gamma + beta) * (1 + scale[cls]) + shift[cls]  (+ add[r % add_rows,:])
// Replaces nn.LayerNorm + the modulation arithmetic of diffusers CogVideoXLayerNormZero / AdaLayerNorm
// (models/transformer.py:233, :251, :944-948), router / audio / face LayerNorms (router.py:247-248, :380-393,
// :475-491; audio_model.py:249).
#define BYA_LN_PARAMS                                                                                              \
  const __nv_bfloat16 *__restrict__ x, int ldx, __nv_bfloat16 *__restrict__ out, int ldo, int rows, float eps,     \
      const __nv_bfloat16 *__restrict__ gamma, const __nv_bfloat16 *__restrict__ beta,                             \
      const float *__restrict__ scale_a, const float *__restrict__ shift_a, const float *__restrict__ scale_b,     \
      const float *__restrict__ shift_b, int split_row, const __nv_bfloat16 *__restrict__ add, int add_rows
#define BYA_LN_ARGS x, ldx, out, ldo, rows, eps, gamma, beta, scale_a, shift_a, scale_b, shift_b, split_row, add, add_rows

// FLAGS >= 0 fixes at compile time which optional operands exist (bit 0: gamma AND beta, bit 1: modulation, bit 2:
// additive table) — predicated-off code for absent operands still costs issue slots, and this kernel is half
// issue-bound at D = 3072; FLAGS < 0 tests the pointers at run time (uncommon combinations).
template <int NV, int FLAGS>  // NV = D / 256 : 16-byte vectors per lane
BYA_DEVICE void ln_mod_row(BYA_LN_PARAMS) {
  constexpr int D = NV * 256;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + size_t(warp) * ldx);
  float v[NV * 8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint4 u = src[i * 32 + lane];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + 2 * j] = bf16_lo(w[j]);
      v[i * 8 + 2 * j + 1] = bf16_hi(w[j]);
      sum += v[i * 8 + 2 * j] + v[i * 8 + 2 * j + 1];
    }
  }
  const float mean = warp_sum(sum) * (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) {
    const float d = v[i] - mean;
    var += d * d;
  }
  const float rstd = rsqrtf(warp_sum(var) * (1.f / D) + eps);
  const float* sc = (warp < split_row) ? scale_a : scale_b;
  const float* sh = (warp < split_row) ? shift_a : shift_b;
  const bool has_gamma = FLAGS < 0 ? gamma != nullptr : (FLAGS & 1) != 0;
  const bool has_beta = FLAGS < 0 ? beta != nullptr : (FLAGS & 1) != 0;
  const bool has_mod = FLAGS < 0 ? sc != nullptr : (FLAGS & 2) != 0;
  const bool has_add = FLAGS < 0 ? add != nullptr : (FLAGS & 4) != 0;
  const uint4* addp = has_add ? reinterpret_cast<const uint4*>(add + size_t(warp % add_rows) * D) : nullptr;
  uint4* dst = reinterpret_cast<uint4*>(out + size_t(warp) * ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = (v[i * 8 + j] - mean) * rstd;
    if (has_gamma) {
      const uint4 g = *reinterpret_cast<const uint4*>(gamma + c0);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[2 * j] *= bf16_lo(gw[j]);
        y[2 * j + 1] *= bf16_hi(gw[j]);
      }
    }
    if (has_beta) {
      const uint4 g = *reinterpret_cast<const uint4*>(beta + c0);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[2 * j] += bf16_lo(gw[j]);
        y[2 * j + 1] += bf16_hi(gw[j]);
      }
    }
    if (has_mod) {
      const float4 s0 = *reinterpret_cast<const float4*>(sc + c0), s1 = *reinterpret_cast<const float4*>(sc + c0 + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(sh + c0), h1 = *reinterpret_cast<const float4*>(sh + c0 + 4);
      const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = y[j] * (1.f + ss[j]) + hh[j];
    }
    if (has_add) {
      const uint4 a = addp[i * 32 + lane];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[2 * j] += bf16_lo(aw[j]);
        y[2 * j + 1] += bf16_hi(aw[j]);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    dst[i * 32 + lane] = o;
  }
}

// D = 3072 keeps the row in 96 fp32 registers per lane: 4-warp blocks, three per SM (168 registers, 12 rows in flight per
// SM): 53.9 us at 17 776 rows against 74.1 us as one 8-warp block per SM (175 registers) — tools/gpu_time_ln.py.
template <int NV>
struct LnLaunch {
  static constexpr int kThreads = NV >= 12 ? 128 : 256;
  static constexpr int kMinBlocks = NV >= 12 ? 3 : 2;
};
template <int NV, int FLAGS>
__global__ void __launch_bounds__(LnLaunch<NV>::kThreads, LnLaunch<NV>::kMinBlocks) ln_mod_kernel(BYA_LN_PARAMS) {
  ln_mod_row<NV, FLAGS>(BYA_LN_ARGS);
}

// ---------------------------------------------------------------------------------------------------------------
// y[b, n] = bias[n] + sum_k W[n, k] * act(x[b, k])     (B <= 4, K % 256 == 0), one warp per output n.
// All 2*L+1 adaLN linears of a step share the same input (temb), so they run as ONE call over the row-stacked
// weight; also serves time_embedding.linear_1/2 (transformer.py:686).
template <int B>
__global__ void __launch_bounds__(256) gemv_kernel(const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ bias,
                                                   const float* __restrict__ x, float* __restrict__ y, int N, int K,
                                                   int in_act, int out_act) {
  extern __shared__ float xs[];  // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    float t = x[i];
    if (in_act == 1) t = t / (1.f + __expf(-t));
    xs[i] = t;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + size_t(n) * K);
    float acc[B];
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = 0.f;
    for (int c = lane; c < K / 8; c += 32) {
      const uint4 u = wr[c];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w0 = bf16_lo(w[j]), w1 = bf16_hi(w[j]);
#pragma unroll
        for (int b = 0; b < B; ++b) acc[b] += w0 * xs[b * K + c * 8 + 2 * j] + w1 * xs[b * K + c * 8 + 2 * j + 1];
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float r = warp_sum(acc[b]);
      if (lane == 0) {
        if (bias) r += __bfloat162float(bias[n]);
        if (out_act == 1) r = r / (1.f + __expf(-r));
        y[size_t(b) * N + n] = r;
      }
    }
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t w_i) | sin(t w_i)], w_i = 10000^(-i/half)
__global__ void timestep_features_kernel(const long long* __restrict__ t, float* __restrict__ out, int B, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  const float w = expf(-9.210340371976184f * float(k) / float(half));
  const float a = float(t[b]) * w;
  // kept in fp32 (the reference rounds to the model dtype at transformer.py:685; fp32 is closer to its fp32 truth)
  out[size_t(b) * dim + k] = cosf(a);
  out[size_t(b) * dim + half + k] = sinf(a);
}

// latents [F, C, H, W] (one batch element) -> rows [(f, h/2, w/2)], columns (c, dy, dx): the im2col of
// CogVideoXPatchEmbed's Conv2d(k=2, s=2) (transformer.py:690).  K padded with zeros up to ldo.
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ lat, __nv_bfloat16* __restrict__ out, int F, int C, int H,
                                int W, int ldo) {
  const int gh = H / 2, gw = W / 2;
  const size_t total = size_t(F) * gh * gw * ldo;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int col = int(i % ldo);
    const size_t row = i / ldo;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (col < C * 4) {
      const int c = col >> 2, dy = (col >> 1) & 1, dx = col & 1;
      const int w = int(row % gw), h = int((row / gw) % gh), f = int(row / (size_t(gw) * gh));
      v = lat[((size_t(f) * C + c) * H + 2 * h + dy) * W + 2 * w + dx];
    }
    out[i] = v;
  }
}

// y [F*gh*gw, ldy] (columns (c, dy, dx)) -> out [F, C, 2gh, 2gw]   (transformer.py:955-957)
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ y, int ldy, __nv_bfloat16* __restrict__ out, int F, int C,
                                  int gh, int gw) {
  const int H = 2 * gh, W = 2 * gw;
  const size_t total = size_t(F) * C * H * W;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W), yy = int((i / W) % H), c = int((i / (size_t(W) * H)) % C);
    const int f = int(i / (size_t(W) * H * C));
    const size_t row = (size_t(f) * gh + (yy >> 1)) * gw + (x >> 1);
    out[i] = y[row * ldy + c * 4 + (yy & 1) * 2 + (x & 1)];
  }
}

// r[n, c] = sigmoid(w . x[c*rows + n, :] + b)   — router output head (router.py:408-411), output [rows, C] fp32
__global__ void __launch_bounds__(256) router_head_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                          const __nv_bfloat16* __restrict__ b, float* __restrict__ r, int rows,
                                                          int C, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows * C) return;
  const int c = warp / rows, n = warp % rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + size_t(warp) * D);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  float acc = 0.f;
  for (int i = lane; i < D / 8; i += 32) {
    const uint4 a = xr[i], bb = wr[i];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += bf16_lo(aw[j]) * bf16_lo(bw[j]) + bf16_hi(aw[j]) * bf16_hi(bw[j]);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += __bfloat162float(b[0]);
    // the reference's Linear emits bf16 before the sigmoid, and the sigmoid emits bf16 (bf16 module)
    r[size_t(n) * C + c] = 1.f / (1.f + __expf(-acc));
  }
}

}  // namespace bya

using namespace bya;

extern "C" int bya_layernorm_modulate(void* stream, const void* x, int ldx, void* out, int ldo, int rows, int dim,
                                      float eps, const void* gamma, const void* beta, const float* scale_a,
                                      const float* shift_a, const float* scale_b, const float* shift_b, int split_row,
                                      const void* add, int add_rows) {
  if (!x || !out || rows <= 0) return BYA_ERR_SHAPE;
  if (ldx % 8 || ldo % 8) return BYA_ERR_ALIGN;
  if ((scale_a == nullptr) != (shift_a == nullptr) || (scale_b == nullptr) != (shift_b == nullptr)) return BYA_ERR_SHAPE;
  if (add && add_rows <= 0) return BYA_ERR_SHAPE;
  if (!scale_a && scale_b) split_row = 0;
  if (scale_a && !scale_b) split_row = rows;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int flags = ((gamma && beta) ? 1 : 0) | ((scale_a || scale_b) ? 2 : 0) | (add ? 4 : 0);
  if ((gamma == nullptr) != (beta == nullptr)) flags = -1;
#define BYA_LN_LAUNCH(NV, FLAGS)                                                                                      \
  ln_mod_kernel<NV, FLAGS><<<(rows * 32 + LnLaunch<NV>::kThreads - 1) / LnLaunch<NV>::kThreads, LnLaunch<NV>::kThreads, \
                             0, s>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)out, ldo, rows, eps,               \
                                     (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, scale_a, shift_a,       \
                                     scale_b, shift_b, split_row, (const __nv_bfloat16*)add, add_rows)
#define BYA_LN_CASE(NV)                                                                                               \
  case NV * 256:                                                                                                      \
    switch (flags) {                                                                                                  \
      case 0: BYA_LN_LAUNCH(NV, 0); break;  /* plain */                                                               \
      case 1: BYA_LN_LAUNCH(NV, 1); break;  /* gamma, beta */                                                         \
      case 2: BYA_LN_LAUNCH(NV, 2); break;  /* modulation only */                                                     \
      case 3: BYA_LN_LAUNCH(NV, 3); break;  /* gamma, beta, modulation (adaLN) */                                     \
      case 5: BYA_LN_LAUNCH(NV, 5); break;  /* gamma, beta, additive table (router positions) */                      \
      default: BYA_LN_LAUNCH(NV, -1); break;                                                                          \
    }                                                                                                                 \
    break;
  switch (dim) {
    BYA_LN_CASE(2)
    BYA_LN_CASE(3)
    BYA_LN_CASE(4)
    BYA_LN_CASE(8)
    BYA_LN_CASE(12)
    default:
      return BYA_ERR_SHAPE;
  }
#undef BYA_LN_CASE
#undef BYA_LN_LAUNCH
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_gemv(void* stream, const void* W, const void* bias, const float* x, float* y, int batch, int N, int K,
                        int in_act, int out_act) {
  if (!W || !x || !y || N <= 0 || K <= 0 || K % 256 || batch < 1 || batch > 4) return BYA_ERR_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = min((N + 7) / 8, bya_host::num_sms() * 8);
  const size_t sm = size_t(batch) * K * sizeof(float);
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  const __nv_bfloat16* b = (const __nv_bfloat16*)bias;
  switch (batch) {
    case 1: gemv_kernel<1><<<blocks, 256, sm, s>>>(w, b, x, y, N, K, in_act, out_act); break;
    case 2: gemv_kernel<2><<<blocks, 256, sm, s>>>(w, b, x, y, N, K, in_act, out_act); break;
    case 3: gemv_kernel<3><<<blocks, 256, sm, s>>>(w, b, x, y, N, K, in_act, out_act); break;
    default: gemv_kernel<4><<<blocks, 256, sm, s>>>(w, b, x, y, N, K, in_act, out_act); break;
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

namespace bya {
__global__ void rope_pack_kernel(const float* __restrict__ cos, const float* __restrict__ sin, float* __restrict__ packed,
                                 int* __restrict__ mismatch, int rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, pair)
  if (idx >= (long long)rows * 32) return;
  const long long r = idx >> 5;
  const int i = int(idx & 31);
  const float2 c = *reinterpret_cast<const float2*>(cos + r * 64 + 2 * i);
  const float2 s = *reinterpret_cast<const float2*>(sin + r * 64 + 2 * i);
  packed[r * 64 + i] = c.x;
  packed[r * 64 + 32 + i] = s.x;
  if (__float_as_uint(c.x) != __float_as_uint(c.y) || __float_as_uint(s.x) != __float_as_uint(s.y)) *mismatch = 1;
}
}  // namespace bya

extern "C" int bya_rope_pack(void* stream, const float* cos, const float* sin, float* packed, int* mismatch, int rows) {
  if (!cos || !sin || !packed || !mismatch || rows <= 0) return BYA_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(cos) | reinterpret_cast<uintptr_t>(sin) | reinterpret_cast<uintptr_t>(packed)) & 15) return BYA_ERR_ALIGN;
  const long long n = (long long)rows * 32;
  bya::rope_pack_kernel<<<unsigned((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cos, sin, packed, mismatch, rows);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_timestep_features(void* stream, const int64_t* t, float* out, int batch, int dim) {
  if (!t || !out || batch <= 0 || dim <= 0 || dim % 2) return BYA_ERR_SHAPE;
  const int n = batch * dim / 2;
  timestep_features_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(t), out, batch, dim);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_patchify(void* stream, const void* latents, void* out, int frames, int channels, int height, int width,
                            int ldo) {
  if (!latents || !out || height % 2 || width % 2 || ldo < channels * 4) return BYA_ERR_SHAPE;
  const size_t total = size_t(frames) * (height / 2) * (width / 2) * ldo;
  size_t nb = (total + 255) / 256;
  const size_t cap = size_t(bya_host::num_sms()) * 16;
  const int blocks = int(nb < cap ? nb : cap);
  patchify_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)latents, (__nv_bfloat16*)out, frames, channels, height, width, ldo);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_unpatchify(void* stream, const void* y, int ldy, void* out, int frames, int channels, int grid_h,
                              int grid_w) {
  if (!y || !out || ldy < channels * 4) return BYA_ERR_SHAPE;
  const size_t total = size_t(frames) * channels * grid_h * grid_w * 4;
  size_t nb = (total + 255) / 256;
  const size_t cap = size_t(bya_host::num_sms()) * 16;
  const int blocks = int(nb < cap ? nb : cap);
  unpatchify_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)out, frames, channels, grid_h, grid_w);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_router_head(void* stream, const void* x, const void* w, const void* b, float* r, int rows, int chars,
                               int dim) {
  if (!x || !w || !b || !r || rows <= 0 || chars <= 0 || dim % 8) return BYA_ERR_SHAPE;
  const int warps = rows * chars;
  router_head_kernel<<<(warps + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, r, rows, chars, dim);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
