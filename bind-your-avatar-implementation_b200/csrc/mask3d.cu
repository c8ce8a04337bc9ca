// 3-D tracking masks -> routing logits (bit-exact), frame-OR, and audio blend weights.
// Replaces util/utils.py:871-936 (process_masks_to_routing_logits: trilinear resize, > 0.5, index_mask with
// "later character wins", one-hot logits), models/transformer.py:815-818 (OR over frames) and :860-863, :899-900
// (audio weights w = 1 - swap(af @ r)).
//
// The resize reproduces ATen's CPU upsample_trilinear3d bit for bit (fp32; source index fma(scale, dst+0.5, -0.5)
// clamped at 0; each 2-tap sum rounded as fma(t0, w0, rn(t1*w1)), depth(height(width)) nesting) — pinned by
// oracle/mask_oracle.c against torch and against the reference-produced goldens.  Only the 8 taps a token needs are
// read (2 of every ~16 rows/columns), so the kernel touches ~13 % of the mask sectors.
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

struct Tap {
  int i0, i1;
  float l0, l1;
};

BYA_DEVICE Tap make_tap(int in, int out, int d) {
  Tap t;
  if (in == out) {
    t.i0 = t.i1 = d;
    t.l0 = 1.f;
    t.l1 = 0.f;
    return t;
  }
  const float scale = __fdiv_rn(float(in), float(out));
  float src = __fmaf_rn(scale, float(d) + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  int a = int(floorf(src));
  a = a > in - 1 ? in - 1 : a;
  float lam = __fsub_rn(src, float(a));
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  t.i0 = a;
  t.i1 = a + (a < in - 1 ? 1 : 0);
  t.l1 = lam;
  t.l0 = __fsub_rn(1.f, lam);
  return t;
}

BYA_DEVICE float lerp2(float t0, float w0, float t1, float w1) { return __fmaf_rn(t0, w0, __fmul_rn(t1, w1)); }

__global__ void __launch_bounds__(256)
masks_to_routing_kernel(const uint8_t* __restrict__ masks, int C, int T, int H, int W, int F, int gh, int gw,
                        long long* __restrict__ index_mask, float* __restrict__ logits) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = F * gh * gw;
  if (n >= total) return;
  const int w = n % gw, h = (n / gw) % gh, f = n / (gw * gh);
  const Tap tt = make_tap(T, F, f), th = make_tap(H, gh, h), tw = make_tap(W, gw, w);
  int label = -1;
  for (int c = 0; c < C; ++c) {
    const uint8_t* m = masks + size_t(c) * T * H * W;
    float plane[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ti = a ? tt.i1 : tt.i0;
      float row[2];
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int hi = b ? th.i1 : th.i0;
        const uint8_t* p = m + (size_t(ti) * H + hi) * W;
        const float x0 = p[tw.i0] > 0 ? 1.f : 0.f;
        const float x1 = p[tw.i1] > 0 ? 1.f : 0.f;
        row[b] = lerp2(x0, tw.l0, x1, tw.l1);
      }
      plane[a] = lerp2(row[0], th.l0, row[1], th.l1);
    }
    const float v = lerp2(plane[0], tt.l0, plane[1], tt.l1);
    if (v > 0.5f) label = c;
  }
  if (index_mask) index_mask[n] = label;
  for (int c = 0; c < C; ++c) logits[size_t(n) * C + c] = (label == c) ? 1.f : 0.f;
}

// logits_out[f, s, c] = max_f' logits_in[f', s, c]      (values are 0/1 for hard masks; works for any floats)
__global__ void frame_or_kernel(const float* __restrict__ in, float* __restrict__ out, int F, int hw, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw * C) return;
  float mx = -INFINITY;
  for (int f = 0; f < F; ++f) mx = fmaxf(mx, in[size_t(f) * hw * C + i]);
  for (int f = 0; f < F; ++f) out[size_t(f) * hw * C + i] = mx;
}

// w[n,c] = 1 - max_{c' != c} av[n,c'],  av[n,:] = af @ r[n,:];  wsum[n] = sum_c w[n,c]
template <int C>
__global__ void audio_weights_kernel(const float* __restrict__ af, const float* __restrict__ r, float* __restrict__ w,
                                     float* __restrict__ wsum, int n_tok) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_tok) return;
  float rv[C], av[C];
#pragma unroll
  for (int c = 0; c < C; ++c) rv[c] = r[size_t(n) * C + c];
#pragma unroll
  for (int a = 0; a < C; ++a) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += af[a * C + c] * rv[c];
    av[a] = s;
  }
  float tot = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float mx = -INFINITY;
#pragma unroll
    for (int o = 0; o < C; ++o)
      if (o != c) mx = fmaxf(mx, av[o]);
    const float v = (C == 1) ? 1.f : 1.f - mx;
    w[size_t(n) * C + c] = v;
    tot += v;
  }
  if (wsum) wsum[n] = tot;
}

}  // namespace bya

using namespace bya;

extern "C" int bya_masks_to_routing(void* stream, const uint8_t* masks, int chars, int T, int H, int W, int frames,
                                    int grid_h, int grid_w, int64_t* index_mask, float* logits) {
  if (!masks || !logits || chars <= 0 || T <= 0 || H <= 0 || W <= 0 || frames <= 0 || grid_h <= 0 || grid_w <= 0)
    return BYA_ERR_SHAPE;
  const int total = frames * grid_h * grid_w;
  masks_to_routing_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      masks, chars, T, H, W, frames, grid_h, grid_w, reinterpret_cast<long long*>(index_mask), logits);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_routing_frame_or(void* stream, const float* logits, float* out, int frames, int tokens_per_frame,
                                    int chars) {
  if (!logits || !out || frames <= 0 || tokens_per_frame <= 0 || chars <= 0) return BYA_ERR_SHAPE;
  const int n = tokens_per_frame * chars;
  frame_or_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, out, frames,
                                                                                      tokens_per_frame, chars);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_audio_weights(void* stream, const float* af, const float* routing, float* w, float* wsum, int tokens,
                                 int chars) {
  if (!af || !routing || !w || tokens <= 0) return BYA_ERR_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (tokens + 255) / 256;
  switch (chars) {
    case 1: audio_weights_kernel<1><<<blocks, 256, 0, s>>>(af, routing, w, wsum, tokens); break;
    case 2: audio_weights_kernel<2><<<blocks, 256, 0, s>>>(af, routing, w, wsum, tokens); break;
    case 3: audio_weights_kernel<3><<<blocks, 256, 0, s>>>(af, routing, w, wsum, tokens); break;
    case 4: audio_weights_kernel<4><<<blocks, 256, 0, s>>>(af, routing, w, wsum, tokens); break;
    default: return BYA_ERR_SHAPE;
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
