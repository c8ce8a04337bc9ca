/* mask_oracle.c — plain-C restatement of the reference's 3-D mask -> routing-logit path (TEST INFRASTRUCTURE).
 *
 * Follows /root/reference/util/utils.py:871-936 (process_masks_to_routing_logits) and :481-514 (resize_mask,
 * process_first_frame_only=False): per character, (mask > 0) as float -> F.interpolate(size=(F,gh,gw),
 * mode='trilinear', align_corners=False) -> > 0.5 -> index_mask (-1 background, later character wins) -> one-hot
 * logits; optional frame-OR of models/transformer.py:815-818.
 * The interpolation restates ATen's CPU upsample_trilinear3d (UpSampleKernel.cpp, generic N-d path with
 * HelperInterpLinear; UpSample.h area_pixel_compute_source_index / guard_index_and_lambda): fp32 throughout,
 * src = scale*(dst+0.5)-0.5 clamped at 0, scale = in/out, i0 = floor(src), i1 = i0 + (i0 < in-1), l1 = src - i0,
 * l0 = 1 - l1; value = sum_d w_d * (sum_h w_h * (sum_w w_w * x)).  Rounding, pinned empirically against this image's
 * torch CPU build (0 mismatching floats over random volumes, see tests/test_mask_oracle.py): the source index is
 * fma(scale, dst + 0.5, -0.5) and every 2-term sum is fma(t0, w0, round(t1 * w1)) — how GCC contracts ATen's
 * "output = t0*w0; output += t1*w1".
 * parity unpinned by the reference (no golden vectors there); pinned here against tests/golden/masks_*.pt which the
 * reference's own function produced, and against torch.nn.functional.interpolate on random float volumes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef BYA_FMA
#define BYA_FMA 1
#endif

/* t0*w0 + t1*w1 the way ATen's compiled kernel rounds it */
static inline float lerp2(float t0, float w0, float t1, float w1) {
#if BYA_FMA
  volatile float p = t1 * w1;
  return fmaf(t0, w0, p);
#else
  volatile float p0 = t0 * w0;
  volatile float p1 = t1 * w1;
  return p0 + p1;
#endif
}

static void axis_tables(int in, int out, int* i0, int* i1, float* l0, float* l1) {
  const float scale = (float)in / (float)out;
  for (int d = 0; d < out; ++d) {
    if (in == out) {
      i0[d] = i1[d] = d;
      l0[d] = 1.f;
      l1[d] = 0.f;
      continue;
    }
#if BYA_FMA
    float src = fmaf(scale, (float)d + 0.5f, -0.5f);
#else
    volatile float pr = scale * ((float)d + 0.5f);
    float src = pr - 0.5f;
#endif
    if (src < 0.f) src = 0.f;
    int a = (int)floorf(src);
    if (a > in - 1) a = in - 1;
    float lam = src - (float)a;
    if (lam < 0.f) lam = 0.f;
    if (lam > 1.f) lam = 1.f;
    i0[d] = a;
    i1[d] = a + (a < in - 1 ? 1 : 0);
    l1[d] = lam;
    l0[d] = 1.f - lam;
  }
}

/* x: [T,H,W] fp32 -> y: [F,gh,gw] fp32 */
void bya_oracle_trilinear(const float* x, int T, int H, int W, float* y, int F, int gh, int gw) {
  int *t0 = malloc(sizeof(int) * F), *t1 = malloc(sizeof(int) * F);
  int *h0 = malloc(sizeof(int) * gh), *h1 = malloc(sizeof(int) * gh);
  int *w0 = malloc(sizeof(int) * gw), *w1 = malloc(sizeof(int) * gw);
  float *lt0 = malloc(sizeof(float) * F), *lt1 = malloc(sizeof(float) * F);
  float *lh0 = malloc(sizeof(float) * gh), *lh1 = malloc(sizeof(float) * gh);
  float *lw0 = malloc(sizeof(float) * gw), *lw1 = malloc(sizeof(float) * gw);
  axis_tables(T, F, t0, t1, lt0, lt1);
  axis_tables(H, gh, h0, h1, lh0, lh1);
  axis_tables(W, gw, w0, w1, lw0, lw1);
  for (int f = 0; f < F; ++f)
    for (int h = 0; h < gh; ++h)
      for (int w = 0; w < gw; ++w) {
        float plane[2];
        const int tt[2] = {t0[f], t1[f]};
        for (int a = 0; a < 2; ++a) {
          float row[2];
          const int hh[2] = {h0[h], h1[h]};
          for (int b = 0; b < 2; ++b) {
            const float* p = x + ((size_t)tt[a] * H + hh[b]) * W;
            row[b] = lerp2(p[w0[w]], lw0[w], p[w1[w]], lw1[w]);
          }
          plane[a] = lerp2(row[0], lh0[h], row[1], lh1[h]);
        }
        y[((size_t)f * gh + h) * gw + w] = lerp2(plane[0], lt0[f], plane[1], lt1[f]);
      }
  free(t0); free(t1); free(h0); free(h1); free(w0); free(w1);
  free(lt0); free(lt1); free(lh0); free(lh1); free(lw0); free(lw1);
}

/* masks: uint8 [C,T,H,W] (>0 = inside).  index_mask: int64 [F*gh*gw]; logits: float [F*gh*gw, C].
 * frame_or != 0 additionally applies the OR over frames (transformer.py:815-818) to `logits`. */
void bya_oracle_masks_to_routing(const uint8_t* masks, int C, int T, int H, int W, int F, int gh, int gw,
                                 int64_t* index_mask, float* logits, int frame_or) {
  const size_t nin = (size_t)T * H * W, nout = (size_t)F * gh * gw;
  float* xf = malloc(sizeof(float) * nin);
  float* yf = malloc(sizeof(float) * nout);
  for (size_t i = 0; i < nout; ++i) index_mask[i] = -1;
  for (int c = 0; c < C; ++c) {
    const uint8_t* m = masks + (size_t)c * nin;
    for (size_t i = 0; i < nin; ++i) xf[i] = m[i] > 0 ? 1.f : 0.f;
    bya_oracle_trilinear(xf, T, H, W, yf, F, gh, gw);
    for (size_t i = 0; i < nout; ++i)
      if (yf[i] > 0.5f) index_mask[i] = c;
  }
  memset(logits, 0, sizeof(float) * nout * C);
  for (size_t i = 0; i < nout; ++i)
    if (index_mask[i] >= 0) logits[i * C + index_mask[i]] = 1.f;
  if (frame_or) {
    const size_t hw = (size_t)gh * gw;
    for (size_t s = 0; s < hw; ++s)
      for (int c = 0; c < C; ++c) {
        float mx = 0.f;
        for (int f = 0; f < F; ++f) mx = fmaxf(mx, logits[((size_t)f * hw + s) * C + c]);
        for (int f = 0; f < F; ++f) logits[((size_t)f * hw + s) * C + c] = mx;
      }
  }
  free(xf);
  free(yf);
}
