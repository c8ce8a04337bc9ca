// The step either side of the hot path: classifier-free-guidance combine, CogVideoXDPMScheduler.step
// and the write of x_{t-1} into the next step's model input, as one elementwise kernel.
// Replaces models/pipeline_bindyouravatar.py:897-906 (torch.cat of the latents with the conditioning latents,
// scale_model_input), :921-931 (.float(), guidance), :934-945 (scheduler.step, cast back to the prompt dtype).  The
// scheduler is diffusers 0.34.0.dev0 `scheduling_dpm_cogvideox.py` (DPM-Solver++ SDE, 2nd order multistep):
//   pred = sqrt(a_t) x - sqrt(1 - a_t) v
//   x'   = mult0 x - mult1 pred + mult_noise n0                                     (first step / last step)
//   x'   = mult0 x - mult1 (mult2 pred - mult3 old_pred) + mult_noise n1            (otherwise)
// HBM-bound: per latent element 2 x 2 B model output + 2 B x + 4 B old pred + 2 B noise in, (1 + in_batch) x 2 B x' +
// 4 B pred out = 22 B at CFG batch 2; 1.12 M elements (13 x 16 x 60 x 90) -> 24.7 MB per step.
//
// Rounding follows torch's type promotion on the reference expression exactly: a 0-dim coefficient times a bf16 tensor
// is a bf16 tensor (fp32 product rounded to bf16), times an fp32 tensor an fp32 tensor; sums are fp32; nothing is fused
// into an FMA.  The parity tests evaluate the same expression in torch on the same draws.
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

BYA_DEVICE float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct alignas(16) Bf16x8 {
  __nv_bfloat16 v[8];
};

BYA_DEVICE void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const Bf16x8 t = *reinterpret_cast<const Bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __bfloat162float(t.v[i]);
}
BYA_DEVICE void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

__global__ void __launch_bounds__(256) cfg_dpm_step_kernel(const ByaDpmStepArgs a) {
  const long long n = (long long)a.frames * a.channels * a.hw;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const int step = a.step_index ? *a.step_index : 0;
  const float* c = a.coef + (size_t)step * BYA_DPM_NCOEF;
  const float g = c[BYA_DPM_GUIDANCE], sa = c[BYA_DPM_SQRT_ALPHA], sb = c[BYA_DPM_SQRT_BETA];
  const float m0 = c[BYA_DPM_MULT0], m1 = c[BYA_DPM_MULT1], m2 = c[BYA_DPM_MULT2], m3 = c[BYA_DPM_MULT3];
  const float mn = c[BYA_DPM_MULT_NOISE];
  const bool second = c[BYA_DPM_SECOND_ORDER] != 0.f;

  float v[8], x[8], nz[8], old[8];
  if (a.model_out_f32) {
    load8(a.model_out_f32 + i, v);
  } else {
    load8(a.model_out + i, v);
    if (a.cfg_batch == 2) {   // uncond + g * (cond - uncond), each op rounded to fp32 (:929-931)
      float t[8];
      load8(a.model_out + n + i, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __fadd_rn(v[k], __fmul_rn(g, __fsub_rn(t[k], v[k])));
    }
  }
  load8(a.sample + i, x);
  load8(a.noise + ((size_t)step * 2 + (second ? 1 : 0)) * n + i, nz);
  if (second) load8(a.old_pred + i, old);

  // torch divides a CUDA tensor by a CPU scalar as a multiplication by the reciprocal, taken in the scalar's own
  // precision (float64 table) and then rounded to fp32 (measured: 1.f / fp32(sa) is one ulp off for some timesteps)
  const float inv_sa = c[BYA_DPM_INV_SQRT_ALPHA];
  float pred[8];
  Bf16x8 out;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float p;
    if (a.prediction_type == BYA_PRED_V)
      p = __fsub_rn(bf16_round(__fmul_rn(sa, x[k])), __fmul_rn(sb, v[k]));
    else if (a.prediction_type == BYA_PRED_EPSILON)
      p = __fmul_rn(__fsub_rn(x[k], __fmul_rn(sb, v[k])), inv_sa);
    else
      p = v[k];
    pred[k] = p;
    const float d = second ? __fsub_rn(__fmul_rn(m2, p), __fmul_rn(m3, old[k])) : p;
    const float xs = bf16_round(__fmul_rn(m0, x[k]));
    const float ns = bf16_round(__fmul_rn(mn, nz[k]));
    out.v[k] = __float2bfloat16_rn(__fadd_rn(__fsub_rn(xs, __fmul_rn(m1, d)), ns));
  }
  *reinterpret_cast<float4*>(a.pred_out + i) = make_float4(pred[0], pred[1], pred[2], pred[3]);
  *reinterpret_cast<float4*>(a.pred_out + i + 4) = make_float4(pred[4], pred[5], pred[6], pred[7]);
  *reinterpret_cast<Bf16x8*>(a.prev_sample + i) = out;
  if (a.model_input) {   // latents [f, c, hw] -> model input [b, f, in_channels, hw], channels [0, channels)
    const long long per_frame = (long long)a.channels * a.hw;
    const long long f = i / per_frame, rem = i - f * per_frame;
    const long long in_frame = (long long)a.in_channels * a.hw;
    for (int b = 0; b < a.in_batch; ++b)
      *reinterpret_cast<Bf16x8*>(a.model_input + ((long long)b * a.frames + f) * in_frame + rem) = out;
  }
}

__global__ void select_step_kernel(const long long* __restrict__ timesteps, int n_steps, long long* __restrict__ t_out,
                                   int batch, int* counter, int* step_index) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int i = *counter;
  i = i < 0 ? 0 : (i >= n_steps ? n_steps - 1 : i);
  for (int b = 0; b < batch; ++b) t_out[b] = timesteps[i];
  *step_index = i;
  *counter = i + 1;
}

}  // namespace bya

using namespace bya;

extern "C" int bya_cfg_dpm_step(void* stream, const ByaDpmStepArgs* p) {
  if (!p) return BYA_ERR_SHAPE;
  const ByaDpmStepArgs& a = *p;
  if (a.frames <= 0 || a.channels <= 0 || a.hw <= 0 || (a.cfg_batch != 1 && a.cfg_batch != 2)) return BYA_ERR_SHAPE;
  if (a.prediction_type < BYA_PRED_EPSILON || a.prediction_type > BYA_PRED_V) return BYA_ERR_SHAPE;
  if ((!a.model_out && !a.model_out_f32) || (a.model_out_f32 && a.cfg_batch != 1)) return BYA_ERR_SHAPE;
  if (!a.sample || !a.prev_sample || !a.pred_out || !a.noise || !a.coef || !a.old_pred) return BYA_ERR_SHAPE;
  if (a.model_input && (a.in_batch <= 0 || a.in_channels < a.channels)) return BYA_ERR_SHAPE;
  if (((long long)a.channels * a.hw) % 8 != 0) return BYA_ERR_SHAPE;
  // the model-input write uses 16-byte stores at a frame stride of in_channels * hw elements
  if (a.model_input && ((long long)a.in_channels * a.hw) % 8 != 0) return BYA_ERR_SHAPE;
  const uintptr_t bits = (uintptr_t)a.model_out | (uintptr_t)a.model_out_f32 | (uintptr_t)a.sample |
                         (uintptr_t)a.prev_sample | (uintptr_t)a.old_pred | (uintptr_t)a.pred_out |
                         (uintptr_t)a.noise | (uintptr_t)a.model_input;
  if (bits & 15) return BYA_ERR_ALIGN;
  const long long vecs = (long long)a.frames * a.channels * a.hw / 8;
  cfg_dpm_step_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_denoise_select_step(void* stream, const int64_t* timesteps, int n_steps, int64_t* timestep_out,
                                       int batch, int* counter, int* step_index) {
  if (!timesteps || !timestep_out || !counter || !step_index || n_steps <= 0 || batch <= 0) return BYA_ERR_SHAPE;
  select_step_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(timesteps), n_steps, reinterpret_cast<long long*>(timestep_out), batch, counter,
      step_index);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
