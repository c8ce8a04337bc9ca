"""TEST INFRASTRUCTURE ONLY — CPU restatement of the step either side of the hot path.

Two pieces:
  * `DPMSchedulerOracle`: `CogVideoXDPMScheduler` as the reference uses it (`infer.py:202,289`;
    `models/pipeline_bindyouravatar.py:868-870` set_timesteps, `:898` scale_model_input, `:934-944` step).
    The class lives in a third-party dependency that is NOT in /root/reference and NOT in this image:
    diffusers==0.34.0.dev0 (`requirements.txt:23`), `schedulers/scheduling_dpm_cogvideox.py`.  This file restates
    its published algorithm (DPM-Solver++ SDE multistep, Lu et al. 2022, with CogVideoX's SNR shift and zero-terminal
    -SNR rescale) in plain torch, double-precision table.  **parity unpinned** for the table / coefficients: there is
    no diffusers here to run and the reference holds no golden vectors for it; what pins it are the solver's
    analytical identities (tests/test_denoise_glue_cpu.py: mean / variance preservation, zero terminal SNR, trailing
    timesteps) — a mistake in the restatement that still satisfied those would go unnoticed.
  * `denoise_loop_oracle`: lines :893-945 of the pipeline (CFG batch build, channel concat, guidance, scheduler step,
    cast) verbatim in torch, taking the transformer as a callable and the randn draws as a callable, so the CUDA loop
    can be compared bit for bit on the same draws.

dtype notes that the bit-exact comparison relies on (torch semantics, not diffusers'): a 0-dim tensor coefficient times
a bf16 tensor yields bf16; times an fp32 tensor yields fp32; see `DPMSchedulerOracle.step` for the two places where
torch's CUDA kernels (what the reference runs) and its CPU kernels round differently.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch


def _rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    s = alphas_cumprod.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = (s - sT) * (s0 / (s0 - sT))
    return s**2


class DPMSchedulerOracle:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.0120, beta_schedule="scaled_linear",
                 set_alpha_to_one=True, rescale_betas_zero_snr=True, snr_shift_scale=1.0, prediction_type="v_prediction",
                 timestep_spacing="trailing", steps_offset=0):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float64) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            ac = _rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else ac[0]
        self.num_train_timesteps, self.prediction_type = num_train_timesteps, prediction_type
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int):
        n, T = num_inference_steps, self.num_train_timesteps
        self.num_inference_steps = n
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, timestep: int, timestep_back: Optional[int]):
        """(alpha_t, mult list, mult_noise, prev_timestep) as 0-dim tensors, `get_variables` / `get_mult` of the class."""
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        a_b = self.alphas_cumprod[timestep_back] if timestep_back is not None else None
        lamb = ((a_t / (1 - a_t)) ** 0.5).log()
        lamb_next = ((a_p / (1 - a_p)) ** 0.5).log()
        h = lamb_next - lamb
        mult = [((1 - a_p) / (1 - a_t)) ** 0.5 * (-h).exp(), (-2 * h).expm1() * a_p**0.5]
        if a_b is not None:
            lamb_prev = ((a_b / (1 - a_b)) ** 0.5).log()
            r = (lamb - lamb_prev) / h
            mult += [1 + 1 / (2 * r), 1 / (2 * r)]
        mult_noise = (1 - a_p) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        return a_t, mult, mult_noise, prev_timestep

    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample, randn: Callable,
             device_semantics: bool = True):
        """`randn(shape, dtype)` stands for diffusers' `randn_tensor(sample.shape, generator=..., dtype=sample.dtype)`.

        device_semantics=False evaluates the published expressions as written; that is bit-exact to the reference only
        when the tensors live on a CUDA device, where the reference runs.  device_semantics=True spells out what the
        CUDA kernels of torch compute for `0-dim coefficient (op) tensor`, so the CPU gives the same bits:
          * coefficient * bf16 tensor: fp32(coefficient) * fp32(element), rounded to bf16 (the CPU kernel instead
            rounds a left-hand scalar to bf16 first — measured on this image, tests/test_gpu_denoise_glue.py);
          * tensor / coefficient: tensor * fp32(1 / coefficient), the reciprocal taken in the coefficient's own precision
            (float64 table) before rounding — ATen's cpu-scalar case of the CUDA true-division kernel; measured:
            1 / fp32(coefficient) is one ulp off at t = 199 of the 5-step schedule and every element then differs."""
        timestep = int(timestep)
        timestep_back = None if timestep_back is None else int(timestep_back)
        a_t, mult, mult_noise, prev_timestep = self.coefficients(timestep, timestep_back)
        b_t = 1 - a_t

        def times(coef, t):
            if not device_semantics:
                return coef * t
            if t.dtype in (torch.bfloat16, torch.float16):
                return (coef.to(torch.float32) * t.float()).to(t.dtype)
            return coef.to(t.dtype) * t

        if self.prediction_type == "epsilon":
            num = sample - times(b_t**0.5, model_output)
            if device_semantics:
                pred = num * (1.0 / a_t**0.5).to(num.dtype)
            else:
                pred = num / a_t**0.5
        elif self.prediction_type == "sample":
            pred = model_output
        elif self.prediction_type == "v_prediction":
            pred = times(a_t**0.5, sample) - times(b_t**0.5, model_output)
        else:
            raise ValueError(self.prediction_type)
        noise = randn(sample.shape, sample.dtype)
        prev_sample = times(mult[0], sample) - times(mult[1], pred) + times(mult_noise, noise)
        if old_pred_original_sample is None or prev_timestep < 0:
            return prev_sample, pred
        denoised_d = times(mult[2], pred) - times(mult[3], old_pred_original_sample)
        noise = randn(sample.shape, sample.dtype)
        return times(mult[0], sample) - times(mult[1], denoised_d) + times(mult_noise, noise), pred


def dynamic_guidance(guidance_scale: float, num_inference_steps: int, t: int) -> float:
    """pipeline_bindyouravatar.py:925-928"""
    return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)


def denoise_loop_oracle(transformer: Callable, scheduler: DPMSchedulerOracle, latents: torch.Tensor,
                        image_latents: torch.Tensor, image_bg_latents: Optional[torch.Tensor], num_inference_steps: int,
                        guidance_scale: float, randn: Callable, do_cfg: bool = True, use_dynamic_cfg: bool = False,
                        zero2cond_cfg_flag: bool = False, out_dtype=torch.bfloat16, trace: Optional[List] = None):
    """pipeline_bindyouravatar.py:893-945.  `transformer(latent_model_input, timestep[B] int64, i)` returns the noise
    prediction [B, F, 16, H, W]; every other argument of the real call is step-invariant and bound by the caller."""
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    old_pred = None
    for i, t in enumerate(timesteps):
        x = torch.cat([latents] * 2) if do_cfg else latents
        x = scheduler.scale_model_input(x, t)
        if do_cfg:
            img = torch.cat([image_latents] * 2) if not zero2cond_cfg_flag else \
                torch.cat([torch.zeros_like(image_latents), image_latents], dim=0)
        else:
            img = image_latents
        if image_bg_latents is not None:
            bg = torch.cat([image_bg_latents, image_bg_latents]) if do_cfg else image_bg_latents
            img = torch.cat([img, bg], dim=2)
        x = torch.cat([x, img], dim=2)
        noise_pred = transformer(x, t.expand(x.shape[0]), i).float()
        g = guidance_scale
        if use_dynamic_cfg:
            g = dynamic_guidance(guidance_scale, num_inference_steps, int(t))
        if do_cfg:
            u, c = noise_pred.chunk(2)
            noise_pred = u + g * (c - u)
        latents, old_pred = scheduler.step(noise_pred, old_pred, t, timesteps[i - 1] if i > 0 else None, latents, randn)
        latents = latents.to(out_dtype)
        if trace is not None:
            trace.append((latents.clone(), old_pred.clone()))
    return latents
