"""`CogVideoXDPMScheduler` as the reference uses it (`infer.py:202,289`; `models/pipeline_bindyouravatar.py:868-870`,
`:898`, `:934-944`), with `step()` running on the fused CUDA kernel `bya_cfg_dpm_step`.

The original class is diffusers 0.34.0.dev0 `schedulers/scheduling_dpm_cogvideox.py` (not vendored by the reference and
not in this image): DPM-Solver++ SDE multistep on a scaled-linear beta table with CogVideoX's SNR shift and
zero-terminal-SNR rescale.  The table and the per-step coefficients are a few hundred scalars and are computed on the
host in float64 with the same torch expressions the published class uses; everything proportional to the latent size
runs in the kernel.  Same method names, argument order and return values as the class it stands in for; what is not
supported raises instead of silently differing (`eta`, `use_clipped_model_output`, `variance_noise`, non-bf16 samples).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops

_PRED = {"epsilon": ops.PRED_EPSILON, "sample": ops.PRED_SAMPLE, "v_prediction": ops.PRED_V}


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor for a single generator: a CPU generator draws on the CPU and the
    result is moved, a device generator draws on the device."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    draw_on = device
    if generator is not None and generator.device.type != device.type:
        if generator.device.type != "cpu":
            raise ValueError(f"cannot draw {device} noise from a {generator.device} generator")
        draw_on = torch.device("cpu")
    return torch.randn(tuple(shape), generator=generator, device=draw_on, dtype=dtype).to(device)


# What the CogVideoX-5B (and 5B-I2V) `scheduler/scheduler_config.json` supplies — the values the reference pipeline
# actually runs with (`infer.py:202` loads them with `from_pretrained(model_path, subfolder="scheduler")`).  The
# constructor DEFAULTS below are the published class's own (epsilon / leading / no rescale / snr_shift_scale 3.0).
COGVIDEOX_5B_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
    set_alpha_to_one=True, steps_offset=0, prediction_type="v_prediction", clip_sample_range=1.0, sample_max_value=1.0,
    timestep_spacing="trailing", rescale_betas_zero_snr=True, snr_shift_scale=1.0)


class CogVideoXDPMScheduler:
    """NOTE for integrators: `pipeline_bindyouravatar.py:936` picks the (model_output, old_pred_original_sample, timestep,
    timestep_back, sample) call form with `isinstance(self.scheduler, CogVideoXDPMScheduler)` against the class IT
    imported — import this class under that name in the pipeline module, otherwise the DDIM-style
    call form is used and the arguments bind wrongly."""

    order = 1
    init_noise_sigma = 1.0
    config_name = "scheduler_config.json"

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.0120,
                 beta_schedule: str = "scaled_linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, snr_shift_scale: float = 3.0, **unused):
        self.config = _Config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                              beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                              set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                              prediction_type=prediction_type, clip_sample_range=clip_sample_range,
                              sample_max_value=sample_max_value, timestep_spacing=timestep_spacing,
                              rescale_betas_zero_snr=rescale_betas_zero_snr, snr_shift_scale=snr_shift_scale)
        if prediction_type not in _PRED:
            raise ValueError(f"prediction_type {prediction_type!r} must be one of {sorted(_PRED)}")
        if trained_betas is not None:
            betas = torch.as_tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float64) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__.__name__}")
        self.betas = betas
        self.alphas = 1.0 - betas
        ac = torch.cumprod(self.alphas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)      # SNR shift
        if rescale_betas_zero_snr:                                       # zero terminal SNR, on sqrt(alphas_cumprod)
            root = ac.sqrt()
            first, last = root[0].clone(), root[-1].clone()
            root = (root - last) * (first / (first - last))
            ac = root**2
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else ac[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        cfg.pop("variance_type", None)   # infer.py:283-287 forwards it; the DPM class has no use for it
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        """`infer.py:202`: CogVideoXDPMScheduler.from_pretrained(model_path, subfolder="scheduler") — reads
        `<path>/<subfolder>/scheduler_config.json` (local directories only: there is no hub access here)."""
        import json
        import os

        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        f = d if os.path.isfile(d) else os.path.join(d, cls.config_name)
        if not os.path.isfile(f):
            raise OSError(f"bya_b200: no {cls.config_name} under {d}")
        with open(f, "r") as fh:
            return cls.from_config(json.load(fh), **kwargs)

    @classmethod
    def cogvideox_5b(cls, **kwargs):
        """The scheduler of the CogVideoX-5B lineage without a checkpoint directory at hand."""
        return cls(**dict(COGVIDEOX_5B_SCHEDULER_CONFIG, **kwargs))

    def __len__(self):
        return self.config.num_train_timesteps

    # ------------------------------------------------------------------------------------------------ schedule
    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {T}")
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (T // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / num_inference_steps)).astype(np.int64) - 1
        else:
            raise ValueError(f"{spacing} is not supported; choose one of 'leading', 'linspace', 'trailing'")
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    # ------------------------------------------------------------------------------------------------ coefficients
    def step_coefficients(self, timestep: int, timestep_back: Optional[int], have_old_pred: bool,
                          guidance_scale: float = 1.0) -> List[float]:
        """One row of the kernel's coefficient table (include/bya.h BYA_DPM_*), from the float64 table."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        log_snr_t = ((a_t / (1 - a_t)) ** 0.5).log()
        log_snr_prev = ((a_prev / (1 - a_prev)) ** 0.5).log()
        h = log_snr_prev - log_snr_t
        mult0 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        mult1 = (-2 * h).expm1() * a_prev**0.5
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        second = have_old_pred and prev_timestep >= 0
        mult2 = mult3 = torch.tensor(0.0)
        if second:
            if timestep_back is None:
                raise ValueError("a second-order step needs timestep_back")
            a_back = self.alphas_cumprod[timestep_back]
            r = (log_snr_t - ((a_back / (1 - a_back)) ** 0.5).log()) / h
            mult2, mult3 = 1 + 1 / (2 * r), 1 / (2 * r)
        row = [guidance_scale, a_t**0.5, (1 - a_t) ** 0.5, mult0, mult1, mult2, mult3, mult_noise, float(second),
               1 / a_t**0.5]   # (inf at alpha_t = 0: epsilon prediction is undefined there, in the reference too)
        row = [float(torch.as_tensor(v, dtype=torch.float64).to(torch.float32)) for v in row]
        return row + [0.0] * (ops.DPM_NCOEF - len(row))

    # ------------------------------------------------------------------------------------------------ one step
    @torch.no_grad()
    def step(self, model_output: torch.Tensor, old_pred_original_sample: Optional[torch.Tensor], timestep,
             timestep_back, sample: torch.Tensor, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = False):
        """Returns `(prev_sample, pred_original_sample)`; `prev_sample` comes back in the sample's dtype (bf16) — the
        reference casts it there on the next line (`pipeline_bindyouravatar.py:945`), and the rounding happens once
        from the same fp32 value either way."""
        if eta != 0.0 or use_clipped_model_output or variance_noise is not None:
            raise NotImplementedError("bya_b200 CogVideoXDPMScheduler.step: eta / use_clipped_model_output / "
                                      "variance_noise are unused by the published class and not supported")
        if return_dict:
            raise NotImplementedError("bya_b200 CogVideoXDPMScheduler.step: the reference calls with return_dict=False")
        if sample.dtype != torch.bfloat16 or not sample.is_cuda:
            raise NotImplementedError("bya_b200 CogVideoXDPMScheduler.step: bf16 CUDA samples only")
        timestep = int(timestep)
        timestep_back = None if timestep_back is None else int(timestep_back)
        row = self.step_coefficients(timestep, timestep_back, old_pred_original_sample is not None)
        second = row[8] != 0.0
        dev, n = sample.device, sample.numel()
        noise = torch.empty(1, 2, n, dtype=torch.bfloat16, device=dev)
        noise[0, 0] = randn_tensor(sample.shape, generator, dev, sample.dtype).reshape(-1)
        if second:
            noise[0, 1] = randn_tensor(sample.shape, generator, dev, sample.dtype).reshape(-1)
        coef = torch.tensor([row], dtype=torch.float32, device=dev)
        pred = torch.empty(sample.shape, dtype=torch.float32, device=dev)
        prev = torch.empty_like(sample)
        old = old_pred_original_sample if second else pred
        if model_output.dtype not in (torch.float32, torch.bfloat16) or model_output.numel() != n:
            raise ValueError("model_output must be one fp32 / bf16 prediction of the sample's shape")
        ops.cfg_dpm_step(model_output.contiguous(), sample.contiguous(), prev, old.contiguous(), pred, noise, coef,
                         prediction_type=_PRED[self.config.prediction_type])
        return prev, pred
