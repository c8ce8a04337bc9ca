"""The denoising loop around the hot path, resident on the device.

`DenoiseLoop.run` is `models/pipeline_bindyouravatar.py:893-945`: build the CFG batch of the model input (latents ‖
conditioning-image latents ‖ background latents on the channel axis), call the transformer, combine the two
guidance branches, take one `CogVideoXDPMScheduler.step`, cast back to bf16 — for every timestep.  The reference
spends ≈15 torch launches, two `torch.cat` copies of the 48-channel input, an fp32 round trip and (with dynamic CFG) a
`.item()` sync per step on it.  Here:

  * the model input is one static buffer `[B, F, 48, H, W]`; its conditioning channels are written once per
    generation, and `bya_cfg_dpm_step` writes x_{t-1} straight into channels [0, 16) of both CFG entries;
  * timesteps, guidance scales and solver coefficients of all steps sit in device tables, selected by a device-side
    step counter (`bya_denoise_select_step`), so nothing about a step depends on the host;
  * [select step, transformer step, guidance + solver step] is captured once as ONE CUDA graph and replayed
    `num_inference_steps` times back to back — no host work, sync or allocation between steps;
  * the randn draws of all steps are taken up front from the caller's generator, in the order the reference takes
    them (one per step, two on second-order steps), so a seed gives the same latents as the step-by-step loop.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops
from .scheduler import _PRED, CogVideoXDPMScheduler, randn_tensor


class DenoiseLoop:
    def __init__(self, transformer, scheduler: CogVideoXDPMScheduler, guidance_scale: float = 6.0,
                 use_dynamic_cfg: bool = False, do_classifier_free_guidance: bool = True,
                 zero2cond_cfg_flag: bool = False, cuda_graph: bool = True):
        self.transformer, self.scheduler = transformer, scheduler
        self.guidance_scale, self.use_dynamic_cfg = float(guidance_scale), use_dynamic_cfg
        self.do_cfg, self.zero2cond = do_classifier_free_guidance, zero2cond_cfg_flag
        self.cuda_graph = cuda_graph
        self._graph = None   # (signature, graph, state)

    # ------------------------------------------------------------------------------------------------ tables
    def guidance_at(self, t: int, num_inference_steps: int) -> float:
        if not self.use_dynamic_cfg:   # pipeline_bindyouravatar.py:925-928
            return self.guidance_scale
        return 1 + self.guidance_scale * (
            (1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)

    def coefficient_table(self, num_inference_steps: int) -> torch.Tensor:
        ts = [int(t) for t in self.scheduler.timesteps]
        rows = [self.scheduler.step_coefficients(t, ts[i - 1] if i > 0 else None, i > 0,
                                                 self.guidance_at(t, num_inference_steps)) for i, t in enumerate(ts)]
        return torch.tensor(rows, dtype=torch.float32)

    def draw_noise(self, shape, coef: torch.Tensor, generator, device) -> torch.Tensor:
        """[steps, 2, numel] bf16: slot 0 = the step's first draw, slot 1 = the second (second-order steps only)."""
        n = math.prod(shape)
        noise = torch.zeros(coef.shape[0], 2, n, dtype=torch.bfloat16, device=device)
        for i in range(coef.shape[0]):
            noise[i, 0] = randn_tensor(shape, generator, device, torch.bfloat16).reshape(-1)
            if coef[i, 8] != 0:
                noise[i, 1] = randn_tensor(shape, generator, device, torch.bfloat16).reshape(-1)
        return noise

    # ------------------------------------------------------------------------------------------------ the loop
    @torch.no_grad()
    def run(self, latents: torch.Tensor, image_latents: torch.Tensor, image_bg_latents: Optional[torch.Tensor],
            prompt_embeds: torch.Tensor, image_rotary_emb, id_cond, id_vit_hidden, audio_embeds, af_matrix,
            num_inference_steps: int = 50, generator=None, routing_logits_forcing=None, per_frame_forcing: bool = False,
            noise: Optional[torch.Tensor] = None, trace: Optional[list] = None,
            conditions_cfg_batched: bool = True) -> torch.Tensor:
        """latents / image_latents / image_bg_latents: bf16 [1, F, 16, H, W]; the other conditions already carry the
        CFG batch the way the pipeline prepares them (:877-884), or pass conditions_cfg_batched=False to have
        id_cond / id_vit_hidden / audio_embeds / af_matrix batched here (`conditions.prepare_cfg_conditions`;
        prompt_embeds is always [negative | positive] as `encode_prompt` returns it).
        Returns the final latents bf16 [1, F, 16, H, W]."""
        if not conditions_cfg_batched:
            from .conditions import prepare_cfg_conditions

            id_cond, id_vit_hidden, audio_embeds, af_matrix = prepare_cfg_conditions(
                id_cond, id_vit_hidden, audio_embeds, af_matrix, self.do_cfg, self.zero2cond)
        m, sch = self.transformer, self.scheduler
        # Multi-GPU (bya_b200.sp.enable): every rank runs the same loop on the same inputs and the same noise draws (seed the
        # generators identically).  Sequence parallel: the step returns the full prediction on every rank, the solver step
        # is replicated.  Batch-parallel CFG: this rank computes ONE guidance branch of the static model input and the two
        # predictions are exchanged (`sp.cfg_gather`) before the fused guidance + solver kernel.
        cfgp = getattr(m, "_cfg", None)
        multi = getattr(m, "_sp_group", None) is not None or cfgp is not None
        if cfgp is not None and not self.do_cfg:
            raise ValueError("bya_b200.DenoiseLoop: cfg_parallel needs classifier-free guidance (two branches)")
        dev = m.device
        bf = torch.bfloat16
        B = 2 if self.do_cfg else 1
        if latents.shape[0] != 1 or latents.dtype != bf:
            raise ValueError("latents must be bf16 [1, F, C, H, W]")
        _, F, C, H, W = latents.shape
        sch.set_timesteps(num_inference_steps, device="cpu")
        steps = len(sch.timesteps)
        coef_host = self.coefficient_table(num_inference_steps)
        if noise is None:
            noise = self.draw_noise(latents.shape, coef_host, generator, dev)
        # ---- static model input: [latents | image | background] on the channel axis, CFG batch outermost (:897-906)
        cond = [image_latents.to(dev, bf)]
        if self.do_cfg:
            cond = [torch.zeros_like(cond[0]) if self.zero2cond else cond[0], cond[0]]
        cond = torch.cat(cond, 0)
        if image_bg_latents is not None:
            cond = torch.cat([cond, image_bg_latents.to(dev, bf).expand(B, -1, -1, -1, -1)], dim=2)
        x = torch.empty(B, F, C + cond.shape[2], H, W, dtype=bf, device=dev)
        x[:, :, C:] = cond
        x[:, :, :C] = latents.to(dev)
        st = dict(x=x, lat=latents.to(dev, copy=True).contiguous(), pred=torch.zeros(latents.shape, dtype=torch.float32, device=dev),
                  coef=coef_host.to(dev), noise=noise, ts=sch.timesteps.to(dev, torch.int64).contiguous(),
                  t=torch.zeros(B, dtype=torch.int64, device=dev), counter=torch.zeros(1, dtype=torch.int32, device=dev),
                  idx=torch.zeros(1, dtype=torch.int32, device=dev))
        eng = m.engine()
        use_router = routing_logits_forcing is None
        aud = audio_embeds if m.is_train_audio else None
        kw = dict(hidden_states=x, encoder_hidden_states=prompt_embeds.to(dev), timestep=st["t"],
                  image_rotary_emb=tuple(t.to(dev) for t in image_rotary_emb), id_cond=[t.to(dev) for t in id_cond],
                  id_vit_hidden=[[v.to(dev) for v in l] for l in id_vit_hidden],
                  audio_embeds=None if aud is None else aud.to(dev), af_matrix=None if af_matrix is None else af_matrix.to(dev),
                  routing_logits_forcing=None if use_router else routing_logits_forcing.to(dev),
                  per_frame_forcing=per_frame_forcing, cache_prologue=False)
        if cfgp is not None:   # this rank's guidance branch: views into the static buffers (x is rewritten every step)
            from .sp import cfg_gather, cfg_slice

            br = cfgp["branch"]
            for k in ("hidden_states", "encoder_hidden_states", "timestep", "id_cond", "id_vit_hidden", "audio_embeds", "af_matrix"):
                kw[k] = cfg_slice(kw[k], br)
        kw["_pro"] = eng.prologue(kw["id_cond"], kw["id_vit_hidden"], kw["audio_embeds"], F, use_router)
        pt = _PRED[sch.config.prediction_type]

        def one_step():
            ops.denoise_select_step(st["ts"], st["t"], st["counter"], st["idx"])
            out = eng.step(**kw)
            if cfgp is not None:
                out = cfg_gather(out, cfgp)
            ops.cfg_dpm_step(out, st["lat"], st["lat"], st["pred"], st["pred"], st["noise"], st["coef"],
                             prediction_type=pt, step_index=st["idx"], model_input=st["x"])

        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.cuda_graph and trace is None:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):   # warm-up of the transformer step only: workspaces, tensor maps
                eng.step(**kw)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = ops.LAUNCHES
            # thread_local: the NCCL watchdog thread of a multi-GPU run may query events while we capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if multi else "global"):
                one_step()
            per_replay = ops.LAUNCHES - l0
            ops.LAUNCHES = l0
            t0.record()
            for _ in range(steps):
                graph.replay()
                ops.LAUNCHES += per_replay
            t1.record()
            self._graph = graph   # keeps the captured buffers alive until the result has been read
        else:
            t0.record()
            for _ in range(steps):
                one_step()
                if trace is not None:
                    trace.append((st["lat"].clone(), st["pred"].clone()))
            t1.record()
        self.state, self.loop_events = st, (t0, t1)   # elapsed_time(*loop_events): device time of the steps alone
        return st["lat"]
