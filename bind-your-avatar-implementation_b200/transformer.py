"""Drop-in for the reference's `BindyouravatarTransformer3DModel` (models/transformer.py:265-1093).

Same constructor kwargs, `.config`, parameter names (= checkpoint keys), loaders, and the same
`forward(hidden_states, encoder_hidden_states, timestep, ..., routing_logits_forcing=None)` returning the 5-tuple
`(output, None, None, None, None)` that `pipeline_bindyouravatar.py:910-923` consumes.  The forward runs entirely on
hand-written sm_100a kernels through `engine.StepEngine`; it raises if the model is not on a CUDA device in bf16 or
if libbya.so is missing — there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import glob
import inspect
import json
import os
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from .modules import (AdaLayerNorm, AudioAwareModel, CogVideoXBlock, LocalFacialExtractor, MultiIPRouter, PatchEmbed,
                      PerceiverCrossAttention, TimestepEmbedding)


class _Config(dict):
    """ctor kwargs with attribute access (`model.config.patch_size`), like diffusers' FrozenDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class FusedKernelAttnProcessor:
    """What `attn_processors` reports for every attention site: the computation is fused into the step engine
    (QKV GEMM epilogue + tcgen05 flash attention + out-GEMM epilogue), so there is no per-module Python hook."""

    def __call__(self, *a, **k):
        raise RuntimeError("bya_b200: attention runs inside the fused step engine; call the model's forward")


class BindyouravatarTransformer3DModel(nn.Module):
    _supports_gradient_checkpointing = False
    config_name = "config.json"

    def __init__(
        self,
        num_attention_heads: int = 48,
        attention_head_dim: int = 64,
        in_channels: int = 16,
        out_channels: Optional[int] = 16,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        time_embed_dim: int = 512,
        text_embed_dim: int = 4096,
        num_layers: int = 30,
        dropout: float = 0.0,
        attention_bias: bool = True,
        sample_width: int = 90,
        sample_height: int = 60,
        sample_frames: int = 49,
        patch_size: int = 2,
        temporal_compression_ratio: int = 4,
        max_text_seq_length: int = 226,
        activation_fn: str = "gelu-approximate",
        timestep_activation_fn: str = "silu",
        norm_elementwise_affine: bool = True,
        norm_eps: float = 1e-5,
        spatial_interpolation_scale: float = 1.875,
        temporal_interpolation_scale: float = 1.0,
        use_rotary_positional_embeddings: bool = False,
        use_learned_positional_embeddings: bool = False,
        is_train_face: bool = True,
        is_kps: bool = False,
        cross_attn_interval: int = 1,
        LFE_num_tokens: int = 32,
        LFE_output_dim: int = 768,
        LFE_heads: int = 12,
        local_face_scale: float = 1.0,
        is_train_audio: bool = False,
        audio_attn_interval: int = 1,
        draw_routing_logits: bool = False,
        draw_routing_logits_suffix: str = "default",
        draw_routing_logits_video_save_dir: str = None,
        draw_routing_logits_use_softmax: bool = True,
        debug_routing_logits: bool = False,
        debug_routing_logits_zeros: bool = False,
        debug_routing_logits_ones: bool = False,
        is_teacher_forcing: bool = False,
    ):
        super().__init__()
        frame = inspect.currentframe()
        names = list(inspect.signature(self.__init__).parameters)
        self._config = _Config({k: frame.f_locals[k] for k in names})
        inner_dim = num_attention_heads * attention_head_dim
        if not use_rotary_positional_embeddings and use_learned_positional_embeddings:
            raise ValueError(   # same condition and wording as the reference (transformer.py:370-375)
                "There are no CogVideoX checkpoints available with disable rotary embeddings and learned positional "
                "embeddings. If you're using a custom model and/or believe this should be supported, please open an "
                "issue at https://github.com/huggingface/diffusers/issues.")
        if patch_size != 2:
            raise NotImplementedError("bya_b200: the patchify / unpatchify kernels are built for patch_size == 2 "
                                      "(every CogVideoX / ConsisID / Bind-Your-Avatar checkpoint)")
        if activation_fn != "gelu-approximate" or timestep_activation_fn != "silu" or flip_sin_to_cos is not True or freq_shift != 0:
            raise NotImplementedError("bya_b200: only the reference's activation / timestep-embedding configuration is built")

        self.patch_embed = PatchEmbed(patch_size, in_channels, inner_dim, text_embed_dim, sample_width=sample_width,
                                      sample_height=sample_height, sample_frames=sample_frames,
                                      temporal_compression_ratio=temporal_compression_ratio,
                                      max_text_seq_length=max_text_seq_length,
                                      spatial_interpolation_scale=spatial_interpolation_scale,
                                      temporal_interpolation_scale=temporal_interpolation_scale,
                                      use_positional_embeddings=not use_rotary_positional_embeddings,
                                      use_learned_positional_embeddings=use_learned_positional_embeddings)
        self.time_embedding = TimestepEmbedding(inner_dim, time_embed_dim)
        self.transformer_blocks = nn.ModuleList([
            CogVideoXBlock(inner_dim, num_attention_heads, attention_head_dim, time_embed_dim, norm_elementwise_affine,
                           norm_eps, attention_bias) for _ in range(num_layers)])
        self.norm_final = nn.LayerNorm(inner_dim, norm_eps, norm_elementwise_affine)
        self.norm_out = AdaLayerNorm(time_embed_dim, 2 * inner_dim, norm_elementwise_affine, norm_eps)
        self.proj_out = nn.Linear(inner_dim, patch_size * patch_size * out_channels)
        self.gradient_checkpointing = False

        self.is_train_face = is_train_face
        self.is_kps = is_kps
        if is_train_face:
            self.inner_dim = inner_dim
            self.cross_attn_interval = cross_attn_interval
            self.num_ca = num_layers // cross_attn_interval
            self.LFE_final_output_dim = int(inner_dim / 3 * 2)
            self.local_face_scale = local_face_scale
            self._init_face_inputs()
        self.is_train_audio = is_train_audio
        if is_train_audio:
            self.audio_attn_interval = audio_attn_interval
            self.audio_model = AudioAwareModel(dim=inner_dim, norm_elementwise_affine=norm_elementwise_affine,
                                               norm_eps=norm_eps, num_layers=num_layers // audio_attn_interval)
        self.is_teacher_forcing = is_teacher_forcing
        self._engine_obj = None
        self._engine_sig = None
        self._processors: Dict[str, Any] = {}
        self.cache_prologue = True
        self.bounded_attention = True  # joint self-attention without a running max when the qk-LayerNorm bounds |q.k|
        self.sp_cuda_graph = False   # also capture sequence-parallel steps (NCCL exchanges inside the graph)
        self.fused_router_links = os.environ.get("BYA_ROUTER_FUSED", "1") != "0"  # router chain: GEMM + LayerNorm + GEMM links as one kernel each
        self.packed_rope = os.environ.get("BYA_ROPE_PACKED", "1") != "0"  # QKV epilogue reads the rotary table with every pair stored once
        self.use_cuda_graph = False  # replay the step as one CUDA graph per input geometry (engine.step_graphed)
        self._sp_group = None  # set by bya_b200.sp.enable(): Ulysses sequence parallelism over this process group
        self._cfg = None       # set by bya_b200.sp.enable(cfg_parallel=True): which CFG branch this rank computes

    # ------------------------------------------------------------------ reference surface: config / device / dtype
    @property
    def config(self):
        return self._config

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        kw = {k: v for k, v in dict(config).items() if k in sig}
        kw.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**kw)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def _init_face_inputs(self):
        self.local_facial_extractor = LocalFacialExtractor()
        self.perceiver_cross_attention = nn.ModuleList([
            PerceiverCrossAttention(dim=self.inner_dim, dim_head=128, heads=16, kv_dim=self.LFE_final_output_dim)
            for _ in range(self.num_ca)])
        self.router = MultiIPRouter(num_layers=self.num_ca)

    # ------------------------------------------------------------------ reference surface: module save / load
    def save_audio_modules(self, path: str):
        torch.save(self.audio_model.state_dict(), path)

    def load_audio_modules(self, path: str, strict: bool = True):
        try:
            sd = torch.load(path, map_location=self.device)
            missing, unexpected = self.audio_model.load_state_dict(sd, strict=strict)
            print(f"audio_model Missing keys: {missing}")
            print(f"audio_model Unexpected keys: {unexpected}")
        except Exception as e:  # same leniency as the reference (transformer.py:470-472)
            print(f"Error loading audio modules: {e}")
            print(f"path: {path}")

    def save_face_modules(self, path: str):
        torch.save({"local_facial_extractor": self.local_facial_extractor.state_dict(),
                    "perceiver_cross_attention": [ca.state_dict() for ca in self.perceiver_cross_attention]}, path)

    def load_face_modules(self, path: str, strict: bool = True):
        ck = torch.load(path, map_location=self.device)
        missing, unexpected = self.local_facial_extractor.load_state_dict(ck["local_facial_extractor"], strict=strict)
        print(f"local_facial_extractor Missing keys: {missing}")
        print(f"local_facial_extractor Unexpected keys: {unexpected}")
        for i, (ca, sd) in enumerate(zip(self.perceiver_cross_attention, ck["perceiver_cross_attention"])):
            missing, unexpected = ca.load_state_dict(sd, strict=strict)
            print(f"ca {i} Missing keys: {missing}")
            print(f"ca {i} Unexpected keys: {unexpected}")

    def save_router_modules(self, path: str):
        self.router.save(path)

    def load_router_modules(self, path: str, strict: bool = True):
        self.router.load(path, strict=strict)

    # ------------------------------------------------------------------ reference surface: attention processors
    @property
    def attn_processors(self) -> Dict[str, Any]:
        """One entry per attention site the reference exposes (transformer.py:517-538): 42 joint self-attentions,
        12 router attentions, 42 audio cross-attentions."""
        procs = {}
        for name, mod in self.named_modules():
            if hasattr(mod, "get_processor") and hasattr(mod, "to_q"):
                procs[f"{name}.processor"] = self._processors.get(f"{name}.processor") or FusedKernelAttnProcessor()
        return procs

    def set_attn_processor(self, processor):
        count = len(self.attn_processors)
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(
                f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                f" number of attention layers: {count}. Please make sure to pass {count} processor classes.")
        for k in self.attn_processors:
            p = processor.pop(k) if isinstance(processor, dict) else processor
            if not isinstance(p, FusedKernelAttnProcessor):
                raise RuntimeError("bya_b200: attention is fused into the sm_100a step engine and cannot be replaced by a "
                                   "Python attention processor (the reference's infer.py never installs one)")
            self._processors[k] = p

    def fuse_qkv_projections(self):
        """No-op: the engine always runs Q, K and V as one GEMM (packed at first forward)."""
        self.original_attn_processors = self.attn_processors

    def unfuse_qkv_projections(self):
        self.original_attn_processors = None

    # ------------------------------------------------------------------ engine
    def _signature(self):
        ver = 0
        n = 0
        for p in self.parameters():
            ver += p._version
            n += 1
        p0 = next(self.parameters())
        return (ver, n, p0.device, p0.dtype, p0.data_ptr())

    def engine(self):
        """The packed step engine; rebuilt when any parameter changed in place (e.g. after `pipe.fuse_lora`)."""
        from .engine import StepEngine

        sig = self._signature()
        if self._engine_obj is None or sig != self._engine_sig:
            self._engine_obj = StepEngine(self)
            if self._sp_group is not None:
                self._engine_obj.enable_sequence_parallel(self._sp_group)
            self._engine_sig = sig
        return self._engine_obj

    def invalidate(self):
        self._engine_obj = None

    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        encoder_hidden_states: torch.Tensor,
        timestep: Union[int, float, torch.LongTensor],
        timestep_cond: Optional[torch.Tensor] = None,
        image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        id_cond: Optional[torch.Tensor] = None,
        id_vit_hidden: Optional[torch.Tensor] = None,
        index_mask: Optional[torch.Tensor] = None,
        return_dict: bool = True,
        audio_embeds: Optional[torch.Tensor] = None,
        af_matrix: Optional[torch.Tensor] = None,
        denoise_step: Optional[int] = None,
        index_mask_drop_prob: Optional[float] = 0.0,
        routing_logits_zeros_flag: Optional[bool] = False,
        routing_logits_forcing: Optional[torch.Tensor] = None,
        per_frame_forcing: bool = False,
        taps: Optional[dict] = None,
    ):
        if self.is_train_face:
            assert id_cond is not None and id_vit_hidden is not None
        if index_mask is not None or self.training and self.is_teacher_forcing:
            raise NotImplementedError("bya_b200 is the inference hot path; teacher forcing / router losses "
                                      "(transformer.py:741-774, :963-1021) are training-only and out of scope")
        if timestep_cond is not None:
            raise NotImplementedError("bya_b200: timestep_cond is not used by the reference pipeline")
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep] * hidden_states.shape[0], dtype=torch.int64)
        cfg = self._cfg
        if cfg is not None:   # batch-parallel classifier-free guidance: this rank computes one of the two branches
            from .sp import cfg_gather, cfg_slice

            if hidden_states.shape[0] != 2:
                raise RuntimeError("bya_b200: cfg_parallel expects the two CFG branches as a batch of 2")
            br = cfg["branch"]
            hidden_states, encoder_hidden_states, timestep = (cfg_slice(t, br) for t in (hidden_states, encoder_hidden_states, timestep))
            id_cond, id_vit_hidden = cfg_slice(id_cond, br), cfg_slice(id_vit_hidden, br)
            audio_embeds, af_matrix = cfg_slice(audio_embeds, br), cfg_slice(af_matrix, br)
        eng = self.engine()
        if denoise_step == 0:
            eng.new_generation()  # recompute the timestep-invariant prologue (eager cache and every captured graph)
        # (sequence parallel stays eager: capturing the NCCL all-to-alls hung on the 2-GPU box in round 1)
        if self.use_cuda_graph and taps is None and (self._sp_group is None or self.sp_cuda_graph):
            out = eng.step_graphed(hidden_states, encoder_hidden_states, timestep, image_rotary_emb, id_cond, id_vit_hidden,
                                   audio_embeds if self.is_train_audio else None, af_matrix,
                                   routing_logits_forcing=routing_logits_forcing, per_frame_forcing=per_frame_forcing,
                                   cache_prologue=self.cache_prologue)
            if cfg is not None:
                out = cfg_gather(out, cfg)
            return (out, None, None, None, None)
        out = eng.step(hidden_states, encoder_hidden_states, timestep, image_rotary_emb, id_cond, id_vit_hidden,
                       audio_embeds if self.is_train_audio else None, af_matrix,
                       routing_logits_forcing=routing_logits_forcing, per_frame_forcing=per_frame_forcing,
                       cache_prologue=self.cache_prologue, taps=taps)
        if cfg is not None:
            out = cfg_gather(out, cfg)
        return (out, None, None, None, None)

    # ------------------------------------------------------------------ reference surface: checkpoint loading
    @classmethod
    def from_pretrained_cus(cls, pretrained_model_path, subfolder=None, config_path=None, transformer_additional_kwargs={}):
        """models/transformer.py:1024-1093: config.json + (sharded) safetensors / .bin, size-mismatched keys skipped,
        extra input channels of patch_embed.proj zero-filled."""
        if subfolder:
            config_path = config_path or pretrained_model_path
            config_file = os.path.join(config_path, subfolder, "config.json")
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        else:
            config_file = os.path.join(config_path or pretrained_model_path, "config.json")
        print(f"Loading 3D transformer's pretrained weights from {pretrained_model_path} ...")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"Configuration file '{config_file}' does not exist")
        with open(config_file, "r") as f:
            config = json.load(f)
        model = cls.from_config(config, **transformer_additional_kwargs)
        model_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.bin")
        st_file = model_file.replace(".bin", ".safetensors")
        if os.path.exists(model_file):
            state_dict = torch.load(model_file, map_location="cpu")
        else:
            from safetensors.torch import load_file

            files = [st_file] if os.path.exists(st_file) else glob.glob(os.path.join(pretrained_model_path, "*.safetensors"))
            state_dict = {}
            for fpath in files:
                state_dict.update(load_file(fpath))
        own = model.state_dict()
        key = "patch_embed.proj.weight"
        if key in state_dict and own[key].size() != state_dict[key].size():
            new = torch.zeros_like(own[key])
            c = min(new.shape[1], state_dict[key].shape[1])
            new[:, :c] = state_dict[key][:, :c]
            state_dict[key] = new
        kept = {}
        for k, v in state_dict.items():
            if k in own and own[k].size() == v.size():
                kept[k] = v
            else:
                print(k, "Size don't match, skip")
        missing, unexpected = model.load_state_dict(kept, strict=False)
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(unexpected)};")
        print(missing)
        return model
