"""Seeded synthetic weights and inputs for the denoising hot path.

There is no network for checkpoints or datasets, so parity tests and `bench.py` use random-init weights of
the reference architecture and synthetic latents/embeddings of the reference shapes.  Every parameter is drawn
from its own generator seeded by (seed, crc32(name)), so the values do not depend on enumeration order, on the
number of layers, or on which object (reference module, oracle, this package) owns the tensor.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import torch


@dataclass
class PathConfig:
    """Shape of one denoising step (reference ctor kwargs `models/transformer.py:322-366` + input geometry)."""

    num_layers: int = 42
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 48
    out_channels: int = 16
    patch_size: int = 2
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    text_len: int = 226
    frames: int = 13  # latent frames F
    grid_h: int = 30  # tokens per column (latent height / patch)
    grid_w: int = 45
    chars: int = 2
    cross_attn_interval: int = 2
    audio_attn_interval: int = 1
    local_face_scale: float = 1.0
    batch: int = 1
    use_rotary_positional_embeddings: bool = True
    use_learned_positional_embeddings: bool = False   # CogVideoX-5B-I2V lineage: checkpoint key patch_embed.pos_embedding

    @property
    def dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def n_video(self) -> int:
        return self.frames * self.grid_h * self.grid_w

    @property
    def n_tokens(self) -> int:
        return self.n_video + self.text_len

    @property
    def audio_frames(self) -> int:
        return 4 * (self.frames - 1) + 5

    def ctor_kwargs(self) -> dict:
        return dict(
            num_attention_heads=self.num_attention_heads, attention_head_dim=self.attention_head_dim,
            in_channels=self.in_channels, out_channels=self.out_channels, time_embed_dim=self.time_embed_dim,
            text_embed_dim=self.text_embed_dim, num_layers=self.num_layers, patch_size=self.patch_size,
            max_text_seq_length=self.text_len, sample_width=self.grid_w * self.patch_size,
            sample_height=self.grid_h * self.patch_size, sample_frames=4 * (self.frames - 1) + 1,
            use_rotary_positional_embeddings=self.use_rotary_positional_embeddings,
            use_learned_positional_embeddings=self.use_learned_positional_embeddings,
            is_train_face=True, cross_attn_interval=self.cross_attn_interval, local_face_scale=self.local_face_scale,
            is_train_audio=True, audio_attn_interval=self.audio_attn_interval,
        )


CONFIGS = {
    # BASELINE.json configs[0]: reference's CPU-runnable case
    "c1": PathConfig(num_layers=1, frames=13, grid_h=8, grid_w=12, cross_attn_interval=1),
    # configs[1]: the headline (49 frames 480x720, 2 characters, one step)
    "c2": PathConfig(),
    "c3": PathConfig(batch=2),
    "c4": PathConfig(chars=3),
    "c5": PathConfig(frames=25),
}


def _gen(seed: int, name: str, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


_TILE = 1 << 22  # elements drawn per tile for very large tensors


@torch.no_grad()
def fill_parameter(name: str, p: torch.Tensor, seed: int = 0) -> None:
    """Deterministic init: matrices ~ N(0, 0.02^2); 1-D `*.weight` of norms ~ 1 + 0.1 N; biases ~ 0.02 N."""
    dev = p.device
    g = _gen(seed, name, dev)
    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("pos_emb"):
        return  # router positional buffer is analytic, not random
    if p.ndim <= 1:
        r = torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32)
        is_norm_gain = leaf == "weight"
        p.copy_((1.0 + 0.1 * r) if is_norm_gain else 0.02 * r)
        return
    n = p.numel()
    if n <= 8 * _TILE:
        r = torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32)
        p.copy_(0.02 * r)
    else:
        # huge tensors (the 1.2 B-param audio Conv1d): draw one tile, repeat it with per-repeat scale
        base = 0.02 * torch.randn(_TILE, generator=g, device=dev, dtype=torch.float32)
        flat = p.view(-1)
        reps = (n + _TILE - 1) // _TILE
        scales = 0.5 + torch.rand(reps, generator=g, device=dev, dtype=torch.float32)
        for i in range(reps):
            lo, hi = i * _TILE, min(n, (i + 1) * _TILE)
            flat[lo:hi].copy_(base[: hi - lo] * scales[i])
    if name in ("local_facial_extractor.latents", "local_facial_extractor.proj_out"):
        p.mul_((1024 ** -0.5) / 0.02)  # reference init scale (router.py:111-116)


@torch.no_grad()
def fill_module(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> None:
    for name, p in module.named_parameters():
        fill_parameter(prefix + name, p.data, seed)
    # the LEARNED positional table is a persistent buffer (a checkpoint key): give it "trained" values, not the sincos init
    persistent = set(module.state_dict().keys())
    for name, b in module.named_buffers():
        if name.endswith("patch_embed.pos_embedding") and name in persistent:
            g = _gen(seed, prefix + name, b.device)
            T = module.patch_embed.max_text_seq_length
            b[:, T:].copy_(0.5 * torch.randn(b[:, T:].shape, generator=g, device=b.device, dtype=torch.float32))


@torch.no_grad()
def make_inputs(cfg: PathConfig, seed: int = 1234, device="cpu", dtype=torch.float32, forced_masks: bool = False):
    """Synthetic step inputs with the kwargs of `pipeline_bindyouravatar.py:910-923`."""
    from .rope import rope_3d_tables

    B, C, F = cfg.batch, cfg.chars, cfg.frames
    g = torch.Generator(device="cpu").manual_seed(seed)

    def rn(*shape, s=1.0):
        return (s * torch.randn(*shape, generator=g)).to(device=device, dtype=dtype)

    H, W = cfg.grid_h * cfg.patch_size, cfg.grid_w * cfg.patch_size
    out = dict(
        hidden_states=rn(B, F, cfg.in_channels, H, W),
        encoder_hidden_states=rn(B, cfg.text_len, cfg.text_embed_dim, s=0.2),
        timestep=torch.full((B,), 500, dtype=torch.int64, device=device),
        image_rotary_emb=tuple(t.to(device) for t in rope_3d_tables(cfg.attention_head_dim, cfg.frames, cfg.grid_h, cfg.grid_w))
        if cfg.use_rotary_positional_embeddings else None,
        id_cond=[rn(B, 1280) for _ in range(C)],
        id_vit_hidden=[[rn(B, 577, 1024) for _ in range(5)] for _ in range(C)],
        audio_embeds=rn(B, C, cfg.audio_frames, 12, 768, s=0.27),
        af_matrix=torch.eye(C).expand(B, C, C).contiguous().to(device=device, dtype=dtype),
    )
    if forced_masks:
        out["routing_logits_forcing"] = box_routing_logits(cfg).to(device=device, dtype=dtype)
    return out


@torch.no_grad()
def box_routing_logits(cfg: PathConfig) -> torch.Tensor:
    """Hard 0/1 routing logits [1, Nv, C]: character c owns a vertical strip of the grid (gaps are background)."""
    r = torch.zeros(1, cfg.frames, cfg.grid_h, cfg.grid_w, cfg.chars)
    strip = cfg.grid_w // cfg.chars
    for c in range(cfg.chars):
        lo = c * strip + (1 if c else 0)
        hi = (c + 1) * strip - 1
        r[:, :, 1:-1, lo:hi, c] = 1.0
    return r.reshape(1, cfg.n_video, cfg.chars)


def tracking_masks(kind: str, T: int = 49, H: int = 480, W: int = 720):
    """Synthetic SAM-2-style tracking masks, uint8 [2,T,H,W] (object id > 0 inside), the input format of
    `util/utils.py:853-868`.  Deterministic; used for the bit-exact mask -> routing-logit parity cases."""
    import numpy as np

    m = np.zeros((2, T, H, W), np.uint8)
    rng = np.random.RandomState(7)
    for t in range(T):
        if kind == "moving":  # 240x200 boxes translating 8 px / frame (exact-0.5 ties occur)
            x0 = 20 + 8 * t
            m[0, t, 100:340, x0:x0 + 200] = 1
            x1 = 700 - 8 * t
            m[1, t, 150:390, max(0, x1 - 200):x1] = 2
        elif kind == "static":
            m[0, t, 60:420, 40:320] = 1
            m[1, t, 60:420, 400:680] = 2
        elif kind == "overlap":  # later character wins on overlap; partial-clip presence
            if t < 30:
                m[0, t, 90:400, 100:450] = 1
            if t > 10:
                m[1, t, 200:470, 300:650] = 2
        elif kind == "speckle":
            m[0, t] = (rng.rand(H, W) > 0.5).astype(np.uint8)
            m[1, t] = (rng.rand(H, W) > 0.7).astype(np.uint8) * 3
        else:
            raise ValueError(kind)
    return m
