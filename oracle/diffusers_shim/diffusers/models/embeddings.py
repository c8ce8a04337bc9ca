"""Timestep / patch / rotary embeddings restated from the published diffusers algorithm."""
import math
import torch
from torch import nn


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale=1):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift
        self.scale = scale

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift, scale=self.scale)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None,
                 sample_proj_bias=True):
        super().__init__()
        assert act_fn == "silu" and post_act_fn is None and cond_proj_dim is None
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, sample_proj_bias)
        self.cond_proj = None
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim, sample_proj_bias)

    def forward(self, sample, condition=None):
        if condition is not None and self.cond_proj is not None:
            sample = sample + self.cond_proj(condition)
        return self.linear_2(self.act(self.linear_1(sample)))


class CogVideoXPatchEmbed(nn.Module):
    def __init__(self, patch_size=2, patch_size_t=None, in_channels=16, embed_dim=1920, text_embed_dim=4096, bias=True,
                 sample_width=90, sample_height=60, sample_frames=49, temporal_compression_ratio=4,
                 max_text_seq_length=226, spatial_interpolation_scale=1.875, temporal_interpolation_scale=1.0,
                 use_positional_embeddings=True, use_learned_positional_embeddings=True):
        super().__init__()
        assert patch_size_t is None
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.use_positional_embeddings = use_positional_embeddings
        self.use_learned_positional_embeddings = use_learned_positional_embeddings
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.text_proj = nn.Linear(text_embed_dim, embed_dim)
        if use_positional_embeddings or use_learned_positional_embeddings:
            raise NotImplementedError("shim: only the RoPE configuration (5B lineage, SURVEY.md §7) is restated")

    def forward(self, text_embeds, image_embeds):
        text_embeds = self.text_proj(text_embeds)
        b, f, c, h, w = image_embeds.shape
        x = image_embeds.reshape(-1, c, h, w)
        x = self.proj(x)
        x = x.view(b, f, *x.shape[1:])
        x = x.flatten(3).transpose(2, 3)  # [b, f, h*w, D]
        x = x.flatten(1, 2)  # [b, f*h*w, D]
        return torch.cat([text_embeds, x], dim=1).contiguous()


def get_1d_rotary_pos_embed(dim, pos, theta=10000.0, use_real=True):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    freqs = torch.outer(pos, freqs)
    cos = freqs.cos().repeat_interleave(2, dim=1).float()
    sin = freqs.sin().repeat_interleave(2, dim=1).float()
    return cos, sin


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size, theta=10000, use_real=True,
                            grid_type="linspace", max_size=None, device=None):
    assert use_real and grid_type == "linspace"
    start, stop = crops_coords
    gh, gw = grid_size
    grid_h = torch.linspace(start[0], stop[0] * (gh - 1) / gh, gh, dtype=torch.float32)
    grid_w = torch.linspace(start[1], stop[1] * (gw - 1) / gw, gw, dtype=torch.float32)
    grid_t = torch.arange(temporal_size, dtype=torch.float32)
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    t_cos, t_sin = get_1d_rotary_pos_embed(dim_t, grid_t, theta)
    h_cos, h_sin = get_1d_rotary_pos_embed(dim_h, grid_h, theta)
    w_cos, w_sin = get_1d_rotary_pos_embed(dim_w, grid_w, theta)

    def combine(ft, fh, fw):
        ft = ft[:, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1)

    return combine(t_cos, h_cos, w_cos), combine(t_sin, h_sin, w_sin)


def apply_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1):
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)
