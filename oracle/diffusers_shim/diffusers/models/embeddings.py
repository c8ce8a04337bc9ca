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
        self.sample_height, self.sample_width, self.sample_frames = sample_height, sample_width, sample_frames
        self.temporal_compression_ratio = temporal_compression_ratio
        self.max_text_seq_length = max_text_seq_length
        self.spatial_interpolation_scale = spatial_interpolation_scale
        self.temporal_interpolation_scale = temporal_interpolation_scale
        if use_positional_embeddings or use_learned_positional_embeddings:
            # sincos table; a PERSISTENT buffer (= checkpoint key `patch_embed.pos_embedding`) only when learned
            pos = self._get_positional_embeddings(sample_height, sample_width, sample_frames)
            self.register_buffer("pos_embedding", pos, persistent=use_learned_positional_embeddings)

    def _get_positional_embeddings(self, sample_height, sample_width, sample_frames, device=None):
        ph, pw = sample_height // self.patch_size, sample_width // self.patch_size
        pf = (sample_frames - 1) // self.temporal_compression_ratio + 1
        pos = get_3d_sincos_pos_embed(self.embed_dim, (pw, ph), pf, self.spatial_interpolation_scale,
                                      self.temporal_interpolation_scale, device=device).flatten(0, 1)
        joint = pos.new_zeros(1, self.max_text_seq_length + ph * pw * pf, self.embed_dim)
        joint[:, self.max_text_seq_length:].copy_(pos)
        return joint

    def forward(self, text_embeds, image_embeds):
        text_embeds = self.text_proj(text_embeds)
        b, f, c, h, w = image_embeds.shape
        x = image_embeds.reshape(-1, c, h, w)
        x = self.proj(x)
        x = x.view(b, f, *x.shape[1:])
        x = x.flatten(3).transpose(2, 3)  # [b, f, h*w, D]
        x = x.flatten(1, 2)  # [b, f*h*w, D]
        embeds = torch.cat([text_embeds, x], dim=1).contiguous()
        if self.use_positional_embeddings or self.use_learned_positional_embeddings:
            if self.use_learned_positional_embeddings and (self.sample_width != w or self.sample_height != h):
                raise ValueError("It is currently not possible to generate videos at a different resolution that the "
                                 "defaults. This should only be the case with 'THUDM/CogVideoX-5b-I2V'.")
            pre_frames = (f - 1) * self.temporal_compression_ratio + 1
            if self.sample_height != h or self.sample_width != w or self.sample_frames != pre_frames:
                pos = self._get_positional_embeddings(h, w, pre_frames, device=embeds.device)
            else:
                pos = self.pos_embedding
            embeds = embeds + pos.to(dtype=embeds.dtype)
        return embeds


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    omega = torch.arange(embed_dim // 2, device=pos.device, dtype=torch.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = torch.outer(pos.reshape(-1), omega)
    return torch.concat([torch.sin(out), torch.cos(out)], dim=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    emb_h = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    emb_w = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return torch.concat([emb_h, emb_w], dim=1)


def get_3d_sincos_pos_embed(embed_dim, spatial_size, temporal_size, spatial_interpolation_scale=1.0,
                            temporal_interpolation_scale=1.0, device=None):
    """Published diffusers algorithm: [T, H*W, D] = [temporal D/4 | spatial 3D/4 (first half from the w-grid, second
    from the h-grid: meshgrid(grid_w, grid_h, indexing='xy'))], each half [sin | cos]."""
    assert embed_dim % 4 == 0
    if isinstance(spatial_size, int):
        spatial_size = (spatial_size, spatial_size)
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    grid_h = torch.arange(spatial_size[1], device=device, dtype=torch.float32) / spatial_interpolation_scale
    grid_w = torch.arange(spatial_size[0], device=device, dtype=torch.float32) / spatial_interpolation_scale
    grid = torch.stack(torch.meshgrid(grid_w, grid_h, indexing="xy"), dim=0)
    grid = grid.reshape([2, 1, spatial_size[1], spatial_size[0]])
    pos_sp = get_2d_sincos_pos_embed_from_grid(d_sp, grid)
    grid_t = torch.arange(temporal_size, device=device, dtype=torch.float32) / temporal_interpolation_scale
    pos_t = get_1d_sincos_pos_embed_from_grid(d_t, grid_t)
    pos_sp = pos_sp[None].repeat_interleave(temporal_size, dim=0)
    pos_t = pos_t[:, None].repeat_interleave(spatial_size[0] * spatial_size[1], dim=1)
    return torch.concat([pos_t, pos_sp], dim=-1).float()


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
