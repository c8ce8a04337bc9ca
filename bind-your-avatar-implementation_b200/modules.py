"""Host-side mirrors of the reference's sub-modules: same constructor arguments, same parameter names (the
checkpoint contract) and same call signatures, so `face_modules.pt` / `router_modules.pt` /
`audio_modules.pt` / the base safetensors load unchanged.

These classes are parameter containers.  The per-step hot path never runs their PyTorch math: the step engine
(`engine.py`) packs their weights once and drives the sm_100a kernels.  The only PyTorch compute kept here is the
per-generation, timestep-invariant prologue (`LocalFacialExtractor`, `AudioProjModel`),
which runs once per video on the GPU through torch.

Reference: models/router.py (LocalFacialExtractor :78-193, PerceiverCrossAttention :196-275, MultiIPRouter :280-423,
SpatialTemporalAttentionBlock :425-493), models/audio_model.py (AudioProjModel :43-114, AudioAwareModel :130-261),
and the diffusers leaves named in SURVEY.md §8c.
"""
from __future__ import annotations


import torch
import torch.nn.functional as F
from torch import nn


# --------------------------------------------------------------------------------------- diffusers-shaped leaves
class Attention(nn.Module):
    """Parameter layout of diffusers `Attention` (to_q/to_k/to_v/to_out.0 [+ norm_q/norm_k])."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False, qk_norm=None, eps=1e-5,
                 out_bias=True):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.is_cross_attention = cross_attention_dim is not None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])
        if qk_norm == "layer_norm":
            self.norm_q = nn.LayerNorm(dim_head, eps=eps)
            self.norm_k = nn.LayerNorm(dim_head, eps=eps)
        else:
            self.norm_q = self.norm_k = None
        self.processor = None  # the reference's plug-in point (transformer.py:541-573); see set_attn_processor

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor


class _GELUProj(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)


class FeedForward(nn.Module):
    """diffusers FeedForward(activation_fn='gelu-approximate', final_dropout=True): net.0.proj, net.2"""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim), nn.Dropout(0.0)])


class LayerNormZero(nn.Module):
    def __init__(self, cond_dim, dim, affine=True, eps=1e-5):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=affine)


class AdaLayerNorm(nn.Module):
    def __init__(self, cond_dim, out_dim, affine=True, eps=1e-5):
        super().__init__()
        self.linear = nn.Linear(cond_dim, out_dim)
        self.norm = nn.LayerNorm(out_dim // 2, eps=eps, elementwise_affine=affine)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)


def sincos_table_1d(dim, pos):
    """[len(pos), dim] = [sin(pos * w) | cos(pos * w)], w_k = 10000^(-k / (dim/2)) evaluated in float64 (the published
    diffusers `get_1d_sincos_pos_embed_from_grid`)."""
    w = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
    a = torch.outer(pos.reshape(-1), w)
    return torch.cat([a.sin(), a.cos()], dim=1)


def sincos_positional_embedding(dim, frames, grid_h, grid_w, text_len, spatial_scale=1.875, temporal_scale=1.0):
    """CogVideoX's joint positional table [1, text_len + frames*grid_h*grid_w, dim] (diffusers `CogVideoXPatchEmbed.
    _get_positional_embeddings` / `get_3d_sincos_pos_embed`, restated from the published algorithm): text rows are
    zero; a video row (f, h, w) is [temporal dim/4 | w-axis 3dim/8 | h-axis 3dim/8], each [sin | cos]."""
    d_sp, d_t = 3 * dim // 4, dim // 4
    ph = torch.arange(grid_h, dtype=torch.float32) / spatial_scale
    pw = torch.arange(grid_w, dtype=torch.float32) / spatial_scale
    ew = sincos_table_1d(d_sp // 2, pw)[None, :, :].expand(grid_h, -1, -1)
    eh = sincos_table_1d(d_sp // 2, ph)[:, None, :].expand(-1, grid_w, -1)
    sp = torch.cat([ew, eh], dim=-1).reshape(1, grid_h * grid_w, d_sp).expand(frames, -1, -1)
    et = sincos_table_1d(d_t, torch.arange(frames, dtype=torch.float32) / temporal_scale)
    et = et[:, None, :].expand(-1, grid_h * grid_w, -1)
    video = torch.cat([et, sp], dim=-1).reshape(frames * grid_h * grid_w, dim).float()
    out = torch.zeros(1, text_len + video.shape[0], dim)
    out[0, text_len:] = video
    return out


class PatchEmbed(nn.Module):
    """Parameter layout of diffusers `CogVideoXPatchEmbed`: proj (Conv2d k=s=patch), text_proj, and — for the
    non-RoPE (`use_positional_embeddings`) and the CogVideoX-5B-I2V (`use_learned_positional_embeddings`)
    configurations — the joint positional table `pos_embedding`, a checkpoint key only when learned
    (models/transformer.py:370-392)."""

    def __init__(self, patch, in_ch, dim, text_dim, sample_width=90, sample_height=60, sample_frames=49,
                 temporal_compression_ratio=4, max_text_seq_length=226, spatial_interpolation_scale=1.875,
                 temporal_interpolation_scale=1.0, use_positional_embeddings=False, use_learned_positional_embeddings=False):
        super().__init__()
        self.patch_size, self.embed_dim = patch, dim
        self.proj = nn.Conv2d(in_ch, dim, kernel_size=(patch, patch), stride=patch)
        self.text_proj = nn.Linear(text_dim, dim)
        self.sample_width, self.sample_height, self.sample_frames = sample_width, sample_height, sample_frames
        self.temporal_compression_ratio, self.max_text_seq_length = temporal_compression_ratio, max_text_seq_length
        self.spatial_interpolation_scale = spatial_interpolation_scale
        self.temporal_interpolation_scale = temporal_interpolation_scale
        self.use_positional_embeddings = use_positional_embeddings
        self.use_learned_positional_embeddings = use_learned_positional_embeddings
        if use_positional_embeddings or use_learned_positional_embeddings:
            self.register_buffer("pos_embedding", self.positional_table(sample_height, sample_width, sample_frames),
                                 persistent=use_learned_positional_embeddings)

    def positional_table(self, height, width, pixel_frames):
        frames = (pixel_frames - 1) // self.temporal_compression_ratio + 1
        return sincos_positional_embedding(self.embed_dim, frames, height // self.patch_size, width // self.patch_size,
                                           self.max_text_seq_length, self.spatial_interpolation_scale,
                                           self.temporal_interpolation_scale)

    def table_for(self, latent_frames, height, width):
        """The table the reference adds for an input of this geometry, or None (RoPE-only configuration).  The learned
        table cannot change resolution (diffusers raises the same ValueError)."""
        if not (self.use_positional_embeddings or self.use_learned_positional_embeddings):
            return None
        if self.use_learned_positional_embeddings and (self.sample_width != width or self.sample_height != height):
            raise ValueError("It is currently not possible to generate videos at a different resolution that the defaults. "
                             "This should only be the case with 'THUDM/CogVideoX-5b-I2V'.")
        pixel_frames = (latent_frames - 1) * self.temporal_compression_ratio + 1
        if (self.sample_height, self.sample_width, self.sample_frames) != (height, width, pixel_frames):
            return self.positional_table(height, width, pixel_frames).to(self.pos_embedding.device, self.pos_embedding.dtype)
        return self.pos_embedding


class CogVideoXBlock(nn.Module):
    """models/transformer.py:143-262 (parameters only; computed by engine.StepEngine.dit_block)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, time_embed_dim, norm_elementwise_affine=True,
                 norm_eps=1e-5, attention_bias=True):
        super().__init__()
        self.norm1 = LayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps)
        self.attn1 = Attention(dim, heads=num_attention_heads, dim_head=attention_head_dim, bias=attention_bias,
                               qk_norm="layer_norm", eps=1e-6, out_bias=True)
        self.norm2 = LayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps)
        self.ff = FeedForward(dim)


# --------------------------------------------------------------------------------------- face branch
def _split_heads(x, h):
    b, n, _ = x.shape
    return x.view(b, n, h, -1).transpose(1, 2)


class _PerceiverAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, latents):
        x, latents = self.norm1(x), self.norm2(latents)
        q = _split_heads(self.to_q(latents), self.heads)
        k, v = self.to_kv(torch.cat((x, latents), dim=-2)).chunk(2, dim=-1)
        k, v = _split_heads(k, self.heads), _split_heads(v, self.heads)
        s = self.dim_head ** -0.25
        w = torch.softmax(((q * s) @ (k * s).transpose(-2, -1)).float(), dim=-1).to(q.dtype)
        return self.to_out((w @ v).transpose(1, 2).flatten(2))


def _lfe_ffn(dim, mult):
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim * mult, bias=False), nn.GELU(),
                         nn.Linear(dim * mult, dim, bias=False))


def _lfe_mapping(in_dim, out_dim):
    return nn.Sequential(nn.Linear(in_dim, 1024), nn.LayerNorm(1024), nn.LeakyReLU(), nn.Linear(1024, 1024),
                         nn.LayerNorm(1024), nn.LeakyReLU(), nn.Linear(1024, out_dim))


class LocalFacialExtractor(nn.Module):
    """Per-generation prologue (timestep-invariant): id embedding + 5 ViT feature maps -> 32 face tokens of 2048.
    models/router.py:78-193."""

    def __init__(self, dim=1024, depth=10, dim_head=64, heads=16, num_id_token=5, num_queries=32, output_dim=2048,
                 ff_mult=4):
        super().__init__()
        self.num_id_token, self.dim, self.num_queries = num_id_token, dim, num_queries
        self.depth = depth // 5
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) * dim ** -0.5)
        self.proj_out = nn.Parameter(dim ** -0.5 * torch.randn(dim, output_dim))
        self.layers = nn.ModuleList(
            [nn.ModuleList([_PerceiverAttention(dim, dim_head, heads), _lfe_ffn(dim, ff_mult)]) for _ in range(depth)])
        for i in range(5):
            setattr(self, f"mapping_{i}", _lfe_mapping(1024, dim))
        self.id_embedding_mapping = _lfe_mapping(1280, dim * num_id_token)

    def forward(self, x, y):
        lat = self.latents.repeat(x.size(0), 1, 1)
        x = self.id_embedding_mapping(x).reshape(-1, self.num_id_token, self.dim)
        lat = torch.cat((lat, x), dim=1)
        for i in range(5):
            ctx = torch.cat((x, getattr(self, f"mapping_{i}")(y[i])), dim=1)
            for attn, ff in self.layers[i * self.depth:(i + 1) * self.depth]:
                lat = attn(ctx, lat) + lat
                lat = ff(lat) + lat
        return lat[:, :self.num_queries] @ self.proj_out


class PerceiverCrossAttention(nn.Module):
    """models/router.py:196-275.  Parameters: norm1 (kv_dim), norm2 (dim), to_q, to_kv, to_out (no biases)."""

    def __init__(self, *, dim=3072, dim_head=128, heads=16, kv_dim=2048):
        super().__init__()
        self.dim, self.dim_head, self.heads, self.kv_dim = dim, dim_head, heads, kv_dim
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(kv_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(kv_dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        self.return_weight_out = False   # forward(): also return the pre-softmax logits (dead in the reference flow)

    @torch.no_grad()
    def face_kv(self, x):
        """x [C,32,kv_dim] -> (k, v) each [C,heads,32,dim_head]; per-generation constant."""
        k, v = self.to_kv(self.norm1(x)).chunk(2, dim=-1)
        return _split_heads(k, self.heads).contiguous(), _split_heads(v, self.heads).contiguous()

    @torch.no_grad()
    def forward(self, x, latents):
        """Reference signature: (face tokens [C,32,kv_dim], latents [C,Nv,dim]) -> (out [C,Nv,dim], weight_out,
        q_out [C,heads,Nv,dh], k_out [C,heads,32,dh]).  `weight_out` (the pre-softmax logits [C,heads,Nv,32],
        router.py:262-263) is dead in the reference flow (MultiIPRouter.forward ignores it, router.py:364): it is None unless
        `self.return_weight_out` is set, in which case it is computed the way the router's score GEMM is (one dense GEMM
        against the block-structured keys)."""
        from . import ops

        C, Nv, _ = latents.shape
        k, v = self.face_kv(x)
        vt = v.transpose(-1, -2).contiguous()
        out = torch.empty_like(latents)
        q_all = []
        for c in range(C):
            lat = latents[c].contiguous()
            xn = torch.empty_like(lat)
            ops.layernorm_modulate(lat, xn, eps=self.norm2.eps, gamma=self.norm2.weight, beta=self.norm2.bias)
            q = torch.empty(Nv, self.heads * self.dim_head, device=lat.device, dtype=lat.dtype)
            ops.gemm(xn, self.to_q.weight, q)
            a = torch.empty_like(q)
            ops.xattn_kv32(q, k[c:c + 1].contiguous(), vt[c:c + 1].contiguous(), None, a, self.heads, self.dim_head, 1, 1,
                           self.dim_head ** -0.5)
            ops.gemm(a, self.to_out.weight, out[c])
            q_all.append(q.view(Nv, self.heads, self.dim_head).transpose(0, 1))
        w_out = None
        if self.return_weight_out:
            H, dh = self.heads, self.dim_head
            w_out = torch.empty(C, H, Nv, 32, device=latents.device, dtype=latents.dtype)
            for c in range(C):
                kf = k[c].transpose(0, 1).reshape(32, H * dh).contiguous()          # [token, head-major features]
                mat = torch.empty(32 * H, H * dh, device=kf.device, dtype=kf.dtype)
                ops.router_keys_scatter(kf, mat, 1, H, dh)                           # row (token * H + h): key of head h
                q_nat = q_all[c].transpose(0, 1).reshape(Nv, H * dh).contiguous()
                sc = torch.empty(Nv, 32 * H, device=kf.device, dtype=kf.dtype)
                ops.gemm(q_nat, mat, sc)
                w_out[c] = (sc.float() * (dh ** -0.5)).view(Nv, 32, H).permute(2, 0, 1).to(w_out.dtype)
        return out, w_out, torch.stack(q_all), k


# --------------------------------------------------------------------------------------- router
class SpatialTemporalAttentionBlock(nn.Module):
    """models/router.py:425-493 (parameters only)."""

    def __init__(self, dim, num_heads=8, mlp_ratio=4):
        super().__init__()
        kw = dict(query_dim=dim, heads=num_heads, dim_head=dim // num_heads, bias=True)
        self.spatial_attn = Attention(**kw)
        self.temporal_attn = Attention(**kw)
        self.multi_id_attn = Attention(**kw)
        self.norm1, self.norm2, self.norm3, self.norm4 = (nn.LayerNorm(dim) for _ in range(4))
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim))


class MultiIPRouter(nn.Module):
    """models/router.py:280-423.  The dead `layer_merge` sub-modules are kept so router_modules.pt loads strictly."""

    def __init__(self, *, num_id_token=32, num_heads=16, inner_dim1=256, inner_dim2=128, inner_dim3=32, addtional_dim=3,
                 num_layers=21, q_k_dim=2048, frames=13, height=45, width=30):
        super().__init__()
        weight_dim = num_id_token * num_heads
        self.heads = num_heads
        self.norm = nn.LayerNorm(weight_dim)
        self.norm_q = nn.LayerNorm(q_k_dim)
        self.norm_k = nn.LayerNorm(q_k_dim)
        self.to_q = nn.ModuleList([nn.Linear(q_k_dim, q_k_dim, bias=False) for _ in range(num_layers)])
        self.to_k = nn.ModuleList([nn.Linear(q_k_dim, q_k_dim, bias=False) for _ in range(num_layers)])
        self.layer_merge = nn.ModuleList([nn.Sequential(nn.Linear(weight_dim + addtional_dim, inner_dim1, bias=True), nn.ReLU(),
                                                        nn.Linear(inner_dim1, inner_dim2, bias=False), nn.ReLU())
                                          for _ in range(num_layers)])
        # the reference names the (f, h, w) token grid (frames, "height"=grid_w, "width"=grid_h): router.py:312-314
        self.frames, self.height, self.width = frames, height, width
        self.feat_dim = weight_dim
        self.register_buffer("pos_emb", self._create_positional_embedding())
        self.spatial_temporal_layers = nn.ModuleList(
            [SpatialTemporalAttentionBlock(dim=self.feat_dim, num_heads=8, mlp_ratio=1) for _ in range(4)])
        self.final_proj = nn.Sequential(nn.Linear(self.feat_dim, 1), nn.Sigmoid())

    def _create_positional_embedding(self):
        d3 = self.feat_dim // 3
        div = torch.pow(10000, torch.arange(0, d3, 2).float() / d3)

        def axis(n):
            a = torch.arange(n).float()[:, None] / div
            return torch.stack([a.sin(), a.cos()], dim=-1).flatten(-2)

        T, H, W = self.frames, self.height, self.width
        t = axis(T)[:, None, None].expand(-1, H, W, -1)
        h = axis(H)[None, :, None].expand(T, -1, W, -1)
        w = axis(W)[None, None, :].expand(T, H, -1, -1)
        pe = torch.cat([t, h, w], dim=-1)
        pad = self.feat_dim - pe.shape[-1]
        return torch.cat([pe, torch.zeros(T, H, W, pad)], dim=-1) if pad else pe

    def set_grid(self, frames, grid_h, grid_w):
        """Re-derive the positional buffer for a non-default latent grid (height:=grid_w, width:=grid_h)."""
        if (self.frames, self.height, self.width) != (frames, grid_w, grid_h):
            self.frames, self.height, self.width = frames, grid_w, grid_h
            self.pos_emb = self._create_positional_embedding().to(self.pos_emb.device, self.pos_emb.dtype)

    @torch.no_grad()
    def router_keys(self, k_out, layer_idx):
        """k_out [C,16,32,128] -> block-structured score matrix [C*512, 2048]: row (c, tok*16+h) holds the routed key
        of head h in columns h*128.. (router.py:377-393 K side; per-generation constant)."""
        C = k_out.shape[0]
        k = k_out.permute(0, 2, 3, 1).flatten(2)
        k = self.to_k[layer_idx](self.norm_k(k))  # [C,32,2048]
        k = k.view(C, 32, self.heads, -1)  # [C,tok,h,128]
        dh = k.shape[-1]
        mat = torch.zeros(C, 32, self.heads, self.heads, dh, device=k.device, dtype=k.dtype)
        idx = torch.arange(self.heads, device=k.device)
        mat[:, :, idx, idx] = k
        return mat.view(C * 32 * self.heads, self.heads * dh).contiguous()

    def save(self, path: str):
        torch.save(self.state_dict(), path)

    def load(self, path: str, strict: bool = True):
        sd = torch.load(path, map_location=next(self.parameters()).device)
        missing, unexpected = self.load_state_dict(sd, strict=strict)
        print(f"Router Missing keys: {missing}")
        print(f"Router Unexpected keys: {unexpected}")
        return missing, unexpected

    @torch.no_grad()
    def forward(self, weight, q_out, k_out, layer_idx, is_teacher_forcing=False):
        """Reference signature (router.py:364): q_out [C,16,Nv,128], k_out [C,16,32,128] -> [1,Nv,C] soft routing.
        `weight` and `is_teacher_forcing` are dead arguments there too."""
        from .engine import router_forward_standalone

        return router_forward_standalone(self, q_out, k_out, layer_idx)


# --------------------------------------------------------------------------------------- audio branch
class AudioProjModel(nn.Module):
    """Per-generation prologue: windows of 5 audio frames -> 32 context tokens per latent frame.
    models/audio_model.py:43-114."""

    def __init__(self, seq_len=5, blocks=12, channels=768, intermediate_dim=512, output_dim=768, context_tokens=32):
        super().__init__()
        self.context_tokens, self.output_dim = context_tokens, output_dim
        self.proj1 = nn.Linear(seq_len * blocks * channels, intermediate_dim)
        self.proj2 = nn.Linear(intermediate_dim, intermediate_dim)
        self.proj3 = nn.Linear(intermediate_dim, context_tokens * output_dim)
        self.norm = nn.LayerNorm(output_dim)
        self.conv1 = nn.Conv1d(context_tokens * output_dim, context_tokens * output_dim, kernel_size=2, stride=2)

    def _conv_weight_as_matrix(self):
        """conv1 (k=2, s=2) as a [C_out, 2*C_in] matrix over time-major pairs: W2[o, k*C_in + c] = W[o, c, k].
        Built once per weight version (cuDNN's per-call NCHW->NHWC transform of this 1.2 B-element weight cost
        ~10 ms per call; as a skinny GEMM the conv is a single pass over the weight)."""
        w = self.conv1.weight
        key = (w.data_ptr(), w._version, w.dtype, w.device)
        if getattr(self, "_w2_key", None) != key:
            self._w2 = w.detach().permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()
            self._w2_key = key
        return self._w2

    def _halve(self, x):
        """x [R, L, C] time-major -> [R, L/2, C]: Conv1d(k=2, s=2) over the time axis (audio_model.py:98-109)."""
        R, L, C = x.shape
        y = F.linear(x.reshape(R * (L // 2), 2 * C), self._conv_weight_as_matrix(), self.conv1.bias)
        return y.reshape(R, L // 2, C)

    def forward(self, audio_embeds):
        R, L = audio_embeds.shape[:2]
        x = audio_embeds.reshape(R * L, -1)
        x = torch.relu(self.proj1(x))
        x = torch.relu(self.proj2(x))
        x = self.proj3(x).reshape(R, L, -1)
        for _ in range(2):  # 49 -> 25 -> 13: keep frame 0, halve the rest with the k=2,s=2 conv
            if x.shape[1] % 2 == 1:
                x = torch.cat([x[:, :1], self._halve(x[:, 1:].contiguous())], dim=1) if x.shape[1] > 1 else x
            else:
                x = self._halve(x)
        x = x.reshape(R, x.shape[1], self.context_tokens, self.output_dim)
        return self.norm(x)


class AudioAwareModel(nn.Module):
    """models/audio_model.py:130-261."""

    def __init__(self, dim=3072, audio_dim=768, num_attention_heads=48, attention_head_dim=64, window_size=5,
                 window_stride=1, norm_elementwise_affine=True, norm_eps=1e-5, num_layers=42, audio_cross_attn_scale=0.05):
        super().__init__()
        self.dim, self.window_size, self.window_stride, self.num_layers = dim, window_size, window_stride, num_layers
        self.heads, self.head_dim = num_attention_heads, attention_head_dim
        self.learnable_scale = nn.Parameter(torch.tensor([0.01]))
        self.audio_cross_attn_scale = audio_cross_attn_scale
        self.audio_proj_model = AudioProjModel()
        self.layers = nn.ModuleList([
            nn.ModuleDict({
                "norm_q": nn.LayerNorm(dim, norm_eps, norm_elementwise_affine),
                "attn": Attention(query_dim=dim, cross_attention_dim=audio_dim, dim_head=attention_head_dim,
                                  heads=num_attention_heads, bias=True),
            }) for _ in range(num_layers)])
        self.mute_learnable_tokens = nn.Parameter(torch.zeros(1, 32, 768))

    def sliding_windows(self, audio_embeds, hidden_states_num_frames):
        want = 1 + (hidden_states_num_frames - 1) * 4 + (self.window_size - self.window_stride)
        assert want == audio_embeds.shape[1], (
            f"hidden_states_num_frames: {hidden_states_num_frames}, window_size: {self.window_size}, "
            f"window_stride: {self.window_stride}, audio_embeds.shape[1]: {audio_embeds.shape[1]}")
        return audio_embeds.unfold(1, self.window_size, self.window_stride).permute(0, 1, 4, 2, 3)

    def proj_in(self, audio_embeds):
        return self.audio_proj_model(audio_embeds)

    @torch.no_grad()
    def audio_kv(self, ctx, layer_index):
        """ctx [C,F,32,768] -> K [C*F,heads,32,hd], Vt [C*F,heads,hd,32]; per-generation constant."""
        attn = self.layers[layer_index]["attn"]
        a = ctx.reshape(-1, 32, ctx.shape[-1])
        k = _split_heads(attn.to_k(a), self.heads).contiguous()
        v = _split_heads(attn.to_v(a), self.heads)
        return k, v.transpose(-1, -2).contiguous()

    @torch.no_grad()
    def forward(self, audio_embeds, hidden_states, num_frames, layer_index, mask=None):
        """Reference signature (audio_model.py:224): audio ctx [C,F,32,768], hidden [C,Nv,D] -> [C,Nv,D]."""
        from . import ops

        layer = self.layers[layer_index]
        attn = layer["attn"]
        C, Nv, D = hidden_states.shape
        out = torch.empty_like(hidden_states)
        for c in range(C):
            k, vt = self.audio_kv(audio_embeds[c:c + 1], layer_index)
            h = hidden_states[c].contiguous()
            xn = torch.empty_like(h)
            ops.layernorm_modulate(h, xn, eps=layer["norm_q"].eps, gamma=layer["norm_q"].weight, beta=layer["norm_q"].bias)
            q = torch.empty_like(h)
            ops.gemm(xn, attn.to_q.weight, q, bias=attn.to_q.bias)
            a = torch.empty_like(h)
            ops.xattn_kv32(q, k, vt, None, a, self.heads, self.head_dim, 1, num_frames, self.head_dim ** -0.5)
            ops.gemm(a, attn.to_out[0].weight, out[c], bias=attn.to_out[0].bias)
        return out
