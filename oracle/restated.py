"""Own-words restatement (plain torch, any device/dtype) of the reference's denoising step.

TEST INFRASTRUCTURE — never imported by the product package.  parity unpinned by the reference (it ships no golden
vectors); this file is pinned instead against the reference's own forward executed verbatim
(`oracle/reference_harness.py`, `tests/test_oracle_vs_reference.py`) and against `tests/golden/*.pt`, which that
verbatim forward produced (`oracle/make_goldens.py`).

Follows, function by function:
  step()                  /root/reference/models/transformer.py:615-964  (inference branch)
  dit_block()             models/transformer.py:223-262 + diffusers CogVideoXLayerNormZero / CogVideoXAttnProcessor2_0 / FeedForward
  face_cross_attention()  models/router.py:230-275
  router()                models/router.py:364-411, st_block() :468-493
  audio_layer()           models/audio_model.py:224-261
  audio_context()         models/audio_model.py:188-193, :78-114
  facial_extractor()      models/router.py:157-193 (+ PerceiverAttention :46-75)
  routing_from_masks()    util/utils.py:481-514, :871-936 ; frame-OR transformer.py:813-819
Generalisations beyond the reference (C characters, F frames, any grid) reduce to it exactly at C=2, F=13
(SURVEY.md §8c): audio weight for C>2 is w_c = 1 - max_{c'!=c} av_{c'}.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd.get(name + ".weight"), sd.get(name + ".bias"), eps)


def _heads(x, h):  # [b, n, h*d] -> [b, h, n, d]
    b, n, _ = x.shape
    return x.view(b, n, h, -1).transpose(1, 2)


def _sdpa(q, k, v):
    """softmax(q k^T / sqrt(d)) v.  Small problems go through torch's SDPA; large fp32 ones (the full 17 776-token grid
    on the GPU) are evaluated explicitly a few heads at a time, so the truth never depends on which fused backend
    torch would pick for fp32 and never materialises more than ~6 GB of scores."""
    n_q, n_k = q.shape[-2], k.shape[-2]
    if q.dtype != torch.float32 or not q.is_cuda or q.shape[:-2].numel() * n_q * n_k * 4 <= (2 << 30):
        return F.scaled_dot_product_attention(q, k, v)
    lead = q.shape[:-2]
    q2, k2, v2 = (t.reshape(-1, t.shape[-2], t.shape[-1]) for t in (q.expand(*lead, -1, -1), k.expand(*lead, -1, -1), v.expand(*lead, -1, -1)))
    out = torch.empty(q2.shape[0], n_q, v2.shape[-1], device=q.device, dtype=q.dtype)
    step = max(1, int((6 << 30) // (n_q * n_k * 4)))
    scale = q.shape[-1] ** -0.5
    for i in range(0, q2.shape[0], step):
        w = torch.bmm(q2[i:i + step] * scale, k2[i:i + step].transpose(1, 2))
        out[i:i + step] = torch.softmax(w, dim=-1) @ v2[i:i + step]
        del w
    return out.reshape(*lead, n_q, v2.shape[-1])


def rope_rotate(x, cos, sin):
    """Interleaved-pair rotation in fp32 (diffusers apply_rotary_emb, use_real_unbind_dim=-1)."""
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(-2)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


# ----------------------------------------------------------------------------- embeddings / head
def time_embedding(sd, timestep, dim, dtype):
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timestep.device) / half)
    ang = timestep[:, None].float() * freq[None]
    t = torch.cat([ang.cos(), ang.sin()], -1).to(dtype)  # flip_sin_to_cos=True
    return _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t)))


def patch_embed(sd, text, latents, p):
    b, f, c, h, w = latents.shape
    x = F.conv2d(latents.reshape(-1, c, h, w), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)
    x = x.view(b, f, x.shape[1], -1).transpose(2, 3).flatten(1, 2)  # [b, f*gh*gw, D], (f,h,w) row-major
    return _lin(sd, "patch_embed.text_proj", text), x


def positional_table(sd, cfg, frames, gh, gw):
    """The joint table diffusers' CogVideoXPatchEmbed adds after the patch embedding (reference ctor
    models/transformer.py:370-392): the checkpoint buffer when learned (CogVideoX-5B-I2V lineage), the analytic sincos
    table when the model is built without RoPE, nothing for the RoPE-only configuration."""
    if getattr(cfg, "use_learned_positional_embeddings", False):
        return sd["patch_embed.pos_embedding"]
    if getattr(cfg, "use_rotary_positional_embeddings", True):
        return None
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "_bya_shim_embeddings", os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_shim", "diffusers",
                                             "models", "embeddings.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pos = mod.get_3d_sincos_pos_embed(cfg.dim, (gw, gh), frames, 1.875, 1.0).flatten(0, 1)
    return torch.cat([torch.zeros(cfg.text_len, cfg.dim), pos], 0)[None]


# ----------------------------------------------------------------------------- DiT block
def dit_block(sd, pre, h, e, temb, rope, heads):
    T = e.shape[1]

    def norm_zero(nm, h, e):
        mod = _lin(sd, f"{pre}.{nm}.linear", F.silu(temb))
        sh, sc, g, esh, esc, eg = mod.chunk(6, dim=1)
        nh = _ln(sd, f"{pre}.{nm}.norm", h) * (1 + sc)[:, None] + sh[:, None]
        ne = _ln(sd, f"{pre}.{nm}.norm", e) * (1 + esc)[:, None] + esh[:, None]
        return nh, ne, g[:, None], eg[:, None]

    nh, ne, g, eg = norm_zero("norm1", h, e)
    x = torch.cat([ne, nh], 1)
    a = f"{pre}.attn1"
    q, k, v = (_heads(_lin(sd, f"{a}.{n}", x), heads) for n in ("to_q", "to_k", "to_v"))
    q = F.layer_norm(q, (q.shape[-1],), sd[f"{a}.norm_q.weight"], sd[f"{a}.norm_q.bias"], 1e-6)
    k = F.layer_norm(k, (k.shape[-1],), sd[f"{a}.norm_k.weight"], sd[f"{a}.norm_k.bias"], 1e-6)
    if rope is not None:
        cos, sin = rope
        q = torch.cat([q[:, :, :T], rope_rotate(q[:, :, T:], cos, sin)], 2)
        k = torch.cat([k[:, :, :T], rope_rotate(k[:, :, T:], cos, sin)], 2)
    o = _sdpa(q, k, v).transpose(1, 2).flatten(2)
    o = _lin(sd, f"{a}.to_out.0", o)
    h = h + g * o[:, T:]
    e = e + eg * o[:, :T]
    nh, ne, g, eg = norm_zero("norm2", h, e)
    x = torch.cat([ne, nh], 1)
    x = F.gelu(_lin(sd, f"{pre}.ff.net.0.proj", x), approximate="tanh")
    x = _lin(sd, f"{pre}.ff.net.2", x)
    h = h + g * x[:, T:]
    e = e + eg * x[:, :T]
    return h, e


# ----------------------------------------------------------------------------- face cross-attention + router
def face_cross_attention(sd, pre, face, h1, heads=16):
    """face [C,32,2048]; h1 [1,Nv,D] (the reference repeats it C times: transformer.py:784).
    Returns per-character features [C,Nv,D] and the un-scaled q [1,16,Nv,128], k [C,16,32,128]."""
    C = face.shape[0]
    x = _ln(sd, f"{pre}.norm1", face)
    lat = _ln(sd, f"{pre}.norm2", h1)
    q = _heads(F.linear(lat, sd[f"{pre}.to_q.weight"]), heads)  # [1,16,Nv,128]
    k, v = F.linear(x, sd[f"{pre}.to_kv.weight"]).chunk(2, dim=-1)
    k, v = _heads(k, heads), _heads(v, heads)  # [C,16,32,128]
    s = 1.0 / math.sqrt(math.sqrt(q.shape[-1]))
    w = (q * s) @ (k * s).transpose(-2, -1)  # [C,16,Nv,32]
    w = torch.softmax(w.float(), dim=-1).to(w.dtype)
    o = (w @ v).permute(0, 2, 1, 3).flatten(2)  # [C,Nv,2048]
    return F.linear(o, sd[f"{pre}.to_out.weight"]), q, k


def _attn_block(sd, pre, x, heads=8):
    q, k, v = (_heads(_lin(sd, f"{pre}.{n}", x), heads) for n in ("to_q", "to_k", "to_v"))
    o = _sdpa(q, k, v).transpose(1, 2).flatten(2)
    return _lin(sd, f"{pre}.to_out.0", o)


def st_block(sd, pre, x):
    C, T, H, W, D = x.shape
    x = x + _attn_block(sd, f"{pre}.spatial_attn", _ln(sd, f"{pre}.norm1", x.reshape(C * T, H * W, D))).reshape(C, T, H, W, D)
    xt = _ln(sd, f"{pre}.norm2", x.permute(0, 2, 3, 1, 4).reshape(C * H * W, T, D))
    x = x + _attn_block(sd, f"{pre}.temporal_attn", xt).reshape(C, H, W, T, D).permute(0, 3, 1, 2, 4)
    xi = _ln(sd, f"{pre}.norm3", x.permute(2, 3, 1, 0, 4).reshape(H * W * T, C, D))
    x = x + _attn_block(sd, f"{pre}.multi_id_attn", xi).reshape(H, W, T, C, D).permute(3, 2, 0, 1, 4)
    y = _ln(sd, f"{pre}.norm4", x.reshape(-1, D))
    y = _lin(sd, f"{pre}.mlp.2", F.gelu(_lin(sd, f"{pre}.mlp.0", y)))
    return x + y.reshape(C, T, H, W, D)


def router(sd, q_out, k_out, layer, frames, grid_h, grid_w, heads=16):
    """q_out [1 or C,16,Nv,128], k_out [C,16,32,128] -> soft routing [1,Nv,C].
    The router views the (f, h, w) token order as (frames, "height"=grid_w, "width"=grid_h) — router.py:312-314,396."""
    C = k_out.shape[0]
    q = q_out.permute(0, 2, 3, 1).flatten(2)  # feature index d*16+h
    k = k_out.permute(0, 2, 3, 1).flatten(2)
    q = F.linear(_ln(sd, "router.norm_q", q), sd[f"router.to_q.{layer}.weight"])
    k = F.linear(_ln(sd, "router.norm_k", k), sd[f"router.to_k.{layer}.weight"])
    q, k = _heads(q, heads), _heads(k, heads)
    w = (q @ k.transpose(-2, -1)).permute(0, 2, 3, 1).flatten(2)  # [C,Nv,512], feature index tok*16+h
    w = _ln(sd, "router.norm", w)
    x = w.reshape(C, frames, grid_w, grid_h, -1) + sd["router.pos_emb"]
    for i in range(4):
        x = st_block(sd, f"router.spatial_temporal_layers.{i}", x)
    x = torch.sigmoid(_lin(sd, "router.final_proj.0", x.reshape(C, -1, x.shape[-1])))  # [C,Nv,1]
    return x.permute(2, 1, 0)


def router_pos_emb(frames, height, width, dim=512):
    """`MultiIPRouter._create_positional_embedding` (router.py:334-362)."""
    d3 = dim // 3
    div = torch.pow(10000, torch.arange(0, d3, 2).float() / d3)

    def axis(n):
        a = torch.arange(n).float()[:, None] / div
        return torch.stack([a.sin(), a.cos()], -1).flatten(-2)

    t = axis(frames)[:, None, None].expand(-1, height, width, -1)
    h = axis(height)[None, :, None].expand(frames, -1, width, -1)
    w = axis(width)[None, None, :].expand(frames, height, -1, -1)
    pe = torch.cat([t, h, w], -1)
    pad = dim - pe.shape[-1]
    return torch.cat([pe, torch.zeros(frames, height, width, pad)], -1) if pad else pe


# ----------------------------------------------------------------------------- audio
def audio_context(sd, audio, frames, pre="audio_model.audio_proj_model"):
    """audio [R, 4(F-1)+5, 12, 768] -> context tokens [R, F, 32, 768]."""
    assert audio.shape[1] == 1 + (frames - 1) * 4 + 4
    x = audio.unfold(1, 5, 1).permute(0, 1, 4, 2, 3)  # [R,49,5,12,768]
    R, L = x.shape[:2]
    x = x.reshape(R * L, -1)
    x = torch.relu(_lin(sd, f"{pre}.proj1", x))
    x = torch.relu(_lin(sd, f"{pre}.proj2", x))
    x = _lin(sd, f"{pre}.proj3", x).reshape(R, L, -1)  # [R,49,24576]
    cw, cb = sd[f"{pre}.conv1.weight"], sd[f"{pre}.conv1.bias"]
    for _ in range(2):
        x = x.permute(0, 2, 1)
        if x.shape[-1] % 2 == 1:
            first, rest = x[..., :1], x[..., 1:]
            if rest.shape[-1] > 0:
                rest = F.conv1d(rest, cw, cb, stride=2)
            x = torch.cat([first, rest], -1)
        else:
            x = F.conv1d(x, cw, cb, stride=2)
        x = x.permute(0, 2, 1)
    x = x.reshape(R, x.shape[1], 32, -1)
    return _ln(sd, f"{pre}.norm", x)


def audio_layer(sd, layer, ctx, h1, frames, heads=48):
    """ctx [C,F,32,768]; h1 [1,Nv,D] -> per-character audio features [C,Nv,D] (token n sees frame n // (Nv/F))."""
    pre = f"audio_model.layers.{layer}"
    C = ctx.shape[0]
    D = h1.shape[-1]
    hw = h1.shape[1] // frames
    x = _ln(sd, f"{pre}.norm_q", h1).reshape(frames, hw, D)
    q = _heads(_lin(sd, f"{pre}.attn.to_q", x), heads)  # [F,48,hw,64]
    a = ctx.reshape(C * frames, 32, -1)
    k = _heads(_lin(sd, f"{pre}.attn.to_k", a), heads).view(C, frames, heads, 32, -1)
    v = _heads(_lin(sd, f"{pre}.attn.to_v", a), heads).view(C, frames, heads, 32, -1)
    o = _sdpa(q[None].expand(C, -1, -1, -1, -1), k, v)  # [C,F,48,hw,64]
    o = o.transpose(2, 3).flatten(3)  # [C,F,hw,D]
    return _lin(sd, f"{pre}.attn.to_out.0", o).reshape(C, frames * hw, D)


def audio_weights(af, r):
    """af [C,C], r [Nv,C] (routing of the most recent cross-attention layer) -> [Nv,C]  (transformer.py:860-863,899-900)."""
    av = (af @ r.transpose(0, 1)).transpose(0, 1)  # [Nv,C]
    C = av.shape[1]
    if C == 2:
        return 1 - av[:, [1, 0]]
    w = []
    for c in range(C):
        others = torch.cat([av[:, :c], av[:, c + 1:]], 1)
        w.append(1 - others.max(dim=1).values)
    return torch.stack(w, 1)


# ----------------------------------------------------------------------------- LocalFacialExtractor
def _perceiver(sd, pre, x, lat, heads=16):
    x = _ln(sd, f"{pre}.norm1", x)
    lat = _ln(sd, f"{pre}.norm2", lat)
    q = _heads(F.linear(lat, sd[f"{pre}.to_q.weight"]), heads)
    k, v = F.linear(torch.cat([x, lat], -2), sd[f"{pre}.to_kv.weight"]).chunk(2, -1)
    k, v = _heads(k, heads), _heads(v, heads)
    s = 1.0 / math.sqrt(math.sqrt(q.shape[-1]))
    w = torch.softmax(((q * s) @ (k * s).transpose(-2, -1)).float(), -1).to(q.dtype)
    return F.linear((w @ v).permute(0, 2, 1, 3).flatten(2), sd[f"{pre}.to_out.weight"])


def facial_extractor(sd, id_cond, vit, pre="local_facial_extractor"):
    def mlp(nm, x):
        x = F.leaky_relu(_ln(sd, f"{nm}.1", _lin(sd, f"{nm}.0", x)))
        x = F.leaky_relu(_ln(sd, f"{nm}.4", _lin(sd, f"{nm}.3", x)))
        return _lin(sd, f"{nm}.6", x)

    B = id_cond.shape[0]
    lat = sd[f"{pre}.latents"].repeat(B, 1, 1)
    x = mlp(f"{pre}.id_embedding_mapping", id_cond).reshape(B, 5, -1)
    lat = torch.cat([lat, x], 1)
    for i in range(5):
        ctx = torch.cat([x, mlp(f"{pre}.mapping_{i}", vit[i])], 1)
        for j in (2 * i, 2 * i + 1):
            lat = _perceiver(sd, f"{pre}.layers.{j}.0", ctx, lat) + lat
            y = _ln(sd, f"{pre}.layers.{j}.1.0", lat)
            y = F.linear(F.gelu(F.linear(y, sd[f"{pre}.layers.{j}.1.1.weight"])), sd[f"{pre}.layers.{j}.1.3.weight"])
            lat = y + lat
    return lat[:, :32] @ sd[f"{pre}.proj_out"]


# ----------------------------------------------------------------------------- 3-D masks -> routing logits
def routing_from_masks(masks: torch.Tensor, frames: int, grid_h: int, grid_w: int, resize_dtype=torch.float32):
    """masks [C, T_px, H_px, W_px] (any dtype, >0 = inside) -> (index_mask int64 [1,Nv], logits fp32 [1,Nv,C]).
    util/utils.py:871-936: trilinear resize (align_corners=False), > 0.5, later character wins on overlap."""
    C = masks.shape[0]
    idx = torch.full((1, 1, frames, grid_h, grid_w), -1, dtype=torch.long)
    for c in range(C):
        m = (masks[c] > 0).to(resize_dtype)[None, None]
        r = F.interpolate(m, size=(frames, grid_h, grid_w), mode="trilinear", align_corners=False)
        idx = torch.where((r > 0.5).long() == 1, torch.tensor(c, dtype=torch.long), idx)
    idx = idx.reshape(1, -1)
    logits = torch.zeros(1, idx.shape[1], C)
    for c in range(C):
        logits[0, idx[0] == c, c] = 1
    return idx, logits


def frame_or(logits, frames, grid_h, grid_w):
    """transformer.py:815-818: OR over the frame axis, broadcast back to every frame."""
    C = logits.shape[-1]
    x = logits.view(1, frames, grid_h, grid_w, C)
    return x.max(dim=1).values.unsqueeze(1).repeat(1, frames, 1, 1, 1).reshape(1, -1, C)


# ----------------------------------------------------------------------------- the step
@torch.no_grad()
def step(sd: Dict[str, torch.Tensor], cfg, hidden_states, encoder_hidden_states, timestep, image_rotary_emb=None,
         id_cond=None, id_vit_hidden=None, audio_embeds=None, af_matrix=None, routing_logits_forcing=None,
         per_frame_forcing: bool = False, taps: Optional[dict] = None, **unused):
    """One denoising-step forward.  Returns the noise prediction [B,F,out_ch,H,W]; fills `taps` with sub-module
    boundary tensors when a dict is given."""
    assert id_cond is not None and id_vit_hidden is not None
    B, Fr, _, Hl, Wl = hidden_states.shape
    p, C, heads = cfg.patch_size, cfg.chars, cfg.num_attention_heads
    gh, gw = Hl // p, Wl // p
    dtype = hidden_states.dtype
    tap = taps if callable(taps) else (
        (lambda k, v: taps.__setitem__(k, v.detach().float().cpu())) if taps is not None else (lambda k, v: None))

    face = torch.stack([facial_extractor(sd, id_cond[c], id_vit_hidden[c]) for c in range(C)], 1)  # [B,C,32,2048]
    tap("face_tokens", face)
    actx = None
    if audio_embeds is not None:
        a = audio_embeds.to(dtype)
        actx = audio_context(sd, a.reshape(B * C, *a.shape[2:]), Fr).reshape(B, C, Fr, 32, -1)
        tap("audio_ctx", actx)

    temb = time_embedding(sd, timestep, cfg.dim, dtype)
    tap("temb", temb)
    e, h = patch_embed(sd, encoder_hidden_states, hidden_states, p)
    pos = positional_table(sd, cfg, Fr, gh, gw)
    if pos is not None:
        pos = pos.to(h.device, h.dtype)
        e, h = e + pos[:, :e.shape[1]], h + pos[:, e.shape[1]:]
    tap("embed_video", h)
    routing = [torch.zeros(1, h.shape[1], C, dtype=dtype, device=h.device) for _ in range(B)]
    ca = 0
    for i in range(cfg.num_layers):
        h, e = dit_block(sd, f"transformer_blocks.{i}", h, e, temb, image_rotary_emb, heads)
        tap(f"block{i}.video", h)
        tap(f"block{i}.text", e)
        if i % cfg.cross_attn_interval == 0:
            adds = []
            for b in range(B):
                feat, q_out, k_out = face_cross_attention(sd, f"perceiver_cross_attention.{ca}", face[b], h[b:b + 1])
                r = router(sd, q_out, k_out, ca, Fr, gh, gw)
                if b == 0:
                    tap(f"ca{ca}.router", r)
                    tap(f"ca{ca}.id_feat", feat)
                if routing_logits_forcing is not None:
                    r = routing_logits_forcing.to(dtype)
                    if not per_frame_forcing:
                        r = frame_or(r, Fr, gh, gw)
                routing[b] = r
                adds.append(torch.einsum("nc,cnd->nd", r[0], feat)[None])
            h = h + cfg.local_face_scale * torch.cat(adds)
            tap(f"ca{ca}.video", h)
            ca += 1
        if actx is not None and i % cfg.audio_attn_interval == 0:
            adds = []
            for b in range(B):
                w = audio_weights(af_matrix[b].to(dtype), routing[b][0].to(dtype))
                feat = audio_layer(sd, i // cfg.audio_attn_interval, actx[b], h[b:b + 1], Fr)
                if b == 0:
                    tap(f"audio{i}.weights", w)
                adds.append(torch.einsum("nc,cnd->nd", w, feat)[None])
            h = h + torch.cat(adds)
            tap(f"audio{i}.video", h)

    T = e.shape[1]
    x = _ln(sd, "norm_final", torch.cat([e, h], 1))[:, T:]
    mod = _lin(sd, "norm_out.linear", F.silu(temb))
    shift, scale = mod.chunk(2, dim=1)
    x = _ln(sd, "norm_out.norm", x) * (1 + scale)[:, None] + shift[:, None]
    x = _lin(sd, "proj_out", x)
    out = x.reshape(B, Fr, gh, gw, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    return out
