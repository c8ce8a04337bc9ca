"""Attention processors restated from the published diffusers algorithm."""
import torch
import torch.nn.functional as F


class AttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw):
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        hd = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(b, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(b, -1, attn.heads, hd).transpose(1, 2)
        if attn.norm_q is not None:
            q = attn.norm_q(q)
        if attn.norm_k is not None:
            k = attn.norm_k(k)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * hd).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o


AttentionProcessor = AttnProcessor2_0


class CogVideoXAttnProcessor2_0:
    """Joint [text; video] self-attention: qk LayerNorm, RoPE on the video rows only, SDPA, to_out."""

    def __call__(self, attn, hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None):
        from .embeddings import apply_rotary_emb

        t = encoder_hidden_states.size(1)
        x = torch.cat([encoder_hidden_states, hidden_states], dim=1)
        b = x.shape[0]
        q, k, v = attn.to_q(x), attn.to_k(x), attn.to_v(x)
        hd = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(b, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(b, -1, attn.heads, hd).transpose(1, 2)
        if attn.norm_q is not None:
            q = attn.norm_q(q)
        if attn.norm_k is not None:
            k = attn.norm_k(k)
        if image_rotary_emb is not None:
            q[:, :, t:] = apply_rotary_emb(q[:, :, t:], image_rotary_emb)
            if not attn.is_cross_attention:
                k[:, :, t:] = apply_rotary_emb(k[:, :, t:], image_rotary_emb)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * hd)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        e, h = o.split([t, o.size(1) - t], dim=1)
        return h, e


class FusedCogVideoXAttnProcessor2_0(CogVideoXAttnProcessor2_0):
    """Never installed by infer.py; kept only so the reference's import succeeds."""
